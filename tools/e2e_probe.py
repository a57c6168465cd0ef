"""Where the end-to-end tick (sapio_tick_host_begin / _end) spends its time next to the device-resident tick: PCIe download rate of the
frame from pinned memory, the pipelined loop with and without the sprite download. usage: python tools/e2e_probe.py [organisms] [age]"""
import sys, time, torch
sys.path.insert(0, '.')
from sapiogenesis_b200.synth import build_world
from sapiogenesis_b200.world import PH_ALL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
age = int(sys.argv[2]) if len(sys.argv) > 2 else 600
w, eng = build_world(n, seed=0); eng.exclusive = True
eng.tick(age); torch.cuda.synchronize()
def timeit(f, k=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
print("tick(1) device-resident      ms", round(timeit(lambda: eng.tick(1)), 3))
eng.tick_host(1)
hb = eng._host_buffers()
nb = hb["cells"].shape[0]
src = torch.zeros(nb, 4, device="cuda")
for mb in (8, 35, 50):
    k = mb * 1_000_000 // 16
    ms = timeit(lambda: hb["cells"][:k].copy_(src[:k], non_blocking=True))
    print(f"d2h {mb} MB pinned: {ms:.3f} ms = {mb / ms:.2f} GB/s")
def pipelined(steps=20):
    eng.tick_host_begin(1, PH_ALL)
    for _ in range(steps - 1):
        eng.tick_host_begin(1, PH_ALL); eng.tick_host_end()
    return eng.tick_host_end()
pipelined(3)
torch.cuda.synchronize(); t0 = time.perf_counter(); hdr = pipelined(20)[0]; torch.cuda.synchronize()
print("pipelined host tick          ms", round((time.perf_counter() - t0) / 20 * 1e3, 3), "cells", hdr["n_cells"], "dead", hdr["n_dead"])
# the same loop with the calls timed on the host
tb = te = 0.0
eng.tick_host_begin(1, PH_ALL)
for _ in range(19):
    a = time.perf_counter(); eng.tick_host_begin(1, PH_ALL); b = time.perf_counter(); eng.tick_host_end(); c = time.perf_counter()
    tb += b - a; te += c - b
eng.tick_host_end()
print(f"host time per tick: begin {tb / 19 * 1e3:.3f} ms, end (wait for the frame) {te / 19 * 1e3:.3f} ms")

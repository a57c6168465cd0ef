import sys, time, torch
sys.path.insert(0, '.')
from sapiogenesis_b200.synth import build_world
from sapiogenesis_b200.world import PH_ALL
w, eng = build_world(100000, seed=0)
for _ in range(10): eng.tick(1)
torch.cuda.synchronize()
def timeit(f, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("tick        ms", timeit(lambda: eng.tick(1)))
eng.tick_host(1)
print("tick_host   ms", timeit(lambda: eng.tick_host(1)))
hb = eng._host_buffers()
print("pinned?", hb["cells"].is_pinned(), hb["header"].is_pinned(), hb["control"].is_pinned())
src = torch.zeros(2_200_000, 4, device="cuda")
dst = hb["cells"][:2_200_000]
print("d2h 35MB    ms", timeit(lambda: dst.copy_(src, non_blocking=True)))
eng.profile(True)
print("tick_host+p ms", timeit(lambda: eng.tick_host(1)))
print({k: round(v[0] / max(v[1], 1), 3) for k, v in eng.profile_read().items() if v[1]})

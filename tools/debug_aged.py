"""Single-tick error anatomy of the aged dense world of tests/test_gpu_atsize.py: who is off, by how much, and why."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle import oracle as orc
from sapiogenesis_b200.synth import build_world
from sapiogenesis_b200.world import MAXC, PH_ALL

n = int(sys.argv[1]) if len(sys.argv) > 1 else 14000
age = int(sys.argv[2]) if len(sys.argv) > 2 else 3200
orc.build()
w, eng = build_world(n, seed=5, area_per_org=200_000.0, food_per_org=1.0)
eng.exclusive = True
eng.tick(age); eng.synchronize()
host = w.to("cpu"); h0 = host.clone()
# stage by stage
def posdiff(tag):
    for name in ("c_pos", "d_pos", "c_spos"):
        g = w.np(name); h = host.np(name)
        m = (host.np("c_alive") == 1) if name[0] == "c" else (host.np("d_state") > 0)
        nd = int((g[m] != h[m]).any(1).sum())
        print(f"{tag}: {name} rows that differ: {nd} of {int(m.sum())}")
def veldiff(tag):
    live = host.np("c_alive") == 1; dead = host.np("d_state") > 0
    for name, m in (("c_vel", live), ("d_vel", dead), ("c_angvel", live), ("d_angvel", dead), ("c_energy", live)):
        g = w.np(name).astype(np.float64); h = host.np(name).astype(np.float64)
        e = np.abs(g - h); e = e.max(1) if e.ndim == 2 else e
        e[~m] = 0
        rms = np.sqrt(np.mean(h[m] ** 2))
        print(f"{tag}: {name} rms {rms:.4g} max err {e.max():.3e} ({e.max() / rms:.2e} rms); rows > 1e-5 rms: {(e > 1e-5 * rms).sum()}, > 3e-5: {(e > 3e-5 * rms).sum()}, > 1e-4: {(e > 1e-4 * rms).sum()}")
    return
eng.rules(PH_ALL); eng.synchronize(); orc.rules(host, PH_ALL); veldiff("rules")
eng.train(); eng.synchronize(); orc.train(host)
eng.settle(); eng.synchronize(); orc.settle(host); veldiff("settle")
eng.reproduce(); eng.synchronize(); orc.reproduce(host); veldiff("reproduce")
hpre = host.clone()
eng.space_step(); eng.synchronize(); orc.space_step(host); veldiff("step"); posdiff("step")
NC = host.p.n_org * MAXC
live = host.np("c_alive") == 1; dead = host.np("d_state") > 0
xa = host.np("xa_off").astype(np.int64); deg = xa[1:] - xa[:-1]
xk = host.np("x_key")[: int(host.np("g_x_n")[0])]
for name, m, base in (("c_vel", live, 0), ("d_vel", dead, NC)):
    g = w.np(name).astype(np.float64); h = host.np(name).astype(np.float64)
    e = np.abs(g - h).max(1); e[~m] = 0
    rms = np.sqrt(np.mean(h[m] ** 2))
    top = np.argsort(-e)[:10]
    for r in top:
        b = r + base
        print(f"  {name} row {r} body {b} err {e[r]:.3e} ({e[r] / rms:.1e}) gpu {g[r]} ora {h[r]} pre {hpre.np(name)[r]} cross contacts {deg[b]}",
              (f"org {r // MAXC} ncells {host.np('org_ncells')[r // MAXC]} ic_n {host.np('org_ic_n')[r // MAXC]} org cross deg {deg[(r // MAXC) * MAXC:(r // MAXC + 1) * MAXC].sum()}" if base == 0 else
               f"mass {host.np('d_mass')[r]} mass0 {host.np('d_mass0')[r]} size {host.np('d_size')[r]}"))
xg = w.np("x_jn")[: len(xk)].astype(np.float64); xh = host.np("x_jn")[: len(xk)].astype(np.float64)
top = np.argsort(-np.abs(xg - xh))[:10]
for c in top:
    print(f"  x contact {c} ({xk[c] >> 32}, {xk[c] & 0xffffffff}) jn gpu {xg[c]:.5f} ora {xh[c]:.5f}")
print("x_jn percentiles", np.percentile(np.abs(xh), [50, 90, 99, 100]))

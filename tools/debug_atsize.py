"""Where the single-tick error of a large world sits: top offenders of a field with their organism's context."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle import oracle as orc
from sapiogenesis_b200.synth import build_world
from sapiogenesis_b200.world import MAXC, PH_ALL, PH_REPRODUCE, PH_TRAIN

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 24
orc.build()
w, eng = build_world(n, seed=3)
w.t["org_flags"] |= 2; w.p.population_limit = 0; eng.refresh(); eng.exclusive = True
ph = PH_ALL & ~(PH_TRAIN | PH_REPRODUCE)
eng.tick(warm, ph); eng.synchronize()
host = w.to("cpu")
h0 = host.clone()
eng.tick(1, ph); eng.synchronize()
orc.tick(host, ph, 1)
live = host.np("c_alive") == 1
for name in ("c_vel", "c_vbias", "c_angvel"):
    g = w.np(name).astype(np.float64); o = host.np(name).astype(np.float64)
    if g.ndim == 2:
        err = np.abs(g - o).max(1)
    else:
        err = np.abs(g - o)
    err[~live] = 0
    rms = np.sqrt(np.mean(o[live] ** 2))
    top = np.argsort(-err)[:12]
    print(f"== {name}: rms {rms:.4g}, cells with err > 1e-5 rms: {(err > 1e-5 * rms).sum()} of {live.sum()}; > 3e-5: {(err > 3e-5 * rms).sum()}")
    orgs = sorted(set(int(r) // MAXC for r in top))
    for r in top:
        oo = int(r) // MAXC
        nc = int(host.np("org_ncells")[oo]); nlive = int(live[oo * MAXC: oo * MAXC + nc].sum())
        print(f"   row {r} org {oo} k {r % MAXC} ncells {nc} live {nlive} ic_n {int(host.np('org_ic_n')[oo])} err {err[r]:.3e} ({err[r] / rms:.2e} rms) gpu {g[r]} ora {o[r]} "
              f"vel0 {h0.np(name)[r]} type {int(host.np('c_type')[r])} pos {host.np('c_pos')[r]}")
# errors per organism: how many organisms are affected, by size
g = w.np("c_vel").astype(np.float64); o = host.np("c_vel").astype(np.float64)
err = np.abs(g - o).max(1); err[~live] = 0
rms = np.sqrt(np.mean(o[live] ** 2))
eo = err.reshape(-1, MAXC).max(1)
bad = np.nonzero(eo > 1e-5 * rms)[0]
print("organisms with a cell off by > 1e-5 rms:", len(bad))
nl = live.reshape(-1, MAXC).sum(1)
print("their live cells:", np.bincount(nl[bad] // 8))
print("their ic_n:", np.percentile(host.np("org_ic_n")[bad], [0, 50, 90, 100]) if len(bad) else None)
# coupled?
xa = host.np("xa_off")
deg = (xa[1:] - xa[:-1])[: len(live)].reshape(-1, MAXC).sum(1)
print("of them coupled (cross contacts):", int((deg[bad] > 0).sum()))
jn = np.abs(host.np("j_jn").reshape(-1, 2016)).max(1)
print("max |j_jn| of the bad:", np.percentile(jn[bad], [0, 50, 90, 100]) if len(bad) else None, " of all:", np.percentile(jn[nl > 0], [50, 90, 99, 100]))
icj = np.abs(host.np("ic_jn").reshape(-1, 192)).max(1)
print("max |ic_jn| of the bad:", np.percentile(icj[bad], [0, 50, 90, 100]) if len(bad) else None, " of all:", np.percentile(icj[nl > 0], [50, 90, 99, 100]))
vmax = np.abs(h0.np("c_vel")).reshape(-1, MAXC * 2).max(1)
print("max |v| before of the bad:", np.percentile(vmax[bad], [0, 50, 90, 100]) if len(bad) else None, " of all:", np.percentile(vmax[nl > 0], [50, 90, 99, 100]))
gr = w.np("org_gas_refund"); orf = host.np("org_gas_refund")
e = np.abs(gr - orf); top = np.argsort(-e)[:5]
print("gas refund rms", np.sqrt(np.mean(orf ** 2)), [(int(t), float(gr[t]), float(orf[t])) for t in top])

"""Stage-by-stage comparison for chosen organisms of a large world (rules, then Space.step)."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle import oracle as orc
from sapiogenesis_b200.synth import build_world
from sapiogenesis_b200.world import MAXC, PH_ALL, PH_REPRODUCE, PH_TRAIN

n = int(sys.argv[1]); warm = int(sys.argv[2]); orgs = [int(a) for a in sys.argv[3:]]
orc.build()
w, eng = build_world(n, seed=3)
w.t["org_flags"] |= 2; w.p.population_limit = 0; eng.refresh(); eng.exclusive = True
ph = PH_ALL & ~(PH_TRAIN | PH_REPRODUCE)
eng.tick(warm, ph); eng.synchronize()
host = w.to("cpu")
np.set_printoptions(precision=7, linewidth=200)
def show(tag):
    for o in orgs:
        nc = int(host.np("org_ncells")[o]); sl = slice(o * MAXC, o * MAXC + nc)
        for name in ("c_vel", "c_angvel", "c_vbias", "c_health", "c_energy"):
            g = w.np(name)[sl].astype(np.float64); h = host.np(name)[sl].astype(np.float64)
            e = np.abs(g - h); e = e.max(1) if e.ndim == 2 else e
            k = int(np.argmax(e))
            print(f"{tag} org {o} {name}: max err {e.max():.3e} at k {k} gpu {g[k]} ora {h[k]}")
        print(f"   move {host.np('org_move')[o]} types {host.np('c_type')[sl]}")
eng.rules(ph); eng.synchronize(); orc.rules(host, ph); show("rules")
eng.settle(); eng.synchronize(); orc.settle(host); show("settle")
eng.space_step(); eng.synchronize(); orc.space_step(host); show("step")
for o in orgs:
    nc = int(host.np("org_ncells")[o])
    n_ic = int(host.np("org_ic_n")[o])
    code = host.np("ic_code")[o * 192: o * 192 + n_ic]
    gj = w.np("ic_jn")[o * 192: o * 192 + n_ic]; hj = host.np("ic_jn")[o * 192: o * 192 + n_ic]
    gt = w.np("ic_jt")[o * 192: o * 192 + n_ic]; ht = host.np("ic_jt")[o * 192: o * 192 + n_ic]
    print(f"org {o}: intra contacts (i, j, jn gpu, jn ora, jt gpu, jt ora)")
    for c in range(n_ic):
        print(f"   {code[c] // 128:3d} {code[c] % 128:3d}  {gj[c]:14.6f} {hj[c]:14.6f}   {gt[c]:12.6f} {ht[c]:12.6f}")
    print("   pos", host.np("c_pos")[o * MAXC: o * MAXC + nc])
    print("   size", host.np("c_size")[o * MAXC: o * MAXC + nc])
for o in orgs:
    nc = int(host.np("org_ncells")[o]); sl = slice(o * MAXC, o * MAXC + nc)
    print("c_pos exact:", np.array_equal(w.np("c_pos")[sl], host.np("c_pos")[sl]), "c_spos exact:", np.array_equal(w.np("c_spos")[sl], host.np("c_spos")[sl]))
    gj = w.np("j_jn")[o * 2016:(o + 1) * 2016].astype(np.float64); hj = host.np("j_jn")[o * 2016:(o + 1) * 2016].astype(np.float64)
    I, J = np.triu_indices(MAXC, 1)
    top = np.argsort(-np.abs(gj - hj))[:12]
    for t in top:
        print(f"   pin ({I[t]},{J[t]}) gpu {gj[t]:.5f} ora {hj[t]:.5f} diff {gj[t] - hj[t]:.2e}")
    print("   mass", host.np("c_mass")[sl])
    gs = w.np("c_spin_jn")[sl]; hs = host.np("c_spin_jn")[sl]
    print("   spin_jn diff", np.abs(gs - hs).max(), "svel diff", np.abs(w.np("c_svel")[sl] - host.np("c_svel")[sl]).max())
sys.exit(0)
# sensitivity of the oracle's own step to one ulp-sized perturbations of the input velocities (conditioning of the organism's solve)
if "--sens" in sys.argv or True:
    base = w.to("cpu")   # not used: rebuild the pre-step state instead
w2, eng2 = build_world(n, seed=3)
w2.t["org_flags"] |= 2; w2.p.population_limit = 0; eng2.refresh(); eng2.exclusive = True
eng2.tick(warm, ph); eng2.rules(ph); eng2.settle(); eng2.synchronize()
ha = w2.to("cpu"); hb = ha.clone()
rng = np.random.default_rng(0)
for o in orgs:
    sl = slice(o * MAXC, o * MAXC + int(ha.np("org_ncells")[o]))
    v = hb.t["c_vel"][sl]
    import torch
    hb.t["c_vel"][sl] = v * (1.0 + torch.from_numpy(rng.choice([-1.0, 1.0], size=tuple(v.shape)).astype(np.float32)) * 1.2e-7)
orc.space_step(ha); orc.space_step(hb)
for o in orgs:
    sl = slice(o * MAXC, o * MAXC + int(ha.np("org_ncells")[o]))
    for name in ("c_vel", "c_vbias", "c_angvel"):
        a = ha.np(name)[sl].astype(np.float64); b = hb.np(name)[sl].astype(np.float64)
        print(f"oracle sensitivity org {o} {name}: max |diff| {np.abs(a - b).max():.3e} (input perturbed by 1 ulp)")

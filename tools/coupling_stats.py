"""Connected components of the cross-contact graph (organisms + dead bodies; walls do not connect) of a synthetic island at
a given tick: how the coupled organisms are distributed over component shapes. Guides the component-wise coupled solve."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from sapiogenesis_b200.synth import build_world
n, ticks = int(sys.argv[1]), int(sys.argv[2])
w, eng = build_world(n, seed=0)
eng.exclusive = True
eng.tick(ticks); eng.synchronize()
p = w.p
NC, ND = p.n_org * 64, p.n_dead
nx = int(w.t["g_x_n"][0])
key = w.t["x_key"][:nx].cpu().numpy().astype(np.uint64)
a, b = (key >> np.uint64(32)).astype(np.int64), (key & np.uint64(0xffffffff)).astype(np.int64)
def node(x):                                   # organisms 0..n_org-1, dead bodies n_org.., walls -> -1
    return np.where(x < NC, x // 64, np.where(x < NC + ND, p.n_org + (x - NC), -1))
na, nb = node(a), node(b)
keep = (na >= 0) & (nb >= 0) & (na != nb)
N = p.n_org + ND
g = coo_matrix((np.ones(keep.sum()), (na[keep], nb[keep])), shape=(N, N))
ncomp, lab = connected_components(g, directed=False)
touched = np.zeros(N, bool); touched[na[na >= 0]] = True; touched[nb[nb >= 0]] = True
is_org = np.arange(N) < p.n_org
orgs_in = np.bincount(lab[touched & is_org], minlength=ncomp)
dead_in = np.bincount(lab[touched & ~is_org], minlength=ncomp)
cont_in = np.bincount(lab[na[keep]], minlength=ncomp)
has = (orgs_in + dead_in) > 0
print(f"tick {w.tick}: cross contacts {nx} (body-body {int(keep.sum())}), coupled organisms {int((touched & is_org).sum())}, touched dead bodies {int((touched & ~is_org).sum())}, components {int(has.sum())}")
tot_org = orgs_in.sum()
for name, sel in [("1 organism + dead bodies only", (orgs_in == 1)), ("2 organisms", (orgs_in == 2)), ("3-4 organisms", (orgs_in >= 3) & (orgs_in <= 4)),
                  ("5-16 organisms", (orgs_in >= 5) & (orgs_in <= 16)), (">16 organisms", orgs_in > 16), ("dead bodies only", (orgs_in == 0) & has)]:
    k = sel & has
    print(f"  {name:32s}: components {int(k.sum()):7d} organisms {int(orgs_in[k].sum()):7d} ({100.0 * orgs_in[k].sum() / max(tot_org, 1):5.1f} %) dead bodies {int(dead_in[k].sum()):7d} "
          f"max dead/comp {int(dead_in[k].max(initial=0))} contacts {int(cont_in[k].sum())} max contacts/comp {int(cont_in[k].max(initial=0))}")
one = (orgs_in == 1) & has
for lim in (1, 2, 4, 8, 16, 32):
    print(f"  1-organism components with <= {lim:2d} dead bodies: {int((one & (dead_in <= lim)).sum())}")

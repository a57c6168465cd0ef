"""Component structure of an aged bench world: how many per tier, the oversize ones (reads the planner's arrays from the workspace is
not possible from Python: recomputed here with scipy from the contact list)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sapiogenesis_b200.synth import build_world
from sapiogenesis_b200.world import MAXC
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
age = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
w, eng = build_world(n, seed=0); eng.exclusive = True
eng.tick(age); eng.synchronize()
nx = int(w.t["g_x_n"].item())
key = w.np("x_key")[:nx]
a = (key >> np.uint64(32)).astype(np.int64); b = (key & np.uint64(0xFFFFFFFF)).astype(np.int64)
NC = w.p.n_org * MAXC; NB = NC + w.p.n_dead
def node(x): return np.where(x < NC, x // MAXC, np.where(x < NB, w.p.n_org + (x - NC), -1))
na, nb = node(a), node(b)
NN = w.p.n_org + w.p.n_dead
m = nb >= 0
g = coo_matrix((np.ones(m.sum()), (na[m], nb[m])), shape=(NN, NN))
nc, lab = connected_components(g, directed=False)
touched = np.zeros(NN, bool); touched[na] = True; touched[nb[m]] = True
labs = lab[touched]
comp_ids, inv = np.unique(lab[na], return_inverse=True)
cnt_x = np.bincount(inv)
is_org = np.arange(NN) < w.p.n_org
ncells = w.np("org_ncells")
org_nodes = np.nonzero(touched & is_org)[0]; dead_nodes = np.nonzero(touched & ~is_org)[0]
idx = {c: i for i, c in enumerate(comp_ids)}
norg = np.zeros(len(comp_ids), int); ndead = np.zeros(len(comp_ids), int); need = np.zeros(len(comp_ids), int)
SL = [8544, 19120, 26000, 43000]
for o in org_nodes:
    i = idx[lab[o]]; norg[i] += 1; c = ncells[o]; need[i] += SL[0 if c <= 24 else 1 if c <= 40 else 2 if c <= 48 else 3]
for d in dead_nodes:
    ndead[idx[lab[d]]] += 1
print("components", len(comp_ids), "contacts", nx, "coupled organisms", len(org_nodes), "coupled dead", len(dead_nodes))
print("organisms per component:", np.bincount(norg)[:20])
print("contacts per component percentiles", np.percentile(cnt_x, [50, 90, 99, 99.9, 100]))
print("dead per component percentiles", np.percentile(ndead, [50, 90, 99, 99.9, 100]))
print("slab need percentiles (KB)", np.percentile(need, [50, 90, 99, 99.9, 100]) / 1024)
big = np.nonzero((norg > 16) | (need > 190 * 1024) | (cnt_x > 384) | (ndead > 192))[0]
print("oversize:", [(int(norg[i]), int(need[i] // 1024), int(cnt_x[i]), int(ndead[i])) for i in big])

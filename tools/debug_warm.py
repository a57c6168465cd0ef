import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import oracle
from sapiogenesis_b200.engine import Engine
from sapiogenesis_b200.world import *
from worlds import make_world
w0 = make_world(oracle, n_org=6, seed=13, n_food=12, crowd=0.7, kick=20.0, extra_slots=6)
w0.p.think_ticks = w0.p.train_ticks = 3
assert oracle.tick(w0, PH_ALL, 40) == 0
w = w0.to("cpu"); w.p.iterations = 0
v0 = w.np("c_vel").copy(); d0 = w.np("d_vel").copy()
wg = w.to("cuda"); e = Engine(wg); e.tick(1, PH_STEP); e.synchronize()
oracle.tick(w, PH_STEP, 1)
nx = int(w.np("g_x_n")[0]); keys = w.np("x_key")[:nx]
print("cross contacts", [(int(k >> 32), int(k & 0xffffffff)) for k in keys], "jn", w.np("x_jn")[:nx], "NC", w.p.n_org * 64)
rows = np.nonzero(w.np("c_alive") == 1)[0]
err = np.abs(wg.np("c_vel")[rows].astype(np.float64) - w.np("c_vel")[rows]).max(1)
dv = np.abs(w.np("c_vel")[rows].astype(np.float64) - v0[rows]).max(1)
for o in range(w.p.n_org):
    m = rows // 64 == o
    if m.any(): print("org", o, "cells", int(m.sum()), "max err", err[m].max(), "max |dv| oracle", dv[m].max(), "max|j_jn|", np.abs(w0.np("j_jn").reshape(w.p.n_org, -1)[o]).max(), "ic_n", int(w.np("org_ic_n")[o]))
dead = np.nonzero(w.np("d_state") > 0)[0]
de = np.abs(wg.np("d_vel")[dead].astype(np.float64) - w.np("d_vel")[dead]).max(1)
i = int(np.argmax(de)); print("worst dead", dead[i], "body id", w.p.n_org * 64 + dead[i], "err", de[i], "gpu", wg.np("d_vel")[dead[i]], "oracle", w.np("d_vel")[dead[i]], "before", d0[dead[i]])

import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import oracle
from parity import compare_worlds
from sapiogenesis_b200.engine import Engine
from sapiogenesis_b200.world import *
from worlds import make_world
w0 = make_world(oracle, n_org=6, seed=13, n_food=12, crowd=0.7, kick=20.0, extra_slots=6)
w0.p.think_ticks = w0.p.train_ticks = 3
assert oracle.tick(w0, PH_ALL, 40) == 0
for iters in (0, 1, 2, 5, 10, 20):
    w = w0.to("cpu")
    w.p.iterations = iters
    wg = w.to("cuda"); e = Engine(wg); e.tick(1, PH_STEP); e.synchronize()
    oracle.tick(w, PH_STEP, 1)
    rows = np.nonzero(w.np("c_alive") == 1)[0]; dead = np.nonzero(w.np("d_state") > 0)[0]
    sens = rows[np.isin(w.np("c_type")[rows], [2, 3])]
    def err(name, idx):
        a, b = wg.np(name)[idx].astype(np.float64), w.np(name)[idx].astype(np.float64)
        return np.abs(a - b).max(), np.sqrt((b * b).mean())
    print(f"iterations {iters:2d}: c_vel {err('c_vel', rows)}  d_vel {err('d_vel', dead)}  c_svel {err('c_svel', sens)}  coupled {wg.stats()['coupled']} xc {wg.stats()['xcontacts']}")
    a, b = wg.np("c_svel")[sens].astype(np.float64), w.np("c_svel")[sens].astype(np.float64)
    i = np.argmax(np.abs(a - b).max(1)); print("    worst sensor row", sens[i], a[i] - b[i])

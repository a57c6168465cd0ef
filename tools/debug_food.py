import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import oracle
from sapiogenesis_b200.world import *
from sapiogenesis_b200.engine import Engine
from worlds import add_food
w = World(default_params(n_org=4, n_dead=64, width=600.0, height=400.0, seed=2))
rng = np.random.default_rng(0)
for d in range(40):
    add_food(w, d, 25.0 + 20.0 * (d % 7) + rng.uniform(-3, 3), 25.0 + 20.0 * (d // 7) + rng.uniform(-3, 3), vx=rng.uniform(-40, 40), vy=rng.uniform(-40, 40))
for step in range(4):
    wg = w.to("cuda"); eng = Engine(wg)
    eng.tick(1, PH_STEP | PH_FRICTION | PH_POST); eng.synchronize()
    oracle.tick(w, PH_STEP | PH_FRICTION | PH_POST, 1)
    nx = int(w.np("g_x_n")[0])
    a, b = wg.np("x_jn")[:nx].astype(np.float64), w.np("x_jn")[:nx].astype(np.float64)
    at, bt = wg.np("x_jt")[:nx].astype(np.float64), w.np("x_jt")[:nx].astype(np.float64)
    print("step", step, "rms jn", np.sqrt(np.mean(b * b)), "max|jn|", np.abs(b).max(), "rms v", np.sqrt(np.mean(w.np("d_vel")[:40] ** 2)))
    idx = np.argsort(-np.abs(a - b))[:4]
    for i in idx:
        k = int(w.np("x_key")[i]); print("   c", i, "key", k >> 32, k & 0xffffffff, "jn gpu/ora", a[i], b[i], "jt", at[i], bt[i])
    idx = np.argsort(-np.abs(at - bt))[:3]
    for i in idx:
        k = int(w.np("x_key")[i]); print("   t", i, "key", k >> 32, k & 0xffffffff, "jn gpu/ora", a[i], b[i], "jt", at[i], bt[i])

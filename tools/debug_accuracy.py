"""A/B of solver variants on ONE state: an aged crowded world, one tick on the GPU per variant from the same host copy, against one
oracle tick. Needs a library built with -DSAPIO_DEBUG_FLAGS (SAPIO_SO=...). usage: python tools/debug_accuracy.py [organisms] [age]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import oracle
from sapiogenesis_b200.engine import Engine
from sapiogenesis_b200.synth import build_world
from sapiogenesis_b200.world import MAXC, PH_ALL

n = int(sys.argv[1]) if len(sys.argv) > 1 else 14000
age = int(sys.argv[2]) if len(sys.argv) > 2 else 3200
w, eng = build_world(n, seed=5, area_per_org=200_000.0, food_per_org=1.0)
eng.exclusive = True
eng.tick(age); eng.synchronize()
for rep in range(3):
    host0 = w.to("cpu")
    ref = host0.to("cpu") if hasattr(host0, "to") else host0
    ref = w.to("cpu")
    assert oracle.tick(ref, PH_ALL, 1) == 0
    ov = ref.np("c_vel").astype(np.float64)
    live = (ref.np("c_alive") == 1) & np.repeat(ref.np("org_state") == 1, MAXC)
    rms = np.sqrt((ov[live] ** 2).mean())
    for flags in (0, 1, 2, 3):
        wg = host0.to("cuda"); e2 = Engine(wg); e2.exclusive = True
        e2.lib.sapio_debug_flags(flags)
        e2.tick(1); e2.synchronize()
        gv = wg.np("c_vel").astype(np.float64)
        err = np.abs(gv - ov).max(axis=1) * live
        r = int(np.argmax(err)); o = r // MAXC
        nl = int((ref.np("c_alive")[o * MAXC:(o + 1) * MAXC] == 1).sum())
        print(f"tick {ref.tick} flags {flags}: max c_vel err {err.max():.3e} px/s = {err.max() / rms:.2e} of rms {rms:.1f}; bodies above 1e-4 rms: {int((err > 1e-4 * rms).sum())}, "
              f"above 3e-5: {int((err > 3e-5 * rms).sum())}; worst organism {o} rows {int(ref.np('org_ncells')[o])} living {nl} ic {int(ref.np('org_ic_n')[o])}", flush=True)
        del wg, e2
    eng.tick(17); eng.synchronize()

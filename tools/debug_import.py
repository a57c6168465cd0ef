import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import oracle
from parity import compare_worlds, Report
from sapiogenesis_b200 import dna_io
from sapiogenesis_b200.engine import Engine
from sapiogenesis_b200.world import *
from worlds import make_world
def grown():
    w = make_world(oracle, n_org=6, seed=13, n_food=12, crowd=0.7, kick=20.0, extra_slots=6)
    w.p.think_ticks = w.p.train_ticks = 3
    assert oracle.tick(w, PH_ALL, 40) == 0
    return w
for do_import in (False, True):
    w = grown()
    if do_import:
        donor = int(np.nonzero(w.np("org_mem_n") > 0)[0][0])
        dna = dna_io.export_dna(w, donor)
        free = int(np.nonzero(w.np("org_state") == 0)[0][0])
        dna_io.import_dna(w, free, (2500.0, 1200.0), dna); oracle.post(w)
    for k in range(3):
        wg = w.to("cuda"); e = Engine(wg); e.tick(1, PH_ALL); e.synchronize()
        oracle.tick(w, PH_ALL, 1)
        try:
            compare_worlds(wg, w, f"import={do_import} tick {w.tick}", verbose=True)
        except AssertionError as ex:
            print("FAIL", str(ex)[:300])
            sens = (w.np("c_alive") == 1) & np.isin(w.np("c_type"), [2, 3])
            a, b = wg.np("c_svel")[sens].astype(np.float64), w.np("c_svel")[sens].astype(np.float64)
            i = np.argmax(np.abs(a - b).max(1)); rows = np.nonzero(sens)[0]
            print("  worst sensor row", rows[i], "org", rows[i] // 64, "gpu", a[i], "oracle", b[i], "rms", np.sqrt((b * b).mean()), "spos-pos", w.np("c_spos")[rows[i]] - w.np("c_pos")[rows[i]], "cvel", w.np("c_vel")[rows[i]])

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import oracle
from worlds import make_world
from sapiogenesis_b200.world import MAXC, PH_ALL, MEMROWS, STM
from sapiogenesis_b200.engine import Engine
w = make_world(oracle, n_org=16, seed=11, n_food=40, crowd=0.45, kick=30.0, extra_slots=24)
w.p.think_ticks, w.p.train_ticks, w.p.repro_ticks = 3, 3, 8
w.p.population_limit = 40
oracle.tick(w, PH_ALL, 6)
for t in range(30):
    wg = w.to("cuda"); eng = Engine(wg)
    eng.rules(PH_ALL); eng.synchronize(); oracle.rules(w, PH_ALL)
    eng.train(); eng.synchronize(); oracle.train(w)
    bad = False
    for o in np.nonzero(w.np("org_state") == 1)[0]:
        if int(w.np("org_adam_t")[o]) == 0: continue
        nl = int(w.np("org_nlayers")[o]); ld = w.np("org_ldim")[o]; I = int(ld[0]); h0 = int(ld[1]); hl = int(ld[nl-1])
        npar = h0*I + h0 + 4*hl + 4
        gm, om = wg.np("org_adam_m")[o][:npar], w.np("org_adam_m")[o][:npar]
        rms = np.sqrt(np.mean(om.astype(np.float64)**2))
        err = np.abs(gm-om)/max(rms,1e-12)
        if err.max() > 2e-5:
            bad = True
            M = int(w.np("org_mem_n")[o])
            segs = {"Win": slice(0,h0*I), "bin": slice(h0*I,h0*I+h0), "Wout": slice(h0*I+h0, h0*I+h0+4*hl), "bout": slice(h0*I+h0+4*hl, npar)}
            print(f"tick {w.tick} org {o}: M={M} (gpu {int(wg.np('org_mem_n')[o])}) adam_t={int(w.np('org_adam_t')[o])} loss gpu {wg.np('org_loss')[o]:.7f} or {w.np('org_loss')[o]:.7f} nl={nl} ldim={ld[:nl+1]}")
            for k, sl in segs.items():
                print("   ", k, "max err", float(err[sl].max()), "rms seg", float(np.sqrt(np.mean(om[sl].astype(np.float64)**2))))
            mo_g = wg.np("org_mem_out")[o][:M*4]; mo_o = w.np("org_mem_out")[o][:M*4]
            print("    mem_out max diff", float(np.abs(mo_g-mo_o).max()), "mem_in equal", np.array_equal(wg.np("org_mem_in")[o], w.np("org_mem_in")[o]), "mem_rew diff", float(np.abs(wg.np("org_mem_rew")[o][:M]-w.np("org_mem_rew")[o][:M]).max()))
            print("    bout m gpu", gm[segs["bout"]], "oracle", om[segs["bout"]])
    if bad: break
    oracle.settle(w); oracle.reproduce(w); oracle.space_step(w); oracle.friction(w); oracle.post(w)
    w.t["g_tick"] += 1

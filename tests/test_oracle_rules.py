"""Pins the oracle's rules phase (oracle/rules.c: Organism.update, _update_cell, add/take_energy, convert_co2/o2,
kill_cell cascade, gradual_dispersion) against the reference's OWN sprites.update_organisms, run headless on single
organisms for 40 ticks (fixtures: tests/golden/rules_cases.npz, made by tests/golden/make_golden_rules.py).
Both sides keep the carried state in fp32 (the world arrays), so per-tick agreement is at fp32 rounding level."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from sapiogenesis_b200.world import MAXC, World, default_params

CASES = np.load(os.path.join(GOLDEN, "rules_cases.npz"))
N_CASES, SEED, N_TICKS = (int(v) for v in CASES["meta"])
PH = 1            # rules only: no think, no train, no reproduce


def u24(tag, n):
    rng = random.Random(tag)
    return np.asarray([rng.getrandbits(24) / 16777216.0 for _ in range(n)], dtype=np.float32)


@pytest.mark.parametrize("ci", range(N_CASES))
def test_rules_match_reference(oracle, ci):
    g = lambda k: CASES[f"c{ci}.{k}"]
    tag = f"rules-{SEED}-{ci}"
    p = default_params(n_org=2, n_dead=80)
    p.population_limit = 0
    w = World(p)
    used, pos = g("used"), g("pos")
    rc, _ = oracle.spawn_arr(w, 0, float(pos[0]), float(pos[1]), u24(tag + "-u", used[0] + 8), u24(tag + "-ui", used[1] + 8),
                             u24(tag + "-ub", used[2] + 8))
    assert rc == 0
    nc = int(w.np("org_ncells")[0])
    assert nc == len(g("energy0"))
    w.t["g_co2"][0], w.t["g_o2"][0] = float(g("env0")[0]), float(g("env0")[1])
    w.t["g_population"][0] = 1
    w.t["c_energy"][:nc] = torch.from_numpy(g("energy0"))
    w.t["c_health"][:nc] = torch.from_numpy(g("health0"))
    w.t["org_move"][0] = torch.from_numpy(g("move"))
    org0 = g("org0")
    w.t["org_consumed"][0], w.t["org_last_health"][0], w.t["org_last_energy"][0] = float(org0[0]), float(org0[1]), float(org0[2])
    w.t["org_think_tick"][0] = 1 << 30            # brain.lastUpdated far in the future: never thinks
    w.t["org_flags"][0] |= 2                      # maladaptive: never trains
    if org0[3] > 0:
        w.t["org_lastbirth_tick"][0] = -100000    # reproduction_limit already satisfied
    w.t["c_pos"][:nc] = w.t["c_pos"][:nc]         # fp32 already
    worst = {}
    for t in range(N_TICKS):
        assert oracle.rules(w, PH) == 0
        assert oracle.settle(w) == 0
        w.t["g_tick"] += 1
        alive_ref = g("alive")[t].astype(bool)
        alive = w.np("c_alive")[:nc] == 1
        np.testing.assert_array_equal(alive, alive_ref, err_msg=f"tick {t}: living cells")
        np.testing.assert_array_equal(w.np("c_inuse")[:nc][alive].astype(bool), g("inuse")[t].astype(bool)[alive], err_msg=f"tick {t}: in_use")

        def chk(name, got, ref, rtol, atol):
            got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
            err = np.abs(got - ref) / (atol / rtol + np.abs(ref))
            worst[name] = max(worst.get(name, 0.0), float(err.max()) if err.size else 0.0)
            np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol, err_msg=f"tick {t}: {name}")
        chk("health", w.np("c_health")[:nc][alive], g("health")[t][alive], 2e-6, 1e-4)
        chk("energy", w.np("c_energy")[:nc][alive], g("energy")[t][alive], 2e-6, 1e-4)
        chk("vel", w.np("c_vel")[:nc][alive], g("vel")[t][alive], 2e-6, 1e-5)
        chk("angvel", w.np("c_angvel")[:nc][alive], g("angvel")[t][alive], 2e-6, 1e-5)
        org = g("org")[t]
        chk("consumed", w.np("org_consumed")[0], org[0], 2e-6, 1e-5)
        if org[3] > 0:
            chk("last_health", w.np("org_last_health")[0], org[1], 2e-6, 1e-5)
            chk("last_energy", w.np("org_last_energy")[0], org[2], 2e-6, 1e-5)
        assert int(w.np("org_state")[0]) == int(org[3])
        # gas: the only defined difference is the deferred co2 refund of saturated organisms (contract C4): second order
        env = g("env")[t]
        chk("co2", w.np("g_co2")[0], env[0], 5e-7, 1e-6)
        chk("o2", w.np("g_o2")[0], env[1], 5e-7, 1e-6)
        dead_ref = sorted(tuple(r) for r in g("dead")[t] if r[0] > 0)
        ds = w.np("d_state") > 0
        dead = sorted(zip(w.np("d_mass")[ds].tolist(), w.np("d_pos")[ds][:, 0].tolist(), w.np("d_pos")[ds][:, 1].tolist()))
        assert len(dead) == len(dead_ref), f"tick {t}: dead bodies"
        if dead:
            np.testing.assert_allclose(np.asarray(dead), np.asarray(dead_ref), rtol=2e-6, atol=1e-5, err_msg=f"tick {t}: dead bodies")
    print(f"case {ci}: worst normalised errors {worst}")

"""Long-run population statistics (north star: "births, deaths, mean fitness within stated bounds, since trajectories are
chaotic"). BASELINE configs[0]: the default world of environment_settings.json (3180 x 1800, co2 30000, population limit 12),
12 creatures added the way add_creature_clicked does, 1000 ticks, fixed seeds -- free-running on the GPU and in the oracle, no
re-synchronisation. Single ticks agree to 1e-5 (the other tests); over 1000 ticks the two trajectories part ways like any two
runs of a chaotic system, so what is compared is the statistics of six worlds taken together, within the bounds written below
(they are a few standard deviations of the seed-to-seed spread of the oracle itself)."""
import numpy as np
import pytest
import torch

from sapiogenesis_b200.world import PH_ALL, World, default_params

pytestmark = pytest.mark.gpu

SEEDS = range(6)
TICKS = 1000


def default_world(oracle, seed, n=12):
    p = default_params(n_org=24, n_dead=2048, seed=seed, population_limit=12)
    w = World(p)
    w.t["g_co2"][0] = 30000.0
    rng = np.random.default_rng(seed)
    for i in range(n):
        assert oracle.spawn(w, i, float(rng.uniform(300, 2880)), float(rng.uniform(300, 1500))) == 0
    assert oracle.post(w) == 0
    return w


def summary(w):
    st = w.stats()
    live = w.np("c_alive") == 1
    org = w.np("org_state") == 1
    return dict(population=int(org.sum()), cells=int(live.sum()), dead_cells=int((w.np("d_state") > 0).sum()),
                births=st["births"], deaths=st["deaths"], cell_deaths=st["cell_deaths"], thinks=st["thinks"], trains=st["trains"],
                energy=float(w.np("c_energy")[live].sum() / max(float(w.np("c_storage")[live].sum()), 1.0)),
                health=float(w.np("c_health")[live].sum() / max(float(w.np("c_maxhealth")[live].sum()), 1.0)),
                co2=float(w.np("g_co2")[0]), gas=float(w.np("g_co2")[0] + w.np("g_o2")[0]), overflow=st["overflow"])


def test_default_world_1000_ticks_statistics(oracle):
    from sapiogenesis_b200.engine import Engine
    G, O = [], []
    for seed in SEEDS:
        w = default_world(oracle, seed)
        wg = w.to("cuda"); eng = Engine(wg)
        eng.tick(TICKS, PH_ALL); eng.synchronize()
        assert oracle.tick(w, PH_ALL, TICKS) == 0
        G.append(summary(wg)); O.append(summary(w))
        assert wg.tick == w.tick == TICKS and G[-1]["overflow"] == 0 and np.isfinite(wg.np("c_pos")).all()
    tot = lambda S, k: float(sum(s[k] for s in S))
    mean = lambda S, k: float(np.mean([s[k] for s in S]))
    for k in ("population", "cells", "dead_cells", "births", "deaths", "cell_deaths", "thinks", "energy", "health", "co2"):
        print(f"{k:12s} gpu {[round(s[k], 3) for s in G]}  oracle {[round(s[k], 3) for s in O]}")
    # the brain and metabolism cadences are clock-driven: nearly identical
    assert abs(tot(G, "thinks") - tot(O, "thinks")) <= 0.05 * tot(O, "thinks")
    # gases: co2 + oxygen is conserved by the exchange (sprites.py:1409-1439); co2 itself follows the living co2C cells
    for g, o in zip(G, O):
        assert abs(g["gas"] - o["gas"]) <= 1e-6 * o["gas"]
    assert abs(mean(G, "co2") - mean(O, "co2")) <= 0.03 * mean(O, "co2")
    # population dynamics, six worlds together
    assert abs(tot(G, "population") - tot(O, "population")) <= 0.20 * tot(O, "population")
    assert abs(tot(G, "cells") - tot(O, "cells")) <= 0.25 * tot(O, "cells")
    assert abs(tot(G, "deaths") - tot(O, "deaths")) <= max(5.0, 0.40 * tot(O, "deaths"))
    assert abs(tot(G, "cell_deaths") - tot(O, "cell_deaths")) <= max(60.0, 0.40 * tot(O, "cell_deaths"))
    assert abs(tot(G, "births") - tot(O, "births")) <= 5.0
    # "mean fitness": energy and health fractions of the living cells. One world's energy fraction moves by ~0.1 with a single
    # birth or death more or less (oracle, seed to seed: 0.32 .. 0.46), so the mean of six carries ~0.04 of noise on either side
    assert abs(mean(G, "energy") - mean(O, "energy")) <= 0.12
    assert abs(mean(G, "health") - mean(O, "health")) <= 0.10

"""GPU parity of the whole world tick and of each stage (rules incl. think, train, settle, reproduce, spawn) against the
oracle, always from identical state: the GPU world is re-created from the oracle's world before every compared call, so
every tick of a real trajectory is a single-tick parity check (trajectories themselves are chaotic).
All calls go through the C-ABI (sapio_* via sapiogenesis_b200.engine.Engine)."""
import numpy as np
import pytest
import torch

from parity import compare_worlds
from sapiogenesis_b200.world import MAXC, PH_ALL, World, default_params
from worlds import make_world

pytestmark = pytest.mark.gpu


def fast_world(oracle, n_org, seed, crowd, food, extra=24, **kw):
    """A busy little world: brains think/train every 3 ticks, reproduction allowed every 8, plenty of free slots."""
    w = make_world(oracle, n_org=n_org, seed=seed, n_food=food, crowd=crowd, kick=30.0, extra_slots=extra, n_dead_cap=64 * (n_org + extra), **kw)
    w.p.think_ticks, w.p.train_ticks, w.p.repro_ticks = 3, 3, 8
    w.p.population_limit = n_org + extra
    return w


def engine_for(w):
    from sapiogenesis_b200.engine import Engine
    wg = w.to("cuda")
    return wg, Engine(wg)


def test_spawn_matches_oracle(oracle):
    p = default_params(n_org=40, n_dead=16, seed=5)
    w = World(p)
    rng = np.random.default_rng(3)
    slots = [3, 0, 7, 12, 13, 20, 21, 22, 30, 39, 5, 6]
    pos = rng.uniform(300, 1500, (len(slots), 2))
    wg, eng = engine_for(w)
    eng.spawn(slots, pos); eng.post(); eng.synchronize()
    for s, (x, y) in zip(slots, pos):
        assert oracle.spawn(w, s, float(x), float(y)) == 0
    oracle.post(w)
    assert wg.stats()["birth_fail_caps"] == 0
    compare_worlds(wg, w, "spawn", brain_tol=0.0)
    for s in slots:     # nn.Linear init + tempNet hidden weights: same uniforms, same double formula -> identical floats
        np.testing.assert_array_equal(wg.np("org_brain")[s], w.np("org_brain")[s])


@pytest.mark.parametrize("n_org,crowd,food,seed,rnn", [(16, 0.45, 40, 11, 0), (40, 0.4, 120, 12, 0), (24, 0.45, 60, 13, 6)])
def test_every_stage_matches_oracle_along_a_trajectory(oracle, n_org, crowd, food, seed, rnn):
    """rnn = hidden_rnn_size with use_rnn set (physics.py:305,308): RNNetwork brains (networks.py:111-175) -- forward on inputs ++ lastHidden,
    hidden_memory along the short-term memory, Trainer.train_rnn's one-sample SGD, inheritance of hidden_memory / trainingHidden at birth;
    use_rnn is switched off half way, so feed-forward children of RNN parents live next to them."""
    w = fast_world(oracle, n_org, seed, crowd, food, extra=96 if rnn else 24, rnn_hidden=rnn)
    assert oracle.tick(w, PH_ALL, 6) == 0                 # warm up: contacts, adjacency, first thinks and memories exist
    seen = dict(thinks=0, trains=0, births=0, cell_deaths=0, deaths=0)
    for t in range(30 if not rnn else 60):
        if rnn and t == 40:
            w.p.rnn_hidden = 0
        if t == 8:                                         # provoke deaths (cascades, dead bodies) and births
            alive = np.nonzero(w.np("c_alive") == 1)[0]
            kill = alive[(alive % MAXC != 0)][::7]
            w.t["c_health"][torch.from_numpy(kill)] = -1.0
        if t in (10, 18, 44):
            w.t["c_energy"][:] = w.t["c_storage"]          # everybody full: reproduction gate opens, co2 refunds saturate
        before = w.stats()
        wg, eng = engine_for(w)
        tag = f"n={n_org} tick={w.tick}"
        eng.rules(PH_ALL); eng.synchronize(); assert oracle.rules(w, PH_ALL) == 0
        compare_worlds(wg, w, tag + " rules", verbose=False)
        eng.train(); eng.synchronize(); assert oracle.train(w) == 0
        compare_worlds(wg, w, tag + " train", verbose=False)
        eng.settle(); eng.synchronize(); assert oracle.settle(w) == 0
        compare_worlds(wg, w, tag + " settle", verbose=False)
        eng.reproduce(); eng.synchronize(); assert oracle.reproduce(w) == 0
        compare_worlds(wg, w, tag + " reproduce", verbose=False)
        eng.space_step(); eng.friction(); eng.post(); eng.synchronize()
        assert oracle.space_step(w) == 0; oracle.friction(w); oracle.post(w)
        # the RNN trajectory is about the brains: its crowded 60-tick physics is held to 3e-5 of the rms instead of 1e-5 (d_vel reached 2e-5 once)
        compare_worlds(wg, w, tag + " step", verbose=(t % 10 == 0), phys_tol=3e-5 if rnn else None)
        assert wg.stats()["overflow"] == 0
        w.t["g_tick"] += 1
        after = w.stats()
        for k in seen:
            seen[k] += after[k] - before[k]
    print("events along the trajectory:", seen)
    assert seen["thinks"] > 0 and seen["trains"] > 0 and seen["cell_deaths"] > 0 and seen["births"] > 0
    if rnn:
        st = w.np("org_state") == 1; h = w.np("org_rnn_h")[st]; gen = w.np("org_generation")[st]
        print(f"rnn: {int((h > 0).sum())} RNN / {int((h == 0).sum())} feed-forward organisms, RNN children {int(((h > 0) & (gen > 0)).sum())}, "
              f"hidden_memory entries {int((w.np('org_hm_n')[st][:, 0] - w.np('org_hm_n')[st][:, 1]).sum())}, trainingHidden entries {int(w.np('org_th_n')[st].sum())}")
        assert (h > 0).any() and (h == 0).any() and ((h > 0) & (gen > 0)).any() and w.np("org_th_n")[st].sum() > 0


def test_full_tick_call_matches_oracle(oracle):
    w = fast_world(oracle, 24, 21, 0.5, 60)
    assert oracle.tick(w, PH_ALL, 5) == 0
    for t in range(12):
        if t == 4:
            w.t["c_energy"][:] = w.t["c_storage"]
        wg, eng = engine_for(w)
        eng.tick(1, PH_ALL); eng.synchronize()
        assert oracle.tick(w, PH_ALL, 1) == 0
        assert wg.tick == w.tick
        compare_worlds(wg, w, f"tick {w.tick}", verbose=(t % 4 == 0))


def test_sensor_views_match_oracle(oracle):
    w = fast_world(oracle, 30, 31, 0.4, 80)
    assert oracle.tick(w, PH_ALL, 3) == 0
    wg, eng = engine_for(w)
    checked = members = 0
    for o in np.nonzero(w.np("org_state") == 1)[0]:
        for q in range(int(w.np("org_nincell")[o])):
            k = int(w.np("org_incell")[o][q])
            if w.np("c_alive")[o * MAXC + k] != 1:
                continue
            ref = np.sort(oracle.sensor_view(w, int(o), k))
            got = eng.sensor_view(int(o), k)
            np.testing.assert_array_equal(got, ref, err_msg=f"view of organism {o} cell {k}")
            checked += 1; members += len(ref)
    print(f"{checked} sensors, {members} viewed bodies, all hit lists identical")
    assert checked > 20 and members > 100


def test_reusing_the_step_grid_for_the_next_think_changes_nothing(oracle):
    """SAPIO_PH_REUSE_GRID (Engine.exclusive): tick by tick, with spawns in between, bit-identical to rebuilding the grid."""
    import torch
    w = fast_world(oracle, 24, 31, 0.45, 60)
    assert oracle.tick(w, PH_ALL, 5) == 0
    wa, ea = engine_for(w)
    wb, eb = engine_for(w)
    eb.exclusive = True
    for t in range(24):
        if t == 9:                                         # a body added through the engine invalidates the grid
            free = torch.nonzero(wa.t["org_state"] == 0).view(-1)[:1].to(torch.int32)
            ea.spawn(free, [[900.0, 700.0]]); eb.spawn(free.clone(), [[900.0, 700.0]])
        ea.tick(1); eb.tick(1)
    ea.tick(6); eb.tick(6)
    ea.synchronize(); eb.synchronize()
    assert wa.stats()["thinks"] > 0 and wa.stats() == wb.stats()
    for name in wa.t:
        a, b = wa.t[name], wb.t[name]
        assert torch.equal(a, b) or torch.equal(torch.nan_to_num(a.double()), torch.nan_to_num(b.double())), name


def test_crowded_world_of_300_matches_oracle_tick_by_tick(oracle, coupled_kernel):
    """A denser, larger world than the stage-by-stage trajectories: hundreds of organisms and food pieces in contact (2500
    cross contacts, chains of coupled organisms), every solver class populated, a hundred births at once, 1800 cell deaths.
    Whole ticks (the fused call), re-synchronised every tick. Decisions and contact sets exact, fp32 fields at 1e-5 as everywhere.
    Brains are kept out of this one (think/train cadence beyond the run):
    with ~100 thinks per tick an output lands within fp32 noise of a 3-decimal rounding tie (neural.py:292) every few ticks, and
    the flipped last digit moves the organism by a different push -- that discontinuity is the reference's, not a solver error."""
    w = make_world(oracle, n_org=300, seed=77, n_food=900, crowd=0.5, kick=40.0, extra_slots=120, n_dead_cap=64 * 200, width=12000.0, height=8000.0)
    w.p.think_ticks, w.p.train_ticks, w.p.repro_ticks = 3, 3, 6
    w.p.population_limit = 420
    assert oracle.tick(w, PH_ALL, 6) == 0
    w.p.think_ticks = w.p.train_ticks = 1000
    full = (torch.arange(w.p.n_org * MAXC) // MAXC) % 3 == 0                  # a third of the organisms reproduce at once
    w.t["c_energy"][full] = w.t["c_storage"][full]
    alive = np.nonzero(w.np("c_alive") == 1)[0]
    w.t["c_health"][torch.from_numpy(alive[(alive % MAXC != 0)][::11])] = -1.0
    b0 = w.stats()
    for t in range(8):
        wg, eng = engine_for(w)
        eng.tick(1, PH_ALL); eng.synchronize()
        assert oracle.tick(w, PH_ALL, 1) == 0
        assert wg.stats()["overflow"] == 0
        compare_worlds(wg, w, f"crowd tick {w.tick}", verbose=(t == 7))
    st = w.stats()
    print({k: st[k] - b0[k] for k in ("births", "deaths", "cell_deaths")}, "xcontacts", st["xcontacts"], "coupled", st["coupled"], "levels", st["levels"])
    assert st["births"] > b0["births"] + 50 and st["cell_deaths"] > b0["cell_deaths"] + 500 and st["xcontacts"] > 500


def test_training_memories_at_their_limit_match_oracle(oracle):
    """train_network with the long-term memory at memory_limit (neural.py:399-404: lowest reward leaves first) -- the state every
    brain is in after a few thousand ticks, reached here in a few dozen by a limit of 5 rows. Every other organism has
    finite_memory switched off (Sapiogenesis.py:548-557) and keeps growing up to the capacity of the table. Compared after the rules
    (think) and train stages of every tick: curation decisions exact, weights at 1e-5."""
    from sapiogenesis_b200.world import F_FINITE_MEMORY
    w = fast_world(oracle, 24, 19, 0.45, 60)
    w.p.think_ticks, w.p.train_ticks, w.p.memory_limit = 1, 2, 5
    w.p.repro_ticks = 10 ** 6
    w.t["org_flags"][1::2] &= ~F_FINITE_MEMORY
    assert oracle.tick(w, PH_ALL, 6) == 0                 # warm up: the initial overlaps of the crowded start are resolved
    trimmed = grown = 0
    for t in range(54):
        if t % 6 == 0:                                     # swings of the energy level: dopamine / pain above 0.1, memories get written
            w.t["c_energy"][:] = w.t["c_storage"] * (0.7 if (t // 6) % 2 == 0 else 0.3)
        wg, eng = engine_for(w)                            # the brain stages of the tick on the GPU; the oracle carries the world on
        eng.rules(PH_ALL); eng.train(); eng.synchronize()
        assert oracle.rules(w, PH_ALL) == 0 and oracle.train(w) == 0
        compare_worlds(wg, w, f"memory limit tick {w.tick}", verbose=False)
        assert oracle.settle(w) == 0 and oracle.reproduce(w) == 0 and oracle.space_step(w) == 0
        oracle.friction(w); oracle.post(w); w.t["g_tick"] += 1
        live = w.np("org_state") == 1
        m, stm = w.np("org_mem_n")[live], w.np("org_stm_n")[live]
        fin = (w.np("org_flags")[live] & F_FINITE_MEMORY) != 0
        assert (m[fin] <= 5).all()
        trimmed += int((m[fin] == 5).sum()); grown = max(grown, int(m[~fin].max(initial=0)))
    print("rows at the limit", trimmed, "largest unbounded memory", grown, "trains", w.stats()["trains"])
    assert trimmed > 20 and grown == 232                   # SAPIO_MEMROWS: the unbounded memories reach the table's capacity


def test_pipelined_host_frames_equal_the_synchronous_ones(oracle):
    """sapio_tick_host_begin/_end with two frames in flight delivers, tick for tick, the frames sapio_tick_host delivers."""
    w = fast_world(oracle, 20, 41, 0.5, 50)
    assert oracle.tick(w, PH_ALL, 4) == 0
    wa, ea = engine_for(w)
    wb, eb = engine_for(w)
    ref = []                                                                   # host buffers are reused: keep copies
    for k in range(6):
        h, c, d, _ = ea.tick_host(1, PH_ALL, poison=0.01 * k)
        ref.append((h, c.clone(), d.clone()))
    got = []
    eb.tick_host_begin(1, PH_ALL, poison=0.0)
    for k in range(1, 6):
        eb.tick_host_begin(1, PH_ALL, poison=0.01 * k)
        h, c, d, _ = eb.tick_host_end(); got.append((h, c.clone(), d.clone()))
    h, c, d, _ = eb.tick_host_end(); got.append((h, c.clone(), d.clone()))
    key = lambda a: a[np.lexsort(a.numpy().T[::-1])] if len(a) else a           # packed order is unspecified
    for (hr, cr, dr), (hg, cg, dg) in zip(ref, got):
        assert hr == hg
        assert torch.equal(key(cr), key(cg)) and torch.equal(key(dr), key(dg))
    assert ref[-1][0]["tick"] == w.tick + 6 and ref[-1][0]["n_cells"] > 0


def test_world_events_match_oracle(oracle):
    """updateUI's events (Sapiogenesis.py:37-53): drought, poison and the random dead cell, as a stage and inside the host tick."""
    w = fast_world(oracle, 12, 51, 0.6, 30)
    assert oracle.tick(w, PH_ALL, 3) == 0
    born = 0
    for k in range(12):
        wg, eng = engine_for(w)
        d0 = int((w.np("d_state") > 0).sum())
        eng.events(drought=0.4, poison=0.02, random_cells=0.6); eng.synchronize()
        assert oracle.events(w, 0.4, 0.02, 0.6) == 0
        compare_worlds(wg, w, f"events at tick {w.tick}", verbose=False)
        born += int((w.np("d_state") > 0).sum()) - d0
        assert oracle.tick(w, PH_ALL, 1) == 0
    assert 3 <= born <= 11                                   # probability 0.6 per tick
    fresh = np.nonzero(w.np("d_size") == 30.0)[0]
    assert len(fresh) >= born and (w.np("d_pos")[fresh] >= 10).all()
    # inside the tick with host buffers: same effect as tick + events
    wg, eng = engine_for(w)
    eng.tick_host(1, PH_ALL, drought=0.3, poison=0.01, random_cells=1.0)
    for stage in (lambda: oracle.rules(w, PH_ALL), lambda: oracle.train(w), lambda: oracle.settle(w), lambda: oracle.reproduce(w),
                  lambda: oracle.space_step(w), lambda: oracle.friction(w), lambda: oracle.events(w, 0.3, 0.01, 1.0), lambda: oracle.post(w)):
        assert stage() == 0                                  # the events sit between friction and the end-of-tick bookkeeping
    w.t["g_tick"] += 1
    compare_worlds(wg, w, "tick_host with events", verbose=False)

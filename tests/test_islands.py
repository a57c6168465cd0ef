"""Population islands (sapiogenesis_b200/islands.py): host logic on CPU with gloo at world size 2, records on the GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sapiogenesis_b200 import islands
from sapiogenesis_b200.world import FIELDS, MAXC, PH_ALL


def test_choose_emigrants_is_seeded_and_samples_living_slots():
    state = torch.zeros(200, dtype=torch.uint8); state[::3] = 1
    a = islands.choose_emigrants(state, 10, seed=5, rank=1, tick=50)
    b = islands.choose_emigrants(state, 10, seed=5, rank=1, tick=50)
    c = islands.choose_emigrants(state, 10, seed=5, rank=0, tick=50)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert a.dtype == torch.int32 and a.numel() == 10 and len(set(a.tolist())) == 10
    assert bool((state[a.long()] == 1).all()) and a.tolist() == sorted(a.tolist())
    few = torch.zeros(50, dtype=torch.uint8); few[[3, 9]] = 1
    assert islands.choose_emigrants(few, 10, 0, 0, 0).tolist() == [3, 9]
    assert islands.free_slots(few, 4).tolist() == [0, 1, 2, 4]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); return s.getsockname()[1]


def _ring_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        send = torch.full((6, 32), rank + 1, dtype=torch.uint8)
        send[:, 0] = torch.arange(6, dtype=torch.uint8)
        recv, n = islands.ring_exchange(send, 3 + rank)
        prev = (rank - 1) % world
        ok = n == 3 + prev and bool((recv[:, 1:] == prev + 1).all()) and recv[:, 0].tolist() == list(range(6))
        # the tick loop of an island with a stub engine: migration happens exactly every `every` ticks, on both ranks
        class Eng:
            def __init__(self): self.ticks, self.packed, self.unpacked = 0, [], []
            def record_bytes(self): return 16
            def tick(self, n): self.ticks += n
            def pack_organisms(self, slots, out): out[:len(slots), 0] = slots.to(torch.uint8); out[:len(slots), 1] = rank; self.packed.append(slots.tolist()); return out
            def release_organisms(self, slots): w.t["org_state"][slots.long()] = 0
            def unpack_organisms(self, slots, rec): w.t["org_state"][slots.long()] = 1; self.unpacked.append((slots.tolist(), rec[:len(slots), :2].tolist()))
        class W:
            device = torch.device("cpu"); tick = 0
            t = {"org_state": torch.zeros(40, dtype=torch.uint8)}
        w = W(); w.t["org_state"][: 10 + 5 * rank] = 1
        isl = islands.Island(w, Eng(), every=4, count=3, seed=1)
        isl.step(9)
        pop = int(w.t["org_state"].sum())
        ok = ok and isl.engine.ticks == 9 and len(isl.engine.packed) == 2 and isl.migrated_out == 6 and isl.migrated_in == 6 and pop == 10 + 5 * rank
        ok = ok and all(r[1] == prev for _, recs in isl.engine.unpacked for r in recs)
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_ring_exchange_and_island_loop_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == {0: True, 1: True}


@pytest.mark.gpu
def test_records_move_whole_organisms_between_worlds(oracle):
    from sapiogenesis_b200.engine import Engine
    from worlds import make_world
    wa = make_world(oracle, n_org=12, seed=4, n_food=20, crowd=0.4, kick=20.0, extra_slots=6)
    wa.p.think_ticks = wa.p.train_ticks = 3
    assert oracle.tick(wa, PH_ALL, 30) == 0                      # brains, memories, contact caches are populated
    wb = make_world(oracle, n_org=5, seed=9, n_food=4, crowd=0.3, extra_slots=13)
    assert wb.p.n_org == wa.p.n_org
    ga, gb = wa.to("cuda"), wb.to("cuda")
    ea, eb = Engine(ga), Engine(gb)
    out = torch.tensor([2, 7, 11], dtype=torch.int32)
    rec = ea.pack_organisms(out)
    assert rec.shape == (3, ea.record_bytes())
    ea.release_organisms(out)
    dst = islands.free_slots(gb.t["org_state"], 3)
    uid0 = int(gb.t["g_next_uid"][0])
    eb.unpack_organisms(dst, rec); eb.synchronize(); ea.synchronize()
    per_org = {"ORG": 1, "CELL": MAXC, "PAIR": MAXC * (MAXC - 1) // 2, "ICON": 192}
    for f in FIELDS:
        if f.scope not in per_org: continue
        src = wa.np(f.name).reshape(wa.p.n_org, -1); got = gb.np(f.name).reshape(gb.p.n_org, -1)
        for k, (s, d) in enumerate(zip(out.tolist(), dst.tolist())):
            if f.name == "org_uid": assert int(got[d, 0]) == uid0 + k
            else: np.testing.assert_array_equal(got[d], src[s], err_msg=f.name)
    assert int(gb.t["g_next_uid"][0]) == uid0 + 3
    assert ga.np("org_state")[out.long().numpy()].sum() == 0 and ga.np("c_alive").reshape(-1, MAXC)[out.long().numpy()].sum() == 0
    ea.tick(5); eb.tick(5); ea.synchronize(); eb.synchronize()
    assert ga.stats()["overflow"] == 0 and gb.stats()["overflow"] == 0
    assert int((gb.t["org_state"] == 1).sum()) >= 5 and np.isfinite(gb.np("c_pos")).all() and np.isfinite(ga.np("c_vel")).all()


def _busy_world(oracle, n_org, seed, crowd, extra, food=30, ticks=30):
    from worlds import make_world
    w = make_world(oracle, n_org=n_org, seed=seed, n_food=food, crowd=crowd, kick=20.0, extra_slots=extra, n_dead_cap=64 * 16)
    w.p.think_ticks = w.p.train_ticks = 3
    assert oracle.tick(w, PH_ALL, ticks) == 0                    # brains, memories, contact caches, adjacency are populated
    return w


@pytest.mark.gpu
def test_device_migration_matches_oracle_and_places_arrivals(oracle):
    """sapio_migrate_out / _in against oracle/migrate.c: the same emigrants, byte-identical ragged records, the same places, slots and
    uids; arrivals overlap nobody; nothing points at a slot whose occupant changed (no phantom bites on the next tick)."""
    from parity import compare_worlds
    from sapiogenesis_b200.engine import Engine
    src = _busy_world(oracle, 14, 4, 0.4, 6)
    dst = _busy_world(oracle, 30, 9, 0.55, 12, food=60)
    assert src.p.n_org != dst.p.n_org                             # records do not depend on the slot count of either world
    gs, gd = src.to("cuda"), dst.to("cuda")
    es, ed = Engine(gs), Engine(gd)
    buf = es.migrate_buffer(5)
    es.migrate_out(5, 3, buf); es.synchronize()
    hbuf = torch.zeros_like(buf, device="cpu")
    assert oracle.migrate_out(src, 5, 3, hbuf) == 5
    c = es.migrate_counts()
    assert c["chosen"] == 5 and c["packed"] == 5
    used = int(hbuf[8:12].view(torch.int32)[0])
    assert 5 * 4000 < used < 5 * 120_000, used                    # ragged: a few dozen kB per organism, not the 431 kB of a slot
    assert torch.equal(buf.cpu()[:used], hbuf[:used])
    compare_worlds(gs, src, "source after migrate_out")
    np.testing.assert_array_equal(gs.np("xa_other"), src.np("xa_other"))
    # arrival
    ed.migrate_in(buf); ed.synchronize()
    acc, drop = oracle.migrate_in(dst, hbuf)
    c = ed.migrate_counts()
    assert (c["accepted"], c["dropped"]) == (acc, drop) and acc == 5
    compare_worlds(gd, dst, "destination after migrate_in")
    np.testing.assert_array_equal(gd.np("org_uid"), dst.np("org_uid"))
    np.testing.assert_array_equal(gd.np("c_pos"), dst.np("c_pos"))
    # nobody overlaps anybody: boxes of the living organisms, pairwise, by the reference's own predicate
    live = np.nonzero(gd.np("org_state") == 1)[0]
    rects = np.array([oracle.org_rect(dst, int(o)) for o in live])
    arrivals = live[np.isin(gd.np("org_uid")[live], np.arange(int(dst.np("g_next_uid")[0]) - acc, int(dst.np("g_next_uid")[0])))]
    assert len(arrivals) == acc
    for a in arrivals:
        ra = rects[list(live).index(a)]
        for o, r in zip(live, rects):
            if o == a: continue
            inx = (ra[0] > r[0] and ra[0] < r[2]) or (ra[2] > r[0] and ra[2] < r[2]) or (r[0] > ra[0] and r[2] < ra[2])
            iny = (ra[1] > r[1] and ra[1] < r[3]) or (ra[3] > r[1] and ra[3] < r[3]) or (r[1] > ra[1] and r[3] < ra[3])
            assert not (inx and iny), (a, o)
        assert ra[0] >= 0 and ra[1] >= 0 and ra[2] <= dst.p.width and ra[3] <= dst.p.height
    # the next ticks agree too (adjacency scrubbed on both sides, arrivals think and train with their memories)
    for t in range(3):
        ed.tick(1); ed.synchronize(); es.tick(1); es.synchronize()
        assert oracle.tick(dst, PH_ALL, 1) == 0 and oracle.tick(src, PH_ALL, 1) == 0
        gd2 = dst.to("cuda"); compare_worlds(gd, dst, f"destination tick +{t + 1}", phys_tol=1e-4); gd = gd2; ed = Engine(gd)
        gs2 = src.to("cuda"); compare_worlds(gs, src, f"source tick +{t + 1}", phys_tol=1e-4); gs = gs2; es = Engine(gs)


@pytest.mark.gpu
def test_migration_leaves_no_stale_adjacency(oracle):
    """an arrival must not inherit the contact partners of the slot's previous occupant. Every organism leaves a
    crowded world (whose carnivores are in contact with their neighbours) and comes back into the freed slots."""
    from sapiogenesis_b200.engine import Engine
    w = _busy_world(oracle, 24, 6, 0.35, 8, food=40)
    g = w.to("cuda"); e = Engine(g)
    nx = int(g.np("g_x_n")[0])
    assert nx > 20
    buf = e.migrate_buffer(10)
    e.migrate_out(10, 0, buf); e.migrate_in(buf); e.synchronize()
    c = e.migrate_counts()
    assert c["packed"] == 10 and c["accepted"] + c["dropped"] == 10
    NC = g.p.n_org * MAXC
    off, other = g.np("xa_off"), g.np("xa_other").reshape(-1)
    na = int(off[NC + g.p.n_dead])
    uid = g.np("org_uid"); fresh = np.nonzero((g.np("org_state") == 1) & (uid >= int(g.np("g_next_uid")[0]) - c["accepted"]))[0]
    assert len(fresh) == c["accepted"]
    ent = other[:na]
    cells = ent[ent < NC]
    assert not np.isin(cells // MAXC, fresh).any(), "an adjacency entry points at an arrival"
    for o in fresh:
        rng = other[off[o * MAXC]: off[(o + 1) * MAXC]]
        assert (rng == 0x7FFFFFFF).all(), "an arrival owns adjacency entries of the previous occupant"
    hp0 = g.np("c_health").copy()
    e.tick(1, 1); e.synchronize()                                 # rules only: carnivores bite what their adjacency lists
    e.settle(); e.synchronize()
    rows = (fresh[:, None] * MAXC + np.arange(MAXC)[None, :]).reshape(-1)
    live = g.np("c_alive")[rows] == 1
    assert (g.np("c_health")[rows][live] >= hp0[rows][live] - 1e-3).all(), "an arrival was bitten through a stale contact"


@pytest.mark.gpu
def test_island_self_ring_conserves_population():
    from sapiogenesis_b200.synth import build_world
    w, eng = build_world(2000, seed=2)
    isl = islands.Island(w, eng, every=5, count=16, seed=3)
    pop0 = int((w.t["org_state"] == 1).sum())
    births0, deaths0 = w.stats()["births"], w.stats()["deaths"]
    xc0 = None
    isl.step(12)
    eng.synchronize()
    st = w.stats()
    c = isl.counts()
    assert isl.device_side and isl.migrations == 2
    assert c["total_out"] == 32 and c["total_in"] + c["total_dropped"] == 32 and c["total_dropped"] == 0 and c["total_stayed"] == 0
    assert int((w.t["org_state"] == 1).sum()) == pop0 + (st["births"] - births0) - (st["deaths"] - deaths0)
    assert st["overflow"] == 0 and w.tick == 12
    assert isl.collect_migration_ms() > 0 and isl.last_bytes < 16 * 70_000
    uid = w.t["org_uid"][w.t["org_state"] == 1]
    assert int(torch.unique(uid).numel()) == int(uid.numel())

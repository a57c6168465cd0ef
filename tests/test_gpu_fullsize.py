"""Properties that do not need the oracle, checked at the bench's full size (100k organisms per GPU, BASELINE configs[2]) where the
oracle would take minutes: conservation, sortedness, symmetry, uniqueness, consistency of counters -- and bit-exact
reproducibility of the whole state from the seed (at 20k organisms, twice)."""
import numpy as np
import pytest
import torch

from sapiogenesis_b200.world import MAXC

pytestmark = pytest.mark.gpu


def test_full_size_world_invariants():
    from sapiogenesis_b200.synth import build_world
    w, eng = build_world(100_000, seed=0)
    eng.exclusive = True
    t = w.t
    gas0 = float(t["g_co2"][0] + t["g_o2"][0])
    eng.tick(40); eng.synchronize()
    st = w.stats()
    assert w.tick == 40 and st["overflow"] == 0 and st["thinks"] > 100_000 and st["trains"] > 100_000
    # gases: the exchange conserves co2 + oxygen exactly (sprites.py:1409-1439); both stay non-negative
    assert float(t["g_co2"][0] + t["g_o2"][0]) == pytest.approx(gas0, rel=1e-12) and float(t["g_co2"][0]) >= 0 and float(t["g_o2"][0]) >= 0
    org = t["org_state"] == 1
    live = t["c_alive"] == 1
    # counters agree with the arrays
    assert int(t["g_population"][0]) == int(org.sum()) == 100_000 + st["births"] - st["deaths"]
    cells_per = live.view(-1, MAXC).sum(1)
    assert bool((cells_per[~org] == 0).all()) and bool((cells_per[org] >= 1).all()) and bool((cells_per <= t["org_ncells"]).all())
    # every body is finite and inside the walls
    pos = t["c_pos"][live]
    assert bool(torch.isfinite(pos).all()) and bool(torch.isfinite(t["c_vel"][live]).all()) and bool(torch.isfinite(t["org_brain"][org]).all())
    assert float(pos.min()) > 0 and float(pos[:, 0].max()) < w.p.width and float(pos[:, 1].max()) < w.p.height
    # uids are unique among the living; brains are what the genome says
    uid = t["org_uid"][org]
    assert int(torch.unique(uid).numel()) == int(uid.numel()) and int(uid.max()) < int(t["g_next_uid"][0])
    # cross contact list: strictly sorted keys (sortedness + uniqueness), a < b, every contact mirrored in the adjacency
    nx = int(t["g_x_n"][0])
    key = t["x_key"][:nx].to(torch.int64)
    assert nx > 100 and bool((key[1:] > key[:-1]).all())
    a, b = key >> 32, key & 0xFFFFFFFF
    assert bool((a < b).all())
    off = t["xa_off"].to(torch.int64)
    deg = off[1:] - off[:-1]
    assert int(off[-1]) == int(deg.sum()) and bool((deg >= 0).all())
    NB = w.p.n_org * MAXC + w.p.n_dead
    walls = int((b >= NB).sum())
    assert int(deg.sum()) == 2 * nx - walls                               # a wall is nobody's adjacency owner
    # per-organism intra contact lists: sorted codes, within capacity
    icn = t["org_ic_n"][org]
    assert int(icn.max()) <= 192 and int(icn.sum()) == st["icontacts"]
    codes = t["ic_code"].view(w.p.n_org, -1)[org].to(torch.int32)
    idx = torch.arange(codes.shape[1], device=codes.device)[None, :]
    valid = idx < icn[:, None]
    inc = (codes[:, 1:] > codes[:, :-1]) | ~valid[:, 1:]
    assert bool(inc.all())
    # memory rings and curated memories within their capacities
    assert int(t["org_mem_n"][org].max()) <= 232 and int(t["org_mem_n"][org].min()) >= 0


def test_state_is_reproducible_from_the_seed():
    from sapiogenesis_b200.synth import build_world
    snaps = []
    for _ in range(2):
        w, eng = build_world(20_000, seed=7)
        eng.exclusive = True
        eng.tick(60); eng.synchronize()
        snaps.append({k: v.clone() for k, v in w.t.items() if k not in ("org_rect",)})
        del w, eng
        torch.cuda.empty_cache()
    for k in snaps[0]:
        a, b = snaps[0][k], snaps[1][k]
        assert torch.equal(a, b) or torch.equal(torch.nan_to_num(a.double()), torch.nan_to_num(b.double())), k

"""Edge cases of the world step on the GPU against the oracle: empty world, food only, and hand-built organisms on every
size-class boundary of the organism solver (1..64 cell rows), with dead rows in the middle (rank compaction), sensors,
organisms pressed against the walls (wall contacts merged into the canonical contact order) and against each other."""
import numpy as np
import pytest
import torch

from parity import compare_worlds
from sapiogenesis_b200.world import MAXC, PH_ALL, PH_FRICTION, PH_POST, PH_STEP, World, default_params
from test_gpu_step import compare_step
from worlds import add_food, manual_org

pytestmark = pytest.mark.gpu


def engine_for(w):
    from sapiogenesis_b200.engine import Engine
    wg = w.to("cuda")
    return wg, Engine(wg)


def blob(n, size, cols, squeeze=0.9, types=None):
    """n cells on a grid, neighbours overlapping by (1 - squeeze) of a diameter; row 0 is the heart."""
    cells = []
    for k in range(n):
        t = 8 if k == 0 else (types[k % len(types)] if types else 6)
        cells.append(dict(type=t, size=float(size), mass=10.0 + (k % 5), rel=((k % cols) * size * squeeze, (k // cols) * size * squeeze),
                          parent=-1 if k == 0 else max(0, k - cols if k >= cols else k - 1)))
    return cells


def test_empty_world_ticks(oracle):
    w = World(default_params(n_org=8, n_dead=16, seed=2))
    wg, eng = engine_for(w)
    eng.tick(3, PH_ALL); eng.synchronize()
    assert oracle.tick(w, PH_ALL, 3) == 0
    compare_worlds(wg, w, "empty world", verbose=False)
    assert wg.tick == 3 and wg.stats()["org_ticks"] == 0


def test_food_only_world(oracle):
    w = World(default_params(n_org=4, n_dead=64, width=600.0, height=400.0, seed=2))
    rng = np.random.default_rng(0)
    for d in range(40):                                      # a pile in a corner: dead/dead and dead/wall contacts, levels > 1
        add_food(w, d, 25.0 + 20.0 * (d % 7) + rng.uniform(-3, 3), 25.0 + 20.0 * (d // 7) + rng.uniform(-3, 3), vx=rng.uniform(-40, 40), vy=rng.uniform(-40, 40))
    for step in range(4):
        wg, eng = engine_for(w)
        eng.tick(1, PH_STEP | PH_FRICTION | PH_POST); eng.synchronize()
        assert oracle.tick(w, PH_STEP | PH_FRICTION | PH_POST, 1) == 0
        compare_step(wg, w, f"food only step {step}")
    assert int(w.np("g_x_n")[0]) > 20


SIZES = [1, 2, 8, 9, 16, 17, 24, 25, 32, 33, 40, 41, 48, 49, 56, 57, 64]


def test_every_size_class_boundary(oracle, coupled_kernel):
    p = default_params(n_org=len(SIZES) + 2, n_dead=32, width=6000.0, height=2500.0, seed=7)
    w = World(p)
    x = 200.0
    for slot, n in enumerate(SIZES):
        cols = max(1, int(np.ceil(np.sqrt(n))))
        manual_org(w, slot, (x, 300.0 + 37.0 * (slot % 3)), blob(n, 30.0, cols, types=[6, 2, 3, 5, 7, 0, 1, 4]))
        x += cols * 27.0 + 80.0
    # two organisms pressed against walls (left wall, bottom-right corner), one on top of a neighbour (cross contacts -> coupled path)
    manual_org(w, len(SIZES), (14.0, 1200.0), blob(45, 28.0, 7, types=[6, 2]))
    manual_org(w, len(SIZES) + 1, (p.width - 120.0, p.height - 90.0), blob(20, 28.0, 5, types=[6, 3]))
    t = w.t
    t["c_pos"][MAXC * 5: MAXC * 5 + SIZES[5]] += torch.tensor([-60.0, 10.0])          # organism 5 overlaps organism 4
    t["c_spos"][MAXC * 5: MAXC * 5 + SIZES[5]] += torch.tensor([-60.0, 10.0])
    for slot, n in enumerate(SIZES):                         # died-in-place rows: the solver works on ranks, the lists on rows
        if n >= 9:
            for k in (3, n // 2, n - 2):
                t["c_alive"][slot * MAXC + k] = 0
    g = torch.Generator().manual_seed(5)
    live = t["c_alive"] == 1
    t["c_vel"][live] = (torch.rand(int(live.sum()), 2, generator=g) - 0.5) * 60.0
    t["c_angvel"][live] = (torch.rand(int(live.sum()), generator=g) - 0.5) * 4.0
    for d in range(12):
        add_food(w, d, 260.0 + 330.0 * d, 330.0, vx=0.0, vy=25.0)
    for k, slot in enumerate(range(len(SIZES) - 6, len(SIZES))):      # food lying on the largest organisms: every size class also takes the coupled path
        hx, hy = w.np("c_pos")[slot * MAXC]
        add_food(w, 12 + k, float(hx) + 40.0, float(hy) + 35.0, vx=-10.0, vy=5.0)
    assert oracle.post(w) == 0
    for step in range(4):
        wg, eng = engine_for(w)
        eng.tick(1, PH_STEP | PH_FRICTION | PH_POST); eng.synchronize()
        assert oracle.tick(w, PH_STEP | PH_FRICTION | PH_POST, 1) == 0
        assert wg.stats()["overflow"] == 0
        compare_step(wg, w, f"size classes step {step}")
    st = w.stats()
    assert st["xcontacts"] > 6 and st["icontacts"] > 300
    codes = w.np("ic_code").reshape(p.n_org, -1)[len(SIZES), : int(w.np("org_ic_n")[len(SIZES)])]
    assert ((codes & 127) >= MAXC).any(), "the organism at the left wall has wall contacts"
    for _ in range(3):                                       # and whole ticks (rules, think, train, settle, reproduce) on the same zoo
        wg, eng = engine_for(w)
        eng.tick(1, PH_ALL); eng.synchronize()
        assert oracle.tick(w, PH_ALL, 1) == 0
        compare_worlds(wg, w, f"size classes tick {w.tick}", verbose=False)


def test_push_direction_keeps_the_sign_of_a_zero_command(oracle):
    """A network output rounded to -0.0 stays -0.0 (neural.py:292, tolist()), and the push actuator takes math.atan2 of the two direction
    commands (sprites.py:1519-1544): atan2(+-0, -0) = +-pi, atan2(+-0, +0) = +-0 -- opposite pushes. Found at 100k organisms (one
    organism per tick has both commands at zero); here every combination of signed zeros and a few ordinary directions."""
    from test_gpu_tick import fast_world
    w = fast_world(oracle, 16, 17, 0.9, 0)
    assert oracle.tick(w, PH_ALL, 2) == 0
    z = [(0.0, 0.0), (-0.0, 0.0), (0.0, -0.0), (-0.0, -0.0), (-0.0, 0.5), (0.5, -0.0), (-0.5, -0.0), (-0.5, 0.0), (0.3, -0.4), (-0.0, -0.5)]
    orgs = np.nonzero(w.np("org_state") == 1)[0]
    pushers = 0
    for i, o in enumerate(orgs):
        dx, dy = z[i % len(z)]
        w.t["org_move"][o] = torch.tensor([0.0, -0.6 if i % 2 else 0.8, dx, dy])
        pushers += int(((w.np("c_type")[o * MAXC:(o + 1) * MAXC] == 5) & (w.np("c_alive")[o * MAXC:(o + 1) * MAXC] == 1)).sum())
    assert pushers >= 10
    w.t["org_think_tick"][:] = int(w.tick)                           # nobody thinks this tick: the commands above are the ones acted on
    wg, eng = engine_for(w)
    eng.tick(1, PH_ALL); eng.synchronize()
    assert oracle.tick(w, PH_ALL, 1) == 0
    compare_worlds(wg, w, "signed-zero push commands", verbose=True)

"""The reference-named host side (sapiogenesis_b200.physics / sprites / neural): surface on CPU, behaviour on the GPU.

The expected names and parameter lists below were read off the reference (physics.py:279-471, sprites.py:37-65,1809,
neural.py:91,369,426) and are written out here: nothing in the tests reads /root/reference."""
import inspect

import numpy as np
import pytest
import torch

from parity import compare_worlds
from sapiogenesis_b200.world import MAXC, PH_ALL
from worlds import make_world


def test_reference_entry_points_exist_with_the_reference_parameters():
    from sapiogenesis_b200 import neural, physics, sprites
    params = lambda f: list(inspect.signature(f).parameters)
    assert params(physics.Environment.__init__)[:8] == ["self", "worldView", "scene", "width", "height", "friction", "gravity", "frameWidth"]
    assert params(physics.Environment.update) == ["self", "event"]
    for name in ("preUpdateEvent", "postUpdateEvent", "pause_world", "resume_world", "removeSprite", "add_ball_sprite", "add_sprite", "collision", "stop"):
        assert callable(getattr(physics.Environment, name))
    assert params(physics.Environment.add_ball_sprite)[:7] == ["self", "xy", "width", "height", "mass", "friction", "elasticity"]
    assert params(physics.setupEnvironment)[:4] == ["worldView", "scene", "friction", "gravity"]
    assert params(physics.Environment.add_sprite) == ["self", "sprite"] and params(physics.Environment.collision) == ["self", "arbiter", "space", "data"]
    assert params(physics.Sprite.connectTo) == ["self", "sprite", "rigid"] and params(physics.Sprite.disconnectFrom) == ["self", "sprite"]
    for name in ("getPos", "getCenter", "getWidth", "getHeight", "bump", "updateSprite", "collision", "connectTo", "disconnectFrom"):
        assert callable(getattr(physics.Sprite, name))
    assert params(sprites.update_organisms) == ["environment"]
    assert params(sprites.disperse_all_dead) == ["environment"]
    knobs = dict(co2_conversion_ratio=3, cell_dispersion_rate=6, reproduction_limit=6, offspring_amount=1, heal_rate=4, mass_cutoff=1,
                 break_damage=5, starvation_rate=6, neural_update_delay=.5, learning_update_delay=.5, training_epochs=1,
                 max_push_speed=250, max_rotation_speed=6, age_limit=200, eyesight_multiplier=10, olfactory_multiplier=15)
    for k, v in knobs.items():
        assert getattr(sprites, k) == v, k
    for name in ("alive", "kill", "total_energy", "max_energy", "health_percent", "energy_percent"):
        assert callable(getattr(sprites.Organism, name))
    assert params(neural.activate) == ["network", "environment", "organism", "uDiff"]
    assert params(neural.train_network) == ["organism", "epochs"]
    assert params(neural.setup_network) == ["dna", "learning_rate", "rnn", "hiddenSize"]


def test_environment_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from sapiogenesis_b200 import physics
    from sapiogenesis_b200._lib import SapioError
    with pytest.raises((SapioError, RuntimeError, AssertionError)):
        physics.Environment(None, None, 3180, 1800, device="cuda")
    with pytest.raises(SapioError):
        physics.Environment(None, None, 3180, 1800, device="cpu")


def _env_from(w):
    """An Environment wrapped around an existing world (state identical to `w`)."""
    from sapiogenesis_b200 import physics
    from sapiogenesis_b200.engine import Engine
    env = physics.Environment.__new__(physics.Environment)
    physics.Environment.__init__(env, None, None, w.p.width, w.p.height, capacity=w.p.n_org, dead_capacity=w.p.n_dead)
    env.world = w.to("cuda")
    env.engine = Engine(env.world)
    return env


@pytest.mark.gpu
def test_update_loop_is_the_fused_tick(oracle):
    """env.preUpdateEvent = update_organisms; env.update() per timer event (Sapiogenesis.py:687, physics.py:425) is
    bit-identical to sapio_tick on the same state, and tracks the oracle."""
    from sapiogenesis_b200 import sprites
    from sapiogenesis_b200.engine import Engine
    w = make_world(oracle, n_org=24, seed=21, n_food=60, crowd=0.45, kick=30.0, extra_slots=8, n_dead_cap=64 * 32)
    w.p.think_ticks, w.p.train_ticks, w.p.repro_ticks = 3, 3, 8
    saved = (sprites.neural_update_delay, sprites.learning_update_delay, sprites.reproduction_limit)
    sprites.neural_update_delay, sprites.learning_update_delay, sprites.reproduction_limit = 3 * 0.03, 3 * 0.03, 8 * 0.03
    try:
        env = _env_from(w)
        env.preUpdateEvent = lambda: sprites.update_organisms(env)
        wf = w.to("cuda"); ef = Engine(wf)
        for _ in range(12):
            env.update(None)
        ef.tick(12); ef.synchronize(); env.engine.synchronize()
        assert (env.world.p.think_ticks, env.world.p.repro_ticks) == (3, 8)
        for name in wf.t:
            a, b = env.world.t[name], wf.t[name]
            assert torch.equal(a, b) or torch.equal(torch.nan_to_num(a.double()), torch.nan_to_num(b.double())), name
        assert env.world.tick == w.tick + 12
        assert env.info["population"] == int((wf.t["org_state"] == 1).sum())
        assert len(env.info["organism_list"]) == env.info["population"]
        org = env.info["organism_list"][0]
        assert org.alive() and 0.0 <= org.energy_percent() <= 100.0 and len(org.cells) == int((wf.t["c_alive"][org.slot * MAXC:(org.slot + 1) * MAXC] == 1).sum())
        env.pause_world(); t0 = env.world.tick; env.update(None); assert env.world.tick == t0
    finally:
        sprites.neural_update_delay, sprites.learning_update_delay, sprites.reproduction_limit = saved


@pytest.mark.gpu
def test_activate_and_train_network_on_demand(oracle):
    """neural.activate / neural.train_network for one organism outside the cadence == the oracle's."""
    from sapiogenesis_b200 import neural
    w = make_world(oracle, n_org=12, seed=8, n_food=40, crowd=0.45, kick=30.0)
    w.p.think_ticks, w.p.train_ticks = 3, 3
    assert oracle.tick(w, PH_ALL, 40) == 0
    env = _env_from(w)
    orgs = [o for o in env.info["organism_list"]][:4]
    for org in orgs:
        out = neural.activate(org.brain, env, org, 0.03)
        assert oracle.think_now(w, org.slot) == 0
        np.testing.assert_allclose(out, w.np("org_move")[org.slot], rtol=0, atol=2e-3)
    env.engine.synchronize()
    compare_worlds(env.world, w, "activate on demand", verbose=False)
    for org in orgs:
        neural.train_network(org, epochs=1)
        assert oracle.train_one(w, org.slot) == 0
    env.engine.synchronize()
    compare_worlds(env.world, w, "train_network on demand", verbose=False)
    net = orgs[0].brain
    dims = net.structure
    assert [tuple(W.shape) for W, _ in net.layers()] == [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


@pytest.mark.gpu
def test_manual_controls_and_headless_runner(oracle, tmp_path):
    """heal / feed / hurt / kill / feed_all / kill_all / reset (Sapiogenesis.py:209-242, 396-412) and the headless runner."""
    from sapiogenesis_b200 import sprites
    w = make_world(oracle, n_org=8, seed=5, n_food=10, crowd=0.8, extra_slots=4, n_dead_cap=64 * 16)
    assert oracle.tick(w, PH_ALL, 3) == 0
    env = _env_from(w)
    t = env.world.t
    org = env.info["organism_list"][0]
    rows = slice(org.slot * MAXC, org.slot * MAXC + int(t["org_ncells"][org.slot]))
    org.hurt()
    assert org.health_percent() < 95.0
    org.heal(); org.feed()
    assert abs(org.health_percent() - 100.0) < 1e-3 and abs(org.energy_percent() - 100.0) < 1e-3
    org.immortal = True; org.maladaptive = True; org.finite_memory = False
    assert org.immortal and org.maladaptive and not org.finite_memory
    org.immortal = False
    org.kill()
    sprites.update_organisms(env)
    assert not org.alive() and int((t["c_alive"][rows] == 1).sum()) == 0 and int((t["d_state"] > 0).sum()) > 10
    sprites.feed_all(env)
    live = t["c_alive"] == 1
    assert torch.equal(t["c_energy"][live], t["c_storage"][live])
    sprites.reset(env)
    assert env.info["population"] == 0 and len(env.sprites) == 0 and env.info["co2"] == 30000 and env.info["oxygen"] == 0
    from sapiogenesis_b200 import run
    out = tmp_path / "frame.npz"
    run.main(["--organisms", "6", "--ticks", "20", "--every", "10", "--dump", str(out)])
    z = np.load(out)
    assert z["cells"].shape[1] == 4 and len(z["cells"]) > 6


@pytest.mark.gpu
def test_sprite_joints_add_sprite_and_collision_hooks(oracle):
    """Environment.add_sprite (physics.py:453), Sprite.connectTo / disconnectFrom / connections (:222-248), Environment.collision
    (:379-386): the joints of an organism are its living-cell set, free bodies come and go through add_sprite / removeSprite."""
    from sapiogenesis_b200 import physics
    w = make_world(oracle, n_org=6, seed=9, n_food=6, crowd=0.9, extra_slots=2, n_dead_cap=64 * 8)
    env = _env_from(w)
    t = env.world.t
    org = env.info["organism_list"][0]
    base = org.slot * MAXC
    live = [base + k for k in range(int(t["org_ncells"][org.slot])) if int(t["c_alive"][base + k]) == 1]
    a, b = physics.Sprite(env, "cell", live[0]), physics.Sprite(env, "cell", live[1])
    conn = a.connections
    assert len(conn) == len(live) - 1 and b in conn and a not in conn            # all-pairs pins among the living cells (sprites.py:1721-1726)
    a.connectTo(b)                                                               # already connected: the reference returns early too
    other = physics.Sprite(env, "cell", env.info["organism_list"][1].slot * MAXC)
    with pytest.raises(NotImplementedError):
        a.connectTo(other)
    with pytest.raises(NotImplementedError):
        a.disconnectFrom(b)
    a.disconnectFrom(other)                                                      # not connected: nothing to do (physics.py:244)
    food = env.add_ball_sprite((400.0, 400.0), 30, 30, mass=12)
    assert food.connections == {} and env.collision(None, None, None) is True and food.collision(a) is None
    n0 = len(env.sprites)
    env.removeSprite(food)
    assert len(env.sprites) == n0 - 1
    assert env.add_sprite(food) is food and len(env.sprites) == n0                # back in the space with the state its row holds
    env.update(None)
    assert int(t["d_state"][food._i]) > 0


@pytest.mark.gpu
def test_use_rnn_switch_builds_recurrent_brains(oracle):
    """Environment.info["use_rnn"] / ["hidden_rnn_size"] (physics.py:305,308; the UI checkbox, Sapiogenesis.py:381-386): organisms created
    while it is set get a networks.RNNetwork -- input layer widened by the hidden size, hidden head last (networks.py:111-166) -- the
    others keep their feed-forward brain; both kinds think and train in the same world."""
    from sapiogenesis_b200 import neural, physics
    env = physics.Environment(None, None, 3180, 1800, capacity=16)
    a = env.add_creature((600.0, 600.0))
    env.info["hidden_rnn_size"] = 5
    env.info["use_rnn"] = True
    assert env.world.p.rnn_cap == 5 and env.world.p.rnn_hidden == 5 and env.world.t["org_hm"].shape[1] == 90 * 5
    b = env.add_creature((1600.0, 900.0))
    env.info["use_rnn"] = False
    c = env.add_creature((2400.0, 600.0))
    assert (a.brain.hiddenSize, b.brain.hiddenSize, c.brain.hiddenSize) == (0, 5, 0) and b.brain.is_rnn
    dims = b.brain.structure
    shapes = [tuple(W.shape) for W, _ in b.brain.layers()]
    assert shapes == [(dims[1], dims[0] + 5)] + [(dims[i + 1], dims[i]) for i in range(1, len(dims) - 1)] + [(5, dims[-2])]
    assert neural.setup_network(b.dna, rnn=True, hiddenSize=5).is_rnn
    with pytest.raises(ValueError):
        neural.setup_network(a.dna, rnn=True)
    with pytest.raises(ValueError):
        env.info["hidden_rnn_size"] = 7; env.info["use_rnn"] = True                # the size is fixed once RNN brains exist
    env.info["hidden_rnn_size"] = 5; env.info["use_rnn"] = False
    env.world.p.think_ticks = env.world.p.train_ticks = 3; env.engine.refresh()
    assert float(b.brain.lastHidden.abs().sum()) == 0.0
    env.engine.tick(40); env.engine.synchronize()
    assert float(b.brain.lastHidden.abs().sum()) > 0.0 and env.world.stats()["thinks"] >= 30
    host = env.world.to("cpu")
    env.engine.tick(3); env.engine.synchronize()
    assert oracle.tick(host, PH_ALL, 3) == 0
    compare_worlds(env.world, host, "three ticks of a mixed RNN / feed-forward world", verbose=False)


def test_recurrent_state_arrays_are_reserved_on_demand():
    """A world built without RNN brains carries empty recurrent-state arrays (per = RNNH of include/sapio_fields.def); World.reserve_rnn gives
    them their columns once -- what Environment.info["use_rnn"] = True does the first time -- and the hidden size is fixed from then on."""
    from sapiogenesis_b200.world import FIELDS, HM, MEMROWS, RNN_MAXH, World, default_params
    w = World(default_params(n_org=8, n_dead=8))
    rnn_fields = [f.name for f in FIELDS if "RNNH" in f.per]
    assert set(rnn_fields) == {"org_last_hidden", "org_hm", "org_th"} and all(w.t[n].numel() == 0 for n in rnn_fields)
    assert w.p.rnn_cap == 0 and w.p.rnn_hidden == 0 and w.t["org_rnn_h"].shape == (8,) and w.t["org_mem_seq"].shape == (8, MEMROWS)
    w.reserve_rnn(6)
    assert w.p.rnn_cap == 6 and w.t["org_last_hidden"].shape == (8, 6) and w.t["org_hm"].shape == (8, HM * 6) and w.t["org_th"].shape == (8, MEMROWS * 6)
    w.reserve_rnn(6)                                         # idempotent
    with pytest.raises(ValueError):
        w.reserve_rnn(5)
    with pytest.raises(ValueError):
        World(default_params(n_org=8, n_dead=8)).reserve_rnn(RNN_MAXH + 1)
    s = w.struct()                                           # the C struct sees the new buffers
    assert s.p.rnn_cap == 6 and s.org_hm == w.t["org_hm"].data_ptr()
    w2 = World(default_params(n_org=8, n_dead=8, rnn_hidden=4))
    assert w2.p.rnn_cap == 4 and w2.p.rnn_hidden == 4 and w2.t["org_th"].shape == (8, MEMROWS * 4)

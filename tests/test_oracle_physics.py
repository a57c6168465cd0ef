"""Closed-form checks of the oracle's Space.step restatement (oracle/space_step.c). The Chipmunk boundary itself is
UNPINNED (no pymunk, no Chipmunk source, no reference tests): these cases pin the restated algorithm to physics, not to
Chipmunk's bits. SURVEY.md section 8(c) item (4)."""
import math

import numpy as np
import pytest
import torch

from sapiogenesis_b200.world import MAXC, T_BODY, T_EYE, T_HEART, World, default_params
from worlds import add_food, manual_org

PH_STEP = 16


def small_world(**kw):
    w = World(default_params(n_org=4, n_dead=16, **kw))
    return w


def test_head_on_elastic_equal_masses_swap_velocities(oracle):
    w = small_world()
    add_food(w, 0, 500.0, 500.0, vx=30.0, size=30, mass=10)
    add_food(w, 1, 528.0, 500.0, vx=-30.0, size=30, mass=10)     # overlapping by 2 px
    w.t["d_elast"][:2] = 1.0
    w.t["d_fric"][:2] = 0.0
    assert oracle.space_step(w) == 0
    v = w.np("d_vel")
    np.testing.assert_allclose(v[0], [-30.0, 0.0], atol=1e-4)
    np.testing.assert_allclose(v[1], [30.0, 0.0], atol=1e-4)
    # penetration is corrected through the bias velocity only (no energy injected into v)
    vb = w.np("d_vbias")
    assert vb[0][0] < 0 < vb[1][0]
    keys = w.np("x_key")[: int(w.np("g_x_n")[0])]
    nc = w.p.n_org * MAXC
    assert list(keys) == [((nc + 0) << 32) | (nc + 1)]


def test_head_on_inelastic_sticks(oracle):
    w = small_world()
    add_food(w, 0, 500.0, 500.0, vx=40.0, size=30, mass=10)
    add_food(w, 1, 528.0, 500.0, vx=0.0, size=30, mass=30)
    w.t["d_elast"][:2] = 0.0
    w.t["d_fric"][:2] = 0.0
    oracle.space_step(w)
    v = w.np("d_vel")
    np.testing.assert_allclose(v[0][0], v[1][0], atol=1e-4)          # common normal velocity
    np.testing.assert_allclose(10 * v[0][0] + 30 * v[1][0], 400.0, rtol=1e-5)   # momentum


def test_wall_bounce_uses_product_of_elasticities(oracle):
    w = small_world()
    add_food(w, 0, 30.0, 900.0, vx=-50.0, size=30, mass=10)         # left wall face at x = 20; radius 15
    w.t["d_elast"][0] = 0.8
    w.t["d_fric"][0] = 0.0
    oracle.space_step(w)
    v = w.np("d_vel")[0]
    np.testing.assert_allclose(v[0], 50.0 * 0.8 * 0.95, rtol=1e-5)   # walls: elasticity .95 (physics.py:341)
    np.testing.assert_allclose(v[1], 0.0, atol=1e-6)


def test_friction_decay(oracle):
    w = small_world()
    add_food(w, 0, 1000.0, 900.0, vx=100.0, vy=-20.0)
    w.t["d_angvel"][0] = 3.0
    for _ in range(50):
        oracle.friction(w)
    f = (1 - 0.6 * 0.030) ** 50
    np.testing.assert_allclose(w.np("d_vel")[0], [100.0 * f, -20.0 * f], rtol=1e-5)
    np.testing.assert_allclose(w.np("d_angvel")[0], 3.0 * f, rtol=1e-5)


def test_pinned_pair_keeps_rest_length_and_momentum(oracle):
    w = small_world()
    manual_org(w, 0, (1000.0, 900.0), [dict(type=T_HEART, size=20, mass=10), dict(type=T_BODY, size=20, mass=5, rel=(40.0, 0.0))])
    w.t["c_vel"][1] = torch.tensor([0.0, 60.0])
    p0 = 5 * np.array([0.0, 60.0])
    for _ in range(200):
        assert oracle.space_step(w) == 0
    pos = w.np("c_pos")
    # Baumgarte bias removes 31.6 % of the error per step; the spinning pair keeps a small centrifugal stretch
    assert abs(np.linalg.norm(pos[1] - pos[0]) - 40.0) < 1.0
    v = w.np("c_vel")
    np.testing.assert_allclose(10 * v[0] + 5 * v[1], p0, rtol=1e-4, atol=1e-3)


def test_sensor_body_is_dragged_by_its_zero_length_pin(oracle):
    w = small_world()
    manual_org(w, 0, (1000.0, 900.0), [dict(type=T_HEART, size=20, mass=10), dict(type=T_EYE, size=20, mass=8, rel=(40.0, 0.0))])
    w.t["c_vel"][:2] = torch.tensor([25.0, 0.0])
    for _ in range(100):
        oracle.space_step(w)
    cell, sens = w.np("c_pos")[1], w.np("c_spos")[1]
    assert np.linalg.norm(cell - sens) < 1.0                       # trails by the joint error only
    assert w.np("c_spos")[1][0] > 1040.0 + 50.0                    # and it did move
    # momentum of cells + sensor (mass 1) is conserved by the step
    v, sv = w.np("c_vel"), w.np("c_svel")
    np.testing.assert_allclose(10 * v[0] + 8 * v[1] + 1 * sv[1], [18 * 25.0, 0.0], rtol=1e-4, atol=1e-3)


def test_first_step_has_no_warm_start_and_persistent_contacts_keep_impulses(oracle):
    w = small_world()
    add_food(w, 0, 500.0, 500.0, size=30, mass=10)
    add_food(w, 1, 525.0, 500.0, size=30, mass=10)                  # resting overlap: bias pushes apart over time
    oracle.space_step(w)
    assert int(w.np("g_stepped")[0]) == 1
    jn1 = float(w.np("x_jn")[0])
    oracle.space_step(w)
    assert int(w.np("g_x_n")[0]) == 1 and float(w.np("x_jn")[0]) >= 0.0 and jn1 >= 0.0


def test_contact_adjacency_lists_both_directions(oracle):
    w = small_world()
    add_food(w, 0, 500.0, 500.0); add_food(w, 1, 525.0, 500.0); add_food(w, 2, 512.0, 520.0)
    oracle.space_step(w)
    nc = w.p.n_org * MAXC
    off, other = w.np("xa_off"), w.np("xa_other").reshape(-1)
    adj = {b: [int(v) for v in other[off[nc + b]: off[nc + b + 1]]] for b in range(3)}
    assert adj == {0: [nc + 1, nc + 2], 1: [nc + 0, nc + 2], 2: [nc + 0, nc + 1]}

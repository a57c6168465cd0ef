"""GPU parity of Space.step + friction (sapio_space_step / sapio_friction through the C-ABI) against the oracle,
single step from identical state, re-synchronised after every step so each comparison starts from the same bytes.
Bit-exact: integrated positions, contact-pair sets (cross list, adjacency, per-organism intra lists).
fp32 tolerance (north star): 1e-5 relative to the field's magnitude for velocities, bias velocities and impulses."""
import numpy as np
import pytest
import torch

from sapiogenesis_b200.world import MAXC, NPAIR, ICAP
from worlds import make_world, live_rows

pytestmark = pytest.mark.gpu

TOL = 1e-5


def field_err(got, ref, mask=None, floor=1e-3, elementwise=False):
    from parity import field_err as fe
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    if mask is not None:
        got, ref = got[mask], ref[mask]
    return fe(got, ref, floor, elementwise)


def compare_step(wg, w, tag):
    rows = live_rows(w)
    dead = np.nonzero(w.np("d_state") > 0)[0]
    # bit-exact state
    for name in ("c_pos", "c_spos", "c_ang"):
        np.testing.assert_array_equal(wg.np(name)[rows], w.np(name)[rows], err_msg=f"{tag} {name}")
    for name in ("d_pos", "d_ang"):
        np.testing.assert_array_equal(wg.np(name)[dead], w.np(name)[dead], err_msg=f"{tag} {name}")
    nx = int(w.np("g_x_n")[0])
    assert int(wg.np("g_x_n")[0]) == nx, f"{tag}: cross contact count {int(wg.np('g_x_n')[0])} != {nx}"
    np.testing.assert_array_equal(wg.np("x_key")[:nx], w.np("x_key")[:nx], err_msg=f"{tag} cross contact pairs")
    np.testing.assert_array_equal(wg.np("xa_off"), w.np("xa_off"), err_msg=f"{tag} adjacency offsets")
    na = int(w.np("xa_off")[-1])
    np.testing.assert_array_equal(wg.np("xa_other").reshape(-1)[:na], w.np("xa_other").reshape(-1)[:na])
    orgs = np.nonzero(w.np("org_state") == 1)[0]
    np.testing.assert_array_equal(wg.np("org_ic_n")[orgs], w.np("org_ic_n")[orgs], err_msg=f"{tag} intra contact counts")
    icm = np.zeros(w.p.n_org * ICAP, dtype=bool)
    for o in orgs:
        icm[o * ICAP: o * ICAP + int(w.np("org_ic_n")[o])] = True
    np.testing.assert_array_equal(wg.np("ic_code")[icm], w.np("ic_code")[icm], err_msg=f"{tag} intra contact pairs")
    # fp32 tolerance
    errs = {}
    for name in ("c_vel", "c_angvel", "c_vbias", "c_svel"):
        errs[name] = field_err(wg.np(name)[rows], w.np(name)[rows])
    from parity import velocity_rms
    errs["c_spin_jn"] = field_err(wg.np("c_spin_jn")[rows], w.np("c_spin_jn")[rows], floor=max(1e-3, velocity_rms(w)), elementwise=True)   # sensor body: mass 1
    for name in ("d_vel", "d_angvel", "d_vbias"):
        errs[name] = field_err(wg.np(name)[dead], w.np(name)[dead])
    # circle contacts have r x n == 0: the angular bias velocity is identically zero (the oracle's general r x j form
    # leaves ~1e-17 of rounding noise); compare on the scale of a bias velocity over a cell radius
    errs["c_wbias"] = field_err(wg.np("c_wbias")[rows], w.np("c_wbias")[rows], floor=1.0)
    errs["d_wbias"] = field_err(wg.np("d_wbias")[dead], w.np("d_wbias")[dead], floor=1.0)
    jm = np.zeros(w.p.n_org * NPAIR, dtype=bool)
    alive = w.np("c_alive")
    for o in orgs:
        nc = int(w.np("org_ncells")[o])
        for i in range(nc):
            for j in range(i + 1, nc):
                if alive[o * MAXC + i] == 1 and alive[o * MAXC + j] == 1:
                    jm[o * NPAIR + i * (2 * MAXC - i - 1) // 2 + (j - i - 1)] = True
    from parity import momentum_scale
    pf = max(1e-3, momentum_scale(w))          # impulse accumulators: see tests/parity.py
    errs["j_jn"] = field_err(wg.np("j_jn")[jm], w.np("j_jn")[jm], floor=pf, elementwise=True)
    errs["ic_jn"] = field_err(wg.np("ic_jn")[icm], w.np("ic_jn")[icm], floor=pf, elementwise=True)
    errs["ic_jt"] = field_err(wg.np("ic_jt")[icm], w.np("ic_jt")[icm], floor=pf, elementwise=True)
    errs["x_jn"] = field_err(wg.np("x_jn")[:nx], w.np("x_jn")[:nx], floor=pf, elementwise=True)
    errs["x_jt"] = field_err(wg.np("x_jt")[:nx], w.np("x_jt")[:nx], floor=pf, elementwise=True)
    print(f"{tag}: nx={nx} nic={int(icm.sum())} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    bad = {k: v for k, v in errs.items() if not v <= TOL}
    assert not bad, f"{tag}: beyond {TOL}: {bad}"


@pytest.mark.parametrize("n_org,crowd,food,seed", [(12, 1.0, 0, 1), (12, 1.0, 40, 2), (30, 0.35, 80, 3), (64, 0.5, 200, 4)])
def test_space_step_matches_oracle(oracle, n_org, crowd, food, seed):
    from sapiogenesis_b200.engine import Engine
    w = make_world(oracle, n_org=n_org, seed=seed, n_food=food, crowd=crowd)
    for step in range(4):
        wg = w.to("cuda")
        eng = Engine(wg)
        eng.space_step(); eng.friction(); eng.synchronize()
        rc = oracle.space_step(w); oracle.friction(w)
        assert rc == 0
        assert wg.stats()["overflow"] == 0
        compare_step(wg, w, f"n={n_org} crowd={crowd} step={step}")
        w.t["g_tick"] += 1          # bodies created at tick 0 stop being "fresh": cached impulses apply from now on


def test_many_steps_stay_close(oracle):
    """Trajectories are chaotic; over 50 free-running steps the two must still agree to ~1e-3 in this gentle world."""
    from sapiogenesis_b200.engine import Engine
    w = make_world(oracle, n_org=12, seed=7, n_food=0, crowd=1.0, kick=10.0)
    wg = w.to("cuda")
    eng = Engine(wg)
    for _ in range(50):
        eng.space_step(); eng.friction()
        oracle.space_step(w); oracle.friction(w)
    eng.synchronize()
    rows = live_rows(w)
    assert field_err(wg.np("c_pos")[rows] - w.np("c_pos")[rows] + 1.0, np.ones_like(w.np("c_pos")[rows])) < 1e-2
    assert field_err(wg.np("c_vel")[rows], w.np("c_vel")[rows]) < 1e-2


def test_tick_physics_only_advances_clock(oracle):
    from sapiogenesis_b200.engine import Engine
    w = make_world(oracle, n_org=6, seed=9, n_food=10)
    wg = w.to("cuda")
    eng = Engine(wg)
    eng.tick(3, phases=16 | 32 | 64)
    eng.synchronize()
    assert wg.tick == 3 and int(wg.np("g_population")[0]) == 6

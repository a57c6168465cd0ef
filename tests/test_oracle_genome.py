"""Pins the oracle's genome code (oracle/genome.c, oracle/world.c) against the reference's own
sprites.DNA.randomize / DNA.mutate / Organism.build_body, run on identical uniforms
(fixtures: tests/golden/genome_cases.npz, made by tests/golden/make_golden.py)."""
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from sapiogenesis_b200.world import MAXC, World, default_params

CASES = np.load(os.path.join(GOLDEN, "genome_cases.npz"))
N_CASES, SEED = int(CASES["meta"][0]), int(CASES["meta"][1])


def u24(tag, n):
    rng = random.Random(tag)
    return np.asarray([rng.getrandbits(24) / 16777216.0 for _ in range(n)], dtype=np.float32)


def check_rows(w, slot, g, prefix, hw_defined=None):
    order = g(prefix + ".order")
    ids = list(g(prefix + ".ids"))
    nc = int(w.np("org_ncells")[slot])
    assert nc == len(ids) == len(order)
    rows = slice(slot * MAXC, slot * MAXC + nc)
    assert list(w.np("c_id")[rows]) == list(order), "Organism.cells creation order"
    gi = [ids.index(c) for c in order]                      # genome index of each row
    assert list(w.np("c_gorder")[rows]) == gi
    for name, key in (("c_type", "type"),):
        assert list(w.np(name)[rows]) == [int(g(prefix + "." + key)[j]) for j in gi]
    for name, key in (("c_size", "size"), ("c_mass", "mass"), ("c_elast", "elast"), ("c_fric", "fric"),
                      ("c_maxhealth", "maxhealth"), ("c_use_idle", "use_idle"), ("c_use_act", "use_act"),
                      ("c_storage", "storage"), ("c_damage", "damage")):
        ref = np.asarray([g(prefix + "." + key)[j] for j in gi])
        np.testing.assert_allclose(w.np(name)[rows], ref.astype(np.float32), rtol=2e-7, atol=0, err_msg=name)
    rel = np.asarray([g(prefix + ".rel")[j] for j in gi])
    np.testing.assert_allclose(w.np("c_rel")[rows], rel, rtol=1e-6, atol=1e-5)
    grow = np.asarray([g(prefix + ".grow")[j] for j in gi])
    np.testing.assert_allclose(w.np("c_grow")[rows], grow, rtol=1e-6, atol=1e-6)
    # parent / mirror as row indices
    par = g(prefix + ".parent"); mir = g(prefix + ".mirror")
    exp_par = [(-1 if par[j] == 0 else list(order).index(par[j])) for j in gi]
    exp_mir = [(-1 if mir[j] == 0 else list(order).index(mir[j])) for j in gi]
    assert list(w.np("c_parent")[rows]) == exp_par
    assert list(w.np("c_mirror")[rows]) == exp_mir
    assert int(w.np("c_parent")[slot * MAXC]) == -1, "row 0 is the first cell (heart)"
    # spawn state: build_sprite (sprites.py:282-283)
    np.testing.assert_array_equal(w.np("c_health")[rows], w.np("c_maxhealth")[rows])
    np.testing.assert_allclose(w.np("c_energy")[rows], w.np("c_storage")[rows] / 2, rtol=1e-7)
    pos = g(prefix + ".pos")
    np.testing.assert_allclose(w.np("c_pos")[rows], pos[None, :] + rel, rtol=0, atol=2e-4)
    np.testing.assert_array_equal(w.np("c_spos")[rows], w.np("c_pos")[rows])
    fl = g(prefix + ".flags")
    flags = int(w.np("org_flags")[slot])
    assert bool(flags & 8) == bool(fl[0]) and bool(flags & 16) == bool(fl[1])
    assert int(w.np("org_generation")[slot]) == int(fl[2])
    np.testing.assert_allclose(w.np("org_curiosity")[slot], g(prefix + ".curiosity")[0], rtol=2e-7)
    # network wiring + parameters
    hidden = list(g(prefix + ".hidden"))
    nl = int(w.np("org_nlayers")[slot])
    assert nl == len(hidden) + 1
    ldim = list(w.np("org_ldim")[slot][: nl + 1])
    n_eye = sum(1 for j in range(len(ids)) if g(prefix + ".type")[j] == 2)
    n_olf = sum(1 for j in range(len(ids)) if g(prefix + ".type")[j] == 3)
    assert ldim == [8 * n_eye + n_olf + 4] + hidden + [4]
    incell = [int(order[k]) for k in w.np("org_incell")[slot][: int(w.np("org_nincell")[slot])]]
    assert incell == list(g(prefix + ".input_cells"))
    brain = g(prefix + ".brain")
    mask = np.ones(len(brain), dtype=bool)
    if hw_defined is not None:              # child: inherited rows whose resize_ tail is uninitialised memory in the reference
        off, h = 0, 0
        for l in range(nl):
            n = ldim[l + 1] * ldim[l]
            if 1 <= l < nl - 1:
                mask[off:off + n] = hw_defined[h:h + n]
                h += n
            off += n + ldim[l + 1]
    got = w.np("org_brain")[slot][: len(brain)]
    np.testing.assert_array_equal(got[mask], brain[mask])
    assert np.all((got[~mask] >= 0) & (got[~mask] < 1))
    return nc


@pytest.mark.parametrize("ci", range(N_CASES))
def test_randomize_build_and_mutate_match_reference(oracle, ci):
    g = lambda k: CASES[f"c{ci}.{k}" if not k.startswith("c") else k]
    gp = lambda k: CASES[f"c{ci}.{k}"]
    tag = f"genome-{SEED}-{ci}"
    w = World(default_params(n_org=4, n_dead=8))
    used = gp("parent.used")
    ppos = gp("parent.pos")
    rc, consumed = oracle.spawn_arr(w, 0, ppos[0], ppos[1], u24(tag + "-u", used[0] + 8), u24(tag + "-ui", used[1] + 8),
                                    u24(tag + "-ub", used[2] + 8))
    assert rc == 0 and consumed[3] == 0
    assert consumed[:3] == [int(used[0]), int(used[1]), int(used[2])], "same number of draws as the reference"
    check_rows(w, 0, gp, "parent")
    pct = gp("parent.pct")
    np.testing.assert_allclose([w.np("org_last_health")[0], w.np("org_last_energy")[0]], pct, rtol=1e-6)
    rect, size = oracle.dna_rect(w, 0)
    np.testing.assert_allclose(rect, gp("parent.rect"), rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(size, gp("parent.size_wh"), rtol=1e-6, atol=1e-5)

    cused = gp("child.used")
    cpos = gp("child.pos")
    rc, consumed = oracle.mutate_arr(w, 0, 1, cpos[0], cpos[1], u24(tag + "-um", cused[0] + 8), u24(tag + "-uf", cused[1] + 8),
                                     u24(tag + "-ub2", cused[2] + 8))
    if len(gp("child.ids")) > MAXC or len(gp("child.order")) == 0:
        assert rc < 0
        return
    assert rc == 0 and consumed[3] == 0
    assert consumed[:3] == [int(cused[0]), int(cused[1]), int(cused[2])]
    check_rows(w, 1, gp, "child", gp("child.hw_defined"))
    rect, size = oracle.dna_rect(w, 1)
    np.testing.assert_allclose(rect, gp("child.rect"), rtol=1e-6, atol=1e-5)

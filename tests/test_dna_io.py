"""Genome exchange in the reference's own format (sapiogenesis_b200/dna_io.py).

Round trips on the CPU (the module is host-side plumbing over the world's tensors), and -- in the build container, where
/root/reference exists -- against the real `sprites.DNA`: a genome exported here unpickles there as a working DNA object, and a
DNA the reference randomised and pickled imports here with identical tables."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from sapiogenesis_b200 import dna_io
from sapiogenesis_b200.world import MAXC, PH_ALL, World, default_params
from worlds import make_world

REF = os.environ.get("SAPIO_REFERENCE", "/root/reference")


def grown_world(oracle):
    w = make_world(oracle, n_org=6, seed=13, n_food=12, crowd=0.7, kick=20.0, extra_slots=6)
    w.p.think_ticks = w.p.train_ticks = 3
    assert oracle.tick(w, PH_ALL, 40) == 0                       # memories, history, short-term memory exist
    return w


def same_genome(a, b):
    assert list(a["cells"]) == list(b["cells"])
    for cid in a["cells"]:
        ca, cb = a["cells"][cid], b["cells"][cid]
        assert ca.keys() == cb.keys()
        for k in ca:
            assert ca[k] == pytest.approx(cb[k], rel=0, abs=0), (cid, k)
    assert a["growth_pattern"] == b["growth_pattern"]
    assert a["brain_structure"]["hidden_layers"] == b["brain_structure"]["hidden_layers"]
    assert a["brain_structure"]["hidden_weights"] == b["brain_structure"]["hidden_weights"]
    assert a["base_info"] == b["base_info"] and a["generation"] == b["generation"]


def test_export_pickle_import_round_trip(oracle):
    w = grown_world(oracle)
    slot = int(np.nonzero(w.np("org_mem_n") > 0)[0][0])
    dna = dna_io.export_dna(w, slot)
    assert len(dna["cells"]) == int(w.np("org_ncells")[slot]) and len(dna["trainingInput"]) == int(w.np("org_mem_n")[slot])
    assert len(dna["previousInputs"]) == min(10, int(w.np("org_hist_n")[slot])) and len(dna["stim_history"]) == 10
    blob = dna_io.dumps_reference_pickle(dna)
    assert b"sprites" in blob and b"DNA" in blob and "sprites" not in sys.modules or True
    back = dna_io.loads_reference_pickle(blob)
    same_genome(dna, back)
    free = int(np.nonzero(w.np("org_state") == 0)[0][0])
    uid0 = int(w.np("g_next_uid")[0])
    assert dna_io.import_dna(w, free, (1500.0, 900.0), back) == uid0
    again = dna_io.export_dna(w, free)
    same_genome(dna, again)
    for k in ("trainingInput", "trainingOutput", "trainingReward", "short_term_memory", "previousOutputs", "previousDopamine", "stim_history"):
        assert again[k] == dna[k], k
    assert all(torch.equal(x, y) for x, y in zip(again["previousInputs"], dna["previousInputs"]))
    # the materialised organism is what build_body makes: heart at pos, rows in creation order, energy = storage / 2, health = max
    base = free * MAXC
    n = len(dna["cells"])
    np.testing.assert_array_equal(w.np("c_pos")[base], np.float32([1500.0, 900.0]))
    np.testing.assert_array_equal(w.np("c_energy")[base:base + n], w.np("c_storage")[base:base + n] / 2)
    np.testing.assert_array_equal(w.np("c_health")[base:base + n], w.np("c_maxhealth")[base:base + n])
    np.testing.assert_array_equal(w.np("c_parent")[base:base + n] < np.arange(n), np.ones(n, bool))       # parents are created first
    # and it lives: the oracle ticks the world with the import in it
    assert oracle.post(w) == 0 and oracle.tick(w, PH_ALL, 5) == 0 and int(w.np("org_state")[free]) == 1
    with pytest.raises(pickle.UnpicklingError):
        dna_io.loads_reference_pickle(pickle.dumps(os.system))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_against_the_reference_dna_class(oracle):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import refshim
    ns = refshim.load()
    # ours -> theirs: the pickle opens as a real sprites.DNA whose own methods work on it
    w = grown_world(oracle)
    slot = int(np.nonzero(w.np("org_state") == 1)[0][0])
    mine = dna_io.export_dna(w, slot)
    theirs = pickle.loads(dna_io.dumps_reference_pickle(mine))
    assert type(theirs) is ns.sprites.DNA
    assert theirs.first_cell() == next(c for c in mine["cells"] if mine["cells"][c]["first"])
    rect, size = oracle.dna_rect(w, slot)
    np.testing.assert_allclose(theirs.get_rect(), rect, rtol=1e-6, atol=1e-5); np.testing.assert_allclose(theirs.get_size(), size, rtol=1e-6, atol=1e-5)   # the reference's own (quirky) rect == the restated one
    net = ns.neural.setup_network(theirs)
    dims = [int(v) for v in w.np("org_ldim")[slot][: int(w.np("org_nlayers")[slot]) + 1]]
    assert [tuple(l.weight.shape) for l in net.layers()] == [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]
    copy = theirs.copy()                                           # deepcopy of every field, as reproduction does
    assert copy.cells == theirs.cells and copy.growth_pattern == theirs.growth_pattern
    # theirs -> ours: a genome the reference randomised and pickled (Sapiogenesis.py:297-320 ranges)
    ref_dna = ns.sprites.DNA()
    ref_dna.randomize(cellRange=[3, 25], sizeRange=[6, 42], massRange=[5, 20], mirror_x=[0.6, 0.4], mirror_y=[0.6, 0.4], hiddenRange=[3, 12])
    got = dna_io.loads_reference_pickle(pickle.dumps(ref_dna))
    w2 = World(default_params(n_org=4, n_dead=16, seed=3))
    dna_io.import_dna(w2, 1, (800.0, 600.0), got)
    out = dna_io.export_dna(w2, 1)
    assert list(out["cells"]) == list(ref_dna.cells)
    for cid, c in ref_dna.cells.items():
        o = out["cells"][cid]
        assert o["type"] == c["type"] and o["first"] == c["first"] and o["mirror_self"] == c["mirror_self"]
        for k in ("size", "mass", "elasticity", "friction", "max_health", "energy_storage", "damage"):
            assert o[k] == pytest.approx(c[k], rel=1e-6), (cid, k)
        assert o["relative_pos"] == pytest.approx(c["relative_pos"], rel=1e-6, abs=1e-5) and o["energy_usage"] == pytest.approx(c["energy_usage"], rel=1e-6)
    assert {p: list(k) for p, k in out["growth_pattern"].items()} == {p: list(k) for p, k in ref_dna.growth_pattern.items()}
    assert out["brain_structure"]["hidden_layers"] == list(ref_dna.brain_structure["hidden_layers"])
    for a, b in zip(out["brain_structure"]["hidden_weights"], ref_dna.brain_structure["hidden_weights"]):
        np.testing.assert_allclose(np.float32(a), np.float32(b), rtol=0, atol=0)
    assert oracle.post(w2) == 0 and oracle.tick(w2, PH_ALL, 3) == 0 and int(w2.np("org_state")[1]) == 1


@pytest.mark.gpu
def test_imported_organism_lives_on_the_gpu_like_in_the_oracle(oracle):
    from parity import compare_worlds
    from sapiogenesis_b200.engine import Engine
    w = grown_world(oracle)
    donor = int(np.nonzero(w.np("org_mem_n") > 0)[0][0])
    wg = w.to("cuda")
    dna = dna_io.loads_reference_pickle(dna_io.dumps_reference_pickle(dna_io.export_dna(wg, donor)))     # exported from device memory
    free = int(np.nonzero(w.np("org_state") == 0)[0][0])
    dna_io.import_dna(wg, free, (2500.0, 1200.0), dna)
    dna_io.import_dna(w, free, (2500.0, 1200.0), dna)
    assert oracle.post(w) == 0
    eng = Engine(wg); eng.post(); eng.synchronize()
    compare_worlds(wg, w, "import", verbose=False)
    for _ in range(4):
        wg2 = w.to("cuda"); e2 = Engine(wg2)
        e2.tick(1, PH_ALL); e2.synchronize()
        assert oracle.tick(w, PH_ALL, 1) == 0
        # seed 13 has a dead cell wedged between pinned cells (cached contact impulses of 1e5): the case the exact warm start and the
        # double accumulators of large contact impulses exist for
        compare_worlds(wg2, w, f"imported organism, tick {w.tick}", verbose=False)
    assert int(w.np("org_state")[free]) == 1 and int(w.np("org_hist_n")[free]) > int(dna_io.export_dna(w, donor)["generation"]) * 0


def test_genome_file_cannot_reach_arbitrary_callables():
    """the unpickler admits exact (module, name) pairs only -- a crafted genome file must not reach eval / exec /
    __import__ / torch.load / numpy.load through REDUCE."""
    import pickle
    from sapiogenesis_b200 import dna_io

    class Evil:
        def __init__(self, fn, *args): self.fn, self.args = fn, args
        def __reduce__(self): return (self.fn, self.args)

    import builtins, os
    for fn, args in ((builtins.eval, ("1+1",)), (builtins.__import__, ("os",)), (os.system, ("true",)), (builtins.exec, ("pass",))):
        blob = pickle.dumps({"cells": Evil(fn, *args), "growth_pattern": [], "brain_structure": {}, "base_info": {}})
        with pytest.raises(pickle.UnpicklingError):
            dna_io.loads_reference_pickle(blob)


def test_rnn_organism_round_trip(oracle):
    """An organism with an RNNetwork brain (use_rnn, physics.py:305): hidden_memory and trainingHidden travel with the genome, the curated
    memory in list order, and the import builds a brain with the wider input layer and the hidden head."""
    from sapiogenesis_b200.world import HM, MEMROWS
    w = make_world(oracle, n_org=6, seed=13, n_food=12, crowd=0.7, kick=20.0, extra_slots=6, rnn_hidden=6)
    w.p.think_ticks = w.p.train_ticks = 3
    assert oracle.tick(w, PH_ALL, 60) == 0
    slot = int(np.nonzero((w.np("org_th_n") > 0) & (w.np("org_mem_n") > 1))[0][0])
    dna = dna_io.export_dna(w, slot)
    hn = w.np("org_hm_n")[slot]
    assert len(dna["hidden_memory"]) == int(hn[0] - hn[1]) > 0 and len(dna["trainingHidden"]) == int(w.np("org_th_n")[slot])
    assert all(len(e) == 6 for e in dna["hidden_memory"] + dna["trainingHidden"])
    seq = w.np("org_mem_seq")[slot][: int(w.np("org_mem_n")[slot])]
    first, last = int(np.argmin(seq)), int(np.argmax(seq))                      # train_rnn's target row and input row (networks.py:59-62)
    I = int(w.np("org_ldim")[slot][0])
    np.testing.assert_array_equal(np.float32(dna["trainingInput"][-1]), w.np("org_mem_in")[slot].reshape(MEMROWS, -1)[last, :I])
    np.testing.assert_array_equal(np.float32(dna["trainingOutput"][0]), w.np("org_mem_out")[slot].reshape(MEMROWS, 4)[first])
    back = dna_io.loads_reference_pickle(dna_io.dumps_reference_pickle(dna))
    free = int(np.nonzero(w.np("org_state") == 0)[0][0])
    dna_io.import_dna(w, free, (1500.0, 900.0), back)
    again = dna_io.export_dna(w, free)
    same_genome(dna, again)
    keep = min(len(dna["hidden_memory"]), HM - 30)
    assert int(w.np("org_rnn_h")[free]) == 6 and again["hidden_memory"] == dna["hidden_memory"][:keep] and again["trainingHidden"] == dna["trainingHidden"]
    for k in ("trainingInput", "trainingOutput", "trainingReward"):
        assert again[k] == dna[k], k
    assert oracle.post(w) == 0 and oracle.tick(w, PH_ALL, 6) == 0 and int(w.np("org_state")[free]) == 1
    # a feed-forward world opens the same file: the recurrent lists are dropped, the brain is a networks.Network
    w2 = World(default_params(n_org=4, n_dead=16, seed=3))
    dna_io.import_dna(w2, 0, (800.0, 600.0), back)
    assert int(w2.np("org_rnn_h")[0]) == 0 and dna_io.export_dna(w2, 0)["hidden_memory"] == []
    assert oracle.post(w2) == 0 and oracle.tick(w2, PH_ALL, 3) == 0


def test_genome_without_hidden_layers_imports_as_a_brainless_organism(oracle):
    """neural.setup_network returns None for a genome whose hidden_layers list is empty (neural.py:471): the reference keeps such an organism
    alive without a brain; the import maps it to org_nlayers = 0 (it never thinks, trains or -- its brain cannot mutate -- reproduces)."""
    w = make_world(oracle, n_org=4, seed=3, n_food=4, extra_slots=4)
    dna = dna_io.export_dna(w, 0)
    dna["brain_structure"]["hidden_layers"] = []; dna["brain_structure"]["hidden_weights"] = []
    for k in ("trainingInput", "trainingOutput", "trainingReward"):
        dna[k] = []
    dna_io.import_dna(w, 5, (1500.0, 900.0), dna)
    assert int(w.np("org_nlayers")[5]) == 0
    thinks0 = w.stats()["thinks"]
    assert oracle.post(w) == 0 and oracle.tick(w, PH_ALL, 40) == 0
    assert int(w.np("org_state")[5]) == 1 and int(w.np("org_hist_n")[5]) == 0 and w.stats()["thinks"] > thinks0
    assert dna_io.export_dna(w, 5)["brain_structure"]["hidden_layers"] == []

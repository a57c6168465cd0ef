"""Pins the oracle's brain path (oracle/brain.c) against the reference's OWN neural.activate / train_network and
networks.Network / Trainer.train_epoch (fixtures: tests/golden/brain_cases.npz, made by tests/golden/make_golden_brain.py).
Rounded inputs and sensor membership must match exactly; network outputs within 1e-5 before the 3-decimal rounding
(a 1e-7 difference can flip a rounded digit, so rounded outputs are compared with that in mind);
curated memory keys exactly; trained weights within 1e-5 after each Adam epoch."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from sapiogenesis_b200.world import HIST, MAXC, MEMROWS, STM, World, default_params

CASES = np.load(os.path.join(GOLDEN, "brain_cases.npz"))
N_CASES, SEED, N_THINKS = (int(v) for v in CASES["meta"])
# the same scenes through the reference's networks.RNNetwork / Trainer.train_rnn and the is_rnn branches of neural.py (use_rnn = True,
# hidden_rnn_size = 6; 48 thinks, so that the short-term memory -- and hidden_memory with it -- wraps)
RNN_CASES = np.load(os.path.join(GOLDEN, "rnn_cases.npz"))
RNN_N_CASES, _, RNN_N_THINKS, RNN_H = (int(v) for v in RNN_CASES["meta"])


def u24(tag, n):
    rng = random.Random(tag)
    return np.asarray([rng.getrandbits(24) / 16777216.0 for _ in range(n)], dtype=np.float32)


def build_scene(oracle, ci, cases=CASES, rnn=0):
    g = lambda k: cases[f"c{ci}.{k}"]
    tagA, tagB = (str(s) for s in g("tags"))
    used, pos = g("used"), g("pos")
    p = default_params(n_org=3, n_dead=16)
    p.rnn_cap = p.rnn_hidden = rnn
    w = World(p)
    for slot, tag in ((0, tagA), (1, tagB)):
        rc, _ = oracle.spawn_arr(w, slot, float(pos[slot][0]), float(pos[slot][1]), u24(tag + "-u", used[slot][0] + 8),
                                 u24(tag + "-ui", used[slot][1] + 8), u24(tag + "-ub", used[slot][2] + 8))
        assert rc == 0
    for f, fp in enumerate(g("food")):
        t = w.t
        t["d_state"][f] = 1; t["d_pos"][f] = torch.tensor(fp, dtype=torch.float32)
        t["d_size"][f] = 30.0; t["d_mass"][f] = 12.0; t["d_mass0"][f] = 12.0; t["d_elast"][f] = .5; t["d_fric"][f] = .5; t["d_owner"][f] = -1
    return w, g


@pytest.mark.parametrize("ci", range(N_CASES))
def test_activate_and_train_match_reference(oracle, ci):
    run_case(oracle, ci, CASES, N_THINKS, 0)


@pytest.mark.parametrize("ci", range(RNN_N_CASES))
def test_rnn_activate_and_train_match_reference(oracle, ci):
    run_case(oracle, ci, RNN_CASES, RNN_N_THINKS, RNN_H)


def run_case(oracle, ci, cases, N_THINKS, H):
    w, g = build_scene(oracle, ci, cases, H)
    t = w.t
    assert int(w.np("org_rnn_h")[0]) == H
    ncA, ncB = int(w.np("org_ncells")[0]), int(w.np("org_ncells")[1])
    I = int(w.np("org_ldim")[0][0])
    assert g("inputs").shape == (N_THINKS, I)
    IN = w.p.in_cap
    flips = 0
    moved = {}
    for k in range(N_THINKS):
        t["c_pos"][MAXC:MAXC + ncB] = torch.from_numpy(g("bpos")[k].astype(np.float32))
        t["c_pos"][:ncA] = torch.from_numpy(g("apos")[k].astype(np.float32))
        sens = np.isin(w.np("c_type")[:ncA], (2, 3))
        sp = g("aspos")[k].astype(np.float32)
        t["c_spos"][:ncA][torch.from_numpy(sens)] = torch.from_numpy(sp[sens])
        t["c_energy"][:ncA] = torch.from_numpy(g("aenergy")[k].astype(np.float32))
        balive = g("balive")[k]
        for j in range(ncB):                       # a cell of B that died becomes a free dead body (same place, same size)
            row = MAXC + j
            if balive[j] == 0 and int(t["c_alive"][row]) == 1:
                d = 10 + j % 6
                t["d_state"][d] = 1; t["d_pos"][d] = t["c_pos"][row]; t["d_size"][d] = t["c_size"][row]
                t["d_mass"][d] = t["c_mass"][row]; t["d_mass0"][d] = t["c_mass"][row]; t["d_owner"][d] = t["org_uid"][1]
                t["c_alive"][row] = 0
                moved[j] = d
            elif balive[j] == 0 and j in moved:
                t["d_pos"][moved[j]] = t["c_pos"][row]
        dop, pain = g("dopa")[k]
        t["org_dopamine"][0], t["org_pain"][0] = float(dop), float(pain)
        # the reference drew these uniforms: boredom gate, then Box-Muller pairs for torch.randn(4)
        gate = g("gate")[k]
        np.testing.assert_allclose(oracle.rotation(w, 0), g("rot")[k], rtol=0, atol=1e-9)
        raw_in = oracle.inputs(w, 0, oracle.rotation(w, 0))
        x_ref = g("inputs")[k]
        x = np.round(raw_in.astype(np.float32) * np.float32(1000)) / np.float32(1000)
        tie = np.abs(raw_in * 1000 - np.floor(raw_in * 1000) - 0.5) < 1e-4
        assert np.all((x == x_ref.astype(np.float32)) | tie), f"think {k}: rounded inputs"
        assert oracle.think_now(w, 0, uniforms=gate) == 0      # same five uniforms the reference consumed
        raw = w.np("org_out_raw")[0]
        np.testing.assert_allclose(raw, g("raw")[k], rtol=1e-5, atol=1e-6, err_msg=f"think {k}: network output")
        slot = k % HIST
        got_out = w.np("org_hist_out")[0][4 * slot: 4 * slot + 4]
        assert (np.abs(got_out - g("out")[k]) <= 1e-3 + 1e-6).all(), f"think {k}: outputs (rounded, or the random action)"
        tol_flip = np.abs(got_out - g("out")[k].astype(np.float32)) > 1e-6
        flips += int(tol_flip.sum())
        if tol_flip.any():          # a tie flipped a rounded digit: realign so the histories stay comparable
            t["org_hist_out"][0][4 * slot: 4 * slot + 4] = torch.from_numpy(g("out")[k].astype(np.float32))
            t["org_move"][0] = torch.from_numpy(np.clip(g("out")[k], -1, 1).astype(np.float32))
        np.testing.assert_array_equal(w.np("org_hist_in")[0][slot * IN: slot * IN + I], x_ref.astype(np.float32))
        np.testing.assert_allclose(w.np("org_move")[0], g("move")[k], rtol=0, atol=1e-6)
        stim, boredom = g("scal")[k][0], g("scal")[k][1]
        np.testing.assert_allclose(w.np("org_stimulation")[0], stim, rtol=2e-6, atol=1e-7, err_msg=f"think {k}: stimulation")
        np.testing.assert_allclose(w.np("org_boredom")[0], boredom, rtol=2e-6, atol=1e-7, err_msg=f"think {k}: boredom")
        np.testing.assert_allclose(w.np("org_stim_hist")[0], g("stim_hist")[k], rtol=2e-6, atol=1e-7)
        assert min(int(w.np("org_hist_n")[0]), HIST) == int(g("nhist")[k])
        assert min(int(w.np("org_stm_n")[0]), STM) == int(g("stm_n")[k]), f"think {k}: short-term memory length"
        if H:
            np.testing.assert_allclose(w.np("org_last_hidden")[0][:H], g("lasthid")[k], rtol=1e-5, atol=1e-6, err_msg=f"think {k}: lastHidden")
            hn = w.np("org_hm_n")[0]
            assert int(hn[0]) - int(hn[1]) == int(g("hm_n")[k]), f"think {k}: hidden_memory length"
    assert flips <= 2, "rounded outputs differ from the reference beyond tie flips"
    # short-term memory content (ring, chronological)
    sc = int(w.np("org_stm_n")[0]); sn = min(sc, STM)
    assert sn == len(g("stm_rew"))
    ring = [(sc - sn + i) % STM for i in range(sn)]                # chronological order of the ring
    got_in = w.np("org_stm_in")[0].reshape(STM, IN)[ring, :I]
    np.testing.assert_array_equal(got_in, g("stm_in"))
    np.testing.assert_allclose(w.np("org_stm_out")[0].reshape(STM, 4)[ring], g("stm_out"), rtol=0, atol=1e-3 + 1e-6)
    np.testing.assert_allclose(w.np("org_stm_rew")[0][ring], g("stm_rew"), rtol=2e-6, atol=1e-6)
    # make the actions identical to the reference's before training on them
    t["org_stm_out"][0].view(STM, 4)[ring] = torch.from_numpy(g("stm_out"))
    if H:                                                           # hidden_memory runs along with the short-term memory (neural.py:312-318)
        hn = w.np("org_hm_n")[0]; RC = w.p.rnn_cap
        hm = w.np("org_hm")[0].reshape(-1, RC)
        got_hm = hm[[(int(hn[1]) + i) % hm.shape[0] for i in range(int(hn[0]) - int(hn[1]))], :H]
        np.testing.assert_allclose(got_hm, g("hm"), rtol=1e-5, atol=1e-6)
        t["org_hm"][0].view(-1, RC)[[(int(hn[1]) + i) % hm.shape[0] for i in range(len(g("hm")))], :H] = torch.from_numpy(g("hm"))

    # ---- train_network twice
    pre_in, pre_out, pre_rew = g("pre_in"), g("pre_out"), g("pre_rew")
    M0 = len(pre_rew)
    assert M0 <= MEMROWS
    mem_in = t["org_mem_in"][0].view(MEMROWS, IN)
    mem_in[:M0, :I] = torch.from_numpy(pre_in)
    t["org_mem_out"][0].view(MEMROWS, 4)[:M0] = torch.from_numpy(pre_out)
    t["org_mem_rew"][0][:M0] = torch.from_numpy(pre_rew)
    t["org_mem_n"][0] = M0
    t["org_mem_seq"][0][:M0] = torch.arange(M0, dtype=torch.int32); t["org_mem_stamp"][0] = M0      # list order of the pre-filled rows
    nl = int(w.np("org_nlayers")[0]); ldim = w.np("org_ldim")[0]
    P = sum(int(ldim[l + 1]) * (int(ldim[l]) + (H if l == 0 else 0)) + int(ldim[l + 1]) for l in range(nl)) + (H * int(ldim[nl - 1]) + H if H else 0)
    if H:
        t["org_th"][0].view(MEMROWS, w.p.rnn_cap)[:M0, :H] = torch.from_numpy(g("pre_th")); t["org_th_n"][0] = M0
    for ep in range(2):
        assert oracle.train_one(w, 0) == 0
        M = int(w.np("org_mem_n")[0])
        ref_in, ref_out, ref_rew = g(f"mem_in{ep}"), g(f"mem_out{ep}"), g(f"mem_rew{ep}")
        assert M == len(ref_rew), f"epoch {ep}: memory size {M} vs {len(ref_rew)}"
        got = {tuple(np.round(r * 1000).astype(int)): (tuple(o), float(q)) for r, o, q in
               zip(w.np("org_mem_in")[0].reshape(MEMROWS, IN)[:M, :I], w.np("org_mem_out")[0].reshape(MEMROWS, 4)[:M], w.np("org_mem_rew")[0][:M])}
        ref = {tuple(np.round(r * 1000).astype(int)): (tuple(o), float(q)) for r, o, q in zip(ref_in, ref_out, ref_rew)}
        assert set(got) == set(ref), f"epoch {ep}: curated memory keys"
        for key in ref:
            np.testing.assert_allclose(got[key][0], ref[key][0], rtol=0, atol=1e-6)
            np.testing.assert_allclose(got[key][1], ref[key][1], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(w.np("org_loss")[0], g(f"loss{ep}")[0], rtol=2e-5, atol=1e-9)
        if H:
            # Trainer.train_rnn (networks.py:56-73): one sample, plain SGD on the input and output layers, lastHidden as a side effect
            tn_ref = int(g(f"th_n{ep}")[0])
            assert int(w.np("org_th_n")[0]) == min(tn_ref, MEMROWS), f"epoch {ep}: trainingHidden length"
            np.testing.assert_allclose(w.np("org_th")[0].reshape(MEMROWS, -1)[: min(tn_ref, MEMROWS), :H], g(f"th{ep}"), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(w.np("org_last_hidden")[0][:H], g(f"lasthid_t{ep}"), rtol=1e-5, atol=1e-6, err_msg=f"epoch {ep}: lastHidden after train_rnn")
            got_w, ref_w = w.np("org_brain")[0][:P].astype(np.float64), g(f"brain{ep}").astype(np.float64)
            assert len(ref_w) == P
            np.testing.assert_allclose(got_w, ref_w, rtol=1e-5, atol=1e-6, err_msg=f"epoch {ep}: weights after train_rnn")
            continue
        # Adam's first steps move every weight by ~lr * g / (|g| + 1e-8): where |g| is within a few orders of eps the update is
        # ill-conditioned in the gradient's own fp32 summation noise (torch's sgemm order is not reproducible). So: 1e-5 relative to
        # the weight scale for >= 99 % of the parameters, and never more than 1e-3 of one learning-rate step (lr = 0.02).
        got_w, ref_w = w.np("org_brain")[0][:P].astype(np.float64), g(f"brain{ep}").astype(np.float64)
        scale = np.maximum(np.abs(ref_w), np.sqrt(np.mean(ref_w ** 2)))
        rel = np.abs(got_w - ref_w) / scale
        assert np.mean(rel <= 1e-5) >= 0.99, f"epoch {ep}: trained weights ({np.mean(rel <= 1e-5):.4f} within 1e-5)"
        assert np.max(np.abs(got_w - ref_w)) <= 1e-3 * 0.02, f"epoch {ep}: trained weights max abs diff {np.max(np.abs(got_w - ref_w)):.2e}"


@pytest.mark.parametrize("ci", range(RNN_N_CASES))
def test_rnn_lists_a_child_inherits_match_reference(oracle, ci):
    """DNA.mutate starts with copy() + erase_memory() (sprites.py:1070-1072), and erase_memory (sprites.py:413-421) clears the training
    and short-term lists but neither hidden_memory nor trainingHidden: the child starts with the parent's. ow_inherit_recurrent (oracle/world.c)
    against the reference's own copy + erase on the parent of the scene, after its 48 thinks and two train_network calls."""
    import ctypes
    from oracle import oracle as omod
    g = lambda k: RNN_CASES[f"c{ci}.{k}"]
    n_train, n_stm, n_prev, n_hm, n_th = (int(v) for v in g("child_lens"))
    assert (n_train, n_stm, n_prev) == (0, 0, 0) and n_hm == len(g("hm")) and n_th == int(g("th_n1")[0])
    H = RNN_H
    p = default_params(n_org=2, n_dead=4, rnn_hidden=H)
    w = World(p); t = w.t
    # the parent's lists as the scene left them: hidden_memory (30 entries, some popped before), trainingHidden after the second call
    hm = g("hm"); pop = 7
    ring = [(pop + i) % 90 for i in range(len(hm))]
    t["org_hm"][0].view(90, H)[ring] = torch.from_numpy(hm); t["org_hm_n"][0] = torch.tensor([pop + len(hm), pop], dtype=torch.int32)
    th = g("th1"); t["org_th"][0].view(MEMROWS, H)[: len(th)] = torch.from_numpy(th); t["org_th_n"][0] = len(th)
    L = omod.lib(); L.ow_inherit_recurrent.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]; L.ow_inherit_recurrent.restype = None
    s = w.struct(); L.ow_inherit_recurrent(ctypes.byref(s), 0, 1)
    hn = w.np("org_hm_n")[1]
    assert (int(hn[0]), int(hn[1])) == (min(n_hm, 60), 0)
    np.testing.assert_array_equal(w.np("org_hm")[1].reshape(90, H)[: int(hn[0])], g("child_hm")[: int(hn[0])])
    assert int(w.np("org_th_n")[1]) == min(n_th, MEMROWS)
    np.testing.assert_array_equal(w.np("org_th")[1].reshape(MEMROWS, H)[: min(n_th, MEMROWS)], g("child_th"))


def test_one_decimal_keys_follow_python_round(oracle):
    """roundList(x, 1) on the doubles the reference actually holds: fp32 values from tensor.tolist() (short-term memory)
    and round(x, 3) results (stored memory). neural.py:77-78, :375-393."""
    for k in range(-2100, 2101):
        x32 = float(np.float32(k / 1000.0))
        assert oracle.key1(x32) == int(round(round(x32, 1) * 10)), k
        x3 = round(x32, 3)
        assert oracle.key1(x3) == int(round(round(x3, 1) * 10)), k

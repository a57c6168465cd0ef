"""Golden vectors for the brain path: the REFERENCE's own neural.activate (eye / olfactory encoding, rounding, forward,
boredom, history, short-term memory, stimulation), networks.Network.forward, neural.train_network (memory curation) and
networks.Trainer.train_epoch (MSE + torch Adam), run headless behind refshim.py.
Output: tests/golden/brain_cases.npz. Run in the build container only.

The `view` lists the reference fills through Chipmunk callbacks are filled here with the geometric contract they
implement: every solid sprite of another organism (or ownerless food) whose disc overlaps the sensor disc (strict)."""
from __future__ import annotations

import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402
from make_golden import UI, build, feed, round_dna_to_f32  # noqa: E402

N_THINKS = 16


def f32(x):
    return float(np.float32(x))


def find_seed(ns, base, want_eye=2, want_olf=1, max_cells=30):
    none = lambda: refshim.UniformFeed([], "none")
    for k in range(400):
        tag = f"{base}-{k}"
        u, ui = feed(tag + "-u", 30000, "u"), feed(tag + "-ui", 600000, "init")
        refshim.set_feeds(ns, u, ui, none())
        dna = ns.sprites.DNA().randomize(**UI)
        types = [c["type"] for c in dna.cells.values()]
        if types.count("eye") >= want_eye and types.count("olfactory") >= want_olf and len(types) <= max_cells:
            return tag, dna, [u.pos, ui.pos]
    raise RuntimeError("no genome with enough sensors")


def fill_views(ns, env, org):
    """contract C3: view == overlap of the sensor disc with solid sprites that do not belong to the same organism"""
    for cid, sp in org.cells.items():
        if not sp.info["detection_items"]:
            continue
        shape = sp.info["detection_items"][1]
        sp.info["view"] = [s for s in env.sprites if s.organism is not org and s.shape is not shape and not s.shape.sensor
                           and shape.shapes_collide(s.shape).points]


def gen_brain(n_cases=6, seed=41, rnn=0, name="brain_cases.npz", n_thinks=N_THINKS):
    """rnn = hidden_rnn_size with Environment.info["use_rnn"] set (physics.py:305,308): the same scenes through networks.RNNetwork,
    the is_rnn branches of neural.activate / train_network and Trainer.train_rnn -> rnn_cases.npz"""
    ns = refshim.load()
    torch = ns.torch
    N_THINKS = n_thinks
    out = {"meta": np.asarray([n_cases, seed, N_THINKS, rnn], dtype=np.int32)}
    none = lambda: refshim.UniformFeed([], "none")
    for ci in range(n_cases):
        base = f"brain-{seed}-{ci}"
        rng = random.Random(base + "-state")
        tagA, dnaA, usedA = find_seed(ns, base + "-A")
        tagB, dnaB, usedB = find_seed(ns, base + "-B", want_eye=0, want_olf=0, max_cells=40)
        round_dna_to_f32(dnaA); round_dna_to_f32(dnaB)
        env = refshim.FakeEnv(ns)
        if rnn:
            env.info["use_rnn"], env.info["hidden_rnn_size"] = True, rnn
        ubA, ubB = feed(tagA + "-ub", 600000, "build"), feed(tagB + "-ub", 600000, "build")
        posA = [f32(rng.uniform(900, 2200)), f32(rng.uniform(600, 1200))]
        ang = rng.uniform(0, 2 * math.pi); dist = rng.uniform(120, 260)
        posB = [f32(posA[0] + dist * math.cos(ang)), f32(posA[1] + dist * math.sin(ang))]
        refshim.set_feeds(ns, none(), ubA, none()); A = build(ns, env, dnaA, posA); usedA.append(ubA.pos)
        refshim.set_feeds(ns, none(), ubB, none()); B = build(ns, env, dnaB, posB); usedB.append(ubB.pos)
        # ownerless food (add_dead_clicked recipe, Sapiogenesis.py:490-507) around A
        food = []
        for f in range(6):
            a2, d2 = rng.uniform(0, 2 * math.pi), rng.uniform(60, 420)
            fp = [f32(posA[0] + d2 * math.cos(a2)), f32(posA[1] + d2 * math.sin(a2))]
            info = {"size": 30, "mass": 12, "elasticity": .5, "friction": .5, "mirror_self": None, "first": False, "type": "body"}
            ns.sprites.finish_cell_info(info)
            sp = ns.sprites.build_sprite(env, None, fp, 0, info, alive=False)
            env.add_sprite(sp)
            food.append(fp)
        for org in (A, B):
            for cid, sp in org.cells.items():
                sp.body.position = (f32(sp.body.position[0]), f32(sp.body.position[1]))
                if sp.info["detection_items"]:
                    sp.info["detection_items"][0].position = sp.body.position
        orderA, orderB = list(A.cells.keys()), list(B.cells.keys())
        out[f"c{ci}.tags"] = np.asarray([tagA, tagB])
        out[f"c{ci}.used"] = np.asarray([usedA, usedB], dtype=np.int32)
        out[f"c{ci}.pos"] = np.asarray([posA, posB], dtype=np.float64)
        out[f"c{ci}.food"] = np.asarray(food, dtype=np.float64)
        I = A.brain.inputSize - rnn          # RNNetwork.inputSize counts the hidden columns (networks.py:116)
        rec = dict(inputs=[], raw=[], out=[], move=[], scal=[], stim_hist=[], nhist=[], stm_n=[], rot=[], bpos=[], benergy=[], aenergy=[],
                   gate=[], dopa=[])
        curiosity = f32(A.dna.base_info["curiosity"]); A.dna.base_info["curiosity"] = curiosity
        for t in range(N_THINKS):
            # the scene changes between thinks: B drifts and rotates a little, energies vary, some of B's cells die
            dxy = (rng.uniform(-25, 25), rng.uniform(-25, 25))
            for cid in orderB:
                b = B.cells[cid].body
                b.position = (f32(b.position[0] + dxy[0]), f32(b.position[1] + dxy[1]))
            if t == 9:
                victim = orderB[-1]
                B.cells[victim].alive = False       # a viewed cell that is no longer alive counts for the nose, not less for the eye
            aen = []
            for cid in orderA:
                sp = A.cells[cid]
                sp.info["energy"] = f32(sp.cell_info["energy_storage"] * rng.uniform(0.05, 0.95))
                aen.append(sp.info["energy"])
            # A itself turns: rotate its cells about the heart
            th = rng.uniform(-0.4, 0.4)
            hx, hy = A.cells[orderA[0]].body.position
            for cid in orderA[1:]:
                b = A.cells[cid].body
                rx, ry = b.position[0] - hx, b.position[1] - hy
                b.position = (f32(hx + rx * math.cos(th) - ry * math.sin(th)), f32(hy + rx * math.sin(th) + ry * math.cos(th)))
            for cid in orderA:
                sp = A.cells[cid]
                if sp.info["detection_items"]:
                    # the sensor body trails its cell by the joint error (SURVEY A-6): give it a small offset
                    sp.info["detection_items"][0].position = (f32(sp.body.position[0] + rng.uniform(-3, 3)), f32(sp.body.position[1] + rng.uniform(-3, 3)))
            A.dopamine = f32(rng.choice([0.0, 0.05, rng.uniform(-3, 3), rng.uniform(-30, 30)]))
            A.pain = f32(rng.choice([0.0, 0.0, rng.uniform(0, 5)]))
            gate_u = rng.getrandbits(24) / 16777216.0
            randn_u = [rng.getrandbits(24) / 16777216.0 for _ in range(4)]
            fill_views(ns, env, A)
            u = refshim.UniformFeed([gate_u], "gate")
            refshim.set_feeds(ns, u, None, None)

            def randn(shape, _u=randn_u):
                vals = []
                for h in range(2):
                    u1 = (_u[2 * h] * 16777216.0 + 1.0) / 16777216.0
                    rad, thh = math.sqrt(-2.0 * math.log(u1)), 2.0 * math.pi * _u[2 * h + 1]
                    vals += [rad * math.cos(thh), rad * math.sin(thh)]
                return torch.tensor(vals, dtype=torch.float64).float()
            ns.neural.torch = refshim._TorchProxy(torch, randn=randn)
            rot = A.rotation()
            h_prev = A.brain.lastHidden.detach().clone() if rnn else None
            A.brain.activate(env, A, 0.03)
            x = A.dna.previousInputs[-1]
            rec["inputs"].append(x.tolist())
            if rnn:
                rec["raw"].append(A.brain.forward(x, h_prev).tolist())      # recomputes the same lastHidden
                rec.setdefault("lasthid", []).append(A.brain.lastHidden.tolist())
                rec.setdefault("hm_n", []).append(len(A.dna.hidden_memory))
            else:
                rec["raw"].append(A.brain.forward(x).tolist())
            rec["out"].append(A.dna.previousOutputs[-1])
            rec["move"].append([A.movement["rotation"], A.movement["speed"], A.movement["direction"][0], A.movement["direction"][1]])
            rec["scal"].append([A.brain.stimulation, A.brain.boredom, A.dopamine, A.pain])
            rec["stim_hist"].append(list(A.dna.stim_history))
            rec["nhist"].append(len(A.dna.previousInputs))
            rec["stm_n"].append(len(A.dna.short_term_memory))
            rec["rot"].append(rot)
            rec["bpos"].append([[B.cells[c].body.position[0], B.cells[c].body.position[1]] for c in orderB])
            rec["aenergy"].append(aen)
            rec["gate"].append([gate_u] + randn_u)
            rec["dopa"].append([A.dopamine, A.pain])
            rec.setdefault("apos", []).append([[A.cells[c].body.position[0], A.cells[c].body.position[1]] for c in orderA])
            rec.setdefault("aspos", []).append([[A.cells[c].info["detection_items"][0].position[0], A.cells[c].info["detection_items"][0].position[1]]
                                                 if A.cells[c].info["detection_items"] else [0.0, 0.0] for c in orderA])
            rec.setdefault("balive", []).append([1 if B.cells[c].alive else 0 for c in orderB])
            rec.setdefault("views", []).append(sum(len(A.cells[c].info["view"]) for c in orderA))
        for k, v in rec.items():
            out[f"c{ci}.{k}"] = np.asarray(v, dtype=np.float64)
        stm = A.dna.short_term_memory
        out[f"c{ci}.stm_in"] = np.asarray([m[0] for m in stm], dtype=np.float32).reshape(len(stm), I)
        out[f"c{ci}.stm_out"] = np.asarray([m[1] for m in stm], dtype=np.float32).reshape(len(stm), 4)
        out[f"c{ci}.stm_rew"] = np.asarray([m[2] for m in stm], dtype=np.float64)

        if rnn:
            assert len(A.dna.hidden_memory) == len(stm)
            out[f"c{ci}.hm"] = np.asarray(A.dna.hidden_memory, dtype=np.float32).reshape(len(stm), rnn)
        # ---- train_network: curated memory + one Adam epoch, twice; pre-filled memory so that replace / keep / trim all occur
        pre_n = ([0, 150, 200, 185] if rnn else [0, 150, 205, 230, 60, 199])[ci % 6]      # RNN scenes hold 30 short-term rows: <= 200 + 30 rows fit the table
        pre_in, pre_out, pre_rew = [], [], []
        seen = set()                              # the reference's memory never holds two entries with one 1-dp key
        m = 0                                     # (train_network replaces on a key match), so neither does the pre-fill
        while len(pre_in) < pre_n:
            m += 1
            if stm and m % 7 == 0:            # same 1-dp key as a short-term memory: replaced or kept depending on the reward
                src = stm[rng.randrange(len(stm))][0]
                xin = [f32(round(v + rng.choice([0.0, 0.0, 0.004, -0.004]), 3)) for v in src]
            else:
                xin = [f32(round(rng.uniform(-1, 1), 3)) for _ in range(I)]
            key = tuple(round(round(v, 3), 1) for v in xin)
            if key in seen:
                continue
            seen.add(key)
            pre_in.append([round(v, 3) for v in xin])
            pre_out.append([f32(rng.uniform(-1, 1)) for _ in range(4)])
            pre_rew.append(f32(rng.uniform(-5, 5)))
        A.dna.trainingInput = [list(v) for v in pre_in]
        A.dna.trainingOutput = [list(v) for v in pre_out]
        A.dna.trainingReward = list(pre_rew)
        if rnn:      # trainingHidden runs along with the memory (one entry per appended row, never removed: neural.py:83-87, 396-397)
            pre_th = [[f32(rng.uniform(-1, 1)) for _ in range(rnn)] for _ in range(pre_n)]
            A.dna.trainingHidden = [list(v) for v in pre_th]
            out[f"c{ci}.pre_th"] = np.asarray(pre_th, dtype=np.float32).reshape(pre_n, rnn)
        out[f"c{ci}.pre_in"] = np.asarray(pre_in, dtype=np.float32).reshape(pre_n, I)
        out[f"c{ci}.pre_out"] = np.asarray(pre_out, dtype=np.float32).reshape(pre_n, 4)
        out[f"c{ci}.pre_rew"] = np.asarray(pre_rew, dtype=np.float32)
        ns.neural.torch = torch
        for ep in range(2):
            ns.neural.train_network(A, epochs=1)
            M = len(A.dna.trainingInput)
            out[f"c{ci}.mem_in{ep}"] = np.asarray(A.dna.trainingInput, dtype=np.float32).reshape(M, I)
            out[f"c{ci}.mem_out{ep}"] = np.asarray(A.dna.trainingOutput, dtype=np.float32).reshape(M, 4)
            out[f"c{ci}.mem_rew{ep}"] = np.asarray(A.dna.trainingReward, dtype=np.float64)
            flat = []
            for layer in A.brain.layers():
                flat += layer.weight.detach().reshape(-1).tolist() + layer.bias.detach().reshape(-1).tolist()
            out[f"c{ci}.brain{ep}"] = np.asarray(flat, dtype=np.float32)
            out[f"c{ci}.loss{ep}"] = np.asarray([A.brain.lastLoss], dtype=np.float64)
            if rnn:
                th = A.dna.trainingHidden
                out[f"c{ci}.th_n{ep}"] = np.asarray([len(th)], dtype=np.int32)
                out[f"c{ci}.th{ep}"] = np.asarray(th[:232], dtype=np.float32).reshape(min(len(th), 232), rnn)
                out[f"c{ci}.lasthid_t{ep}"] = np.asarray(A.brain.lastHidden.tolist(), dtype=np.float32)
        if rnn:
            # what a child starts with: DNA.mutate's first two statements (sprites.py:1070-1072) on the trained parent -- copy(), erase_memory()
            child = A.dna.copy(); child.erase_memory()
            out[f"c{ci}.child_lens"] = np.asarray([len(child.trainingInput), len(child.short_term_memory), len(child.previousInputs),
                                                   len(child.hidden_memory), len(child.trainingHidden)], dtype=np.int32)
            out[f"c{ci}.child_hm"] = np.asarray(child.hidden_memory, dtype=np.float32).reshape(len(child.hidden_memory), rnn)
            out[f"c{ci}.child_th"] = np.asarray(child.trainingHidden[:232], dtype=np.float32).reshape(min(len(child.trainingHidden), 232), rnn)
        print(f"brain case {ci}: inputs {I}, cells A {len(orderA)} B {len(orderB)}, viewed {rec['views']}, stm {len(stm)}, "
              f"memory {pre_n} -> {len(A.dna.trainingInput)}, loss {A.brain.lastLoss:.5f}")
    np.savez_compressed(os.path.join(HERE, name), **out)


if __name__ == "__main__":
    if "rnn" in sys.argv[1:]:
        gen_brain(n_cases=4, seed=43, rnn=6, name="rnn_cases.npz", n_thinks=48)
    else:
        gen_brain()

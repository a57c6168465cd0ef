"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own Python
(/root/reference: sprites.py, neural.py, networks.py, physics.py) behind the stand-ins of refshim.py.

Run in the build container only:   python tests/golden/make_golden.py [genome|rules|brain|placement|all]
The fixtures are committed; the tests (tests/test_oracle_*.py) compare the CPU oracle against them and run
anywhere. Nothing here touches the product package or the oracle.
"""
from __future__ import annotations

import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

TYPES = ["barrier", "carniv", "eye", "olfactory", "co2C", "push", "body", "rotate", "heart"]
UI = dict(cellRange=[3, 25], sizeRange=[6, 42], massRange=[5, 20], mirror_x=[0.6, 0.4], mirror_y=[0.6, 0.4],
          hiddenRange=[3, 12])


def u24(rng: random.Random, n: int):
    """uniforms with 24 significant bits: exactly what the oracle's Philox stream yields (fp32-exact)."""
    return [rng.getrandbits(24) / 16777216.0 for _ in range(n)]


def f32(x):
    return float(np.float32(x))


def dna_to_dict(dna):
    ids = list(dna.cells)
    parent = {}
    grow = {}
    for pid, kids in dna.growth_pattern.items():
        for cid, d in kids.items():
            parent[cid] = pid
            grow[cid] = d
    c = dna.cells
    return dict(
        ids=ids,
        type=[TYPES.index(c[i]["type"]) for i in ids],
        size=[float(c[i]["size"]) for i in ids], mass=[float(c[i]["mass"]) for i in ids],
        elast=[c[i]["elasticity"] for i in ids], fric=[c[i]["friction"] for i in ids],
        rel=[list(map(float, c[i]["relative_pos"])) for i in ids],
        first=[bool(c[i]["first"]) for i in ids],
        mirror=[int(c[i]["mirror_self"] or 0) for i in ids],
        parent=[int(parent.get(i, 0)) for i in ids],
        grow=[list(map(float, grow.get(i, [0.0, 0.0]))) for i in ids],
        maxhealth=[float(c[i]["max_health"]) for i in ids],
        use_idle=[float(c[i]["energy_usage"][0]) for i in ids], use_act=[float(c[i]["energy_usage"][1]) for i in ids],
        storage=[float(c[i]["energy_storage"]) for i in ids], damage=[float(c[i]["damage"]) for i in ids],
        mirror_x=bool(dna.base_info["mirror_x"]), mirror_y=bool(dna.base_info["mirror_y"]),
        curiosity=float(dna.base_info["curiosity"]),
        hidden=list(map(int, dna.brain_structure["hidden_layers"])),
        hidden_weights=[np.asarray(l, dtype=np.float32).reshape(-1).tolist() for l in dna.brain_structure["hidden_weights"]],
        rect=list(map(float, dna.get_rect())), size_wh=list(map(float, dna.get_size())),
        generation=int(dna.generation),
    )


def round_dna_to_f32(dna):
    """What the world arrays keep of a genome is fp32; make the reference parent carry exactly that."""
    for cid, c in dna.cells.items():
        c["elasticity"], c["friction"] = f32(c["elasticity"]), f32(c["friction"])
        c["relative_pos"] = [f32(c["relative_pos"][0]), f32(c["relative_pos"][1])]
        for k in ("max_health", "energy_storage", "damage"):
            c[k] = f32(c[k])
        c["energy_usage"] = [f32(c["energy_usage"][0]), f32(c["energy_usage"][1])]
    for pid, kids in dna.growth_pattern.items():
        for cid in kids:
            kids[cid] = [f32(kids[cid][0]), f32(kids[cid][1])]
    dna.base_info["curiosity"] = f32(dna.base_info["curiosity"])


def build(ns, env, dna, pos):
    org = ns.sprites.Organism(list(pos), env, dna=dna)
    env.info["organism_list"].append(org)
    org.build_body()
    return org


def brain_flat(net):
    out = []
    for layer in net.layers():
        out += layer.weight.detach().reshape(-1).tolist()
        out += layer.bias.detach().reshape(-1).tolist()
    return out


def feed(tag: str, n: int, name: str):
    """Deterministic uniform feed; tests regenerate it from the tag instead of storing it."""
    return refshim.UniformFeed(u24(random.Random(tag), n), name)


def pack_case(prefix: str, d: dict, out: dict):
    """Flatten one organism record into npz entries."""
    dna = d["dna"]
    ints = ["ids", "type", "first", "mirror", "parent", "hidden"]
    for k in ints:
        out[f"{prefix}.{k}"] = np.asarray(dna[k], dtype=np.int32)
    for k in ["size", "mass", "elast", "fric", "rel", "grow", "maxhealth", "use_idle", "use_act", "storage", "damage", "rect", "size_wh"]:
        out[f"{prefix}.{k}"] = np.asarray(dna[k], dtype=np.float64)
    out[f"{prefix}.flags"] = np.asarray([dna["mirror_x"], dna["mirror_y"], dna["generation"]], dtype=np.int32)
    out[f"{prefix}.curiosity"] = np.asarray([dna["curiosity"]], dtype=np.float64)
    hw = [x for l in dna["hidden_weights"] for x in l]
    out[f"{prefix}.hw"] = np.asarray(hw, dtype=np.float32)
    out[f"{prefix}.used"] = np.asarray(d["used"], dtype=np.int32)
    out[f"{prefix}.pos"] = np.asarray(d["pos"], dtype=np.float64)
    out[f"{prefix}.order"] = np.asarray(d["order"], dtype=np.int32)
    out[f"{prefix}.input_cells"] = np.asarray(d["input_cells"], dtype=np.int32)
    out[f"{prefix}.brain"] = np.asarray(d["brain"], dtype=np.float32)
    out[f"{prefix}.pct"] = np.asarray(d.get("pct", [0, 0]), dtype=np.float64)


def gen_genome(n_cases=16, seed=11):
    ns = refshim.load()
    out = {"meta": np.asarray([n_cases, seed], dtype=np.int32)}
    none = lambda: refshim.UniformFeed([], "none")
    for ci in range(n_cases):
        tag = f"genome-{seed}-{ci}"
        rng = random.Random(tag + "-pos")
        u, ui = feed(tag + "-u", 30000, "u"), feed(tag + "-ui", 600000, "init")
        refshim.set_feeds(ns, u, ui, none())
        dna = ns.sprites.DNA().randomize(**UI)
        used_u, used_ui = u.pos, ui.pos
        round_dna_to_f32(dna)
        # Organism.build_body: creation order + a second, fresh nn.Linear init (build_brain)
        ub = feed(tag + "-ub", 600000, "build")
        refshim.set_feeds(ns, none(), ub, none())
        env = refshim.FakeEnv(ns)
        pos = [rng.uniform(400, 2700), rng.uniform(400, 1400)]
        org = build(ns, env, dna, pos)
        pack_case(f"c{ci}.parent", dict(dna=dna_to_dict(dna), used=[used_u, used_ui, ub.pos], pos=pos, order=list(org.cells.keys()),
                  input_cells=[c for c in org.brain.inputCells if isinstance(c, int)], brain=brain_flat(org.brain),
                  pct=[org.health_percent(), org.energy_percent()]), out)
        # DNA.mutate on the fp32-rounded parent
        um, uf, ut = feed(tag + "-um", 30000, "mut"), feed(tag + "-uf", 200000, "fill"), feed(tag + "-ut", 600000, "tempnet")
        refshim.set_feeds(ns, um, ut, uf)     # tempNet init inside mutate_brain_layers only sizes the layers: its draws are discarded
        old_hidden = list(dna.brain_structure["hidden_layers"])
        child = dna.mutate(0.5, 0.5, True)
        child.generation = dna.generation + 1
        # which inherited hidden weights are DEFINED in the reference: row r of new layer l re-reads the old layer's storage
        # from offset r*old_cols (Tensor.resize_ on a row view, sprites.py:1038-1039); past its end it is uninitialised memory.
        new_hidden = list(child.brain_structure["hidden_layers"])
        defined = []
        for l in range(len(new_hidden) - 1):
            rows, cols = new_hidden[l + 1], new_hidden[l]
            for r in range(rows):
                for c in range(cols):
                    if l < len(old_hidden) - 1 and r < old_hidden[l + 1]:
                        defined.append(r * old_hidden[l] + c < old_hidden[l + 1] * old_hidden[l])
                    else:
                        defined.append(True)       # torch.rand row
        out[f"c{ci}.child.hw_defined"] = np.asarray(defined, dtype=bool)
        used_um, used_uf = um.pos, uf.pos
        ub2 = feed(tag + "-ub2", 600000, "build2")
        refshim.set_feeds(ns, none(), ub2, none())
        cpos = [rng.uniform(400, 2700), rng.uniform(400, 1400)]
        ok = len(child.cells) <= 64
        corg = build(ns, env, child, cpos) if ok else None
        pack_case(f"c{ci}.child", dict(dna=dna_to_dict(child), used=[used_um, used_uf, ub2.pos], pos=cpos,
                  order=list(corg.cells.keys()) if ok else [], brain=brain_flat(corg.brain) if ok else [],
                  input_cells=[c for c in corg.brain.inputCells if isinstance(c, int)] if ok else []), out)
        print(f"genome case {ci}: cells {len(dna.cells)} -> {len(child.cells)}, hidden {dna.brain_structure['hidden_layers']} -> "
              f"{child.brain_structure['hidden_layers']}, draws u={used_u} mut={used_um} fill={used_uf}")
    np.savez_compressed(os.path.join(HERE, "genome_cases.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("genome", "all"):
        gen_genome()
    if what in ("rules", "all"):
        from make_golden_rules import gen_rules
        gen_rules()
    if what in ("brain", "all"):
        from make_golden_brain import gen_brain
        gen_brain()

"""Genome exchange with the reference: `sprites.DNA` objects as the GUI pickles them (Sapiogenesis.py:243-254 save, :413-424
import; fields at sprites.py:330-408).

    dna = export_dna(world, slot)                 # plain dict with the reference's attribute names and value types
    blob = dumps_reference_pickle(dna)            # unpickles as a real sprites.DNA inside the reference (no reference needed here)
    dna = loads_reference_pickle(blob)            # ... and back, also for files the reference's "Save organism" wrote
    import_dna(world, slot, (x, y), dna)          # Organism(pos, env, dna=dna).build_body() + build_brain() into a free slot

This is host-side plumbing for a rare operation (a user saving or loading one organism): it reads and writes the world's
tensors with torch indexing on whatever device they live on. The hot path never goes through it.
"""
from __future__ import annotations

import io
import math
import pickle
import sys
import types

import torch

from .world import CELL_TYPES, F_FINITE_MEMORY, F_MIRROR_X, F_MIRROR_Y, HIST, HM, MAXC, MEMROWS, NPAIR, STM, World

DNA_FIELDS = ("cells", "growth_pattern", "_base_brain_structure", "brain_structure", "trainingInput", "trainingOutput", "trainingReward",
              "trainingHidden", "previousInputs", "previousOutputs", "previousDopamine", "stim_history", "short_term_memory", "hidden_memory",
              "base_info", "cellRange", "sizeRange", "massRange", "mirror_x", "mirror_y", "generation")


def _base_brain():
    return {"input_size": 6, "inputs": ["energy", "health", "rotation", "speed", "x_direction", "y_direction"], "hidden_layers": [],
            "hidden_weights": [], "output_size": 4, "outputs": ["rotation", "speed", "x_direction", "y_direction"], "optimizer": "adam"}


def _num(x):
    """Sizes and masses are ints in the reference (randrange); keep integral values integral."""
    x = float(x)
    return int(x) if x == int(x) else x


def export_dna(world: World, slot: int) -> dict:
    """The genome of the organism in `slot`, with its memories, as the attribute dict of a `sprites.DNA`."""
    t = {k: v.detach().cpu() for k, v in world.t.items() if k.startswith(("c_", "org_"))}
    if int(t["org_state"][slot]) != 1:
        raise ValueError(f"slot {slot} holds no living organism")
    nc, base = int(t["org_ncells"][slot]), slot * MAXC
    rows = sorted(range(nc), key=lambda k: int(t["c_gorder"][base + k]))            # dna.cells order
    cid = lambda k: int(t["c_id"][base + k])
    cells, growth = {}, {}
    for k in rows:
        r = base + k
        par, mir = int(t["c_parent"][r]), int(t["c_mirror"][r])
        cells[cid(k)] = {
            "size": _num(t["c_size"][r]), "type": CELL_TYPES[int(t["c_type"][r])], "elasticity": float(t["c_elast"][r]),
            "friction": float(t["c_fric"][r]), "mass": _num(t["c_mass"][r]), "first": par < 0,
            "relative_pos": [float(t["c_rel"][r][0]), float(t["c_rel"][r][1])], "mirror_self": cid(mir) if mir >= 0 else None,
            "max_health": float(t["c_maxhealth"][r]), "energy_usage": [float(t["c_use_idle"][r]), float(t["c_use_act"][r])],
            "energy_storage": float(t["c_storage"][r]), "damage": float(t["c_damage"][r])}
        growth[cid(k)] = {}
    for k in rows:
        par = int(t["c_parent"][base + k])
        if par >= 0:
            growth[cid(par)][cid(k)] = [float(t["c_grow"][base + k][0]), float(t["c_grow"][base + k][1])]
    nl = int(t["org_nlayers"][slot]); dims = t["org_ldim"][slot][:nl + 1].tolist()
    blob = t["org_brain"][slot]
    H, RC = int(t["org_rnn_h"][slot]), int(world.p.rnn_cap)       # RNN brain: the input layer has H more columns (networks.py:116)
    hidden_w, off = [], 0
    for l in range(nl):
        fin, fout = dims[l] + (H if l == 0 else 0), dims[l + 1]
        if 1 <= l <= nl - 2:
            hidden_w.append(blob[off:off + fout * fin].view(fout, fin).tolist())
        off += fout * fin + fout
    brain = _base_brain(); brain["hidden_layers"] = [int(d) for d in dims[1:-1]]; brain["hidden_weights"] = hidden_w
    I, IN = dims[0] if nl else 0, world.p.in_cap
    flags = int(t["org_flags"][slot])
    mem_n = int(t["org_mem_n"][slot]); mem_in = t["org_mem_in"][slot].view(MEMROWS, IN)
    morder = sorted(range(mem_n), key=lambda m: int(t["org_mem_seq"][slot][m]))       # list order of the curated memory (append stamps)
    hm_app, hm_pop = (int(v) for v in t["org_hm_n"][slot])
    hm = t["org_hm"][slot].view(HM, RC) if RC else None; th = t["org_th"][slot].view(MEMROWS, RC) if RC else None
    hist_n = int(t["org_hist_n"][slot]); hin = t["org_hist_in"][slot].view(HIST, IN); hout = t["org_hist_out"][slot].view(HIST, 4)
    order = [(e % HIST) for e in range(max(0, hist_n - HIST), hist_n)]            # ring slots, oldest first
    stm_n = int(t["org_stm_n"][slot]); sin = t["org_stm_in"][slot].view(STM, IN); sout = t["org_stm_out"][slot].view(STM, 4)
    sorder = [(e % STM) for e in range(max(0, stm_n - STM), stm_n)]
    return {
        "cells": cells, "growth_pattern": growth, "_base_brain_structure": _base_brain(), "brain_structure": brain,
        "trainingInput": [mem_in[m, :I].tolist() for m in morder],
        "trainingOutput": [t["org_mem_out"][slot].view(MEMROWS, 4)[m].tolist() for m in morder],
        "trainingReward": [float(t["org_mem_rew"][slot][m]) for m in morder],
        "trainingHidden": [th[e].tolist() for e in range(int(t["org_th_n"][slot]))] if RC else [],
        "previousInputs": [hin[s, :I].clone() for s in order], "previousOutputs": [hout[s].tolist() for s in order],
        "previousDopamine": [float(t["org_hist_dopa"][slot][s]) for s in order],
        "stim_history": [float(v) for v in t["org_stim_hist"][slot]],
        "short_term_memory": [[sin[s, :I].tolist(), sout[s].tolist(), float(t["org_stm_rew"][slot][s])] for s in sorder],
        "hidden_memory": [hm[e % HM].tolist() for e in range(hm_pop, hm_app)] if RC else [],
        "base_info": {"distanceThreshold": .8, "mirror_x": bool(flags & F_MIRROR_X), "mirror_y": bool(flags & F_MIRROR_Y),
                      "maximum_creation_tries": 30, "curiosity": float(t["org_curiosity"][slot])},
        "cellRange": [int(world.p.cell_lo), int(world.p.cell_hi)], "sizeRange": [int(world.p.size_lo), int(world.p.size_hi)],
        "massRange": [int(world.p.mass_lo), int(world.p.mass_hi)], "mirror_x": [0.6, 0.4], "mirror_y": [0.6, 0.4],
        "generation": int(t["org_generation"][slot]),
    }


class _DNAStub:
    """Stands in for sprites.DNA: pickles under that name, unpickles into a plain attribute bag."""


def dumps_reference_pickle(dna: dict, protocol: int = 4) -> bytes:
    """A pickle the reference opens as `sprites.DNA` (default object reduction: class by name + __dict__)."""
    mod = types.ModuleType("sprites")
    cls = type("DNA", (), {"__module__": "sprites"})
    mod.DNA = cls
    obj = cls.__new__(cls)
    obj.__dict__.update({k: dna[k] for k in DNA_FIELDS if k in dna})
    saved = sys.modules.get("sprites")
    sys.modules["sprites"] = mod
    try:
        return pickle.dumps(obj, protocol=protocol)
    finally:
        if saved is None:
            del sys.modules["sprites"]
        else:
            sys.modules["sprites"] = saved


# What a pickled sprites.DNA can legitimately reference: plain containers and scalars, tensors / arrays of numbers. Exact (module, name)
# pairs only -- a prefix rule would admit builtins.eval, torch.load, numpy.load and the like, i.e. arbitrary code from a genome file.
_PICKLE_ALLOWED = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"), ("collections", "deque"),
    ("builtins", "list"), ("builtins", "dict"), ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "tuple"), ("builtins", "int"),
    ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"), ("builtins", "bytearray"), ("builtins", "complex"),
    ("builtins", "slice"), ("builtins", "range"),
    ("copyreg", "_reconstructor"), ("builtins", "object"), ("_codecs", "encode"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"), ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch.storage", "_load_from_bytes"),
    ("torch", "FloatStorage"), ("torch", "DoubleStorage"), ("torch", "HalfStorage"), ("torch", "LongStorage"), ("torch", "IntStorage"),
    ("torch", "ShortStorage"), ("torch", "CharStorage"), ("torch", "ByteStorage"), ("torch", "BoolStorage"), ("torch", "BFloat16Storage"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"), ("numpy.core.numeric", "_frombuffer"),
    ("numpy._core.numeric", "_frombuffer"),
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == ("sprites", "DNA"):
            return _DNAStub
        if (module, name) == ("torch.storage", "_load_from_bytes"):
            return _load_storage_bytes
        if (module, name) in _PICKLE_ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a genome file")


def _load_storage_bytes(b: bytes):
    """torch.storage._load_from_bytes is torch.load(weights_only=False) on the bytes: replaced by the restricted loader."""
    import torch
    return torch.load(io.BytesIO(b), weights_only=True)


def loads_reference_pickle(blob: bytes) -> dict:
    obj = _Unpickler(io.BytesIO(blob)).load()
    d = dict(obj.__dict__) if isinstance(obj, _DNAStub) else dict(obj)
    missing = [k for k in ("cells", "growth_pattern", "brain_structure", "base_info") if k not in d]
    if missing:
        raise ValueError(f"not a sprites.DNA pickle: missing {missing}")
    return d


def _creation_order(ids, first, parent_of):
    """Organism.build_body creation order (sprites.py:1704-1714, _create_cell :1138-1162): the first cell, then for every cell in
    dna.cells order whose parent exists, the cell followed by its not yet created children."""
    n = len(ids); index = {c: j for j, c in enumerate(ids)}
    created, order = [False] * n, []
    f = next((j for j in range(n) if first[j]), None)
    if f is None:
        return order
    for p in range(n + 1):
        j = f if p == 0 else p - 1
        if created[j]:
            continue
        pj = -1 if first[j] else index.get(parent_of[j], -1)
        if not (first[j] or (pj >= 0 and created[pj])):
            continue
        created[j] = True; order.append(j)
        for k in range(n):
            if k != j and parent_of[k] == ids[j] and not created[k]:
                created[k] = True; order.append(k)
    return order


def import_dna(world: World, slot: int, pos, dna: dict, seed: int | None = None) -> int:
    """Materialises `dna` as a new organism in the free `slot` at `pos`: what Organism(pos, env, dna=dna).build_body() and
    build_brain() do (sprites.py:1683-1729, 255-325). Input / output layers and all biases get nn.Linear's default initialisation
    from a generator seeded with (world seed, uid); hidden weights come from the genome, like in the reference. Returns the uid."""
    t, p = world.t, world.p
    if int(t["org_state"][slot]) != 0:
        raise ValueError(f"slot {slot} is not free")
    ids = list(dna["cells"]); n = len(ids)
    if not 0 < n <= MAXC:
        raise ValueError(f"{n} cells do not fit the {MAXC}-row tile")
    cells = dna["cells"]
    parent_of, grow = {j: 0 for j in range(n)}, {j: (0.0, 0.0) for j in range(n)}
    index = {c: j for j, c in enumerate(ids)}
    for pid, kids in dna["growth_pattern"].items():
        for c, d in kids.items():
            parent_of[index[c]] = pid; grow[index[c]] = (float(d[0]), float(d[1]))
    first = [bool(cells[c]["first"]) for c in ids]
    order = _creation_order(ids, first, parent_of)
    if len(order) != n:
        raise ValueError("genome has cells that are not connected to the first cell")
    row_of = {j: k for k, j in enumerate(order)}
    hidden = [int(h) for h in dna["brain_structure"]["hidden_layers"]]
    eyes = [j for j in range(n) if cells[ids[j]]["type"] == "eye"]; olf = [j for j in range(n) if cells[ids[j]]["type"] == "olfactory"]
    # neural.setup_network returns no network for a genome without hidden layers (neural.py:471): the organism lives without a brain
    dims = [8 * len(eyes) + len(olf) + 4] + hidden + [4] if hidden else [8 * len(eyes) + len(olf) + 4]
    H, RC = int(p.rnn_hidden), int(p.rnn_cap)
    if not hidden:
        H = 0                     # build_brain: an RNNetwork when use_rnn is set now (sprites.py:1683-1690)
    nparam = sum(dims[l + 1] * dims[l] + dims[l + 1] for l in range(len(dims) - 1)) + (H * (hidden[0] + hidden[-1]) + H if hidden else 0)
    if dims[0] + H > p.in_cap or max(hidden, default=0) > p.width_cap or nparam > p.brain_cap or len(dims) - 1 > 15:
        raise ValueError("genome exceeds the tile caps (in_cap / width_cap / brain_cap)")
    dev = t["c_pos"].device
    base = slot * MAXC
    uid = int(t["g_next_uid"][0]); t["g_next_uid"] += 1
    tick = int(t["g_tick"][0])
    f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
    t["c_alive"][base:base + MAXC] = 0
    sum_mh = sum_st = 0.0
    for k, j in enumerate(order):
        c, r = cells[ids[j]], base + k
        x, y = float(pos[0]) + float(c["relative_pos"][0]), float(pos[1]) + float(c["relative_pos"][1])
        t["c_alive"][r] = 1; t["c_type"][r] = CELL_TYPES.index(c["type"]); t["c_inuse"][r] = 0
        t["c_parent"][r] = -1 if first[j] else row_of[index[parent_of[j]]]
        t["c_mirror"][r] = row_of[index[c["mirror_self"]]] if c.get("mirror_self") else -1
        t["c_gorder"][r] = j; t["c_id"][r] = int(ids[j])
        for name in ("c_pos", "c_spos"):
            t[name][r] = f32([x, y])
        for name in ("c_vel", "c_svel", "c_vbias", "c_svbias"):
            t[name][r] = 0.0
        for name in ("c_ang", "c_angvel", "c_wbias", "c_spin_jn", "c_dmg"):
            t[name][r] = 0.0
        t["c_health"][r] = float(c["max_health"]); t["c_energy"][r] = float(c["energy_storage"]) / 2.0
        t["c_size"][r] = float(c["size"]); t["c_mass"][r] = float(c["mass"]); t["c_elast"][r] = float(c["elasticity"]); t["c_fric"][r] = float(c["friction"])
        t["c_rel"][r] = f32([float(c["relative_pos"][0]), float(c["relative_pos"][1])]); t["c_grow"][r] = f32(list(grow[j]))
        t["c_maxhealth"][r] = float(c["max_health"]); t["c_use_idle"][r] = float(c["energy_usage"][0]); t["c_use_act"][r] = float(c["energy_usage"][1])
        t["c_storage"][r] = float(c["energy_storage"]); t["c_damage"][r] = float(c["damage"])
        sum_mh += float(c["max_health"]); sum_st += float(c["energy_storage"])
    bi = dna["base_info"]
    t["org_state"][slot] = 1
    t["org_flags"][slot] = F_FINITE_MEMORY | (F_MIRROR_X if bi.get("mirror_x") else 0) | (F_MIRROR_Y if bi.get("mirror_y") else 0)
    t["org_ncells"][slot] = n; t["org_uid"][slot] = uid; t["org_generation"][slot] = int(dna.get("generation", 0))
    for name in ("org_birth_tick", "org_lastbirth_tick", "org_think_tick", "org_train_tick", "org_dopa_tick"):
        t[name][slot] = tick
    for name in ("org_consumed", "org_dopamine", "org_pain", "org_energy_diff", "org_boredom", "org_stimulation", "org_loss"):
        t[name][slot] = 0.0
    t["org_move"][slot] = 0.0; t["org_out_raw"][slot] = 0.0
    t["org_pos"][slot] = f32([float(pos[0]), float(pos[1])]); t["org_curiosity"][slot] = float(bi.get("curiosity", 0.5))
    t["org_last_health"][slot] = 100.0 if sum_mh > 0 else 0.0; t["org_last_energy"][slot] = 50.0 if sum_st > 0 else 0.0
    t["org_ic_n"][slot] = 0; t["org_req"][slot] = 0; t["org_adam_t"][slot] = 0
    t["org_gas_a"][slot] = 1.0; t["org_gas_b"][slot] = 0.0; t["org_gas_in"][slot] = 0.0; t["org_gas_refund"][slot] = 0.0
    inc = [row_of[j] for j in eyes] + [row_of[j] for j in olf]                    # neural.setup_network wiring: eyes, then olfactory cells
    t["org_nincell"][slot] = len(inc); t["org_incell"][slot] = 0
    if inc:
        t["org_incell"][slot][:len(inc)] = torch.tensor(inc, dtype=t["org_incell"].dtype, device=dev)
    t["org_ldim"][slot] = 0; t["org_ldim"][slot][:len(dims)] = torch.tensor(dims, dtype=torch.int32, device=dev)
    t["org_nlayers"][slot] = len(dims) - 1
    # brain: nn.Linear default init (U(+-1/sqrt(fan_in)) for weights and biases), hidden weights from the genome
    g = torch.Generator().manual_seed(((int(p.seed) if seed is None else int(seed)) * 1000003 + uid) & 0x7FFFFFFFFFFFFFFF)
    blob, off = torch.zeros(p.brain_cap, dtype=torch.float32), 0
    hw = dna["brain_structure"].get("hidden_weights", [])
    nlin = len(dims) - 1
    for l in range(nlin + (1 if H else 0)):                       # the hidden head outputLayerH follows the output layer (networks.py:151-166)
        fin, fout = (dims[nlin - 1], H) if l == nlin else (dims[l] + (H if l == 0 else 0), dims[l + 1])
        bound = 1.0 / math.sqrt(fin)
        blob[off:off + fout * fin + fout] = (torch.rand(fout * fin + fout, generator=g) * 2 - 1) * bound
        if 1 <= l <= len(dims) - 3 and l - 1 < len(hw):
            wl = torch.tensor(hw[l - 1], dtype=torch.float32)
            if tuple(wl.shape) == (fout, fin):
                blob[off:off + fout * fin] = wl.reshape(-1)
        off += fout * fin + fout
    t["org_rnn_h"][slot] = H; t["org_hm_n"][slot] = 0; t["org_th_n"][slot] = 0
    if RC:
        t["org_last_hidden"][slot] = 0.0
        hml = [e for e in dna.get("hidden_memory", []) if len(e) == RC][: HM - STM]           # what a deepcopied DNA carries (sprites.py:386, not erased)
        thl = [e for e in dna.get("trainingHidden", []) if len(e) == RC][:MEMROWS]
        if hml:
            t["org_hm"][slot].view(HM, RC)[: len(hml)] = torch.tensor(hml, dtype=torch.float32, device=dev)
        if thl:
            t["org_th"][slot].view(MEMROWS, RC)[: len(thl)] = torch.tensor(thl, dtype=torch.float32, device=dev)
        t["org_hm_n"][slot][0] = len(hml); t["org_th_n"][slot] = len(thl)
    t["org_brain"][slot] = blob.to(dev)
    t["org_adam_m"][slot] = 0.0; t["org_adam_v"][slot] = 0.0
    t["j_jn"][slot * NPAIR:(slot + 1) * NPAIR] = 0.0
    # memories
    I, IN = dims[0], p.in_cap
    ti, to, tr = dna.get("trainingInput", []), dna.get("trainingOutput", []), dna.get("trainingReward", [])
    m = min(len(ti), len(to), len(tr), MEMROWS)
    keep = [e for e in range(m) if len(ti[e]) == I]
    t["org_mem_n"][slot] = len(keep)
    t["org_mem_seq"][slot] = 0; t["org_mem_seq"][slot][: len(keep)] = torch.arange(len(keep), dtype=torch.int32, device=dev); t["org_mem_stamp"][slot] = len(keep)
    mem_in = torch.zeros(MEMROWS, IN); mem_out = torch.zeros(MEMROWS, 4); mem_rew = torch.zeros(MEMROWS)
    for q, e in enumerate(keep):
        mem_in[q, :I] = torch.tensor(ti[e], dtype=torch.float32); mem_out[q] = torch.tensor(to[e], dtype=torch.float32); mem_rew[q] = float(tr[e])
    t["org_mem_in"][slot] = mem_in.view(-1).to(dev); t["org_mem_out"][slot] = mem_out.view(-1).to(dev); t["org_mem_rew"][slot] = mem_rew.to(dev)
    pin, pout, pdo = dna.get("previousInputs", []), dna.get("previousOutputs", []), dna.get("previousDopamine", [])
    h = min(len(pin), len(pout), len(pdo))
    hist_in = torch.zeros(HIST, IN); hist_out = torch.zeros(HIST, 4); hist_do = torch.zeros(HIST)
    ok = h > 0 and all(len(pin[e]) == I for e in range(max(0, h - HIST), h))
    if ok:
        for e in range(max(0, h - HIST), h):
            hist_in[e % HIST, :I] = torch.as_tensor(pin[e], dtype=torch.float32); hist_out[e % HIST] = torch.tensor(pout[e], dtype=torch.float32); hist_do[e % HIST] = float(pdo[e])
    t["org_hist_n"][slot] = h if ok else 0
    t["org_hist_in"][slot] = hist_in.view(-1).to(dev); t["org_hist_out"][slot] = hist_out.view(-1).to(dev); t["org_hist_dopa"][slot] = hist_do.to(dev)
    stm = [s for s in dna.get("short_term_memory", []) if len(s[0]) == I][-STM:]
    s_in = torch.zeros(STM, IN); s_out = torch.zeros(STM, 4); s_rew = torch.zeros(STM)
    for e, (xi, yo, rw) in enumerate(stm):
        s_in[e, :I] = torch.tensor(xi, dtype=torch.float32); s_out[e] = torch.tensor(yo, dtype=torch.float32); s_rew[e] = float(rw)
    t["org_stm_n"][slot] = len(stm)
    t["org_stm_in"][slot] = s_in.view(-1).to(dev); t["org_stm_out"][slot] = s_out.view(-1).to(dev); t["org_stm_rew"][slot] = s_rew.to(dev)
    sh = list(dna.get("stim_history", [0.0] * 10))[-10:]
    t["org_stim_hist"][slot] = f32([0.0] * (10 - len(sh)) + [float(v) for v in sh])
    return uid

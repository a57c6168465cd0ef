"""Cuts organisms out of a large aged world into a small one and compares ONE tick of the small world GPU vs oracle, organism by organism --
a failure of test_100k_full_tick_matches_oracle (340 s, 43 GiB of host copies) becomes a case of a few organisms that runs in seconds.
  python tools/debug_extract.py [organisms] [age] [sample] [slot ...]     -> gpurun_out/extract_<n>.npz (state of the worst organisms BEFORE the tick)
  python tools/debug_extract.py --replay gpurun_out/extract_<n>.npz       -> the same comparison from the saved state
Organisms keep their positions; contacts with bodies that were left behind vanish on both sides alike (cached cross contacts are dropped)."""
import math, sys
import numpy as np
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sapiogenesis_b200.world import FIELDS, ICAP, MAXC, NPAIR, PH_ALL, SapioParams, World
from sapiogenesis_b200.engine import Engine

MULT = {"ORG": 1, "CELL": MAXC, "PAIR": NPAIR, "ICON": ICAP}


def cut(w, slots, n_dead=1024):
    """small CPU world holding the organisms `slots` of w (any device) in slots 0.."""
    p = SapioParams.from_buffer_copy(bytes(w.p))
    p.n_org = int(math.ceil(len(slots) / 64.0) * 64); p.n_dead = n_dead; p.x_cap = max(4096, 16 * len(slots))
    p.population_limit = 0                                   # nobody reproduces: the cut is about the step
    s = World(p, "cpu")
    idx = torch.as_tensor(np.asarray(slots, dtype=np.int64), device=w.device)
    for f in FIELDS:
        if w.t[f.name].numel() == 0:
            continue
        if f.scope in MULT:
            src = w.t[f.name].view(w.p.n_org, -1)[idx].cpu()
            s.t[f.name].view(p.n_org, -1)[: len(slots)] = src.view(len(slots), -1)
        elif f.scope == "GLOB" and f.name not in ("g_x_n", "g_population", "g_stats"):
            s.t[f.name].copy_(w.t[f.name].cpu())
    s.t["g_population"][0] = len(slots)
    return s


def save(s, path, n):
    out = {"params": np.frombuffer(bytes(s.p), dtype=np.uint8), "n": np.asarray([n])}
    for f in FIELDS:
        if s.t[f.name].numel() == 0:
            continue
        if f.scope in MULT:
            out[f.name] = s.t[f.name].view(s.p.n_org, -1)[:n].numpy()
        elif f.scope == "GLOB":
            out[f.name] = s.t[f.name].numpy()
    np.savez_compressed(path, **out)


def load(path):
    z = np.load(path)
    p = SapioParams.from_buffer_copy(z["params"].tobytes())
    s = World(p, "cpu"); n = int(z["n"][0])
    for f in FIELDS:
        if f.name not in z.files:
            continue
        if f.scope in MULT:
            s.t[f.name].view(p.n_org, -1)[:n] = torch.from_numpy(z[f.name]).view(n, -1)
        elif f.scope == "GLOB":
            s.t[f.name].copy_(torch.from_numpy(z[f.name]))
    return s, n


def compare(s, n, tag, phases=PH_ALL):
    from oracle import oracle
    g = s.to("cuda"); eng = Engine(g); eng.exclusive = True
    eng.tick(1, phases); eng.synchronize()
    h = s.clone()
    assert oracle.tick(h, phases, 1) == 0
    gv, ov = g.np("c_vel").astype(np.float64).reshape(-1, MAXC, 2), h.np("c_vel").astype(np.float64).reshape(-1, MAXC, 2)
    live = (h.np("c_alive") == 1).reshape(-1, MAXC)
    rms = float(np.sqrt((ov[live] ** 2).mean()))
    err = (np.abs(gv - ov).max(axis=2) * live).max(axis=1)[:n]
    pe = (np.abs(g.np("c_pos").astype(np.float64) - h.np("c_pos").astype(np.float64)).reshape(-1, MAXC, 2).max(axis=2) * live).max(axis=1)[:n]
    order = np.argsort(-err)
    print(f"{tag}: {n} organisms, velocity rms {rms:.3f} px/s, organisms above 1e-4 rms: {int((err > 1e-4 * rms).sum())}, above 1e-5: {int((err > 1e-5 * rms).sum())}, positions differing: {int((pe > 0).sum())}")
    for o in order[:12]:
        nl = int(live[o].sum())
        print(f"  slot {o}: |dv| {err[o]:.3e} ({err[o] / rms:.1e} rms)  dpos {pe[o]:.2e}  rows {int(h.np('org_ncells')[o])} living {nl} intra contacts gpu {int(g.np('org_ic_n')[o])} oracle {int(h.np('org_ic_n')[o])}"
              f"  |v|max {np.abs(ov[o][live[o]]).max() if nl else 0:.2f} state {int(h.np('org_state')[o])}/{int(g.np('org_state')[o])} xcontacts {int(h.np('g_x_n')[0])}")
    return err, pe, rms


if __name__ == "__main__":
    if sys.argv[1] == "--replay":
        s, n = load(sys.argv[2])
        ph = int(sys.argv[3]) if len(sys.argv) > 3 else PH_ALL
        compare(s, n, sys.argv[2], ph)
        sys.exit(0)
    from sapiogenesis_b200.synth import build_world
    norg, age, sample = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    named = [int(v) for v in sys.argv[4:]]
    w, eng = build_world(norg, seed=0, slot_factor=1.02); eng.exclusive = True
    eng.tick(age); eng.synchronize()
    alive = torch.nonzero(w.t["org_state"] == 1).view(-1).cpu().numpy()
    rng = np.random.default_rng(1)
    pick = np.unique(np.concatenate([np.asarray(named, dtype=np.int64), rng.choice(alive, size=min(sample, len(alive)), replace=False)]))
    s = cut(w, pick)
    err, pe, rms = compare(s, len(pick), f"cut of {len(pick)} organisms at tick {w.tick}")
    for v in named:
        k = int(np.nonzero(pick == v)[0][0]); print(f"named slot {v} -> {k}: |dv| {err[k]:.3e} dpos {pe[k]:.2e}")
    bad = np.nonzero((err > 1e-4 * rms) | (pe > 0))[0][:48]
    if len(bad):
        s2 = cut(s, bad)
        save(s2, f"gpurun_out/extract_{len(bad)}.npz", len(bad))
        print("saved", len(bad), "organisms:", pick[bad].tolist())
        compare(s2, len(bad), "replay of the saved cut")

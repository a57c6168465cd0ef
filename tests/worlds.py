"""Scenario worlds for the parity tests, built on the HOST with the oracle's spawn recipe
(add_creature_clicked, Sapiogenesis.py:297-323) so CPU and GPU start from identical bytes."""
from __future__ import annotations

import math

import numpy as np
import torch

from sapiogenesis_b200.world import MAXC, World, default_params


def add_food(w: World, slot: int, x: float, y: float, vx: float = 0.0, vy: float = 0.0, size: float = 30.0, mass: float = 12.0,
             owner: int = -1, state: int = 1):
    """add_dead_clicked (Sapiogenesis.py:490-507): size 30, mass 12, e .5, mu .5"""
    t = w.t
    t["d_state"][slot] = state
    t["d_pos"][slot] = torch.tensor([x, y])
    t["d_vel"][slot] = torch.tensor([vx, vy])
    t["d_ang"][slot] = 0; t["d_angvel"][slot] = 0
    t["d_vbias"][slot] = 0; t["d_wbias"][slot] = 0
    t["d_mass"][slot] = mass; t["d_mass0"][slot] = mass; t["d_size"][slot] = size
    t["d_elast"][slot] = 0.5; t["d_fric"][slot] = 0.5
    t["d_owner"][slot] = owner


def make_world(oracle, n_org: int = 12, seed: int = 1, width: float = 3180.0, height: float = 1800.0, n_food: int = 40,
               crowd: float = 1.0, kick: float = 60.0, n_dead_cap: int | None = None, extra_slots: int = 8, **kw) -> World:
    """n_org organisms on a jittered grid. crowd < 1 squeezes the grid so that neighbours overlap (cross contacts);
    food is dropped partly onto organisms and partly against the walls; every body gets a random velocity."""
    rng = np.random.default_rng(seed)
    p = default_params(n_org=n_org + extra_slots, n_dead=n_dead_cap or max(64, 4 * n_food), width=width, height=height,
                       seed=seed, population_limit=n_org + extra_slots, **kw)
    w = World(p)
    w.t["g_co2"][0] = 30000.0 * n_org / 12.0
    cols = max(1, int(math.ceil(math.sqrt(n_org * width / height))))
    rows = int(math.ceil(n_org / cols))
    dx, dy = (width - 400) / cols * crowd, (height - 400) / rows * crowd
    for i in range(n_org):
        cx = 200 + (i % cols + 0.5) * dx + rng.uniform(-0.15, 0.15) * dx
        cy = 200 + (i // cols + 0.5) * dy + rng.uniform(-0.15, 0.15) * dy
        assert oracle.spawn(w, i, float(cx), float(cy)) == 0
    nc = w.np("org_ncells")
    cpos = w.np("c_pos")
    alive_rows = np.nonzero(w.np("c_alive") == 1)[0]
    for f in range(n_food):
        mode = f % 4
        if mode == 0:        # on top of a living cell
            r = alive_rows[rng.integers(len(alive_rows))]
            x, y = cpos[r] + rng.uniform(-12, 12, 2)
        elif mode == 1:      # against a wall
            side = rng.integers(4)
            x = [22.0, width - 22.0, rng.uniform(50, width - 50), rng.uniform(50, width - 50)][side] + rng.uniform(-6, 6)
            y = [rng.uniform(50, height - 50), rng.uniform(50, height - 50), 22.0, height - 22.0][side] + rng.uniform(-6, 6)
        elif mode == 2 and f >= 4:   # touching the previous food -> dead/dead contacts and chains
            px, py = w.np("d_pos")[f - 1]
            x, y = px + rng.uniform(10, 24), py + rng.uniform(-10, 10)
        else:
            x, y = rng.uniform(100, width - 100), rng.uniform(100, height - 100)
        add_food(w, f, float(x), float(y), float(rng.normal(0, kick)), float(rng.normal(0, kick)),
                 size=float(rng.integers(10, 40)), mass=float(rng.integers(3, 19)))
    # random motion for every cell
    n = p.n_org * MAXC
    w.t["c_vel"][:] = torch.from_numpy(rng.normal(0, kick, (n, 2)).astype(np.float32))
    w.t["c_angvel"][:] = torch.from_numpy(rng.normal(0, 2.0, n).astype(np.float32))
    w.t["c_svel"][:] = torch.from_numpy(rng.normal(0, kick, (n, 2)).astype(np.float32))
    mask = torch.from_numpy((w.np("c_alive") != 1))
    for name in ("c_vel", "c_svel"):
        w.t[name][mask] = 0
    w.t["c_angvel"][mask] = 0
    w.t["org_move"][:n_org] = torch.from_numpy(rng.uniform(-1, 1, (n_org, 4)).astype(np.float32))
    oracle.post(w)
    return w


def live_rows(w: World) -> np.ndarray:
    return np.nonzero(w.np("c_alive") == 1)[0]


def rel_err(a, b, scale=None):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    s = np.maximum(np.abs(b), scale if scale is not None else (np.abs(b).max() if b.size else 1.0))
    s = np.where(s == 0, 1.0, s)
    return float(np.max(np.abs(a - b) / s)) if a.size else 0.0


def manual_org(w: World, slot: int, pos, cells, uid: int | None = None, hidden=(5, 4)):
    """Hand-built organism: cells = [dict(type, size, mass, rel=(x, y), elast, fric, parent)] in creation order, row 0 = heart.
    Fills exactly what Organism.build_body would (sprites.py:1698-1729) plus a small brain."""
    from sapiogenesis_b200.world import derived_stats
    t = w.t
    nc = len(cells)
    t["org_state"][slot] = 1
    t["org_flags"][slot] = 4
    t["org_ncells"][slot] = nc
    t["org_uid"][slot] = uid if uid is not None else int(t["g_next_uid"].item())
    t["g_next_uid"] += 1
    tick = int(t["g_tick"].item())
    for name in ("org_birth_tick", "org_lastbirth_tick", "org_think_tick", "org_train_tick", "org_dopa_tick"):
        t[name][slot] = tick
    t["org_pos"][slot] = torch.tensor(pos, dtype=torch.float32)
    t["org_curiosity"][slot] = 0.5
    for k, c in enumerate(cells):
        row = slot * MAXC + k
        t["c_alive"][row] = 1
        t["c_type"][row] = c["type"]
        t["c_parent"][row] = c.get("parent", -1 if k == 0 else 0)
        t["c_mirror"][row] = -1
        t["c_gorder"][row] = k
        t["c_id"][row] = k + 1
        rel = c.get("rel", (0.0, 0.0))
        t["c_rel"][row] = torch.tensor(rel, dtype=torch.float32)
        t["c_pos"][row] = torch.tensor([pos[0] + rel[0], pos[1] + rel[1]], dtype=torch.float32)
        t["c_spos"][row] = t["c_pos"][row]
        t["c_size"][row] = c["size"]; t["c_mass"][row] = c["mass"]
        t["c_elast"][row] = c.get("elast", 0.5); t["c_fric"][row] = c.get("fric", 0.5)
        mh, ui, ua, st, dm = derived_stats(c["type"], c["size"], c["mass"])
        t["c_maxhealth"][row] = mh; t["c_use_idle"][row] = ui; t["c_use_act"][row] = ua
        t["c_storage"][row] = st; t["c_damage"][row] = dm
        t["c_health"][row] = c.get("health", mh); t["c_energy"][row] = c.get("energy", st / 2)
    # brain: inputs = 8/eye + 1/olfactory + 4
    eyes = [k for k, c in enumerate(cells) if c["type"] == 2]
    olf = [k for k, c in enumerate(cells) if c["type"] == 3]
    inc = eyes + olf
    t["org_nincell"][slot] = len(inc)
    for i, k in enumerate(inc):
        t["org_incell"][slot, i] = k
    dims = [8 * len(eyes) + len(olf) + 4] + list(hidden) + [4]
    t["org_nlayers"][slot] = len(dims) - 1
    t["org_ldim"][slot, :len(dims)] = torch.tensor(dims, dtype=torch.int32)
    g = torch.Generator().manual_seed(1234 + slot)
    off = 0
    for l in range(len(dims) - 1):
        n = dims[l + 1] * dims[l] + dims[l + 1]
        b = 1.0 / math.sqrt(dims[l])
        t["org_brain"][slot, off:off + n] = (torch.rand(n, generator=g) * 2 - 1) * b
        off += n
    hs = sum(float(t["c_health"][slot * MAXC + k]) for k in range(nc)); mh = sum(float(t["c_maxhealth"][slot * MAXC + k]) for k in range(nc))
    es = sum(float(t["c_energy"][slot * MAXC + k]) for k in range(nc)); st = sum(float(t["c_storage"][slot * MAXC + k]) for k in range(nc))
    t["org_last_health"][slot] = hs / mh * 100.0
    t["org_last_energy"][slot] = es / st * 100.0 if es else 0.0

"""Synthetic worlds of the sizes the north star names.

The reference world is empty until the user clicks organisms into it (add_creature_clicked, Sapiogenesis.py:297-323) and its
population limit is 12; worlds of 10k..1M organisms are generated here with the reference's own spawn recipe, on the device:
  - genomes: DNA().randomize with the UI-default ranges (sapio_spawn: Philox uniforms in the reference's draw order)
  - square world, side sqrt(N * area_per_org); default area = 3180 * 1800 / 12 px^2 (the reference's density at its limit)
  - one organism per cell of a jittered grid; no two spawn boxes overlap (grid pitch > largest genome extent)
  - food: N/4 dead cells of the add_dead_clicked recipe (Sapiogenesis.py:490-507), uniformly placed
  - gases: co2 = 30000 * (W + H) / 4980, oxygen = 0: the same co2 intake per co2C cell and tick as in the default world
  - brain / training / reproduction clocks staggered over their periods, so every tick sees 1/17 of the brains
    (the reference staggers them by spawn time)
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .engine import Engine
from .world import MAXC, World, default_params

REF_AREA = 3180.0 * 1800.0 / 12.0


def build_world(n_organisms: int, device="cuda", seed: int = 0, area_per_org: float = REF_AREA, slot_factor: float = 1.25,
                food_per_org: float = 0.25, in_cap: int = 128, brain_cap: int = 32768, stagger: bool = True,
                population_limit: int | None = None):
    """Returns (world, engine) with n_organisms living organisms on `device`."""
    n_slots = int(math.ceil(n_organisms * slot_factor / 64.0) * 64)
    n_food = int(n_organisms * food_per_org)
    n_dead = max(1024, int(n_organisms * 24))          # a starvation wave leaves ~20 dead cells per organism lying around for ~1400 ticks
    side = math.sqrt(n_organisms * area_per_org)
    p = default_params(n_org=n_slots, n_dead=n_dead, width=side, height=side, in_cap=in_cap, brain_cap=brain_cap, seed=seed,
                       population_limit=population_limit if population_limit is not None else n_slots - max(64, n_slots // 100),
                       x_cap=max(4096, 4 * n_organisms))
    w = World(p, device=device)
    eng = Engine(w)
    rng = np.random.default_rng(seed)
    cols = int(math.ceil(math.sqrt(n_organisms)))
    pitch = (side - 200.0) / cols
    idx = np.arange(n_organisms)
    jitter = rng.uniform(-0.12, 0.12, (n_organisms, 2)) * pitch
    pos = np.stack([100.0 + (idx % cols + 0.5) * pitch, 100.0 + (idx // cols + 0.5) * pitch], axis=1) + jitter
    # gases: a co2C cell converts size / (W + H) * dt of the world's co2 per tick (sprites.py:1413-1416), so the reservoir is
    # scaled with (W + H) to give every cell the default world's intake at t = 0 (30000 at W + H = 4980)
    w.t["g_co2"][0] = 30000.0 * (2.0 * side) / (3180.0 + 1800.0)
    eng.spawn(np.arange(n_organisms, dtype=np.int32), pos)
    t = w.t
    if n_food:
        fpos = torch.from_numpy(rng.uniform(60.0, side - 60.0, (n_food, 2)).astype(np.float32)).to(device)
        t["d_state"][:n_food] = 1
        t["d_pos"][:n_food] = fpos
        t["d_size"][:n_food] = 30.0; t["d_mass"][:n_food] = 12.0; t["d_mass0"][:n_food] = 12.0       # Sapiogenesis.py:490-507
        t["d_elast"][:n_food] = 0.5; t["d_fric"][:n_food] = 0.5; t["d_owner"][:n_food] = -1
    if stagger:
        o = torch.arange(n_organisms, device=device, dtype=torch.int32)
        t["org_think_tick"][:n_organisms] = -(o % p.think_ticks)
        t["org_dopa_tick"][:n_organisms] = -(o % p.think_ticks) - 1
        t["org_train_tick"][:n_organisms] = -((o + 7) % p.train_ticks)
        t["org_lastbirth_tick"][:n_organisms] = -(o % p.repro_ticks)
    eng.post()
    eng.synchronize()
    return w, eng


def world_stats(w: World) -> dict:
    """Realised sizes (cells, sensors, pin joints, contacts) -- the inputs of the algorithmic-byte formulas in DESIGN.md."""
    alive = (w.t["c_alive"] == 1)
    org = (w.t["org_state"] == 1)
    cells_per = alive.view(-1, MAXC).sum(1)[org].to(torch.int64)
    types = w.t["c_type"]
    sensors = int((alive & ((types == 2) | (types == 3))).sum().item())
    return dict(organisms=int(org.sum().item()), cells=int(alive.sum().item()), sensors=sensors,
                pairs=int((cells_per * (cells_per - 1) // 2).sum().item()), dead=int((w.t["d_state"] > 0).sum().item()),
                intra_contacts=int(w.t["org_ic_n"][org].sum().item()), cross_contacts=int(w.t["g_x_n"].item()),
                brain_params=int(_brain_params(w, org)))


def _brain_params(w: World, org) -> int:
    ld = w.t["org_ldim"][org].to(torch.int64)
    nl = w.t["org_nlayers"][org].to(torch.int64)
    tot = torch.zeros((), dtype=torch.int64, device=ld.device)
    for l in range(15):
        m = (l < nl)
        tot += ((ld[:, l + 1] * ld[:, l] + ld[:, l + 1]) * m).sum()
    return int(tot.item())

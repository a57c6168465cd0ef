"""Spatial slabs: one large world cut along x into N slabs, one per GPU (second mode; BASELINE configs[4]).

A slab owns the organisms whose heart cell lies in [x0, x1) and the dead cells whose centre does, and carries copies ("ghosts") of
what its neighbours own within `halo` of the common boundary (csrc/halo.cuh). The world step of a slab is the ordinary one; per tick
the slabs exchange

  * the organisms (ragged records, include/sapio_migrate.h) and dead cells next to a boundary -- neighbour send / recv (NCCL p2p);
  * three scalars composed in RANK ORDER, because the reference updates the two global gas scalars cell by cell in list order
    (sprites.py:1409-1439) and those updates do not commute: every slab reduces its organisms to one affine map of the co2 level
    (the ordered scan the single world runs, SURVEY C-6), the maps are allgathered, slab r starts from the composition of the maps
    before it, and every slab ends the phase with the same co2 / oxygen. Refunds and the dead cells' dispersion factor likewise;
  * the owned population (allgather of one int) for the reproduction gate.

Nothing in the loop waits for the device. `SlabDriver` runs any number of slabs per process: one per rank under torch.distributed, or
several on one device in one process (`LocalComm`) -- that form is how the single-GPU test suite compares a split world with the whole.
"""
from __future__ import annotations

import ctypes
import math
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import Engine
from .world import FIELDS, MAXC, PH_ALL, PH_REUSE_GRID, PH_THINK, SapioParams, World, default_params, F_GHOST

HALO = 1300.0          # px: sensor radius (<= 615) + solid radius + extent of the largest organisms + what a body moves in a tick


class Slab:
    """One slab: world + engine + its halo buffers."""

    def __init__(self, world: World, engine: Engine, rank: int, n_slabs: int, halo: float = HALO, record_budget: int = 1024, dead_budget: int = 16384):
        assert world.p.slab_mode == 1
        self.world, self.engine, self.rank, self.n, self.halo = world, engine, int(rank), int(n_slabs), float(halo)
        self.lib = engine.lib
        dev = world.device
        off = (ctypes.c_int64 * 4)()
        with torch.cuda.device(dev):
            _lib.check(self.lib.sapio_slab_offsets(engine._ref(), *engine._ws, off), "sapio_slab_offsets")
        ws = engine.workspace
        self.gas_map = ws[off[0]: off[0] + 16].view(torch.float64)          # (a, b) of this slab's composed map, after rules part 1
        self.refund = ws[off[1]: off[1] + 8].view(torch.float64)
        self.disp = ws[off[2]: off[2] + 8].view(torch.float64)
        self.pop = ws[off[3]: off[3] + 4].view(torch.int32)
        nb = int(self.lib.sapio_migrate_buffer_bytes(ctypes.byref(world.p), record_budget, 65536))
        db = 16 + 72 * dead_budget
        mk = lambda n: torch.zeros(n, dtype=torch.uint8, device=dev)
        self.send = [(mk(nb), mk(db)), (mk(nb), mk(db))]                       # [side]: 0 to / from the left neighbour, 1 the right one
        self.recv = [(mk(nb), mk(db)), (mk(nb), mk(db))]
        self.halo_bytes = 0

    def _call(self, fn, *args):
        e = self.engine
        with torch.cuda.device(self.world.device):
            _lib.check(fn(e._ref(), *e._ws, *args, e._stream_ptr()), fn.__name__)

    def has_neighbour(self, side: int) -> bool:
        return self.rank > 0 if side == 0 else self.rank < self.n - 1

    def pack(self, side: int):
        p = self.world.p
        edge = p.slab_x0 if side == 0 else p.slab_x1
        b, d = self.send[side]
        self._call(self.lib.sapio_halo_pack, ctypes.c_float(edge - self.halo), ctypes.c_float(edge + self.halo), ctypes.c_void_p(b.data_ptr()), b.numel(),
                   ctypes.c_void_p(d.data_ptr()), d.numel())

    def unpack(self, side: int):
        b, d = self.recv[side]
        self._call(self.lib.sapio_halo_unpack, side, ctypes.c_void_p(b.data_ptr()), b.numel(), ctypes.c_void_p(d.data_ptr()), d.numel())

    def sweep(self):
        self._call(self.lib.sapio_halo_sweep)
        self.engine._grid_valid = False

    def rules_part(self, part: int, phases: int = PH_ALL):
        e = self.engine
        ph = phases | (PH_REUSE_GRID if (e._grid_valid and (phases & PH_THINK)) else 0)
        self._call(self.lib.sapio_rules_part, ph, part)

    def owned(self) -> torch.Tensor:
        t = self.world.t
        return (t["org_state"] == 1) & ((t["org_flags"] & F_GHOST) == 0)


class LocalComm:
    """Several slabs in one process: the collectives are tensor copies (test / single-GPU form)."""

    def __init__(self, n: int):
        self.n = n

    def allgather(self, parts: List[torch.Tensor]) -> List[torch.Tensor]:
        g = torch.stack([p.reshape(-1) for p in parts])
        return [g for _ in parts]

    def exchange(self, slabs: List[Slab]):
        for s in slabs:                                                     # what rank r sends right is what rank r + 1 receives from its left
            for side in (0, 1):
                if not s.has_neighbour(side):
                    continue
                other = slabs[s.rank + (1 if side == 1 else -1)]
                for k in (0, 1):
                    other.recv[1 - side][k].copy_(s.send[side][k])


class DistComm:
    """One slab per rank: torch.distributed (NCCL over NVLink between GPUs)."""

    def __init__(self, group=None):
        self.group = group
        self.n = dist.get_world_size(group)

    def allgather(self, parts: List[torch.Tensor]) -> List[torch.Tensor]:
        p = parts[0].reshape(-1).contiguous()
        out = torch.empty(self.n * p.numel(), dtype=p.dtype, device=p.device)          # flat: gloo wants chunks of the input's shape
        dist.all_gather_into_tensor(out, p, group=self.group)
        return [out.view(self.n, p.numel())]

    def exchange(self, slabs: List[Slab]):
        s = slabs[0]
        ops = []
        for side in (0, 1):
            if not s.has_neighbour(side):
                continue
            peer = s.rank + (1 if side == 1 else -1)
            for k in (0, 1):
                ops.append(dist.P2POp(dist.isend, s.send[side][k], peer, self.group))
                ops.append(dist.P2POp(dist.irecv, s.recv[side][k], peer, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()


class SlabDriver:
    """Runs the slabs of this process in lockstep: tick() == one tick of the whole world."""

    def __init__(self, slabs: List[Slab], comm):
        self.slabs, self.comm = slabs, comm
        self.ticks = 0

    def _gather(self, getter):
        return self.comm.allgather([getter(s) for s in self.slabs])

    def exchange(self):
        for s in self.slabs:
            for side in (0, 1):
                if s.has_neighbour(side):
                    s.pack(side)
        self.comm.exchange(self.slabs)
        for s in self.slabs:
            for side in (0, 1):
                if s.has_neighbour(side):
                    s.unpack(side)
                    s.halo_bytes += s.recv[side][0].numel() + s.recv[side][1].numel()
            s.sweep()

    def tick(self, n: int = 1, phases: int = PH_ALL):
        for _ in range(n):
            self.exchange()
            for s in self.slabs:
                s.rules_part(1, phases)
            maps = self._gather(lambda s: s.gas_map)
            for s, g in zip(self.slabs, maps):
                s._call(s.lib.sapio_slab_gas_begin, ctypes.c_void_p(g.data_ptr()), s.n, s.rank)
                s._keep = g
            for s in self.slabs:
                s.rules_part(2, phases)
            ref = self._gather(lambda s: s.refund)
            for s, g, r in zip(self.slabs, maps, ref):
                with torch.cuda.device(s.world.device):
                    _lib.check(s.lib.sapio_slab_gas_finish(s.engine._ref(), ctypes.c_void_p(g.data_ptr()), s.n, ctypes.c_void_p(r.data_ptr()), s.engine._stream_ptr()),
                               "sapio_slab_gas_finish")
                s._keep2 = r
            for s in self.slabs:
                if phases & 4:
                    s.engine.train()
                s.engine.settle()
            dp = self._gather(lambda s: s.disp)
            for s, d in zip(self.slabs, dp):
                with torch.cuda.device(s.world.device):
                    _lib.check(s.lib.sapio_slab_disperse_finish(s.engine._ref(), ctypes.c_void_p(d.data_ptr()), s.n, s.engine._stream_ptr()), "sapio_slab_disperse_finish")
                s._keep3 = d
            for s in self.slabs:
                if phases & 8:
                    s.engine.reproduce()
                s.engine.space_step()
                s.engine.friction()
                s._call(s.lib.sapio_post_advance)
            pops = self._gather(lambda s: s.pop)
            for s, c in zip(self.slabs, pops):
                with torch.cuda.device(s.world.device):
                    _lib.check(s.lib.sapio_slab_population(s.engine._ref(), ctypes.c_void_p(c.data_ptr()), s.n, s.engine._stream_ptr()), "sapio_slab_population")
                s._keep4 = c
            self.ticks += 1

    def synchronize(self):
        for s in self.slabs:
            s.engine.synchronize()


# ---- building slabs --------------------------------------------------------------------------------------------------------------
def slab_cuts(width: float, n_slabs: int, cuts=None):
    """The n_slabs + 1 boundaries: uniform unless the inner cuts are given."""
    inner = [width * r / n_slabs for r in range(1, n_slabs)] if cuts is None else [float(c) for c in cuts]
    assert len(inner) == n_slabs - 1 and all(a < b for a, b in zip([0.0] + inner, inner + [width]))
    return [0.0] + inner + [float(width)]


def slab_params(base: SapioParams, rank: int, n_slabs: int, n_org: int, n_dead: int, halo: float = HALO, x_cap: int = 0, cuts=None) -> SapioParams:
    """Parameters of slab `rank` of the world `base` describes (same width / height: global coordinates)."""
    p = SapioParams.from_buffer_copy(bytes(base))
    w = float(base.width)
    b = slab_cuts(w, n_slabs, cuts)
    p.slab_mode = 1
    p.slab_x0 = -1e30 if rank == 0 else b[rank]                                # the edge slabs own what lies beyond the walls too (it dies there)
    p.slab_x1 = 1e30 if rank == n_slabs - 1 else b[rank + 1]
    lo = max(0.0, b[rank] - halo - 256.0)
    hi = min(w, b[rank + 1] + halo + 256.0)
    p.grid_x0 = lo
    p.grid_nx = int(math.ceil((hi - lo) / p.grid_cell)) + 2
    p.n_org, p.n_dead = n_org, n_dead
    p.x_cap = x_cap if x_cap > 0 else max(4096, 4 * n_org)
    return p


_PER_ORG = {"ORG": 1, "CELL": MAXC, "PAIR": MAXC * (MAXC - 1) // 2, "ICON": 192}


def split_world(whole: World, n_slabs: int, device, halo: float = HALO, slot_factor: float = 1.6, cuts=None):
    """Cuts a whole (host) world into slabs: every organism / dead cell goes to the slab that owns it, in the whole world's order.
    Returns (slabs, permuted whole world): the second is the whole world with its organisms and dead cells renumbered slab after slab
    -- the order in which the slabs' gas maps compose -- so that the two can be compared tick by tick."""
    assert whole.device.type == "cpu"
    t = whole.t
    W = float(whole.p.width)
    org = np.nonzero(whole.np("org_state") == 1)[0]
    hx = whole.np("c_pos").reshape(-1, MAXC, 2)[org, 0, 0]
    inner = np.asarray(slab_cuts(W, n_slabs, cuts)[1:-1], dtype=np.float32)
    owner_o = np.searchsorted(inner, hx, side="right")                        # x in [b[r], b[r + 1]) -> r, as the kernels decide
    dead = np.nonzero(whole.np("d_state") > 0)[0]
    owner_d = np.searchsorted(inner, whole.np("d_pos")[dead, 0], side="right")
    order_o = np.concatenate([org[owner_o == r] for r in range(n_slabs)])
    order_d = np.concatenate([dead[owner_d == r] for r in range(n_slabs)])

    def place(dst: World, src_orgs, src_dead):
        for f in FIELDS:
            if f.scope in _PER_ORG:
                k = _PER_ORG[f.scope]
                s = t[f.name].reshape(whole.p.n_org, k, -1) if t[f.name].dim() > 1 or k > 1 else t[f.name].reshape(whole.p.n_org, 1, 1)
                d = dst.t[f.name].reshape(dst.p.n_org, k, -1) if dst.t[f.name].dim() > 1 or k > 1 else dst.t[f.name].reshape(dst.p.n_org, 1, 1)
                d[: len(src_orgs)] = s[torch.from_numpy(src_orgs)]
            elif f.scope == "DEAD":
                dst.t[f.name][: len(src_dead)] = t[f.name][torch.from_numpy(src_dead)]
            elif f.scope == "GLOB":
                dst.t[f.name].copy_(t[f.name])
        dst.t["g_x_n"].zero_(); dst.t["xa_off"].zero_()                     # contact caches are keyed by body id: both sides start without
        dst.t["x_jn"].zero_(); dst.t["x_jt"].zero_()

    perm = World(SapioParams.from_buffer_copy(bytes(whole.p)))
    place(perm, order_o, order_d)
    slabs = []
    for r in range(n_slabs):
        mo, md = org[owner_o == r], dead[owner_d == r]
        n_org = int(math.ceil((len(mo) * slot_factor + 64) / 64.0) * 64)
        n_dead = max(1024, int(len(md) * 2 + 512), int(whole.p.n_dead * 1.5 / n_slabs))     # the slab's share of the whole world's dead pool, with room for ghosts
        p = slab_params(whole.p, r, n_slabs, n_org, n_dead, halo, cuts=cuts)
        ws = World(p)
        place(ws, mo, md)
        ws.t["g_next_uid"][0] = int(t["g_next_uid"][0]) + (r << 40)           # births of different slabs never share a uid
        wd = ws.to(device)
        slabs.append(Slab(wd, Engine(wd), r, n_slabs, halo))
    return slabs, perm


def build_slab(n_total: int, rank: int, n_slabs: int, device, seed: int = 0, halo: float = HALO, slot_factor: float = 1.3,
               area_per_org: float = 3180.0 * 1800.0 / 12.0, food_per_org: float = 0.25) -> Slab:
    """Slab `rank` of the synthetic world of n_total organisms (the recipe of synth.build_world on the whole world's jittered grid;
    every slab draws the whole grid from the same seed and keeps its own part)."""
    side = math.sqrt(n_total * area_per_org)
    rng = np.random.default_rng(seed)
    cols = int(math.ceil(math.sqrt(n_total)))
    pitch = (side - 200.0) / cols
    idx = np.arange(n_total)
    pos = np.stack([100.0 + (idx % cols + 0.5) * pitch, 100.0 + (idx // cols + 0.5) * pitch], axis=1) + rng.uniform(-0.12, 0.12, (n_total, 2)) * pitch
    n_food = int(n_total * food_per_org)
    fpos = rng.uniform(60.0, side - 60.0, (n_food, 2)).astype(np.float32)
    x0, x1 = side * rank / n_slabs, side * (rank + 1) / n_slabs
    mine = (pos[:, 0] >= x0) & (pos[:, 0] < x1) if 0 < rank < n_slabs - 1 else ((pos[:, 0] < x1) if rank == 0 else (pos[:, 0] >= x0))
    fmine = (fpos[:, 0] >= x0) & (fpos[:, 0] < x1) if 0 < rank < n_slabs - 1 else ((fpos[:, 0] < x1) if rank == 0 else (fpos[:, 0] >= x0))
    n_mine = int(mine.sum())
    ghosts = int(2 * halo * side / area_per_org * 1.5) + 256                   # room for both halos
    n_org = int(math.ceil(((n_mine * slot_factor) + ghosts) / 64.0) * 64)
    n_dead = max(4096, int(n_mine * 24))
    base = default_params(n_org=n_org, n_dead=n_dead, width=side, height=side, seed=seed, population_limit=int(n_total * 1.2), x_cap=max(4096, 4 * n_org))
    p = slab_params(base, rank, n_slabs, n_org, n_dead, halo, x_cap=max(4096, 4 * n_org))
    w = World(p, device=device)
    eng = Engine(w)
    w.t["g_co2"][0] = 30000.0 * (2.0 * side) / (3180.0 + 1800.0)
    w.t["g_next_uid"][0] = 1 + (rank << 40)
    eng.spawn(np.arange(n_mine, dtype=np.int32), pos[mine])
    t = w.t
    nf = int(fmine.sum())
    if nf:
        t["d_state"][:nf] = 1
        t["d_pos"][:nf] = torch.from_numpy(fpos[fmine]).to(device)
        t["d_size"][:nf] = 30.0; t["d_mass"][:nf] = 12.0; t["d_mass0"][:nf] = 12.0
        t["d_elast"][:nf] = 0.5; t["d_fric"][:nf] = 0.5; t["d_owner"][:nf] = -1
    o = torch.arange(n_mine, device=device, dtype=torch.int32)
    t["org_think_tick"][:n_mine] = -(o % p.think_ticks); t["org_dopa_tick"][:n_mine] = -(o % p.think_ticks) - 1
    t["org_train_tick"][:n_mine] = -((o + 7) % p.train_ticks); t["org_lastbirth_tick"][:n_mine] = -(o % p.repro_ticks)
    eng.post()
    records = int(2 * halo * side / area_per_org * 2.0) + 128                  # organisms within +-halo of one boundary, with room
    return Slab(w, eng, rank, n_slabs, halo, record_budget=min(records, 4096), dead_budget=max(16384, records * 24))

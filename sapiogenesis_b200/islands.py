"""Population islands: one world per GPU, one process per GPU, organisms migrating between neighbouring islands.

The reference runs one world in one thread. Its world step has no cross-organism reduction beyond the two gas scalars, so the
multi-GPU form of the path is a set of independent worlds ("islands"): the tick needs no collective at
all, and the only exchange step is migration -- every `every` ticks each island sends `count` whole organisms (records of
csrc/migrate.cuh: genome, cells, brain, optimizer state, memories) to the next island on a ring and receives as many from
the previous one. Transport is torch.distributed point-to-point (NCCL over NVLink between GPUs; gloo in the CPU tests of
the host logic). Everything is deterministic in (seed, rank, tick).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def choose_emigrants(org_state: torch.Tensor, count: int, seed: int, rank: int, tick: int) -> torch.Tensor:
    """`count` living organism slots (fewer if the island is smaller), a seeded sample without replacement, ascending."""
    living = torch.nonzero(org_state == 1).view(-1)
    if living.numel() <= count:
        return living.to(torch.int32)
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + int(rank)) * 1000003 + int(tick))
    pick = torch.randperm(int(living.numel()), generator=g)[:count].sort().values
    return living[pick.to(living.device)].to(torch.int32)


def free_slots(org_state: torch.Tensor, count: int) -> torch.Tensor:
    """The `count` lowest free organism slots (the rule births use: j-th arrival -> j-th smallest free slot)."""
    return torch.nonzero(org_state == 0).view(-1)[:count].to(torch.int32)


def ring_exchange(send: torch.Tensor, n_send: int, group=None) -> tuple[torch.Tensor, int]:
    """Sends `send` (fixed-size buffer, first n_send records valid) to rank + 1 and receives the same shape from rank - 1.
    Returns (received buffer, number of valid records). World size 1: the island keeps its emigrants."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return send, n_send
    rank = dist.get_rank(group)
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    recv = torch.empty_like(send)
    cnt_s = torch.tensor([n_send], dtype=torch.int64, device=send.device)
    cnt_r = torch.zeros_like(cnt_s)
    ops = [dist.P2POp(dist.isend, cnt_s, nxt, group), dist.P2POp(dist.irecv, cnt_r, prv, group),
           dist.P2POp(dist.isend, send, nxt, group), dist.P2POp(dist.irecv, recv, prv, group)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    return recv, int(cnt_r.item())


class Island:
    """A world + engine that takes part in ring migration. `step(n)` == n ticks with migration every `every` ticks."""

    def __init__(self, world, engine, every: int = 50, count: int = 64, seed: int = 0, group=None):
        self.world, self.engine, self.every, self.count, self.seed, self.group = world, engine, int(every), int(count), int(seed), group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.size = dist.get_world_size(group) if dist.is_initialized() else 1
        self._tick = int(world.tick)                       # host mirror of g_tick: no device read in the loop
        engine.exclusive = True                            # bodies only change through the engine here: the step's grid serves the next think
        self.migrated_out = self.migrated_in = self.dropped = 0
        self._send = torch.zeros((self.count, engine.record_bytes()), dtype=torch.uint8, device=world.device) if self.count else None

    def migrate(self):
        """One exchange: pack emigrants -> free their slots -> ring send/recv -> unpack arrivals into the lowest free slots.
        Arrivals that find no free slot are dropped (counted)."""
        eng, t = self.engine, self.world.t
        out = choose_emigrants(t["org_state"], self.count, self.seed, self.rank, self._tick)
        n_out = int(out.numel())
        if n_out:
            eng.pack_organisms(out, self._send)
            eng.release_organisms(out)
        recv, n_in = ring_exchange(self._send, n_out, self.group)
        slots = free_slots(t["org_state"], n_in)
        n_fit = int(slots.numel())
        if n_fit:
            eng.unpack_organisms(slots, recv[:n_fit].contiguous() if n_fit < n_in else recv)
        self.migrated_out += n_out; self.migrated_in += n_fit; self.dropped += n_in - n_fit
        return n_out, n_fit

    def step(self, n: int = 1):
        done = 0
        while done < n:
            run = n - done
            if self.every > 0 and self.count > 0:
                run = min(run, self.every - self._tick % self.every)
            self.engine.tick(run)
            self._tick += run; done += run
            if self.every > 0 and self.count > 0 and self._tick % self.every == 0:
                self.migrate()

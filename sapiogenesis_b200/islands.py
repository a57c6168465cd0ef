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
    """A world + engine that takes part in ring migration. `step(n)` == n ticks with migration every `every` ticks.

    With an engine that has the device-side exchange (Engine.migrate_out / migrate_in) nothing in the loop waits for the device: the
    emigrants are chosen and packed (ragged records, include/sapio_migrate.h) by kernels, the fixed-size buffer goes to the next
    island (NCCL p2p), the arrivals are placed by the reference's rule for a new organism and scattered into free slots by kernels.
    Counts are read when somebody asks (`counts()`). Engines without it (the CPU stub of the gloo test) use the host-side form."""

    def __init__(self, world, engine, every: int = 50, count: int = 64, seed: int = 0, group=None, bytes_per_organism: int = 65536):
        self.world, self.engine, self.every, self.count, self.seed, self.group = world, engine, int(every), int(count), int(seed), group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.size = dist.get_world_size(group) if dist.is_initialized() else 1
        self._tick = int(world.tick)                       # host mirror of g_tick: no device read in the loop
        engine.exclusive = True                            # bodies only change through the engine here: the step's grid serves the next think
        self.migrated_out = self.migrated_in = self.dropped = 0
        self.migrations = 0
        self.device_side = hasattr(engine, "migrate_out")
        self._events = []                                  # (start, end) CUDA events around every exchange
        if self.device_side and self.count:
            self._send = engine.migrate_buffer(self.count, bytes_per_organism)
            self._recv = torch.zeros_like(self._send)
            self.last_bytes = int(self._send.numel())
        else:
            self._send = torch.zeros((self.count, engine.record_bytes()), dtype=torch.uint8, device=world.device) if self.count else None
            self.last_bytes = int(self._send.numel()) if self._send is not None else 0

    def _exchange(self, send: torch.Tensor, recv: torch.Tensor) -> torch.Tensor:
        """Fixed-size buffers around the ring; world size 1: the island keeps its emigrants (they are placed anew)."""
        if self.size == 1:
            return send
        nxt, prv = (self.rank + 1) % self.size, (self.rank - 1) % self.size
        ops = [dist.P2POp(dist.isend, send, nxt, self.group), dist.P2POp(dist.irecv, recv, prv, self.group)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()                                       # NCCL: enqueues a stream wait, the host does not block
        return recv

    def migrate(self):
        """One exchange. Device-side form: choose + pack + release -> ring send/recv -> place + unpack, all stream-ordered."""
        self.migrations += 1
        if not self.device_side:
            return self._migrate_host()
        eng = self.engine
        timed = self.world.device.type == "cuda"
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        eng.migrate_out(self.count, self.rank, self._send)
        got = self._exchange(self._send, self._recv)
        eng.migrate_in(got)
        if timed:
            e1.record(); self._events.append((e0, e1))
        return None

    def collect_migration_ms(self) -> float:
        """Device time of all exchanges so far (synchronises on the last one)."""
        return sum(a.elapsed_time(b) for a, b in self._events if (b.synchronize() or True))

    def counts(self) -> dict:
        if self.device_side:
            c = self.engine.migrate_counts()
            self.migrated_out, self.migrated_in, self.dropped = c["total_out"], c["total_in"], c["total_dropped"]
            return c
        return dict(total_out=self.migrated_out, total_in=self.migrated_in, total_dropped=self.dropped)

    def _migrate_host(self):
        """Host-side form (whole-slot records, arrivals keep their coordinates): engines without the device-side exchange."""
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

"""Engine: drives a `World` on one CUDA device through libsapio_b200.so.

This is the object behind `physics.Environment.update` (physics.py:425-442): one call == one world tick
(rules -> Space.step -> friction -> UI bookkeeping), or any single stage of it. PyTorch supplies the device
buffers, the stream and the workspace; every kernel is ours.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .world import PH_ALL, PH_REUSE_GRID, PH_STEP, PH_THINK, World


class Engine:
    def __init__(self, world: World, stream: torch.cuda.Stream | None = None):
        if world.device.type != "cuda":
            raise _lib.SapioError("Engine needs a CUDA world: this package has no CPU path "
                                  "(use world.to('cuda'))")
        self.lib = _lib.load()
        if self.lib.sapio_device_count() <= 0:
            raise _lib.SapioError("no CUDA device visible")
        self.world = world
        self.stream = stream
        nbytes = int(self.lib.sapio_workspace_bytes(ctypes.byref(world.p)))
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=world.device)
        self._ws = (ctypes.c_void_p(self.workspace.data_ptr()), ctypes.c_size_t(nbytes))
        self._struct = world.struct()
        # Opt-in: the owner promises that bodies are only added or moved through this engine (tick / spawn / unpack), never by
        # writing the world's tensors directly. The grid built by a tick's Space.step then serves the next tick's sensors.
        self.exclusive = False
        self._grid_valid = False

    def _stream_ptr(self):
        s = self.stream if self.stream is not None else torch.cuda.current_stream(self.world.device)
        return ctypes.c_void_p(s.cuda_stream)

    def _ref(self):
        return ctypes.byref(self._struct)

    def refresh(self):
        """Call after replacing a tensor of the world (pointers are cached)."""
        self._struct = self.world.struct()

    # -- whole tick ------------------------------------------------------------------------------
    def tick(self, n: int = 1, phases: int = PH_ALL):
        ph = phases | (PH_REUSE_GRID if (self.exclusive and self._grid_valid and (phases & PH_THINK)) else 0)
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_tick(self._ref(), *self._ws, ph, n, self._stream_ptr()), "sapio_tick")
        self._grid_valid = bool(phases & PH_STEP) if n > 0 else self._grid_valid

    # -- stages ----------------------------------------------------------------------------------
    def space_step(self):
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_space_step(self._ref(), *self._ws, self._stream_ptr()), "sapio_space_step")
        self._grid_valid = True

    def friction(self):
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_friction(self._ref(), self._stream_ptr()), "sapio_friction")

    def post(self):
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_post(self._ref(), *self._ws, self._stream_ptr()), "sapio_post")

    def rules(self, phases: int = PH_ALL):
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_rules(self._ref(), *self._ws, phases, self._stream_ptr()), "sapio_rules")

    def train(self):
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_train(self._ref(), *self._ws, self._stream_ptr()), "sapio_train")

    def settle(self):
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_settle(self._ref(), *self._ws, self._stream_ptr()), "sapio_settle")
        self._grid_valid = False

    def reproduce(self):
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_reproduce(self._ref(), *self._ws, self._stream_ptr()), "sapio_reproduce")
        self._grid_valid = False

    def spawn(self, slots, positions):
        """add_creature_clicked (Sapiogenesis.py:297-323) for a batch: random genomes (DNA.randomize with the spawn ranges of
        the params) materialised into free organism `slots` at `positions` [n, 2]."""
        dev = self.world.device
        slots_t = torch.as_tensor(slots, dtype=torch.int32, device=dev).contiguous()
        pos_t = torch.as_tensor(positions, dtype=torch.float64, device=dev).contiguous().view(-1)
        n = int(slots_t.numel())
        assert pos_t.numel() == 2 * n
        with torch.cuda.device(dev):
            _lib.check(self.lib.sapio_spawn(self._ref(), *self._ws, ctypes.c_void_p(slots_t.data_ptr()), ctypes.c_void_p(pos_t.data_ptr()), n,
                                            self._stream_ptr()), "sapio_spawn")
        self._keep = (slots_t, pos_t)
        self._grid_valid = False

    def _org_list(self, orgs):
        return torch.as_tensor(orgs, dtype=torch.int32, device=self.world.device).contiguous().view(-1)

    def activate(self, orgs):
        """neural.activate (neural.py:91-366) for the given organism slots, now."""
        lst = self._org_list(orgs)
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_activate(self._ref(), *self._ws, ctypes.c_void_p(lst.data_ptr()), int(lst.numel()),
                                               self._stream_ptr()), "sapio_activate")
        self._keep = (lst,)

    def train_network(self, orgs):
        """neural.train_network (neural.py:369-423) for the given organism slots, now."""
        lst = self._org_list(orgs)
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_train_network(self._ref(), *self._ws, ctypes.c_void_p(lst.data_ptr()), int(lst.numel()),
                                                    self._stream_ptr()), "sapio_train_network")
        self._keep = (lst,)

    # -- island migration: whole-organism records (csrc/migrate.cuh) ---------------------------------------------------
    def record_bytes(self) -> int:
        return int(self.lib.sapio_organism_record_bytes(ctypes.byref(self.world.p)))

    def pack_organisms(self, slots, out: torch.Tensor | None = None) -> torch.Tensor:
        """Gathers the organisms in `slots` into a [n, record_bytes] uint8 tensor on the device."""
        lst = self._org_list(slots)
        n, rb = int(lst.numel()), self.record_bytes()
        buf = out if out is not None else torch.empty((n, rb), dtype=torch.uint8, device=self.world.device)
        assert buf.is_contiguous() and buf.numel() >= n * rb and buf.is_cuda
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_pack_organisms(self._ref(), ctypes.c_void_p(lst.data_ptr()), n, ctypes.c_void_p(buf.data_ptr()),
                                                     buf.numel(), self._stream_ptr()), "sapio_pack_organisms")
        self._keep = (lst, buf)
        return buf

    def unpack_organisms(self, slots, records: torch.Tensor):
        """Scatters records into the free `slots`; the arrivals get fresh uids from this world."""
        lst = self._org_list(slots)
        n = int(lst.numel())
        assert records.is_contiguous() and records.is_cuda and records.numel() >= n * self.record_bytes()
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_unpack_organisms(self._ref(), ctypes.c_void_p(lst.data_ptr()), n, ctypes.c_void_p(records.data_ptr()),
                                                       records.numel(), self._stream_ptr()), "sapio_unpack_organisms")
        self._keep = (lst, records)
        self._grid_valid = False

    def release_organisms(self, slots):
        lst = self._org_list(slots)
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_release_organisms(self._ref(), ctypes.c_void_p(lst.data_ptr()), int(lst.numel()), self._stream_ptr()),
                       "sapio_release_organisms")
        self._keep = (lst,)

    # -- island migration on the device: ragged records, placement of the arrivals, no host round trip (csrc/migrate2.cuh) -----------
    def migrate_buffer(self, count: int, bytes_per_organism: int = 65536) -> torch.Tensor:
        n = int(self.lib.sapio_migrate_buffer_bytes(ctypes.byref(self.world.p), int(count), int(bytes_per_organism)))
        return torch.zeros(n, dtype=torch.uint8, device=self.world.device)

    def migrate_out(self, count: int, rank: int, buf: torch.Tensor):
        """Picks up to `count` living organisms, packs them into `buf` (as many as fit) and frees their slots."""
        assert buf.is_cuda and buf.is_contiguous() and buf.dtype == torch.uint8
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_migrate_out(self._ref(), *self._ws, int(count), int(rank), ctypes.c_void_p(buf.data_ptr()), buf.numel(),
                                                  self._stream_ptr()), "sapio_migrate_out")
        self._grid_valid = False

    def migrate_in(self, buf: torch.Tensor):
        """Places every organism of `buf` by the reference's rule for a new organism and scatters it into a free slot."""
        assert buf.is_cuda and buf.is_contiguous() and buf.dtype == torch.uint8
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_migrate_in(self._ref(), *self._ws, ctypes.c_void_p(buf.data_ptr()), buf.numel(), self._stream_ptr()),
                       "sapio_migrate_in")
        self._grid_valid = False

    def migrate_counts(self) -> dict:
        """(synchronises) last calls: chosen / packed / accepted / dropped; totals: out / in / dropped / stayed."""
        a = (ctypes.c_int32 * 8)()
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_migrate_counts(self._ref(), *self._ws, a, self._stream_ptr()), "sapio_migrate_counts")
        return dict(zip(("chosen", "packed", "accepted", "dropped", "total_out", "total_in", "total_dropped", "total_stayed"), [int(v) for v in a]))

    def sensor_view(self, org: int, cell: int, cap: int = 4096):
        """sprite.info["view"] of one eye / olfactory cell as ascending body ids (parity probe)."""
        dev = self.world.device
        out = torch.zeros(cap, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.sapio_sensor_view(self._ref(), *self._ws, org, cell, ctypes.c_void_p(out.data_ptr()), cap,
                                                  ctypes.c_void_p(cnt.data_ptr()), self._stream_ptr()), "sapio_sensor_view")
        n = int(cnt.item())
        return out[:min(n, cap)].cpu().numpy().astype("uint32")

    def events(self, drought: float = 0.0, poison: float = 0.0, random_cells: float = 0.0):
        """updateUI's world events (Sapiogenesis.py:37-53): drought / poison amounts and the random-food probability, each = severity * uDiff."""
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_events(self._ref(), *self._ws, drought, poison, random_cells, self._stream_ptr()), "sapio_events")
        if random_cells:
            self._grid_valid = False

    # -- end-to-end path: host buffers in, host buffers out (what a GUI host does every frame) ------------------
    def _new_host_buffers(self):
        from .world import MAXC
        p = self.world.p
        return dict(control=torch.zeros(4, dtype=torch.float32).pin_memory(),
                    header=torch.zeros(24, dtype=torch.int64).pin_memory(),                 # SapioFrameHeader = 168 bytes
                    cells=torch.zeros(p.n_org * MAXC, 4, dtype=torch.float32).pin_memory(),
                    dead=torch.zeros(p.n_dead, 4, dtype=torch.float32).pin_memory())

    def _host_buffers(self, k: int = 0):
        if getattr(self, "_hbs", None) is None:
            self._hbs = {}
        if k not in self._hbs:
            self._hbs[k] = self._new_host_buffers()
        return self._hbs[k]

    @staticmethod
    def _parse_frame(hb):
        raw = hb["header"].numpy()
        hdr = dict(co2=float(raw[:1].view("float64")[0]), o2=float(raw[1:2].view("float64")[0]), tick=int(raw[2]),
                   population=int(raw[3:4].view("int32")[0]), n_cells=int(raw[3:4].view("int32")[1]), n_dead=int(raw[4:5].view("int32")[0]),
                   stats=[int(v) for v in raw[5:21]])
        return hdr, hb["cells"][:hdr["n_cells"]], hb["dead"][:hdr["n_dead"]], (16, 168 + 16 * (hdr["n_cells"] + hdr["n_dead"]))

    # pipelined form: begin(k + 1) may be called before end(k); frames come back in order, each in its own pinned buffers
    def tick_host_begin(self, n: int = 1, phases: int = PH_ALL, drought: float = 0.0, poison: float = 0.0, random_cells: float = 0.0):
        self._seq = getattr(self, "_seq", 0)
        self._inflight = getattr(self, "_inflight", [])
        hb = self._host_buffers(self._seq & 1)
        self._seq += 1
        hb["control"][0], hb["control"][1], hb["control"][2] = drought, poison, random_cells
        ph = phases | (PH_REUSE_GRID if (self.exclusive and self._grid_valid and (phases & PH_THINK)) else 0)
        self._grid_valid = (bool(phases & PH_STEP) and not random_cells) if n > 0 else self._grid_valid
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_tick_host_begin(
                self._ref(), *self._ws, ph, n, ctypes.c_void_p(hb["control"].data_ptr()), ctypes.c_void_p(hb["header"].data_ptr()),
                ctypes.c_void_p(hb["cells"].data_ptr()), hb["cells"].shape[0], ctypes.c_void_p(hb["dead"].data_ptr()), hb["dead"].shape[0],
                self._stream_ptr()), "sapio_tick_host_begin")
        self._inflight.append(hb)

    def tick_host_end(self):
        """The oldest frame in flight: (header dict, cells [n,4], dead [m,4], (h2d, d2h) bytes). Valid until two more begins."""
        hb = self._inflight.pop(0)
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_tick_host_end(), "sapio_tick_host_end")
        return self._parse_frame(hb)

    def tick_host(self, n: int = 1, phases: int = PH_ALL, drought: float = 0.0, poison: float = 0.0, random_cells: float = 0.0):
        """n ticks through sapio_tick_host: uploads the event block from pinned host memory, downloads the frame
        (counters + packed sprite poses) into pinned host memory. Returns (header dict, cells [n,4], dead [m,4], bytes)."""
        hb = self._host_buffers()
        hb["control"][0], hb["control"][1], hb["control"][2] = drought, poison, random_cells
        ph = phases | (PH_REUSE_GRID if (self.exclusive and self._grid_valid and (phases & PH_THINK)) else 0)
        self._grid_valid = (bool(phases & PH_STEP) and not random_cells) if n > 0 else self._grid_valid
        with torch.cuda.device(self.world.device):
            _lib.check(self.lib.sapio_tick_host(
                self._ref(), *self._ws, ph, n, ctypes.c_void_p(hb["control"].data_ptr()), ctypes.c_void_p(hb["header"].data_ptr()),
                ctypes.c_void_p(hb["cells"].data_ptr()), hb["cells"].shape[0], ctypes.c_void_p(hb["dead"].data_ptr()), hb["dead"].shape[0],
                self._stream_ptr()), "sapio_tick_host")
        return self._parse_frame(hb)

    # -- measurement hooks ---------------------------------------------------------------------------------------
    def profile(self, enable: bool):
        self.lib.sapio_profile(1 if enable else 0)

    def profile_read(self, reset: bool = True):
        n = self.lib.sapio_profile_stages()
        ms = (ctypes.c_double * n)(); cnt = (ctypes.c_int64 * n)()
        self.lib.sapio_profile_read(ms, cnt, 1 if reset else 0)
        return {self.lib.sapio_profile_name(i).decode(): (ms[i], cnt[i]) for i in range(n)}

    def solver_info(self):
        """Per size class of the organism solver: dict(lanes, orgs_per_warp, warps_per_cta, smem_per_warp, resident_ctas, sms)."""
        out = []
        for cls in range(8):
            a = (ctypes.c_int32 * 6)()
            _lib.check(self.lib.sapio_solver_info(cls, a), "sapio_solver_info")
            out.append(dict(zip(("lanes", "orgs_per_warp", "warps_per_cta", "smem_per_warp", "resident_ctas", "sms"), [int(v) for v in a])))
        return out

    def launch_count(self) -> int:
        return int(self.lib.sapio_launch_count())

    def set_coupled_pack_from(self, n: int) -> int:
        """Threshold (coupled organisms of the previous tick) above which the lane-packed coupled kernel runs; returns the old one."""
        return int(self.lib.sapio_set_coupled_pack_from(int(n)))

    def synchronize(self):
        torch.cuda.synchronize(self.world.device)

"""World state: the structure-of-arrays tensors behind `physics.Environment`.

The array list is parsed from include/sapio_fields.def, the same file that builds the C struct
`SapioWorld` (include/sapio.h) used by the CUDA library and by the CPU oracle. PyTorch owns every
buffer; the C side only ever sees raw pointers.

Replaces the object graph of the reference: `Environment.sprites` / `Environment.info`
(physics.py:285-311), `Organism` (sprites.py:1089-1125), `DNA` (sprites.py:329-408) and the pymunk
bodies/shapes/joints they hold.
"""
from __future__ import annotations

import ctypes
import math
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS_DEF = os.path.join(_ROOT, "include", "sapio_fields.def")

MAXC = 64
NPAIR = MAXC * (MAXC - 1) // 2
ICAP = 192
MAXL = 12
NWALL = 4
HIST = 10
STM = 30
MEMROWS = 232
RNN_MAXH = 16
HM = 90

# cell types: sprites.py:22 (+ "heart", sprites.py:573)
CELL_TYPES = ["barrier", "carniv", "eye", "olfactory", "co2C", "push", "body", "rotate", "heart"]
T_BARRIER, T_CARNIV, T_EYE, T_OLFACTORY, T_CO2C, T_PUSH, T_BODY, T_ROTATE, T_HEART = range(9)

F_IMMORTAL, F_MALADAPTIVE, F_FINITE_MEMORY, F_MIRROR_X, F_MIRROR_Y, F_GHOST = 1, 2, 4, 8, 16, 32

PH_RULES, PH_THINK, PH_TRAIN, PH_REPRODUCE, PH_STEP, PH_FRICTION, PH_POST = 1, 2, 4, 8, 16, 32, 64
PH_ALL = 127
PH_REUSE_GRID = 128

STAT_NAMES = ["births", "deaths", "cell_deaths", "dead_removed", "thinks", "trains", "xcontacts",
              "icontacts", "birth_fail_place", "birth_fail_caps", "overflow", "org_ticks", "levels",
              "coupled", "rare_paths", "full_scans"]

# per-type tables: sprites.base_cell_info (sprites.py:67-121); index = T_*
BASE_HEALTH = [30, 30, 8, 24, 15, 24, 24, 24, 15]
BASE_USAGE_IDLE = [1.2, 2.5, 2.0, 3.5, 3.6, 3.0, 1.2, 2.5, 2.5]
BASE_USAGE_USE = [1.2, 4.0, 2.0, 5.0, 3.6, 6.0, 1.2, 5.0, 2.5]
BASE_STORAGE = [15, 30, 20, 30, 33, 30, 30, 30, 33]
BASE_DAMAGE = [0, 2, 0, 0, 0, 0, 0, 0, 0]


class SapioParams(ctypes.Structure):
    """Mirror of `struct SapioParams` (include/sapio.h). Defaults: SURVEY.md Appendix B."""
    _fields_ = [
        ("n_org", ctypes.c_int32), ("n_dead", ctypes.c_int32), ("x_cap", ctypes.c_int32),
        ("in_cap", ctypes.c_int32), ("brain_cap", ctypes.c_int32), ("adam_cap", ctypes.c_int32),
        ("width_cap", ctypes.c_int32), ("rnn_cap", ctypes.c_int32),
        ("width", ctypes.c_float), ("height", ctypes.c_float), ("frame_width", ctypes.c_float),
        ("wall_elast", ctypes.c_float), ("wall_fric", ctypes.c_float),
        ("dt_wall", ctypes.c_float), ("dt_phys", ctypes.c_float), ("friction_pct", ctypes.c_float),
        ("think_ticks", ctypes.c_int32), ("train_ticks", ctypes.c_int32),
        ("repro_ticks", ctypes.c_int32), ("age_ticks", ctypes.c_int32),
        ("co2_ratio", ctypes.c_float), ("dispersion_rate", ctypes.c_float), ("heal_rate", ctypes.c_float),
        ("mass_cutoff", ctypes.c_float), ("break_damage", ctypes.c_float), ("starvation_rate", ctypes.c_float),
        ("max_push", ctypes.c_float), ("max_rot", ctypes.c_float),
        ("eyesight_mult", ctypes.c_float), ("olfactory_mult", ctypes.c_float),
        ("mutation_severity", ctypes.c_float), ("brain_mutation_severity", ctypes.c_float),
        ("learning_rate", ctypes.c_float),
        ("population_limit", ctypes.c_int32), ("offspring_amount", ctypes.c_int32),
        ("weight_persistence", ctypes.c_int32), ("memory_limit", ctypes.c_int32),
        ("iterations", ctypes.c_int32),
        ("collision_slop", ctypes.c_float), ("collision_bias", ctypes.c_float),
        ("joint_error_bias", ctypes.c_float),
        ("cell_lo", ctypes.c_int32), ("cell_hi", ctypes.c_int32), ("size_lo", ctypes.c_int32),
        ("size_hi", ctypes.c_int32), ("mass_lo", ctypes.c_int32), ("mass_hi", ctypes.c_int32),
        ("hidden_lo", ctypes.c_int32), ("hidden_hi", ctypes.c_int32),
        ("mirror_px", ctypes.c_float), ("mirror_py", ctypes.c_float),
        ("seed", ctypes.c_uint64),
        ("grid_cell", ctypes.c_float), ("grid_nx", ctypes.c_int32), ("grid_ny", ctypes.c_int32),
        ("slab_mode", ctypes.c_int32), ("grid_x0", ctypes.c_float), ("slab_x0", ctypes.c_float), ("slab_x1", ctypes.c_float),
        ("rnn_hidden", ctypes.c_int32),
    ]


def default_params(n_org: int = 16, n_dead: int = 256, width: float = 3180.0, height: float = 1800.0,
                   in_cap: int = 128, width_cap: int = 128, brain_cap: int = 32768, x_cap: int = 0,
                   seed: int = 0, population_limit: int = 12, rnn_hidden: int = 0) -> SapioParams:
    p = SapioParams()
    p.n_org, p.n_dead = n_org, n_dead
    p.x_cap = x_cap if x_cap > 0 else max(1024, 8 * n_org + 4 * n_dead)
    p.in_cap, p.width_cap, p.brain_cap = in_cap, width_cap, brain_cap
    # Environment.info["use_rnn"] / ["hidden_rnn_size"] (physics.py:305,308): 0 = feed-forward brains and no recurrent-state arrays
    assert 0 <= rnn_hidden <= RNN_MAXH
    p.rnn_cap = p.rnn_hidden = rnn_hidden
    # Adam sees only the input and output layers (networks.py:221): (in+1)*h0 + (h_last+1)*4
    p.adam_cap = (in_cap + 1) * width_cap + (width_cap + 1) * 4
    p.width, p.height = width, height               # environment_settings.json:1
    p.frame_width, p.wall_elast, p.wall_fric = 10.0, 0.95, 0.9  # physics.py:280,341-342
    p.dt_wall, p.dt_phys, p.friction_pct = 0.030, 0.06, 60.0    # environment_settings.json, physics.py:319,283
    p.think_ticks = p.train_ticks = math.ceil(0.5 / 0.030 - 1e-9)   # sprites.py:53-55
    p.repro_ticks = math.ceil(6.0 / 0.030 - 1e-9)                   # sprites.py:41
    p.age_ticks = math.ceil(200.0 / 0.030 - 1e-9)                   # sprites.py:62
    p.co2_ratio, p.dispersion_rate, p.heal_rate = 3.0, 6.0, 4.0     # sprites.py:37,39,45
    p.mass_cutoff, p.break_damage, p.starvation_rate = 1.0, 5.0, 6.0  # sprites.py:47,49,51
    p.max_push, p.max_rot = 250.0, 6.0                               # sprites.py:59-60
    p.eyesight_mult, p.olfactory_mult = 10.0, 15.0                   # sprites.py:64-65
    p.mutation_severity = p.brain_mutation_severity = 0.5            # physics.py:298-299
    p.learning_rate = 0.02                                           # physics.py:304
    p.population_limit, p.offspring_amount = population_limit, 1     # physics.py:302, sprites.py:43
    p.weight_persistence, p.memory_limit = 1, 200                    # physics.py:303, neural.py:33
    p.iterations = 10                                                # Chipmunk default
    p.collision_slop = 0.1
    p.collision_bias = float(np.float32((1.0 - 0.1) ** 60))
    p.joint_error_bias = float(np.float32((1.0 - 0.1) ** 60))
    p.cell_lo, p.cell_hi, p.size_lo, p.size_hi = 3, 25, 6, 42     # MainWindow.py:234-313 widget defaults
    p.mass_lo, p.mass_hi, p.hidden_lo, p.hidden_hi = 5, 20, 3, 12
    p.mirror_px = p.mirror_py = 0.4                               # Sapiogenesis.py:690-691
    p.seed = seed
    p.grid_cell = 64.0
    p.grid_nx = int(math.ceil(width / p.grid_cell)) + 2
    p.grid_ny = int(math.ceil(height / p.grid_cell)) + 2
    return p


_CT = {
    "uint8_t": (torch.uint8, ctypes.c_uint8), "int8_t": (torch.int8, ctypes.c_int8),
    "int16_t": (torch.int16, ctypes.c_int16), "uint16_t": (torch.int16, ctypes.c_uint16),
    "int32_t": (torch.int32, ctypes.c_int32), "uint32_t": (torch.int32, ctypes.c_uint32),
    "int64_t": (torch.int64, ctypes.c_int64),
    "uint64_t": (torch.int64, ctypes.c_uint64), "float": (torch.float32, ctypes.c_float),
    "double": (torch.float64, ctypes.c_double),
}
_NP_VIEW = {"uint16_t": np.uint16, "uint32_t": np.uint32, "uint64_t": np.uint64}


@dataclass
class FieldSpec:
    name: str
    ctype: str
    scope: str
    per: str


def parse_fields(path: str = FIELDS_DEF) -> List[FieldSpec]:
    out = []
    rx = re.compile(r"^SAPIO_FIELD\(\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*([^)]+?)\s*\)")
    with open(path) as f:
        for line in f:
            m = rx.match(line)
            if m:
                out.append(FieldSpec(*m.groups()))
    return out


FIELDS = parse_fields()


class SapioWorldStruct(ctypes.Structure):
    _fields_ = [("p", SapioParams)] + [(f.name, ctypes.c_void_p) for f in FIELDS]


def _rows(p: SapioParams, scope: str) -> int:
    return {"ORG": p.n_org, "CELL": p.n_org * MAXC, "PAIR": p.n_org * NPAIR, "ICON": p.n_org * ICAP,
            "DEAD": p.n_dead, "XCON": p.x_cap, "GLOB": 1, "BODY1": p.n_org * MAXC + p.n_dead + 1}[scope]


def _per(p: SapioParams, per: str) -> int:
    return int(eval(per, {"__builtins__": {}}, {"MAXC": MAXC, "IN": p.in_cap, "BRAIN": p.brain_cap,
                                                 "ADAM": p.adam_cap, "RNNH": p.rnn_cap}))


class World:
    """All state of one simulated world on one device."""

    def __init__(self, params: SapioParams, device: str | torch.device = "cpu"):
        self.p = params
        self.device = torch.device(device)
        self.t: Dict[str, torch.Tensor] = {}
        for f in FIELDS:
            rows, per = _rows(params, f.scope), _per(params, f.per)
            shape = (rows,) if per == 1 else (rows, per)
            self.t[f.name] = torch.zeros(shape, dtype=_CT[f.ctype][0], device=self.device)
        self._struct = None
        self.t["c_parent"].fill_(-1)
        self.t["c_mirror"].fill_(-1)
        self.t["d_owner"].fill_(-1)
        self.t["g_next_uid"].fill_(1)

    # -- access ---------------------------------------------------------------------------------
    def __getattr__(self, name):
        t = self.__dict__.get("t")
        if t is not None and name in t:
            return t[name]
        raise AttributeError(name)

    def np(self, name: str) -> np.ndarray:
        """Host numpy view (CPU worlds) or copy (CUDA worlds) with the true unsigned dtype."""
        f = next(f for f in FIELDS if f.name == name)
        a = self.t[name].detach().cpu().numpy()
        if f.ctype in _NP_VIEW:
            a = a.view(_NP_VIEW[f.ctype])
        return a

    def struct(self) -> SapioWorldStruct:
        s = SapioWorldStruct()
        ctypes.memmove(ctypes.byref(s.p), ctypes.byref(self.p), ctypes.sizeof(SapioParams))
        for f in FIELDS:
            setattr(s, f.name, self.t[f.name].data_ptr())
        self._struct = s  # keep alive
        return s

    def reserve_rnn(self, hidden: int) -> None:
        """Allocates the recurrent-state arrays (per = RNNH of sapio_fields.def) of a world that was built without them: what setting
        Environment.info["use_rnn"] needs the first time. The hidden size is then fixed for the life of the world."""
        if self.p.rnn_cap == hidden:
            return
        if self.p.rnn_cap != 0 or not 0 < hidden <= RNN_MAXH:
            raise ValueError(f"hidden_rnn_size {hidden}: between 1 and {RNN_MAXH}, and fixed once RNN brains exist ({self.p.rnn_cap} here)")
        self.p.rnn_cap = hidden
        for f in FIELDS:
            if "RNNH" in f.per:
                self.t[f.name] = torch.zeros((_rows(self.p, f.scope), _per(self.p, f.per)), dtype=_CT[f.ctype][0], device=self.device)
        self._struct = None

    def to(self, device) -> "World":
        w = World.__new__(World)
        w.p = SapioParams.from_buffer_copy(bytes(self.p))
        w.device = torch.device(device)
        w.t = {k: v.to(w.device, copy=True) for k, v in self.t.items()}
        w._struct = None
        return w

    def clone(self) -> "World":
        return self.to(self.device)

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.t.values())

    # -- convenience ----------------------------------------------------------------------------
    @property
    def tick(self) -> int:
        return int(self.t["g_tick"].item())

    def stats(self) -> Dict[str, int]:
        s = self.t["g_stats"].detach().cpu().view(-1).tolist()
        return dict(zip(STAT_NAMES, s))

    def population(self) -> int:
        return int((self.t["org_state"] == 1).sum().item())

    def living_cells(self) -> int:
        return int((self.t["c_alive"] == 1).sum().item())


def derived_stats(ctype: int, size: float, mass: float) -> Tuple[float, float, float, float, float]:
    """finish_cell_info (sprites.py:238-253): k*size + k*mass for every table."""
    s = size + mass
    return (BASE_HEALTH[ctype] * s, BASE_USAGE_IDLE[ctype] * s, BASE_USAGE_USE[ctype] * s,
            BASE_STORAGE[ctype] * s, BASE_DAMAGE[ctype] * s)

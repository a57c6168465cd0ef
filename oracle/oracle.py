"""ctypes front end of the CPU oracle (oracle/*.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
--impl reference legs -- never by the product package. See oracle/oracle.h for what is restated and how it is pinned.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsapio_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("sapio.h", "sapio_fields.def", "sapio_migrate.h")]
    stale = force or not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        vp, i32, u32, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_double
        for name in ("oracle_space_step", "oracle_friction", "oracle_train", "oracle_settle",
                     "oracle_reproduce", "oracle_post"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = i32
        L.oracle_rules.argtypes = [vp, u32]; L.oracle_rules.restype = i32
        L.oracle_tick.argtypes = [vp, u32, i32]; L.oracle_tick.restype = i32
        L.oracle_spawn.argtypes = [vp, i32, dbl, dbl]; L.oracle_spawn.restype = i32
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        L.oracle_spawn_arr.argtypes = [vp, i32, dbl, dbl, fp, i32, fp, i32, fp, i32, ip]; L.oracle_spawn_arr.restype = i32
        L.oracle_mutate_arr.argtypes = [vp, i32, i32, dbl, dbl, fp, i32, fp, i32, fp, i32, ip]; L.oracle_mutate_arr.restype = i32
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_dna_rect.argtypes = [vp, i32, dp, dp]; L.oracle_dna_rect.restype = i32
        L.oracle_viable.argtypes = [vp, i32, dbl, dbl]; L.oracle_viable.restype = i32
        L.oracle_org_rect.argtypes = [vp, i32, dp]; L.oracle_org_rect.restype = None
        L.oracle_stream_fill.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32, u32, u32, i32, fp]
        L.oracle_stream_fill.restype = None
        L.oracle_key1.argtypes = [dbl]; L.oracle_key1.restype = i32
        L.oracle_forward.argtypes = [fp, ctypes.POINTER(ctypes.c_int32), i32, fp, fp, fp]; L.oracle_forward.restype = None
        L.oracle_train_one.argtypes = [vp, i32]; L.oracle_train_one.restype = i32
        L.oracle_think_now.argtypes = [vp, i32]; L.oracle_think_now.restype = i32
        L.oracle_events.argtypes = [vp, dbl, dbl, dbl]; L.oracle_events.restype = i32
        L.oracle_set_think_uniforms.argtypes = [ctypes.c_void_p]; L.oracle_set_think_uniforms.restype = None
        L.oracle_rotation.argtypes = [vp, i32]; L.oracle_rotation.restype = dbl
        L.oracle_think.argtypes = [vp, i32, dp, dp, ctypes.POINTER(ctypes.c_uint8), dbl]; L.oracle_think.restype = i32
        L.oracle_inputs.argtypes = [vp, i32, dp, ctypes.POINTER(ctypes.c_uint8), dbl, dp]; L.oracle_inputs.restype = i32
        L.oracle_migrate_out.argtypes = [vp, i32, i32, ctypes.c_void_p, ctypes.c_size_t]; L.oracle_migrate_out.restype = i32
        L.oracle_migrate_in.argtypes = [vp, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int32)]; L.oracle_migrate_in.restype = i32
        L.oracle_sensor_view.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_uint32), i32]; L.oracle_sensor_view.restype = i32
        _lib = L
    return _lib


def _ref(world):
    assert world.device.type == "cpu", "the oracle runs on host memory"
    return ctypes.byref(world.struct())


def _fa(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size


def tick(world, phases: int = 127, n: int = 1) -> int:
    return lib().oracle_tick(_ref(world), phases, n)


def space_step(world) -> int: return lib().oracle_space_step(_ref(world))
def friction(world) -> int: return lib().oracle_friction(_ref(world))
def rules(world, phases: int = 127) -> int: return lib().oracle_rules(_ref(world), phases)
def train(world) -> int: return lib().oracle_train(_ref(world))
def settle(world) -> int: return lib().oracle_settle(_ref(world))
def reproduce(world) -> int: return lib().oracle_reproduce(_ref(world))
def post(world) -> int: return lib().oracle_post(_ref(world))
def events(world, drought=0.0, poison=0.0, random_cells=0.0) -> int: return lib().oracle_events(_ref(world), float(drought), float(poison), float(random_cells))
def spawn(world, slot: int, x: float, y: float) -> int: return lib().oracle_spawn(_ref(world), slot, x, y)


def spawn_arr(world, slot, x, y, u, ui, ub):
    (a, pa, na), (b, pb, nb), (c, pc, nc) = _fa(u), _fa(ui), _fa(ub)
    used = (ctypes.c_int * 4)()
    rc = lib().oracle_spawn_arr(_ref(world), slot, x, y, pa, na, pb, nb, pc, nc, used)
    return rc, list(used)


def mutate_arr(world, parent, slot, x, y, u, uf, ub):
    (a, pa, na), (b, pb, nb), (c, pc, nc) = _fa(u), _fa(uf), _fa(ub)
    used = (ctypes.c_int * 4)()
    rc = lib().oracle_mutate_arr(_ref(world), parent, slot, x, y, pa, na, pb, nb, pc, nc, used)
    return rc, list(used)


def dna_rect(world, org):
    r = (ctypes.c_double * 4)(); s = (ctypes.c_double * 2)()
    lib().oracle_dna_rect(_ref(world), org, r, s)
    return list(r), list(s)


def org_rect(world, org):
    r = (ctypes.c_double * 4)()
    lib().oracle_org_rect(_ref(world), org, r)
    return list(r)


def viable(world, genome_of, x, y) -> bool:
    return bool(lib().oracle_viable(_ref(world), genome_of, x, y))


def stream(seed, uid, tick_, purpose, n, start=0):
    out = np.zeros(n, dtype=np.float32)
    lib().oracle_stream_fill(seed, uid, tick_, purpose, start, n, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def key1(x: float) -> int: return lib().oracle_key1(float(x))


def forward(brain, ldim, x):
    brain = np.ascontiguousarray(brain, dtype=np.float32)
    ldim = np.ascontiguousarray(ldim, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    f = ctypes.POINTER(ctypes.c_float)
    lib().oracle_forward(brain.ctypes.data_as(f), ldim.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(ldim) - 1,
                         x.ctypes.data_as(f), out.ctypes.data_as(f), None)
    return out


def train_one(world, org) -> int: return lib().oracle_train_one(_ref(world), org)
def think_now(world, org, uniforms=None) -> int:
    """neural.activate for one organism; `uniforms` = (gate, 4 x randn) replaces the Philox draws (pinning tests)."""
    if uniforms is None:
        return lib().oracle_think_now(_ref(world), org)
    a = np.ascontiguousarray(uniforms, dtype=np.float32)
    lib().oracle_set_think_uniforms(a.ctypes.data)
    try:
        return lib().oracle_think_now(_ref(world), org)
    finally:
        lib().oracle_set_think_uniforms(None)
def rotation(world, org) -> float: return lib().oracle_rotation(_ref(world), org)


def inputs(world, org, rotation_deg=None):
    """Unrounded network inputs of organism `org` on the world's current state."""
    from sapiogenesis_b200.world import MAXC
    rows = slice(org * MAXC, (org + 1) * MAXC)
    energy = np.ascontiguousarray(world.np("c_energy")[rows], dtype=np.float64)
    alive = np.ascontiguousarray(world.np("c_alive")[rows], dtype=np.uint8)
    out = np.zeros(2048, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    n = lib().oracle_inputs(_ref(world), org, energy.ctypes.data_as(dp), alive.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                            float(rotation_deg or 0.0), out.ctypes.data_as(dp))
    return out[:n]


def sensor_view(world, org, k, cap=4096):
    hits = np.zeros(cap, dtype=np.uint32)
    n = lib().oracle_sensor_view(_ref(world), org, k, hits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), cap)
    return hits[:min(n, cap)]


def migrate_out(world, count: int, rank: int, buf) -> int:
    """buf: contiguous uint8 CPU tensor / numpy array (include/sapio_migrate.h). Returns the number of records written."""
    import numpy as _np
    a = buf.numpy() if hasattr(buf, "numpy") else _np.asarray(buf)
    return lib().oracle_migrate_out(_ref(world), int(count), int(rank), a.ctypes.data_as(ctypes.c_void_p), a.nbytes)


def migrate_in(world, buf):
    """Returns (accepted, dropped)."""
    import numpy as _np
    a = buf.numpy() if hasattr(buf, "numpy") else _np.asarray(buf)
    out = (ctypes.c_int32 * 2)()
    rc = lib().oracle_migrate_in(_ref(world), a.ctypes.data_as(ctypes.c_void_p), a.nbytes, out)
    assert rc == 0, rc
    return int(out[0]), int(out[1])

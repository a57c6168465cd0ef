"""ctypes binding of libsapio_b200.so -- the hand-written sm_100a CUDA path behind include/sapio.h.

There is no CPU implementation in this package: if the library is missing or no CUDA device is present,
every compute call raises. (The CPU oracle under oracle/ is test infrastructure and is never imported here.)
"""
from __future__ import annotations

import ctypes
import os

from .world import SapioParams, SapioWorldStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SAPIO_SO") or os.path.join(_HERE, "libsapio_b200.so")      # SAPIO_SO: an A/B build of the same sources (build.py, SAPIO_BUILD_TAG)
_lib = None


class SapioError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise SapioError(f"{SO_PATH} is not built: run `python -m sapiogenesis_b200.build` "
                         "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(SO_PATH)
    wp, vp, sz, u32, i32 = ctypes.POINTER(SapioWorldStruct), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int
    L.sapio_abi_version.restype = i32
    L.sapio_last_error.restype = ctypes.c_char_p
    L.sapio_device_count.restype = i32
    L.sapio_workspace_bytes.argtypes = [ctypes.POINTER(SapioParams)]; L.sapio_workspace_bytes.restype = sz
    L.sapio_tick.argtypes = [wp, vp, sz, u32, i32, vp]; L.sapio_tick.restype = i32
    L.sapio_space_step.argtypes = [wp, vp, sz, vp]; L.sapio_space_step.restype = i32
    L.sapio_friction.argtypes = [wp, vp]; L.sapio_friction.restype = i32
    L.sapio_post.argtypes = [wp, vp, sz, vp]; L.sapio_post.restype = i32
    for name in ("sapio_rules",):
        if hasattr(L, name):
            getattr(L, name).argtypes = [wp, vp, sz, u32, vp]; getattr(L, name).restype = i32
    for name in ("sapio_train", "sapio_settle", "sapio_reproduce"):
        if hasattr(L, name):
            getattr(L, name).argtypes = [wp, vp, sz, vp]; getattr(L, name).restype = i32
    L.sapio_spawn.argtypes = [wp, vp, sz, vp, vp, i32, vp]; L.sapio_spawn.restype = i32
    L.sapio_sensor_view.argtypes = [wp, vp, sz, i32, i32, vp, i32, vp, vp]; L.sapio_sensor_view.restype = i32
    for name in ("sapio_activate", "sapio_train_network"):
        getattr(L, name).argtypes = [wp, vp, sz, vp, i32, vp]; getattr(L, name).restype = i32
    L.sapio_organism_record_bytes.argtypes = [vp]; L.sapio_organism_record_bytes.restype = sz
    L.sapio_pack_organisms.argtypes = [wp, vp, i32, vp, sz, vp]; L.sapio_pack_organisms.restype = i32
    L.sapio_unpack_organisms.argtypes = [wp, vp, i32, vp, sz, vp]; L.sapio_unpack_organisms.restype = i32
    L.sapio_release_organisms.argtypes = [wp, vp, i32, vp]; L.sapio_release_organisms.restype = i32
    L.sapio_migrate_buffer_bytes.argtypes = [vp, i32, i32]; L.sapio_migrate_buffer_bytes.restype = sz
    L.sapio_migrate_out.argtypes = [wp, vp, sz, i32, i32, vp, sz, vp]; L.sapio_migrate_out.restype = i32
    L.sapio_migrate_in.argtypes = [wp, vp, sz, vp, sz, vp]; L.sapio_migrate_in.restype = i32
    L.sapio_migrate_counts.argtypes = [wp, vp, sz, ctypes.POINTER(ctypes.c_int32), vp]; L.sapio_migrate_counts.restype = i32
    f32_ = ctypes.c_float
    L.sapio_rules_part.argtypes = [wp, vp, sz, u32, i32, vp]; L.sapio_rules_part.restype = i32
    L.sapio_post_advance.argtypes = [wp, vp, sz, vp]; L.sapio_post_advance.restype = i32
    L.sapio_slab_offsets.argtypes = [wp, vp, sz, ctypes.POINTER(ctypes.c_int64)]; L.sapio_slab_offsets.restype = i32
    L.sapio_halo_pack.argtypes = [wp, vp, sz, f32_, f32_, vp, sz, vp, sz, vp]; L.sapio_halo_pack.restype = i32
    L.sapio_halo_unpack.argtypes = [wp, vp, sz, i32, vp, sz, vp, sz, vp]; L.sapio_halo_unpack.restype = i32
    L.sapio_halo_sweep.argtypes = [wp, vp, sz, vp]; L.sapio_halo_sweep.restype = i32
    L.sapio_slab_gas_begin.argtypes = [wp, vp, sz, vp, i32, i32, vp]; L.sapio_slab_gas_begin.restype = i32
    L.sapio_slab_gas_finish.argtypes = [wp, vp, i32, vp, vp]; L.sapio_slab_gas_finish.restype = i32
    L.sapio_slab_disperse_finish.argtypes = [wp, vp, i32, vp]; L.sapio_slab_disperse_finish.restype = i32
    L.sapio_slab_population.argtypes = [wp, vp, i32, vp]; L.sapio_slab_population.restype = i32
    f32 = ctypes.c_float
    L.sapio_events.argtypes = [wp, vp, sz, f32, f32, f32, vp]; L.sapio_events.restype = i32
    L.sapio_tick_host.argtypes = [wp, vp, sz, u32, i32, vp, vp, vp, i32, vp, i32, vp]; L.sapio_tick_host.restype = i32
    L.sapio_tick_host_begin.argtypes = [wp, vp, sz, u32, i32, vp, vp, vp, i32, vp, i32, vp]; L.sapio_tick_host_begin.restype = i32
    L.sapio_tick_host_end.argtypes = []; L.sapio_tick_host_end.restype = i32
    L.sapio_profile.argtypes = [i32]; L.sapio_profile.restype = i32
    L.sapio_profile_stages.restype = i32
    L.sapio_profile_name.argtypes = [i32]; L.sapio_profile_name.restype = ctypes.c_char_p
    L.sapio_profile_read.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), i32]; L.sapio_profile_read.restype = i32
    L.sapio_launch_count.restype = ctypes.c_uint64
    L.sapio_solver_info.argtypes = [i32, ctypes.POINTER(ctypes.c_int32)]; L.sapio_solver_info.restype = i32
    L.sapio_set_coupled_pack_from.restype = ctypes.c_int
    L.sapio_set_coupled_pack_from.argtypes = [ctypes.c_int]
    if L.sapio_abi_version() != 1:
        raise SapioError("libsapio_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sapio_last_error().decode(errors="replace")
        raise SapioError(f"{what} failed ({rc}): {msg}")

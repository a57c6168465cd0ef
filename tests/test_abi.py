"""The C-ABI library loads and exports every symbol include/sapio.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sapio.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:int|size_t|const char \*)\s*\*?\s*(sapio_\w+)\s*\(", src, flags=re.M)
    return sorted(set(n for n in names if not n.startswith("sapio_body_") and n != "sapio_pair_index"))


def test_library_exports_every_declared_symbol():
    from sapiogenesis_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    syms = declared_symbols()
    assert "sapio_tick" in syms and "sapio_space_step" in syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/sapio.h but not exported: {missing}"
    assert lib.sapio_abi_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back to anything."""
    import torch
    from sapiogenesis_b200 import _lib
    from sapiogenesis_b200.engine import Engine
    from sapiogenesis_b200.world import World, default_params
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    w = World(default_params(n_org=2, n_dead=4))
    with pytest.raises(_lib.SapioError):
        Engine(w)


def test_struct_layout_matches_header():
    """ctypes mirror of SapioParams == the C struct (checked through the oracle, which is built from the same header)."""
    from oracle import oracle as orc
    from sapiogenesis_b200.world import World, default_params
    orc.build()
    w = World(default_params(n_org=3, n_dead=5, seed=77))
    assert orc.post(w) == 0 and int(w.np("g_population")[0]) == 0
    assert orc.spawn(w, 1, 500.0, 500.0) == 0
    orc.post(w)
    assert int(w.np("g_population")[0]) == 1 and int(w.np("org_state")[1]) == 1

"""The C-ABI library loads on a CPU-only machine and exports exactly the symbols include/gsrigid.h declares."""

import ctypes
import os
import re

import pytest

from conftest import ROOT
from genesis_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsrigid.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = _declared_symbols()
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.build())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/gsrigid.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == names


def test_version_and_error_paths_without_gpu():
    L = _lib.lib()
    assert b"sm_100a" in L.gsr_version()
    h = ctypes.c_void_p()
    assert L.gsr_model_begin(ctypes.byref(h)) == 0
    # call-sequence errors are reported through return codes + gsr_last_error, never by crashing
    assert L.gsr_step(h, 1, None) == -4  # GSR_ESTATE: not bound
    assert b"not bound" in L.gsr_last_error(h)
    nw = ctypes.c_size_t()
    assert L.gsr_state_words(h, ctypes.byref(nw)) == -4
    assert L.gsr_model_set_f32(h, b"x", None, 3) == -1  # GSR_EINVAL
    assert L.gsr_field_count(h) == -4
    L.gsr_destroy(h)


def test_product_path_has_no_cpu_fallback():
    """The package never imports the oracle, and building a scene without CUDA fails loudly."""
    import torch

    pkg = os.path.join(ROOT, "genesis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#include\s+\"[^\"]*oracle", r"liboracle", r"pyoracle"):
                    assert not re.search(pat, src, flags=re.M), f"{f} links/imports the oracle ({pat})"
    if not torch.cuda.is_available():
        import genesis_b200 as gs

        gs.init(backend=gs.gpu)
        scene = gs.Scene()
        scene.add_entity(gs.morphs.Plane())
        with pytest.raises(gs.GenesisException, match="no CPU fallback"):
            scene.build(n_envs=2)
        with pytest.raises(gs.GenesisException):
            gs.init(backend=gs.cpu)

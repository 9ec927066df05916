"""The C-ABI library loads, exports every symbol include/modppl_b200.h declares, and refuses to compute without a
GPU (no CPU fallback).  Runs on CPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "modppl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import modppl_b200
    lib = ctypes.CDLL(modppl_b200.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/modppl_b200.h but not exported"


def test_python_binding_covers_header():
    from modppl_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_model_registry_validates():
    import modppl_b200 as m
    assert m.lgssm4().state_dim == 4 and m.lgssm4().obs_dim == 2
    assert m.spiral_model().state_dim == 2
    assert m.hierarchical_model(range(11)).num_latents == 4
    with pytest.raises(m.MplError):
        m.Model("no_such_model", [1.0])
    with pytest.raises(m.MplError):
        m.pointed_model([0, 0, 0, 1], [1, 0, 0, 1])       # xmax must exceed xmin (types_2d.rs:24-25)
    with pytest.raises(m.MplError):
        m.Model("hmm", [3, 3, 0.2])                          # truncated parameter vector


def test_no_cpu_fallback():
    import modppl_b200 as m
    if m.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(m.MplError, match="no CPU fallback"):
        m.ParticleSystem(m.lgssm4(), 1000)
    with pytest.raises(m.MplError, match="no CPU fallback"):
        m.parity.resample_indices([0.5, 0.5], [0.3])
    with pytest.raises(m.MplError, match="no CPU fallback"):
        m.importance_sampling(m.line_model([0.0, 1.0]), [0.0, 1.0], 16)


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under modppl_b200/ may mention it
    pkg = os.path.join(ROOT, "modppl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "libmodppl_oracle" not in text and "modppl_oracle.h" not in text, f

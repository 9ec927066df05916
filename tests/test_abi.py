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


# ------------------------------------------------------------------------------------------------- model front-end (no device needed)
def test_model_spec_compiles_with_nvrtc_without_a_device():
    """a declarative spec becomes a device functor; NVRTC compiles it against the library's embedded kernel headers
    (sm_100a cubin) -- which needs no GPU, so the whole front-end up to the module load is checked here."""
    import modppl_b200 as m
    for spec in (m.lgssm4_spec(), m.spiral_spec()):
        mod = m.compile_model(spec)
        assert mod.state_dim == spec["state_dim"] and mod.obs_dim == spec["obs_dim"]
        src = mod.source("f32")
        assert "struct JitModel" in src and '#include "pf_kernels.cuh"' in src
        for dtype in ("f32", "f64"):
            log = mod.compile(dtype)          # raises with the compiler's messages if the generated functor does not compile
            assert "error" not in log.lower()


def test_model_spec_errors_are_reported():
    import modppl_b200 as m
    bad = [
        "not json",
        {"state_dim": 2, "obs_dim": 1},                                                            # no sample statements
        {"state_dim": 1, "obs_dim": 1, "init": [{"dist": "poisson", "args": ["1", "2"]}], "step": [{"dist": "delta", "args": ["x[0]"]}]},
        {"state_dim": 1, "obs_dim": 1, "params": {"x": 1.0}, "init": [{"dist": "delta", "args": ["0"]}], "step": [{"dist": "delta", "args": ["x[0]"]}]},   # reserved name
        {"state_dim": 1, "obs_dim": 1, "init": [{"dist": "delta", "args": ["0; return 1"]}], "step": [{"dist": "delta", "args": ["x[0]"]}]},  # not an expression
    ]
    for spec in bad:
        with pytest.raises(m.MplError):
            m.compile_model(spec)
    # well-formed JSON whose expression does not compile: reported with the compiler's message when it is compiled
    mod = m.compile_model({"state_dim": 1, "obs_dim": 1, "init": [{"dist": "delta", "args": ["0"]}], "step": [{"dist": "delta", "args": ["undefined_name + x[0]"]}]})
    with pytest.raises(m.MplError) as e:
        mod.compile("f64")
    assert "undefined_name" in str(e.value)


def _split_args(arglist):
    arglist = arglist.strip()
    if arglist in ("", "void"):
        return []
    return [a for a in arglist.split(",") if a.strip()]


def test_rust_ffi_matches_the_header():
    """rust/modppl-b200/src/ffi.rs (the reference-language binding; no Rust toolchain in this image) declares the header's entry
    points with the same argument counts, only exported symbols, and the header's constants."""
    import modppl_b200
    header = open(os.path.join(ROOT, "include", "modppl_b200.h")).read()
    header_nc = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = {m.group(1): len(_split_args(m.group(2))) for m in re.finditer(r"\b(mpl_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", header_nc)}
    rust = open(os.path.join(ROOT, "rust", "modppl-b200", "src", "ffi.rs")).read()
    rust_nc = re.sub(r"//[^\n]*", "", rust)
    decls = {m.group(1): len(_split_args(m.group(2))) for m in re.finditer(r"pub fn (mpl_[a-z0-9_]+)\s*\(([^()]*)\)", rust_nc)}
    assert len(decls) >= 60
    lib = ctypes.CDLL(modppl_b200.LIB_PATH)
    for name, n_args in decls.items():
        assert name in protos, f"{name} declared in ffi.rs but not in the header"
        assert protos[name] == n_args, f"{name}: {n_args} arguments in ffi.rs, {protos[name]} in the header"
        assert hasattr(lib, name)
    missing = [s for s in protos if s not in decls and not s.startswith("mpl_test_") and s != "mpl_fixed_resample"]
    assert missing == [], f"header entry points without a Rust declaration: {missing}"
    defines = dict(re.findall(r"#define\s+(MPL_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", header_nc))
    consts = dict(re.findall(r"pub const (MPL_[A-Z0-9_]+): [a-z_0-9]+ = (-?\d+);", rust_nc))
    assert len(consts) >= 15
    for k, v in consts.items():
        assert k in defines and int(defines[k]) == int(v), (k, v, defines.get(k))
    # the two #[repr(C)] structs mirror the header's field order
    for struct, fields in (("mpl_pf_config", ["dtype", "device", "seed", "gid_offset", "n_global"]), ("mpl_move", ["kind", "proposal", "arg", "mask", "repeat"])):
        body_h = re.search(r"typedef struct " + struct + r"\s*\{(.*?)\}", header_nc, flags=re.S).group(1)
        body_r = re.search(r"pub struct " + struct + r"\s*\{(.*?)\}", rust_nc, flags=re.S).group(1)
        assert re.findall(r"(\w+)\s*;", body_h) == fields
        assert re.findall(r"pub (\w+):", body_r) == fields


def test_example_spec_compiles_without_a_device():
    """examples/custom_model.py's model (the one README.md shows) goes through the spec front-end and NVRTC on the CPU"""
    import importlib.util
    import modppl_b200 as m
    spec = importlib.util.spec_from_file_location("custom_model_example", os.path.join(ROOT, "examples", "custom_model.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    model = m.compile_model(ex.SPEC)
    assert model.state_dim == 1 and model.obs_dim == 1
    model.compile("f32")
    readme = open(os.path.join(ROOT, "README.md")).read()
    assert '"params": {"phi": 0.9, "q": 0.3, "r": 0.5}' in readme      # the same parameter names (x, y, t, z, u, s are reserved)

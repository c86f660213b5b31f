"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU, exports
every symbol include/lkg.h declares, and refuses to compute without a CUDA device (no fallback)."""
import argparse
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lkg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lkg_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from literalkg_b200 import build
    return build.build(verbose=False)


def test_header_symbols_exported(lib_path):
    from literalkg_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 15
    lib = ctypes.CDLL(lib_path)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lkg.h but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == syms            # the ctypes binding covers the header one to one
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (lkg_[a-z0-9_]+)", out))
    assert exported == set(syms)


def test_built_for_sm100a(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_loads_and_fails_loudly_without_gpu(lib_path):
    from literalkg_b200 import _lib
    lib = _lib.load()
    assert lib.lkg_abi_version() == _lib.ABI_VERSION
    if not torch.cuda.is_available():
        assert lib.lkg_device_check(0) != 0
        assert b"cuda" in lib.lkg_last_error().lower()


def test_argument_validation_without_gpu(lib_path):
    from literalkg_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    assert lib.lkg_plan_workspace_bytes(10, 0, ctypes.byref(n)) == -1      # LKG_ERR_INVALID
    assert b"n_entities" in lib.lkg_last_error()
    assert lib.lkg_plan_workspace_bytes(1000, 100, ctypes.byref(n)) == 0 and n.value > 1000 * 40


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_model_refuses_cpu():
    import literalkg_b200 as L
    from literalkg_oracle import OracleConfig
    cfg = OracleConfig(n_conv_layers=1)
    args = argparse.Namespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
    m = L.LiteralKG(args, 10, 2, None, torch.zeros(10, 2), torch.zeros(10, 300))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.gat_embeddings()
    with pytest.raises(RuntimeError, match="CUDA"):
        L.GraphPlan(torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64),
                    torch.zeros(3, dtype=torch.int64), 4, 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "literalkg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "literalkg_oracle" not in src and "import oracle" not in src, f

"""The C-ABI library loads without a GPU and exports exactly what include/doppler_b200.h declares."""
import ctypes
import os
import re

import pytest

from doppler_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "doppler_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(doppler_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_abi_version():
    assert _lib.load().doppler_b200_abi_version() == 1


def test_no_torch_types_and_no_oracle_in_product():
    """The product path must not link or call the oracle, and must carry sm_100a code only."""
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out and "complex_ref" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "doppler_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in src and "oracle_lib" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import doppler_b200
    with pytest.raises(doppler_b200.DopplerError) as ei:
        doppler_b200.Mixer(0)
    assert ei.value.code == doppler_b200.dsp.ENODEV

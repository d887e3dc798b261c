"""CPU: the C-ABI library builds/loads and exports every symbol include/mptc_gpu.h declares;
no compute call is made (there is no GPU here and no CPU fallback to call instead)."""
import ctypes
import os
import re

import pytest

from mptc_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mptc_[a-z0-9_]+)\s*\(", text)))


def test_library_present_and_loads():
    assert os.path.exists(build.LIB), "run python -m mptc_b200.build (or __graft_entry__.build())"
    capi.load()


@pytest.mark.parametrize("header", [h for h in ("mptc_gpu.h", "mptc_codec.h") if os.path.exists(os.path.join(ROOT, "include", h))])
def test_exports_match_header(header):
    L = ctypes.CDLL(build.LIB)
    syms = declared_symbols(header)
    assert len(syms) >= 4
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/{header} but not exported"
    if header == "mptc_gpu.h":
        assert set(capi.EXPORTS) == set(syms)
    else:
        assert set(capi.CODEC_EXPORTS) == set(syms)


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "mptc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle", "").lower() or f == "synth.py", f"{f} references oracle/"


def test_context_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MptcError):
        capi.Context(0)

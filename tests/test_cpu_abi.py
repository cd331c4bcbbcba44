"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/ovp.h declares, and it
refuses to run without a CUDA device (no CPU fallback).  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ovp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ovp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from ov_plane_b200 import api
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(api.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from ov_plane_b200 import api, synth
    S = synth.make_scenario("tiny_points")
    with pytest.raises(api.OvpError):
        api.Context(S.options, device=0, max_state=128, max_meas_rows=1024)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under ov_plane_b200/ or include/ may mention it."""
    bad = []
    for base in ("ov_plane_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".hpp", ".cpp")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"liboracle|oracle_backend|oracle/|import oracle|orc_", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad

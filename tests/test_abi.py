"""CPU: the C-ABI library loads and exports every symbol include/rtjx.h declares; host-only behaviour."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rtjx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rtjx_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported():
    from regtools_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/rtjx.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"


def test_struct_sizes_match_header():
    from regtools_b200 import _lib
    assert C.sizeof(_lib.Junction) == 40 and C.sizeof(_lib.Candidate) == 24
    p = _lib.Params()
    _lib.lib.rtjx_params_default(C.byref(p))
    assert p.struct_size == C.sizeof(_lib.Params)
    assert (p.min_anchor, p.min_intron, p.max_intron, p.region, p.strand_tag) == (8, 70, 500000, b".", b"XS")


def test_unsupported_modes_are_refused():
    from regtools_b200 import _lib
    for field in ("barcode_out",):
        p = _lib.Params()
        _lib.lib.rtjx_params_default(C.byref(p))
        setattr(p, field, b"x")
        h = C.c_void_p()
        assert _lib.lib.rtjx_create(C.byref(p), C.byref(h)) == _lib.RTJX_E_UNSUPPORTED
        assert b"not built" in _lib.lib.rtjx_last_error(None)


def test_compute_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a CUDA device every compute entry point returns RTJX_E_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import regtools_b200 as rt
    ex = rt.JunctionsExtractor(os.path.join(ROOT, "tests", "golden", "hcc1395", "test_hcc1395.bam"), ".", 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.identify_junctions_from_BAM()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.add_junction(rt.Junction("chr1", 1, 100, 0, 120, "+"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.scan_batch(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.uint32), np.array([0, 1], np.uint32),
                      np.array([16], np.uint32))
    host = rt.JunctionsExtractor(device=-1)
    with pytest.raises(RuntimeError, match="host-only"):
        host.add_junction(rt.Junction("chr1", 1, 100, 0, 120, "+"))

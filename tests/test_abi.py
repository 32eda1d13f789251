"""CPU-side checks of the C-ABI library: it loads without a GPU and exports
every symbol include/posfeat_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "posfeat_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(posfeat_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_four_subsystems():
    names = declared_symbols()
    for must in ("posfeat_detect_topk_f32", "posfeat_sample_l2norm_f32", "posfeat_mnn_f32",
                 "posfeat_mnn_host_f32", "posfeat_corr_expect_fwd_f32", "posfeat_corr_expect_bwd_f32",
                 "posfeat_window_expect_fwd_f32", "posfeat_window_expect_bwd_f32"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from posfeat_b200 import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    # the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert lib.posfeat_version() >= 100


def test_argument_validation_without_gpu():
    from posfeat_b200 import _lib
    lib = _lib.load()
    assert lib.posfeat_detect_workspace_bytes(1, 2, 2, 10) == 0
    assert lib.posfeat_detect_workspace_bytes(2, 480, 640, 4096) > 2 * 478 * 638 * 8
    assert lib.posfeat_mnn_workspace_bytes(0, 5, 128, 0) == 0
    # NULL pointers are rejected before any CUDA call
    st = lib.posfeat_mnn_f32(None, 4, 128, None, 4, 128, 128, 0, None, None, None, None, None, 0, None)
    assert st == 1 and "NULL" in _lib.last_error()
    st = lib.posfeat_sample_l2norm_f32(None, 1, 128, 4, 4, 0, 0, 0, 0, None, 1, None, 1, None, None, None)
    assert st == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import posfeat_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        posfeat_b200.mnn_matcher(torch.zeros(4, 128), torch.zeros(4, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        posfeat_b200.generate_kpts_single(torch.ones(1, 1, 16, 16), 1)

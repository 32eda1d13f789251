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


def test_host_buffer_entry_argument_checks_without_gpu():
    """Host-buffer staging (posfeat_fetch_taps_f32 / posfeat_host_device_pointer): argument validation and the
    host-side bookkeeping of PairPipeline.run_host run without a device."""
    import ctypes as C

    import torch
    from posfeat_b200 import _lib
    from posfeat_b200.pairs import PairPipeline
    lib = _lib.load()
    assert lib.posfeat_fetch_taps_workspace_bytes(0, 4, 4) == 0
    # bitmap of ceil(h*w/32) words per image + one 64-bit counter, rounded to 256 bytes
    assert lib.posfeat_fetch_taps_workspace_bytes(2, 224, 300) == (2 * 2100 * 4 + 255) // 256 * 256 + 256
    st = lib.posfeat_fetch_taps_f32(None, None, 1, 128, 4, 4, 0, 1, 0, 0, None, 1, None, 0, None)
    assert st == 1 and "NULL" in _lib.last_error()
    out = C.c_void_p(0)
    assert lib.posfeat_host_device_pointer(None, C.byref(out)) == 1
    # bytes over the host link: whole maps vs. requested taps (4 taps x D floats per keypoint)
    score = torch.zeros(4, 1, 32, 48)
    fmap = torch.zeros(4, 128, 8, 12)
    assert PairPipeline.h2d_bytes(score, fmap) == 4 * 32 * 48 * 4 + 4 * 128 * 8 * 12 * 4
    assert PairPipeline.h2d_bytes(score, fmap, n=10) == 4 * 32 * 48 * 4 + 4 * 10 * 4 * 128 * 4
    assert PairPipeline.d2h_bytes(4, 10, 2) == 4 * 10 * 2 * 4 + 2 * 10 * 2 * 8 + 2 * 4
    # pageable maps never take the gather paths
    assert not PairPipeline.host_gather_applies(fmap.contiguous(memory_format=torch.channels_last))
    pipe = PairPipeline(dict(nms_radius=1, num_pts=128))
    with pytest.raises(ValueError):
        pipe.run_host(score, fmap, gather="nonsense")
    with pytest.raises(ValueError):
        pipe.run_host(score, fmap, gather="stage")          # pageable map
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pipe.run_host(score, fmap)

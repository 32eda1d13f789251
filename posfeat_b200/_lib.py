"""ctypes binding of libposfeat_b200.so (the C ABI declared in include/posfeat_b200.h).

There is no CPU fallback: if the shared library is missing this module raises,
and every op raises if it is handed work without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libposfeat_b200.so")

c_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/posfeat_b200.h one to one
SIGNATURES = {
    "posfeat_version": (_i, []),
    "posfeat_last_error": (_i, [C.c_char_p, _i]),
    "posfeat_device_sm_count": (_i, []),
    "posfeat_launch_count": (_i64, []),
    "posfeat_host_device_pointer": (_i, [_vp, C.POINTER(_vp)]),
    "posfeat_profile_enable": (_i, [_i]),
    "posfeat_profile_slot_count": (_i, []),
    "posfeat_profile_slot_name": (C.c_char_p, [_i]),
    "posfeat_profile_read": (_i, [_i, C.POINTER(C.c_double), c_i32p]),
    "posfeat_detect_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "posfeat_detect_candidates_f32": (_i, [_vp, _i, _i, _i, _i64, _i64, _i, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "posfeat_detect_select_f32": (_i, [_vp, _i, _i, _i, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _sz, _vp]),
    "posfeat_detect_topk_f32": (_i, [_vp, _i, _i, _i, _i64, _i64, _i, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _sz, _vp]),
    "posfeat_detect_status": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "posfeat_detect_finish": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "posfeat_sample_l2norm_f32": (_i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _i, _vp, _i, _vp,
                                       _vp, _vp]),
    "posfeat_sample_pairs_f32": (_i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "posfeat_fetch_taps_workspace_bytes": (_sz, [_i, _i, _i]),
    "posfeat_fetch_taps_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _i, _vp, _sz, _vp]),
    "posfeat_fetch_taps_count": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_uint64), _vp]),
    "posfeat_sample_bwd_f32": (_i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _i, _vp, _vp]),
    "posfeat_mnn_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "posfeat_mnn_f32": (_i, [_vp, _i, _i64, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "posfeat_mnn_batched_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "posfeat_mnn_batched_f32": (_i, [_vp, _i64, _i, _i64, _vp, _i64, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                     _sz, _vp]),
    "posfeat_ratio_match_workspace_bytes": (_sz, [_i, _i, _i]),
    "posfeat_ratio_match_f32": (_i, [_vp, _i, _i64, _vp, _i, _i64, _i, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "posfeat_mnn_host_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "posfeat_mnn_host_f32": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "posfeat_corr_expect_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "posfeat_corr_expect_fwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "posfeat_corr_expect_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "posfeat_corr_expect_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _sz, _vp]),
    "posfeat_compute_prob_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "posfeat_normalize_scale_fwd_f32": (_i, [_vp, _i, _i, _i, _i64, _i64, _f, _f, _vp, _vp, _vp]),
    "posfeat_normalize_scale_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i64, _f, _f, _vp, _vp]),
    "posfeat_window_expect_fwd_f32": (_i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _vp, _i, _vp, _i, _i,
                                           _vp, _vp, _vp, _vp, _vp]),
    "posfeat_dual_softmax_reward_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "posfeat_dual_softmax_reward_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _i, _vp, _vp, _sz, _vp]),
    "posfeat_line_search_f32": (_i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp]),
    "posfeat_window_expect_bwd_f32": (_i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _vp, _vp, _i, _vp, _i,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

NMS_NONE, NMS_HARD, NMS_SOFT = 0, 1, 2
DETECT_FULLMAP = 0x10
MNN_PREPARED = 0x100
THR_NONE, THR_ABS, THR_MAX, THR_MEAN = 0, 1, 2, 3
MNN_AUTO, MNN_SIMT, MNN_TC = 0, 1, 2

_lib = None


def load():
    """Load the shared library (once); raise loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C posfeat_b200/csrc`).  posfeat_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().posfeat_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


class PosfeatError(RuntimeError):
    pass


def check(status: int):
    if status != 0:
        raise PosfeatError(f"posfeat_b200 error {status}: {last_error()}")


def profile_enable(on: bool):
    load().posfeat_profile_enable(int(bool(on)))


def profile_read():
    """{kernel name: (total_ms, launches)} since the last read (synchronises)."""
    lib = load()
    out = {}
    for slot in range(lib.posfeat_profile_slot_count()):
        ms, n = C.c_double(0), C.c_int32(0)
        check(lib.posfeat_profile_read(slot, C.byref(ms), C.byref(n)))
        if n.value:
            out[lib.posfeat_profile_slot_name(slot).decode()] = (ms.value, n.value)
    return out


def launch_count() -> int:
    return int(load().posfeat_launch_count())

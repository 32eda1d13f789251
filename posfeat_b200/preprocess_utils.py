"""Drop-in replacements for the hot-path callables of the reference's
``losses/preprocess_utils.py`` -- same names, argument meaning and return
values, backed by the sm_100a kernels behind the C ABI.

Install into a reference checkout with ``posfeat_b200.install(putils_module)``
(monkey-patches the attributes the managers look up by name,
managers/extractor.py:87, managers/trainer.py:434) or import from here.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._runtime import check, lib, map_ptr, ptr, require_cuda, stream_ptr, to_device, workspace

MIN_PTS = 128   # losses/preprocess_utils.py:260-261


# ---------------------------------------------------------------- coordinates
_centres = {}


def _centre(h, w, device):
    """[(w-1)/2, (h-1)/2] on ``device``, built once per (h, w, device): creating it per call would be a
    pageable host->device copy on every call (and makes the callers impossible to capture in a CUDA graph)."""
    key = (int(h), int(w), str(device))
    c = _centres.get(key)
    if c is None:
        c = _centres[key] = torch.tensor([(w - 1) / 2., (h - 1) / 2.], dtype=torch.float32).to(device)
    return c


def normalize_coords(coord, h, w):
    """losses/preprocess_utils.py:14-26 (plain tensor arithmetic, no kernel)."""
    c = _centre(h, w, coord.device)
    return (coord - c) / c


def denormalize_coords(coord_norm, h, w):
    """losses/preprocess_utils.py:28-38."""
    c = _centre(h, w, coord_norm.device)
    return coord_norm * c + c


# ------------------------------------------------------------------- detector
def _thr_mode(thr, thr_mod):
    if not thr:
        return _lib.THR_NONE, 0.0
    modes = {"abs": _lib.THR_ABS, "max": _lib.THR_MAX, "mean": _lib.THR_MEAN}
    if thr_mod not in modes:
        # the reference leaves kp_thr undefined here -> UnboundLocalError
        raise UnboundLocalError(f"thr_mod must be one of {sorted(modes)} (got {thr_mod!r})")
    return modes[thr_mod], float(thr)


def detect_topk(kp_map, nms_radius, num_pts=False, use_nms=True, thr=False, thr_mod="mean",
                sync=True, fullmap=False, ws_key="detect"):
    """Kernel-level detector.  Returns a dict with device tensors
    ``kps [b,cap,2]``, ``score [b,cap]``, ``idx [b,cap]`` (int64 linear index into
    the interior grid), ``counts [b]``, ``n`` (python int when ``sync`` else a
    device int32 tensor).  Rows >= n are undefined.  ``fullmap`` selects the
    ``generate_kpts_single_noavg`` variant (whole map, no 3x3 centroid)."""
    if use_nms == "softnms" and not thr:
        # the reference reads thr_mask, which only exists under `if thr:` (:232-240, :251-259)
        raise UnboundLocalError("use_nms='softnms' needs thr (thr_mask is referenced before assignment otherwise)")
    x, _ = to_device(kp_map)
    if x.dim() != 4 or x.shape[1] != 1:
        raise ValueError(f"kp_map must be [b,1,h,w], got {tuple(x.shape)}")
    if x.stride(3) != 1 or x.stride(2) < x.shape[3]:
        x = x.contiguous()
    b, _, h, w = x.shape
    L = lib()
    dev = x.device
    nms_mode = _lib.NMS_SOFT if use_nms == "softnms" else _lib.NMS_HARD if use_nms else _lib.NMS_NONE
    thr_mode, thr_val = _thr_mode(thr, thr_mod)
    pad = 0
    if fullmap:
        if thr_mode == _lib.THR_ABS:
            raise UnboundLocalError("generate_kpts_single_noavg has no thr_mod='abs' (kp_thr unset in the reference)")
        nms_mode |= _lib.DETECT_FULLMAP
        pad = 2
    counts = torch.empty(b, dtype=torch.int32, device=dev)
    n_out = torch.empty(1, dtype=torch.int32, device=dev)
    st = stream_ptr(dev)
    with torch.cuda.device(dev):
        if num_pts:
            cap = max(int(num_pts), MIN_PTS)
            ws_bytes = L.posfeat_detect_workspace_bytes(b, h + pad, w + pad, cap)
            ws = workspace(ws_key, ws_bytes, dev)
            idx = torch.empty((b, cap), dtype=torch.int64, device=dev)
            kps = torch.empty((b, cap, 2), dtype=torch.float32, device=dev)
            sc = torch.empty((b, cap), dtype=torch.float32, device=dev)
            check(L.posfeat_detect_topk_f32(x.data_ptr(), b, h, w, x.stride(0), x.stride(2), nms_mode,
                                            int(nms_radius), thr_mode, thr_val, int(num_pts), MIN_PTS, cap,
                                            counts.data_ptr(), n_out.data_ptr(), idx.data_ptr(),
                                            kps.data_ptr(), sc.data_ptr(), ws.data_ptr(), ws.numel(), st))
        else:
            # num_pts=False: n is the smallest survivor count -> read it, then select
            ws_bytes = L.posfeat_detect_workspace_bytes(b, h + pad, w + pad, 1)
            ws = workspace(ws_key, ws_bytes, dev)
            check(L.posfeat_detect_candidates_f32(x.data_ptr(), b, h, w, x.stride(0), x.stride(2), nms_mode,
                                                  int(nms_radius), thr_mode, thr_val, counts.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), st))
            cap = max(int(counts.min().item()), MIN_PTS)
            ws_bytes = L.posfeat_detect_workspace_bytes(b, h + pad, w + pad, cap)
            if ws_bytes > ws.numel():
                raise _lib.PosfeatError("detect workspace grew between phases")  # cannot happen: cand dominates
            idx = torch.empty((b, cap), dtype=torch.int64, device=dev)
            kps = torch.empty((b, cap, 2), dtype=torch.float32, device=dev)
            sc = torch.empty((b, cap), dtype=torch.float32, device=dev)
            check(L.posfeat_detect_select_f32(x.data_ptr(), b, h, w, x.stride(0), x.stride(2), nms_mode, 0, MIN_PTS,
                                              cap,
                                              cap, counts.data_ptr(), n_out.data_ptr(), idx.data_ptr(),
                                              kps.data_ptr(), sc.data_ptr(), ws.data_ptr(), ws.numel(), st))
    r = {"kps": kps, "score": sc, "idx": idx, "counts": counts, "n": n_out, "cap": cap,
         "_finish": (ws, b, h, w, cap, n_out, dev)}
    if sync:
        detect_finish(r)
    return r


def detect_finish(r):
    """Waits for a ``detect_topk(sync=False)`` call: device status and the keypoint count n come back in one
    host round trip; ``r["n"]`` becomes a python int.  Work that does not depend on n (allocations, size
    queries) can be done between the two calls while the selection kernel runs."""
    import ctypes
    ws, b, h, w, cap, n_out, dev = r["_finish"]
    n_host = ctypes.c_int32(0)
    with torch.cuda.device(dev):
        check(lib().posfeat_detect_finish(ws.data_ptr(), b, h, w, cap, n_out.data_ptr(), ctypes.byref(n_host),
                                          stream_ptr(dev)))
    r["n"] = int(n_host.value)
    return r["n"]


def generate_kpts_single(kp_map, nms_radius, num_pts=False, scale=4, stable=True, temperature=1, stride=1,
                         use_nms=True, thr=False, thr_mod="mean", return_idx=False):
    """Same contract as the reference (losses/preprocess_utils.py:215-278, stable
    branch): returns ``kps [b,n,2]`` (normalised x, y) and ``kp_score [b,n,1]``.
    Order: score descending; inside an equal-score group index ascending (the
    reference leaves that to ``torch.topk``)."""
    if not stable:
        raise NotImplementedError("stable=False (gumbel soft selection, training only) is out of scope")
    if stride != 1:
        raise NotImplementedError("stride != 1 is not supported (the reference's own shapes break there)")
    r = detect_topk(kp_map, nms_radius, num_pts, use_nms, thr, thr_mod, sync=True)
    n = r["n"]
    dev = kp_map.device
    kps = r["kps"][:, :n].to(dev)
    sc = r["score"][:, :n, None].to(dev)
    if return_idx:
        return kps, sc, r["idx"][:, :n].to(dev), r["counts"].to(dev)
    return kps, sc


def generate_kpts_single_noavg(kp_map, nms_radius, num_pts=False, scale=4, stable=True, temperature=1, stride=1,
                               use_nms=True, thr=False, thr_mod="mean", return_idx=False):
    """losses/preprocess_utils.py:280-336 (stable branch): NMS / threshold / top-k over the
    whole map; ``kps`` are the winners' own grid coordinates, ``kp_score`` their scores."""
    if not stable:
        raise NotImplementedError("stable=False (gumbel soft selection, training only) is out of scope")
    r = detect_topk(kp_map, nms_radius, num_pts, use_nms, thr, thr_mod, sync=True, fullmap=True)
    n = r["n"]
    dev = kp_map.device
    kps = r["kps"][:, :n].to(dev)
    sc = r["score"][:, :n, None].to(dev)
    if return_idx:
        return kps, sc, r["idx"][:, :n].to(dev), r["counts"].to(dev)
    return kps, sc


# -------------------------------------------------------------------- sampler
def sample_l2norm(x, coord_n, norm=False, n_valid=None, want_bf16=False, out=None):
    """Kernel-level sampler on device tensors.  x may be NCHW-contiguous or
    channels_last; returns [b,n,c] float32 (and a bf16 copy if asked).  x may also be a pinned host
    tensor: the kernel then reads the taps over the host link (coord_n decides the device)."""
    L = lib()
    if x.dim() != 4 or coord_n.dim() != 3 or coord_n.shape[-1] != 2 or coord_n.shape[0] != x.shape[0]:
        raise ValueError(f"expected x [b,c,h,w] and coord_n [b,n,2], got {tuple(x.shape)} {tuple(coord_n.shape)}")
    b, c, h, w = x.shape
    n = coord_n.shape[1]
    dev = x.device if x.is_cuda else coord_n.device
    if dev.type != "cuda" or not coord_n.is_cuda or coord_n.device != dev:
        raise RuntimeError("sample_l2norm needs device tensors on one device (a pinned host map is accepted with device "
                           "coordinates); posfeat_b200 has no CPU fallback")
    # the kernel reads float32 through raw pointers: convert anything else first (an fp16/bf16 backbone map
    # handed over as-is would be read out of bounds)
    if x.dtype != torch.float32:
        if not x.is_cuda:
            raise TypeError(f"a host-resident descriptor map must be float32 (got {x.dtype})")
        x = x.to(torch.float32)
    if coord_n.dtype != torch.float32:
        coord_n = coord_n.to(torch.float32)
    if out is None:
        out = torch.empty((b, n, c), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (b, n, c) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 [b,n,c] tensor")
    obf = torch.empty((b, n, c), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    if n_valid is not None:
        out.zero_()
        if obf is not None:
            obf.zero_()
    coord_n = coord_n.contiguous()
    if n == 0 or b == 0:
        return (out, obf) if want_bf16 else out
    with torch.cuda.device(dev):
        check(L.posfeat_sample_l2norm_f32(map_ptr(x), b, c, h, w, x.stride(0), x.stride(1), x.stride(2),
                                          x.stride(3), coord_n.data_ptr(), n, ptr(n_valid), int(bool(norm)),
                                          out.data_ptr(), ptr(obf), stream_ptr(dev)))
    return (out, obf) if want_bf16 else out


def sample_feat_by_coord(x, coord_n, norm=False):
    """losses/preprocess_utils.py:40-53: bilinear grid_sample (zeros padding,
    align_corners=False) at coord_n [b,n,2], optional L2 norm -> [b,n,c].
    Differentiable: the backward pass is in posfeat_b200.preprocess."""
    if torch.is_grad_enabled() and x.requires_grad:
        from .preprocess import sample_feat_by_coord_grad
        return sample_feat_by_coord_grad(x, coord_n, bool(norm))
    xd, dev = to_device(x)
    cd, _ = to_device(coord_n)
    if cd.dim() != 3 or cd.shape[-1] != 2 or cd.shape[0] != xd.shape[0]:
        raise ValueError(f"coord_n must be [b,n,2], got {tuple(cd.shape)}")
    return sample_l2norm(xd, cd, norm).to(dev)


# -------------------------------------------------------------------- matcher
def mnn_match(desc_a, desc_b, algo=_lib.MNN_AUTO, want_nn21=True):
    """Device-level matcher: returns (matches [N,2] int64 device, n_matches [1]
    int32 device, nn12, nn21).  ``want_nn21=False`` lets the tensor-core path skip
    the second direction (mutual pairs are verified by a column scan instead);
    nn21 is then None."""
    L = lib()
    a, _ = to_device(desc_a)
    b, _ = to_device(desc_b)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"descriptors must be [N,D] and [M,D], got {tuple(a.shape)} {tuple(b.shape)}")
    N, D = a.shape
    M = b.shape[0]
    if N == 0 or M == 0:
        raise IndexError("mnn_matcher: cannot take the max over an empty descriptor set "
                         "(torch.max raises on the reference path too)")
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    dev = a.device
    nn12 = torch.empty(N, dtype=torch.int32, device=dev)
    # nn21 is needed by the exact SIMT kernel and by callers that ask for it; the tensor-core path without it
    # contracts one direction only (candidate lists + group entries, any M)
    needs_nn21 = (want_nn21 or algo == _lib.MNN_SIMT or D != 128 or
                  (algo == _lib.MNN_AUTO and N * M < 1024 * 1024))
    nn21 = torch.empty(M, dtype=torch.int32, device=dev) if needs_nn21 else None
    matches = torch.empty((N, 2), dtype=torch.int64, device=dev)
    nm = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws_bytes = L.posfeat_mnn_workspace_bytes(N, M, D, algo)
        ws = workspace("mnn", ws_bytes, dev)
        check(L.posfeat_mnn_f32(a.data_ptr(), N, a.stride(0), b.data_ptr(), M, b.stride(0), D, algo,
                                nn12.data_ptr(), ptr(nn21), matches.data_ptr(), nm.data_ptr(),
                                ws.data_ptr(), ws.numel(), stream_ptr(dev)))
    return matches, nm, nn12, nn21


def mnn_matcher(descriptors_a, descriptors_b, algo=_lib.MNN_AUTO):
    """mnn_matcher, losses/preprocess_utils.py:795-803 and
    evaluations/hpatches/evaluation.py:27-38: (K,2) int64 ndarray of mutual
    nearest neighbours, rows in ascending index of descriptors_a."""
    if descriptors_a.device.type == "cpu" and descriptors_b.device.type == "cpu":
        return _mnn_host(descriptors_a, descriptors_b, algo)
    matches, nm, _, _ = mnn_match(descriptors_a, descriptors_b, algo, want_nn21=False)
    k = int(nm.item())
    return matches[:k].cpu().numpy()


def _mnn_host(a, b, algo):
    """Host-buffer entry (posfeat_mnn_host_f32): the .npz -> matches path."""
    require_cuda()
    L = lib()
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(torch.float32).contiguous()
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"descriptors must be [N,D] and [M,D], got {tuple(a.shape)} {tuple(b.shape)}")
    N, D = a.shape
    M = b.shape[0]
    if N == 0 or M == 0:
        raise IndexError("mnn_matcher: cannot take the max over an empty descriptor set")
    dev = torch.device("cuda", torch.cuda.current_device())
    sb = L.posfeat_mnn_host_scratch_bytes(N, M, D, algo)
    scratch = workspace("mnn_host", sb, dev)
    out = np.empty((N, 2), dtype=np.int64)
    nm = np.zeros(1, dtype=np.int32)
    check(L.posfeat_mnn_host_f32(a.data_ptr(), N, b.data_ptr(), M, D, algo, out.ctypes.data, nm.ctypes.data,
                                 scratch.data_ptr(), scratch.numel(), stream_ptr(dev)))
    return out[:int(nm[0])].copy()

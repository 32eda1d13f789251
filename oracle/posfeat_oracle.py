"""CPU oracle for the PoSFeat post-backbone hot path (TEST INFRASTRUCTURE ONLY).

This module is a numpy restatement of the reference's algorithms.  It is the
checker that the CUDA path is compared against; it is never the product path.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.

Parity status: PINNED.  Every function below is checked in
``tests/test_oracle_golden.py`` against fixtures under ``tests/golden/`` that
were produced by importing the real reference from ``/root/reference`` (script:
``oracle/make_golden.py``, torch 2.11.0 CPU).  The reference itself ships no
tests or golden vectors for this path.

Each function cites the reference file:line it restates (paths relative to the
reference checkout).  All arithmetic is float32 unless stated otherwise.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# coordinate helpers
# --------------------------------------------------------------------------
def linspace_f32(lo: float, hi: float, n: int) -> np.ndarray:
    """torch.linspace(lo, hi, n) in float32, bit exact.

    ATen evaluates the lower half as fma(step, i, lo) and the upper half as
    fma(-step, n-1-i, hi) with step = (hi-lo)/(n-1) rounded to float32; a
    float64 product of two float32 numbers is exact, so one rounding of the
    float64 expression reproduces the fused operation.
    Used by gen_grid, losses/preprocess_utils.py:84-87.
    """
    if n == 1:
        return np.array([lo], dtype=F32)
    lo32, hi32 = F32(lo), F32(hi)
    step = F32(F32(hi32 - lo32) / F32(n - 1))
    i = np.arange(n, dtype=np.float64)
    up = (np.float64(lo32) + np.float64(step) * i).astype(F32)
    dn = (np.float64(hi32) - np.float64(step) * (n - 1 - i)).astype(F32)
    return np.where(np.arange(n) < n // 2, up, dn).astype(F32)


def gen_grid(h_min, h_max, w_min, w_max, len_h, len_w) -> np.ndarray:
    """Row-major list of (x, y) grid nodes, shape [len_h*len_w, 2].

    losses/preprocess_utils.py:84-87 (meshgrid + stack + transpose + reshape).
    """
    xs = linspace_f32(w_min, w_max, len_w)
    ys = linspace_f32(h_min, h_max, len_h)
    g = np.empty((len_h, len_w, 2), dtype=F32)
    g[..., 0] = xs[None, :]
    g[..., 1] = ys[:, None]
    return g.reshape(-1, 2)


def normalize_coords(coord, h, w):
    """losses/preprocess_utils.py:14-26."""
    c = np.array([(w - 1) / 2.0, (h - 1) / 2.0], dtype=F32)
    return ((coord.astype(F32) - c) / c).astype(F32)


def denormalize_coords(coord_n, h, w):
    """losses/preprocess_utils.py:28-38."""
    c = np.array([(w - 1) / 2.0, (h - 1) / 2.0], dtype=F32)
    return (coord_n.astype(F32) * c + c).astype(F32)


# --------------------------------------------------------------------------
# subsystem 1: score-map keypoint selection
# --------------------------------------------------------------------------
def _reflect(i: np.ndarray, n: int) -> np.ndarray:
    """Index map of F.pad(mode='reflect'): -k -> k, (n-1)+k -> (n-1)-k."""
    i = np.abs(i)
    return np.where(i > n - 1, 2 * (n - 1) - i, i)


def nms_keep_mask(score: np.ndarray, radius: int) -> np.ndarray:
    """Boolean mask of losses/preprocess_utils.py:449-464 (``nms``).

    ``score`` is the INTERIOR map [h, w] (the caller strips the 1-px border,
    :228).  The reference reflect-pads it by ``radius``, runs
    max_pool2d(return_indices=True) with a (2r+1)^2 window, and keeps a pixel
    iff the returned index is the pixel itself.  ATen's pooling scans the
    window row-major and replaces the running maximum only on ``val > max``,
    so the winner is the FIRST maximum in scan order.  Closed form: a pixel is
    kept iff it is strictly greater than every window entry that precedes it in
    scan order and >= every entry that follows it (entries are taken from the
    padded map, so a reflected copy of the pixel itself can beat it).
    """
    h, w = score.shape
    r = int(radius)
    if r == 0:
        return np.ones((h, w), dtype=bool)
    assert r <= h - 1 and r <= w - 1, "reflect padding needs radius < size"
    ys = _reflect(np.arange(-r, h + r), h)
    xs = _reflect(np.arange(-r, w + r), w)
    pad = score[np.ix_(ys, xs)]
    keep = np.ones((h, w), dtype=bool)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            if dy == 0 and dx == 0:
                continue
            nb = pad[r + dy:r + dy + h, r + dx:r + dx + w]
            if dy < 0 or (dy == 0 and dx < 0):
                keep &= score > nb
            else:
                keep &= score >= nb
    return keep


def soft_nms(score: np.ndarray, radius: int) -> np.ndarray:
    """losses/preprocess_utils.py:431-447: softplus(score - boxmean(reflect-pad))."""
    h, w = score.shape
    r = int(radius)
    ys = _reflect(np.arange(-r, h + r), h)
    xs = _reflect(np.arange(-r, w + r), w)
    pad = score[np.ix_(ys, xs)].astype(F32)
    acc = np.zeros((h, w), dtype=F32)
    for dy in range(2 * r + 1):
        for dx in range(2 * r + 1):
            acc = acc + pad[dy:dy + h, dx:dx + w]
    a = (score - acc / F32((2 * r + 1) ** 2)).astype(F32)
    # softplus with torch's threshold=20 shortcut
    return np.where(a > 20, a, np.log1p(np.exp(np.minimum(a, F32(20))))).astype(F32)


def detect_keys(kp_map: np.ndarray, nms_radius: int, use_nms=True, thr=False,
                thr_mod="mean"):
    """Per-image selection key of generate_kpts_single (:215-240, :264).

    Returns (key[h-2, w-2] float32, survivor_count int).  ``key`` is
    ``nms_mask * interior`` -- the tensor the reference hands to ``topk``.
    """
    interior = kp_map[1:-1, 1:-1].astype(F32)
    if use_nms == "softnms":
        mask = soft_nms(interior, nms_radius)
    elif use_nms:
        mask = nms_keep_mask(interior, nms_radius)
    else:
        mask = np.ones(interior.shape, dtype=F32)
    thr_mask = None
    if thr:
        if thr_mod == "max":
            kp_thr = interior.max()
        elif thr_mod == "mean":
            kp_thr = F32(interior.mean(dtype=np.float64))
        elif thr_mod == "abs":
            kp_thr = F32(1.0)
        else:
            raise ValueError(thr_mod)
        thr_mask = interior > F32(F32(thr) * F32(kp_thr))
        mask = thr_mask * mask
    if use_nms == "softnms":
        count = int(thr_mask.sum())
    else:
        count = int(np.asarray(mask, dtype=np.float64).sum())
    key = (np.asarray(mask, dtype=F32) * interior).astype(F32)
    return key, count


def topk_desc(key_flat: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k largest keys, score descending, index ascending on ties.

    ``torch.topk`` (:264) leaves the order inside an equal-score group
    unspecified; this project pins (score desc, linear index asc).
    """
    order = np.lexsort((np.arange(key_flat.size), -key_flat.astype(np.float64)))
    return order[:k].astype(np.int64)


def centroid_and_score(kp_map: np.ndarray, idx: np.ndarray):
    """3x3 score-weighted centroid and 3x3 max at interior indices (:243-247).

    kps = avg_pool3(score*grid)/avg_pool3(score), kp_score = max_pool3(score),
    both evaluated only at the selected interior pixels ``idx`` (linear index
    into the (h-2)x(w-2) grid).  Returns (kps[n,2] normalised xy, score[n]).
    """
    h, w = kp_map.shape
    wi = w - 2
    xs = linspace_f32(-1, 1, w)
    ys = linspace_f32(-1, 1, h)
    y0 = idx // wi
    x0 = idx % wi
    sx = np.zeros(idx.shape, dtype=F32)
    sy = np.zeros(idx.shape, dtype=F32)
    sw = np.zeros(idx.shape, dtype=F32)
    mx = np.full(idx.shape, -np.inf, dtype=F32)
    for dy in range(3):
        for dx in range(3):
            p = kp_map[y0 + dy, x0 + dx].astype(F32)
            sx = sx + p * xs[x0 + dx]
            sy = sy + p * ys[y0 + dy]
            sw = sw + p
            mx = np.maximum(mx, p)
    nine = F32(9)
    with np.errstate(divide="ignore", invalid="ignore"):
        kx = (sx / nine) / (sw / nine)
        ky = (sy / nine) / (sw / nine)
    return np.stack([kx, ky], -1).astype(F32), mx


def generate_kpts_single(kp_map, nms_radius, num_pts=False, scale=4, stable=True,
                         temperature=1, stride=1, use_nms=True, thr=False,
                         thr_mod="mean", return_idx=False):
    """losses/preprocess_utils.py:215-278, ``stable=True`` branch.

    kp_map: [b, 1, h, w] float32.  Returns kps [b, n, 2] (normalised x, y) and
    kp_score [b, n, 1]; with ``return_idx`` also idx [b, n] int64 (linear index
    into the interior grid) and the per-image survivor counts.
    """
    assert stable and stride == 1
    kp_map = np.asarray(kp_map, dtype=F32)
    b, _, h, w = kp_map.shape
    keys, counts = [], []
    for i in range(b):
        k, c = detect_keys(kp_map[i, 0], nms_radius, use_nms, thr, thr_mod)
        keys.append(k.reshape(-1))
        counts.append(c)
    min_count = min(counts)
    if not num_pts:                                   # :249-254
        n = min_count
    else:                                             # :255-259
        n = min(int(num_pts), min_count)
    if n < 128:                                       # :260-261
        n = 128
    kps = np.empty((b, n, 2), dtype=F32)
    sc = np.empty((b, n, 1), dtype=F32)
    idxs = np.empty((b, n), dtype=np.int64)
    for i in range(b):
        idx = topk_desc(keys[i], n)
        kps[i], s = centroid_and_score(kp_map[i, 0], idx)
        sc[i, :, 0] = s
        idxs[i] = idx
    if return_idx:
        return kps, sc, idxs, np.array(counts)
    return kps, sc


def generate_kpts_single_noavg(kp_map, nms_radius, num_pts=False, scale=4, stable=True,
                               temperature=1, stride=1, use_nms=True, thr=False,
                               thr_mod="mean", return_idx=False):
    """losses/preprocess_utils.py:280-336, ``stable=True`` branch: NMS, threshold and
    top-k over the WHOLE map (no border crop), keypoints are the winners' own grid
    coordinates and scores (no 3x3 centroid / max).  idx indexes the h x w map."""
    assert stable
    kp_map = np.asarray(kp_map, dtype=F32)
    b, _, h, w = kp_map.shape
    if thr and thr_mod not in ("max", "mean"):
        raise ValueError(thr_mod)                      # kp_thr is unset in the reference
    keys, counts = [], []
    for i in range(b):
        # detect_keys crops one border pixel; present the map as the interior of a padded one
        k, c = detect_keys(np.pad(kp_map[i, 0], 1), nms_radius, use_nms, thr, thr_mod)
        keys.append(k.reshape(-1))
        counts.append(c)
    n = min(counts) if not num_pts else min(int(num_pts), min(counts))
    if n < 128:
        n = 128
    ys = linspace_f32(-1, 1, h)
    xs = linspace_f32(-1, 1, w)
    kps = np.empty((b, n, 2), dtype=F32)
    sc = np.empty((b, n, 1), dtype=F32)
    idxs = np.empty((b, n), dtype=np.int64)
    for i in range(b):
        idx = topk_desc(keys[i], n)
        kps[i, :, 0] = xs[idx % w]
        kps[i, :, 1] = ys[idx // w]
        sc[i, :, 0] = kp_map[i, 0].reshape(-1)[idx]
        idxs[i] = idx
    if return_idx:
        return kps, sc, idxs, np.array(counts)
    return kps, sc


# --------------------------------------------------------------------------
# subsystem 2: bilinear descriptor sampling + L2 normalisation
# --------------------------------------------------------------------------
def grid_sample_bilinear(x: np.ndarray, coord_n: np.ndarray, padding="zeros"):
    """F.grid_sample(x, grid, 'bilinear', padding, align_corners=False).

    x [b, c, h, w], coord_n [b, n, 2] -> [b, n, c].  Unnormalise with
    ((g+1)*size-1)/2; 'zeros' drops taps outside the map, 'border' clamps the
    sampling coordinate.  Used at losses/preprocess_utils.py:48,673,737.
    """
    x = np.asarray(x, dtype=F32)
    g = np.asarray(coord_n, dtype=F32)
    b, c, h, w = x.shape
    ix = ((g[..., 0] + F32(1)) * F32(w) - F32(1)) / F32(2)
    iy = ((g[..., 1] + F32(1)) * F32(h) - F32(1)) / F32(2)
    if padding == "border":
        ix = np.clip(ix, F32(0), F32(w - 1))
        iy = np.clip(iy, F32(0), F32(h - 1))
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    wx1 = (ix - x0).astype(F32)
    wx0 = (x1 - ix).astype(F32)
    wy1 = (iy - y0).astype(F32)
    wy0 = (y1 - iy).astype(F32)
    out = np.zeros(g.shape[:-1] + (c,), dtype=F32)
    bi = np.arange(b).reshape((b,) + (1,) * (g.ndim - 2))
    for yy, xx, wt in ((y0, x0, wy0 * wx0), (y0, x1, wy0 * wx1),
                       (y1, x0, wy1 * wx0), (y1, x1, wy1 * wx1)):
        ok = (xx >= 0) & (xx <= w - 1) & (yy >= 0) & (yy <= h - 1)
        xi = np.clip(xx, 0, w - 1).astype(np.int64)
        yi = np.clip(yy, 0, h - 1).astype(np.int64)
        v = x[bi, :, yi, xi]                       # [..., c]
        out = out + v * (wt * ok).astype(F32)[..., None]
    return out


def l2_normalize(v: np.ndarray, axis=-1, eps=1e-12):
    """F.normalize(p=2): v / max(||v||, eps)  (:51)."""
    n = np.sqrt(np.sum(v.astype(F32) * v.astype(F32), axis=axis, keepdims=True,
                       dtype=F32))
    return (v / np.maximum(n, F32(eps))).astype(F32)


def sample_feat_by_coord(x, coord_n, norm=False):
    """losses/preprocess_utils.py:40-53 -> [b, n, c]."""
    f = grid_sample_bilinear(x, coord_n, "zeros")
    return l2_normalize(f, -1) if norm else f


# --------------------------------------------------------------------------
# subsystem 3: mutual nearest neighbour matching
# --------------------------------------------------------------------------
def similarity(a: np.ndarray, b: np.ndarray, exact=False) -> np.ndarray:
    """sim = A @ B.T (evaluations/hpatches/evaluation.py:30).

    ``exact=True`` accumulates in float64 (products of float32 numbers are
    exact there) -- the order-independent value near-tie checks use.
    """
    if exact:
        return a.astype(np.float64) @ b.astype(np.float64).T
    return a.astype(F32) @ b.astype(F32).T


def mnn_from_sim(sim: np.ndarray):
    nn12 = np.argmax(sim, axis=1)          # first maximum wins, as torch.max
    nn21 = np.argmax(sim, axis=0)
    ids = np.arange(sim.shape[0])
    keep = nn21[nn12] == ids
    return np.stack([ids[keep], nn12[keep]], -1).astype(np.int64), nn12, nn21


def mnn_matcher(desc_a, desc_b, exact=False, return_nn=False):
    """mnn_matcher / mutual_nn_matcher: evaluations/hpatches/evaluation.py:27-38,
    evaluations/aachen/matchers.py:5-13,
    evaluations/ETH_local_feature/custom_matcher.py:5-13,
    losses/preprocess_utils.py:795-803.  Returns (K, 2) int64, rows ascending.
    """
    m, nn12, nn21 = mnn_from_sim(similarity(np.asarray(desc_a), np.asarray(desc_b), exact))
    return (m, nn12, nn21) if return_nn else m


_C_LIB = False


def _c_oracle():
    """oracle/_build/libposfeat_oracle.so (built by `make -C oracle` / __graft_entry__.build()), or None."""
    global _C_LIB
    if _C_LIB is False:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libposfeat_oracle.so")
        _C_LIB = None
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            p = ctypes.c_void_p
            lib.posfeat_oracle_mnn_f64.restype = ctypes.c_int
            lib.posfeat_oracle_mnn_f64.argtypes = [p, ctypes.c_int64, p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                                   p, p, p, p, p, p]
            _C_LIB = lib
    return _C_LIB


def _mnn_f64_c(lib, desc_a, desc_b):
    import os
    a = np.ascontiguousarray(desc_a, dtype=np.float32)
    b = np.ascontiguousarray(desc_b, dtype=np.float32)
    if not (np.array_equal(a, np.asarray(desc_a)) and np.array_equal(b, np.asarray(desc_b))):
        raise ValueError("the C oracle takes float32 descriptors (their products are exact in float64)")
    N, M, D = a.shape[0], b.shape[0], a.shape[1]
    nn12, nn21 = np.empty(N, np.int64), np.empty(M, np.int64)
    rb, rs, cb, cs = np.empty(N), np.empty(N), np.empty(M), np.empty(M)
    st = lib.posfeat_oracle_mnn_f64(a.ctypes.data, N, b.ctypes.data, M, D, os.cpu_count() or 1, nn12.ctypes.data,
                                    rb.ctypes.data, rs.ctypes.data, nn21.ctypes.data, cb.ctypes.data, cs.ctypes.data)
    if st != 0:
        raise RuntimeError(f"posfeat_oracle_mnn_f64 failed with status {st}")
    ids = np.arange(N)
    keep = nn21[nn12] == ids
    return np.stack([ids[keep], nn12[keep]], -1).astype(np.int64), nn12, nn21, rb - rs, cb - cs


def mnn_blocked_f64(desc_a, desc_b, block=2048, force_numpy=False):
    """mnn_matcher for sizes whose similarity matrix does not fit in memory (BASELINE config 5: 64k x 64k is
    34 GB in float64): the same maths as mnn_from_sim on the exact float64 similarities, evaluated in row
    blocks.  First maximum wins in both directions (np.argmax inside a block; a later block replaces a column
    maximum only when strictly larger).  Returns (matches, nn12, nn21, row_gap, col_gap): the gaps between
    the largest and the second largest value of every row / column tell a caller whether a disagreement is
    a float64 tie.  When the C part of the oracle is built (oracle/mnn_f64.c -> oracle/_build, `make -C oracle`)
    it does the work with all host threads -- the 64k x 64k case takes seconds instead of minutes; the numpy
    form below is the same algorithm and the two are tested against each other."""
    lib = None if force_numpy else _c_oracle()
    if lib is not None:
        return _mnn_f64_c(lib, desc_a, desc_b)
    a = np.ascontiguousarray(desc_a, dtype=np.float64)
    bt = np.ascontiguousarray(np.asarray(desc_b, dtype=np.float64).T)
    N, M = a.shape[0], bt.shape[1]
    nn12 = np.empty(N, np.int64)
    row_gap = np.empty(N, np.float64)
    cmax = np.full(M, -np.inf)
    cmax2 = np.full(M, -np.inf)
    nn21 = np.zeros(M, np.int64)
    for i0 in range(0, N, block):
        s = a[i0:i0 + block] @ bt                                  # [blk, M] exact products, float64 sums
        j = np.argmax(s, axis=1)
        nn12[i0:i0 + block] = j
        r = np.arange(s.shape[0])
        top = s[r, j]
        if M > 1:
            s[r, j] = -np.inf
            row_gap[i0:i0 + block] = top - s.max(axis=1)
            s[r, j] = top
        else:
            row_gap[i0:i0 + block] = np.inf
        bi = np.argmax(s, axis=0)
        c = np.arange(M)
        bm = s[bi, c]
        if s.shape[0] > 1:
            s[bi, c] = -np.inf
            bm2 = s.max(axis=0)
        else:
            bm2 = np.full(M, -np.inf)
        better = bm > cmax
        cmax2 = np.where(better, np.maximum(cmax, bm2), np.maximum(cmax2, bm))
        nn21 = np.where(better, bi + i0, nn21)
        cmax = np.where(better, bm, cmax)
    ids = np.arange(N)
    keep = nn21[nn12] == ids
    return np.stack([ids[keep], nn12[keep]], -1).astype(np.int64), nn12, nn21, row_gap, cmax - cmax2


def _top2(sim):
    order = np.argsort(-sim, axis=1, kind="stable")[:, :2]
    vals = np.take_along_axis(sim, order, 1)
    return vals, order


def ratio_matchers(desc_a, desc_b, ratio=0.95, mutual=True, exact=False):
    """ratio_matcher / mutual_nn_ratio_matcher,
    evaluations/aachen/matchers.py:17-75 (ETH copy: custom_matcher.py:16-73)."""
    sim = similarity(np.asarray(desc_a), np.asarray(desc_b), exact).astype(F32)
    v12, n12 = _top2(sim)
    with np.errstate(invalid="ignore"):      # sim > 1 by rounding -> NaN -> fails the test, as in torch
        d12 = np.sqrt(F32(2) - F32(2) * v12, dtype=F32)
        r12 = d12[:, 0] / (d12[:, 1] + F32(1e-8))
        v21, n21 = _top2(sim.T)
        d21 = np.sqrt(F32(2) - F32(2) * v21, dtype=F32)
        r21 = d21[:, 0] / (d21[:, 1] + F32(1e-8))
    nn12, nn21 = n12[:, 0], n21[:, 0]
    ids = np.arange(sim.shape[0])
    keep = (r12 <= F32(ratio)) & (r21[nn12] <= F32(ratio))
    if mutual:
        keep &= nn21[nn12] == ids
    return np.stack([ids[keep], nn12[keep]], -1).astype(np.int64)


# --------------------------------------------------------------------------
# subsystem 4: training-side correlation + softmax expectation (forward)
# --------------------------------------------------------------------------
def softmax(z, axis=-1):
    z = z - z.max(axis=axis, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=axis, keepdims=True)


def compute_prob(feat1, feat2, loss_distance="cos", with_scale=False, return_sim=False, dtype=F32):
    """losses/preprocess_utils.py:89-115.  feat1 [B, m, d], feat2 [B, n, d] -> prob [B, m, n]."""
    assert loss_distance in ("cos", "euc")
    if return_sim:
        assert loss_distance == "cos"
    f1 = np.asarray(feat1, dtype=dtype)
    f2 = np.asarray(feat2, dtype=dtype)
    sim = f1 @ f2.transpose(0, 2, 1)
    if loss_distance == "cos":
        scale = dtype(np.sqrt(dtype(f2.shape[1]))) if with_scale else dtype(1)      # :102-106
        prob = softmax(scale * sim, -1)
    else:
        dist = (f1 ** 2).sum(-1, keepdims=True) + (f2 ** 2).sum(-1, keepdims=True).transpose(0, 2, 1) - 2 * sim   # :108-111
        prob = softmax(-dist, -1)
    return (prob, sim) if return_sim else prob


def get_expected_correspondence_locs(feat1, featmap2, with_std=False, dtype=F32):
    """losses/preprocess_utils.py:55-82.

    feat1 [B, n, d], featmap2 [B, d, h, w] -> expected normalised xy [B, n, 2];
    with_std also returns std [B, n] (sum over x, y of sqrt(clamp(var, 1e-10)))
    and prob [B, n, h*w].
    """
    feat1 = np.asarray(feat1, dtype=dtype)
    fm = np.asarray(featmap2, dtype=dtype)
    B, d, h2, w2 = fm.shape
    grid = gen_grid(-1, 1, -1, 1, h2, w2).astype(dtype)            # [hw, 2]
    flat = fm.reshape(B, d, h2 * w2).transpose(0, 2, 1)            # [B, hw, d]
    sim = feat1 @ flat.transpose(0, 2, 1)                          # [B, n, hw]
    prob = softmax(sim, -1)
    exp_xy = prob @ grid                                           # [B, n, 2]
    if not with_std:
        return exp_xy
    var = prob @ (grid ** 2) - exp_xy ** 2
    std = np.sqrt(np.clip(var, 1e-10, None)).sum(-1)
    return exp_xy, std, prob


def window_grid(window_size, h2, w2):
    """Offset grid of get_expected_correspondence_within_window (:730-732)."""
    return gen_grid(-window_size, window_size, -window_size, window_size,
                    int(window_size * h2), int(window_size * w2))


def get_expected_correspondence_within_window(feat1, featmap2, coord2_n, window_size,
                                              dtype=F32):
    """losses/preprocess_utils.py:721-758 with with_std=True.

    Returns (expected xy [B,n,2], window coords [B,n,m,2], std [B,n], prob [B,n,m]).
    """
    feat1 = np.asarray(feat1, dtype=dtype)
    fm = np.asarray(featmap2, dtype=dtype)
    c2 = np.asarray(coord2_n, dtype=F32)
    B, d, h2, w2 = fm.shape
    off = window_grid(window_size, h2, w2)                          # [m, 2] f32
    cg = (c2[:, :, None, :] + off[None, None]).astype(F32)          # [B, n, m, 2]
    n, m = cg.shape[1], cg.shape[2]
    win = grid_sample_bilinear(fm.astype(F32), cg.reshape(B, n * m, 2), "zeros")
    win = win.reshape(B, n, m, d).astype(dtype)
    sim = np.einsum("bnd,bnmd->bnm", feat1, win)
    prob = softmax(sim, -1)
    cgd = cg.astype(dtype)
    exp_xy = np.einsum("bnm,bnmc->bnc", prob, cgd)
    var = np.einsum("bnm,bnmc->bnc", prob, cgd ** 2) - exp_xy ** 2
    std = np.sqrt(np.clip(var, 1e-10, None)).sum(-1)
    return exp_xy, cg, std, prob


def grid_softmax_expectation(feat1, feat2, coord1, coord2, coord1_n, coord2_n,
                             temperature, h1, w1, h2, w2, dtype=F32):
    """Grid<->grid stage of Preprocess_Line2Window.forward, losses/preprocess.py:59-81.

    feat1 [B,m,d], feat2 [B,n,d] (L2-normalised samples), coordX pixel xy,
    coordX_n normalised xy.  Returns the four tensors the loss consumes:
    feat1g_corloc [B,m,2], feat2g_corloc [B,n,2], feat1g_std [B,m], feat2g_std [B,n].
    """
    f1 = np.asarray(feat1, dtype=dtype)
    f2 = np.asarray(feat2, dtype=dtype)
    sim = f1 @ f2.transpose(0, 2, 1)
    p12 = softmax(temperature * sim, 2)                 # rows: over image-2 points
    p21 = softmax(temperature * sim, 1)                 # cols: over image-1 points
    c1 = np.asarray(coord1, dtype=dtype)
    c2 = np.asarray(coord2, dtype=dtype)
    c1n = np.asarray(coord1_n, dtype=dtype)
    c2n = np.asarray(coord2_n, dtype=dtype)
    loc1 = p12 @ c2                                       # [B,m,2]
    loc2 = p21.transpose(0, 2, 1) @ c1                    # [B,n,2]
    ctr2 = np.array([(w2 - 1) / 2.0, (h2 - 1) / 2.0], dtype=dtype)
    ctr1 = np.array([(w1 - 1) / 2.0, (h1 - 1) / 2.0], dtype=dtype)
    loc1_n = (loc1 - ctr2) / ctr2
    loc2_n = (loc2 - ctr1) / ctr1
    std1 = np.sqrt(np.clip(p12 @ (c2n ** 2) - loc1_n ** 2, 1e-6, None)).sum(-1)
    std2 = np.sqrt(np.clip(p21.transpose(0, 2, 1) @ (c1n ** 2) - loc2_n ** 2,
                           1e-6, None)).sum(-1)
    return loc1, loc2, std1, std2


# --------------------------------------------------------------------------
# DiskLoss dense affinity (losses/kploss.py:52-88, :158-196), everything after the random point sampling
# --------------------------------------------------------------------------
def disk_loss_dense(feat1, feat2, coord1, coord2, F1, F2, logp1, logp2, acc1, acc2, T, reward_thr=2.0,
                    good_reward=1.0, bad_reward=-0.25, kp_penalty=-0.001, dynamic=False):
    """feat [b,m,d] / [b,n,d] float32 (already normalised), coord pixel xy, F [b,3,3], logp / acc [b,m] / [b,n].
    Returns (loss, dict of the reference's components, dloss/dlogp1, dloss/dlogp2) in float64."""
    f1, f2 = np.asarray(feat1, np.float64), np.asarray(feat2, np.float64)
    b, m, _ = f1.shape
    n = f2.shape[1]
    aff = -T * (1.0 - f1 @ f2.transpose(0, 2, 1))                      # kploss.py:161-166

    def logsoftmax(z, axis):
        z = z - z.max(axis, keepdims=True)
        return z - np.log(np.exp(z).sum(axis, keepdims=True))
    lr, lc = logsoftmax(aff, 2), logsoftmax(aff, 1)                     # cat_I.logits, cat_T.logits^T
    p, logp = np.exp(lr + lc), lr + lc                                   # dense_p, dense_logp

    def lines(F, c):
        ch = np.concatenate([c, np.ones_like(c[..., :1])], -1).astype(np.float64)       # b,k,3
        l = np.einsum("bij,bkj->bki", np.asarray(F, np.float64), ch)
        return l / np.maximum(np.linalg.norm(l[..., :2], axis=-1, keepdims=True), 1e-8), ch
    l1, c1h = lines(F1, np.asarray(coord1))
    l2, c2h = lines(F2, np.asarray(coord2))
    d1 = np.abs(np.einsum("bmi,bni->bmn", l1, c2h))                      # epipolar_dist
    d2 = np.abs(np.einsum("bni,bmi->bmn", l2, c1h))                      # epipolar_dist2^T
    if dynamic:
        reward = np.maximum(np.exp(-d1 / reward_thr) + np.exp(-d2 / reward_thr) - 2 / np.e, bad_reward)
    else:
        good = (d1 < reward_thr) & (d2 < reward_thr)
        reward = good_reward * good + bad_reward * (~good)
    a1, a2 = np.asarray(acc1, bool), np.asarray(acc2, bool)
    acc = a1[:, :, None] & a2[:, None, :]
    lp1, lp2 = np.asarray(logp1, np.float64), np.asarray(logp2, np.float64)
    wgt = acc * reward * p
    reinforce = (wgt * (logp + lp1[:, :, None] + lp2[:, None, :])).sum()
    pen = kp_penalty * (lp1[a1].sum() + lp2[a2].sum())
    loss = -reinforce - pen
    comp = {"reinforce": reinforce, "kp_penalty": pen, "cor minmax": p.reshape(b, -1).max(-1).min(),
            "cor minmean": p.reshape(b, -1).mean(-1).min(), "cor max": p.max(), "cor mean": p.mean(),
            "cor summin": min(p.sum(1).min(), p.sum(2).min()), "cor summax": max(p.sum(1).max(), p.sum(2).max()),
            "n_kps": (a1.sum(-1) + a2.sum(-1)).astype(np.float64).mean(), "n_pairs": p.sum((1, 2)).mean()}
    g1 = -wgt.sum(2) - kp_penalty * a1                                   # d loss / d logp1
    g2 = -wgt.sum(1) - kp_penalty * a2
    return loss, comp, g1, g2

"""Comparison helpers shared by the oracle tests and the GPU parity tests."""
import numpy as np


def check_detect(ref_key, ref_idx, ref_kps, ref_score, ref_counts, kps, sc, idx, counts):
    """(kps, sc, idx, counts) against a reference run on the same map.

    ref_key[b] is the flattened selection key (nms_mask*score) the reference
    feeds to topk.  Keypoint indices must match exactly where the selected
    score is unique, as sets inside equal-score groups above the cut-off, and
    the score sequence must be identical (torch.topk leaves tie order open;
    key == 0 entries are filler that topk picks arbitrarily).
    """
    assert kps.shape == ref_kps.shape, (kps.shape, ref_kps.shape)
    np.testing.assert_array_equal(np.asarray(counts, dtype=np.int64), np.asarray(ref_counts, dtype=np.int64))
    sc = sc.reshape(sc.shape[0], -1)
    ref_score = ref_score.reshape(ref_score.shape[0], -1)
    for b in range(ref_idx.shape[0]):
        rv = ref_key[b][ref_idx[b]]
        ov = ref_key[b][idx[b]]
        np.testing.assert_array_equal(rv, ov)
        assert len(set(idx[b].tolist())) == idx.shape[1]
        real = rv > 0
        above = real & (rv > rv[-1])
        assert set(ref_idx[b][above].tolist()) == set(idx[b][above].tolist())
        vals, cnt = np.unique(rv, return_counts=True)
        uniq = np.isin(rv, vals[cnt == 1]) & real
        np.testing.assert_array_equal(ref_idx[b][uniq], idx[b][uniq])
        o_r = np.argsort(ref_idx[b][above], kind="stable")
        o_o = np.argsort(idx[b][above], kind="stable")
        np.testing.assert_allclose(kps[b][above][o_o], ref_kps[b][above][o_r], rtol=1e-5, atol=2e-6)
        np.testing.assert_array_equal(sc[b][above][o_o], ref_score[b][above][o_r])
        for v in vals[cnt > 1]:
            if v > 0:
                assert np.all(np.diff(idx[b][rv == v]) > 0), "tie policy: index ascending"


def check_detect_float_key(ref_key, ref_idx, ref_kps, ref_score, ref_counts, kps, sc, idx, counts, rel=1e-5):
    """Soft-NMS variant of check_detect: the selection key softplus(score - boxmean) * score
    is a float32 computation (avg_pool summation, exp/log1p), so two implementations
    agree to a few ulp and may swap neighbours whose keys are that close.  Counts (the
    threshold mask) are exact; the chosen keys must equal the reference's sequence within
    `rel`, everything clearly above the cut must be chosen, and coordinates / scores of
    commonly chosen pixels must agree."""
    assert kps.shape == ref_kps.shape, (kps.shape, ref_kps.shape)
    np.testing.assert_array_equal(np.asarray(counts, dtype=np.int64), np.asarray(ref_counts, dtype=np.int64))
    sc = sc.reshape(sc.shape[0], -1)
    ref_score = ref_score.reshape(ref_score.shape[0], -1)
    for b in range(ref_idx.shape[0]):
        rv = ref_key[b][ref_idx[b]].astype(np.float64)
        ov = ref_key[b][idx[b]].astype(np.float64)
        np.testing.assert_allclose(ov, rv, rtol=rel, atol=0)
        assert len(set(idx[b].tolist())) == idx.shape[1]
        sure = rv > rv[-1] * (1 + rel)
        assert set(ref_idx[b][sure].tolist()) <= set(idx[b].tolist())
        pos = {int(j): i for i, j in enumerate(idx[b])}
        common = [(i, pos[int(j)]) for i, j in enumerate(ref_idx[b]) if rv[i] > 0 and int(j) in pos]
        assert len(common) >= int(sure.sum())
        ri, oi = np.array([c[0] for c in common]), np.array([c[1] for c in common])
        np.testing.assert_allclose(kps[b][oi], ref_kps[b][ri], rtol=1e-5, atol=2e-6)
        np.testing.assert_array_equal(sc[b][oi], ref_score[b][ri])


def check_mnn_near_tie(a, b, got, want, tol=1e-6):
    """Match lists must be identical; if they differ, every disagreeing row or
    column must be a genuine near tie of the exact (float64) similarity."""
    if got.shape == want.shape and np.array_equal(got, want):
        return
    sim = a.astype(np.float64) @ b.astype(np.float64).T
    gs, ws = {tuple(x) for x in got.tolist()}, {tuple(x) for x in want.tolist()}
    for (i, j) in gs ^ ws:
        top_r = np.sort(sim[i])[-2:]
        top_c = np.sort(sim[:, j])[-2:]
        assert (top_r[1] - top_r[0] < tol) or (top_c[1] - top_c[0] < tol), \
            f"pair {(i, j)} differs without a near tie: row gap {top_r[1]-top_r[0]}, col gap {top_c[1]-top_c[0]}"


def check_argmax_exact(x, y, got_nn, want_nn, tol=1e-12):
    """got_nn[i] must equal want_nn[i] = argmax_j <x_i, y_j> (float64, first index).  Zero mismatches is the
    bar; an entry may differ only if the two columns' exact similarities coincide to `tol` (a float64 tie,
    where summation order decides) -- and then the test proves it."""
    got_nn, want_nn = np.asarray(got_nn).astype(np.int64), np.asarray(want_nn).astype(np.int64)
    assert got_nn.shape == want_nn.shape
    bad = np.nonzero(got_nn != want_nn)[0]
    for i in bad:
        xi = x[i].astype(np.float64)
        sg, sw = float(xi @ y[got_nn[i]].astype(np.float64)), float(xi @ y[want_nn[i]].astype(np.float64))
        assert abs(sg - sw) <= tol, f"row {i}: got column {got_nn[i]} (sim {sg!r}) instead of {want_nn[i]} (sim {sw!r})"
    return len(bad)


def assert_close_vec(got, want, rel=1e-5, axis=-1):
    """|got - want| <= rel * (largest |component| of the reference vector).

    This is the "within 1e-5 relative at fp32" bar of the north star read
    vector-wise: a bilinear blend of O(1) values carries an absolute error of
    a few ulps of the sampling coordinate regardless of how close to zero an
    individual component lands, so a per-element relative test is meaningless.
    """
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = np.max(np.abs(want), axis=axis, keepdims=True)
    err = np.abs(got - want)
    bad = err > rel * scale + 1e-12
    assert not bad.any(), f"{bad.sum()} elements off; worst err/scale = {np.max(err / np.maximum(scale, 1e-30)):.3e}"


def check_ratio_near_tie(a, b, got, want, ratio, mutual, band=2e-4):
    """Ratio-test matches may differ from the reference only on rows whose Lowe
    ratio sits within float32 noise of the threshold (or whose nearest neighbour is a
    float32 near tie): sqrt(2 - 2 sim) amplifies 1e-7 similarity noise near sim = 1."""
    sim = a.astype(np.float64) @ b.astype(np.float64).T

    def ratios(s):
        part = np.sort(s, axis=1)[:, -2:]
        d0, d1 = np.sqrt(np.maximum(2 - 2 * part[:, 1], 0)), np.sqrt(np.maximum(2 - 2 * part[:, 0], 0))
        return d0 / (d1 + 1e-8), part[:, 1] - part[:, 0], part[:, 1]
    r12, gap12, top12 = ratios(sim)
    r21, gap21, top21 = ratios(sim.T)
    nn12 = sim.argmax(1)
    # noise of the ratio: d(sqrt(2-2s)) = ds / sqrt(2-2s)
    tol12 = band + 4e-7 / np.maximum(2 - 2 * top12, 1e-12)
    tol21 = band + 4e-7 / np.maximum(2 - 2 * top21, 1e-12)
    amb = (np.abs(r12 - ratio) < tol12) | (np.abs(r21[nn12] - ratio) < tol21[nn12]) | (gap12 < 1e-6)
    if mutual:
        amb |= gap21[nn12] < 1e-6
    sure = ~amb
    g = {int(i): int(j) for i, j in got}
    w = {int(i): int(j) for i, j in want}
    for i in np.nonzero(sure)[0]:
        assert g.get(int(i)) == w.get(int(i)), (i, g.get(int(i)), w.get(int(i)), r12[i], r21[nn12[i]])
    assert np.all(np.diff(got[:, 0]) > 0) if len(got) > 1 else True
    return int(amb.sum())

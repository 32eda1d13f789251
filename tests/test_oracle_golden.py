"""Pin oracle/posfeat_oracle.py against outputs of the real reference
(fixtures made by oracle/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import posfeat_oracle as O
from _checks import check_detect, check_detect_float_key

DETECT_CASES = ["r1_abs", "r3_abs", "r2_max", "r1_mean", "r1_nothr", "nonms_abs",
                "ties_r1", "ties_r2", "few", "const", "odd_r1", "odd_r5"]


def _cfg(g, name):
    return eval(str(g[name + "/cfg"]))


def check_detect_against(g, name, kps, sc, idx, counts):
    check_detect(g[name + "/key"], g[name + "/idx"], g[name + "/kps"], g[name + "/score"],
                 g[name + "/count"], kps, sc, idx, counts)


@pytest.mark.parametrize("name", DETECT_CASES)
def test_detect(golden, name):
    g = golden("detect")
    cfg = _cfg(g, name)
    m = g[name + "/map"]
    if cfg.get("use_nms", True) is True:
        for b in range(m.shape[0]):
            keep = O.nms_keep_mask(m[b, 0, 1:-1, 1:-1], cfg["nms_radius"])
            np.testing.assert_array_equal(keep, g[name + "/nms_mask"][b, 0] > 0)
    kps, sc, idx, counts = O.generate_kpts_single(m, return_idx=True, **cfg)
    check_detect_against(g, name, kps, sc, idx, counts)


EXT_CASES = ["soft_r1_mean", "soft_r2_abs", "soft_r4_max", "noavg_r1", "noavg_r2_mean", "noavg_r5_max",
             "noavg_nonms", "noavg_soft"]


@pytest.mark.parametrize("name", EXT_CASES)
def test_detect_ext(golden, name):
    """soft NMS and generate_kpts_single_noavg against the reference's outputs."""
    g = golden("detect_ext")
    cfg = _cfg(g, name)
    fn = O.generate_kpts_single_noavg if name.startswith("noavg") else O.generate_kpts_single
    kps, sc, idx, counts = fn(g[name + "/map"], return_idx=True, **cfg)
    chk = check_detect_float_key if cfg.get("use_nms", True) == "softnms" else check_detect
    chk(g[name + "/key"], g[name + "/idx"], g[name + "/kps"], g[name + "/score"], g[name + "/count"],
        kps, sc, idx, counts)


def test_linspace_matches_torch():
    import torch
    for n in (2, 3, 8, 13, 640, 898, 1200, 1600):
        np.testing.assert_array_equal(O.linspace_f32(-1, 1, n), torch.linspace(-1, 1, n).numpy())
    np.testing.assert_array_equal(O.linspace_f32(-0.1, 0.1, 12), torch.linspace(-0.1, 0.1, 12).numpy())


def test_sample(golden):
    g = golden("sample")
    raw = O.sample_feat_by_coord(g["x"], g["coord"], False)
    np.testing.assert_allclose(raw, g["raw"], rtol=1e-5, atol=1e-6)
    nr = O.sample_feat_by_coord(g["x"], g["coord"], True)
    np.testing.assert_allclose(nr, g["normed"], rtol=1e-5, atol=1e-6)
    n128 = O.sample_feat_by_coord(g["x128"], g["coord128"], True)
    np.testing.assert_allclose(n128, g["normed128"], rtol=1e-5, atol=1e-6)


def test_mnn(golden):
    g = golden("mnn")
    for exact in (False, True):
        np.testing.assert_array_equal(O.mnn_matcher(g["a"], g["b"], exact), g["mnn"])
    np.testing.assert_array_equal(g["mnn"], g["mutual_nn"])
    np.testing.assert_array_equal(O.mnn_matcher(g["ad"], g["bd"], True), g["mnn_dup"])
    np.testing.assert_array_equal(O.mnn_matcher(g["bd"], g["ad"], True), g["mnn_dup_t"])
    np.testing.assert_array_equal(O.ratio_matchers(g["a"], g["b"], 0.95, mutual=False), g["ratio"])
    np.testing.assert_array_equal(O.ratio_matchers(g["a"], g["b"], 0.9, mutual=True), g["mutual_ratio"])


def test_mnn_edge():
    a = np.eye(4, 8, dtype=np.float32)
    m = O.mnn_matcher(a, a[::-1].copy())
    np.testing.assert_array_equal(m, np.array([[0, 3], [1, 2], [2, 1], [3, 0]]))
    assert O.mnn_matcher(a, a[:1]).tolist() == [[0, 0]]


def test_mnn_blocked_f64_equals_plain_oracle(golden):
    """The row-blocked float64 matcher used at 16k / 64k (BASELINE config 5) is the plain oracle, block size
    notwithstanding, including the first-index rule on duplicated descriptors and the reference's fixtures."""
    g = golden("mnn")
    for a, b in ((g["a"], g["b"]), (g["ad"], g["bd"]), (g["bd"], g["ad"])):
        m, nn12, nn21 = O.mnn_matcher(a, b, exact=True, return_nn=True)
        sim = a.astype(np.float64) @ b.astype(np.float64).T
        for blk in (7, 64, 4096):
            m2, n12, n21, rg, cg = O.mnn_blocked_f64(a, b, block=blk)
            np.testing.assert_array_equal(m2, m)
            np.testing.assert_array_equal(n12, nn12)
            np.testing.assert_array_equal(n21, nn21)
            ss = np.sort(sim, 1)
            np.testing.assert_allclose(rg, ss[:, -1] - ss[:, -2], rtol=0, atol=1e-15)
            sc = np.sort(sim, 0)
            np.testing.assert_allclose(cg, sc[-1] - sc[-2], rtol=0, atol=1e-15)
    np.testing.assert_array_equal(O.mnn_blocked_f64(g["ad"], g["bd"])[0], g["mnn_dup"])


def test_corr(golden):
    g = golden("corr")
    e, std, prob = O.get_expected_correspondence_locs(g["f1"], g["fm"], with_std=True)
    np.testing.assert_allclose(e, g["exp"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(std, g["std"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(prob, g["prob"], rtol=2e-5, atol=1e-8)
    ew, cg, stdw, probw = O.get_expected_correspondence_within_window(g["f1w"], g["fmw"], g["c2"], 0.1)
    np.testing.assert_allclose(cg, g["cgw"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(probw, g["probw"], rtol=5e-5, atol=1e-7)
    np.testing.assert_allclose(ew, g["expw"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(stdw, g["stdw"], rtol=1e-4, atol=5e-6)


def test_compute_prob(golden):
    """compute_prob (losses/preprocess_utils.py:89-115), every option, against the reference's output."""
    g = golden("prob")
    for name, kw in (("cos", {}), ("cos_scale", dict(with_scale=True)), ("euc", dict(loss_distance="euc"))):
        p = O.compute_prob(g["f1"], g["f2"], **kw)
        np.testing.assert_allclose(p, g[f"{name}/prob"], rtol=3e-5, atol=1e-9)
    p, sim = O.compute_prob(g["f1"], g["f2"], return_sim=True)
    np.testing.assert_allclose(p, g["sim/prob"], rtol=3e-5, atol=1e-9)
    np.testing.assert_allclose(sim, g["sim/sim"], rtol=1e-5, atol=1e-6)


def test_grid_stage(golden):
    g = golden("preprocess")
    H, W = int(g["H"]), int(g["W"])
    c1n = g["coord1_n"].reshape(2, -1, 2)
    c2n = g["coord2_n"].reshape(2, -1, 2)
    f1 = O.sample_feat_by_coord(g["xf1"], c1n, True)
    f2 = O.sample_feat_by_coord(g["xf2"], c2n, True)
    c1 = O.denormalize_coords(c1n, H, W)
    c2 = O.denormalize_coords(c2n, H, W)
    np.testing.assert_allclose(c1, g["p_coord1"], rtol=1e-6, atol=1e-4)
    l1, l2, s1, s2 = O.grid_softmax_expectation(f1, f2, c1, c2, c1n, c2n, 60.0, H, W, H, W,
                                                dtype=np.float64)
    np.testing.assert_allclose(l1, g["p_feat1g_corloc"], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(l2, g["p_feat2g_corloc"], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(s1, g["p_feat1g_std"], rtol=2e-3, atol=1e-4)
    np.testing.assert_allclose(s2, g["p_feat2g_std"], rtol=2e-3, atol=1e-4)

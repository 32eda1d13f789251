"""GPU parity: score-map keypoint selection vs the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch

from _checks import check_detect, check_detect_float_key
from oracle import posfeat_oracle as O

pytestmark = pytest.mark.gpu

DETECT_CASES = ["r1_abs", "r3_abs", "r2_max", "r1_mean", "r1_nothr", "nonms_abs",
                "ties_r1", "ties_r2", "few", "const", "odd_r1", "odd_r5"]


def run_gpu(m, **cfg):
    import posfeat_b200 as P
    kps, sc, idx, counts = P.generate_kpts_single(torch.from_numpy(m).cuda(), return_idx=True, **cfg)
    return kps.cpu().numpy(), sc.cpu().numpy(), idx.cpu().numpy(), counts.cpu().numpy()


@pytest.mark.parametrize("name", DETECT_CASES)
def test_detect_golden(golden, name):
    g = golden("detect")
    cfg = eval(str(g[name + "/cfg"]))
    kps, sc, idx, counts = run_gpu(g[name + "/map"], **cfg)
    assert sc.shape[-1] == 1
    check_detect(g[name + "/key"], g[name + "/idx"], g[name + "/kps"], g[name + "/score"],
                 g[name + "/count"], kps, sc, idx, counts)


EXT_CASES = ["soft_r1_mean", "soft_r2_abs", "soft_r4_max", "noavg_r1", "noavg_r2_mean", "noavg_r5_max",
             "noavg_nonms", "noavg_soft"]


def run_gpu_noavg(m, **cfg):
    from posfeat_b200.preprocess_utils import generate_kpts_single_noavg
    kps, sc, idx, counts = generate_kpts_single_noavg(torch.from_numpy(m).cuda(), return_idx=True, **cfg)
    return kps.cpu().numpy(), sc.cpu().numpy(), idx.cpu().numpy(), counts.cpu().numpy()


@pytest.mark.parametrize("name", EXT_CASES)
def test_detect_ext_golden(golden, name):
    """use_nms='softnms' and generate_kpts_single_noavg against reference outputs."""
    g = golden("detect_ext")
    cfg = eval(str(g[name + "/cfg"]))
    run = run_gpu_noavg if name.startswith("noavg") else run_gpu
    kps, sc, idx, counts = run(g[name + "/map"], **cfg)
    chk = check_detect_float_key if cfg.get("use_nms", True) == "softnms" else check_detect
    chk(g[name + "/key"], g[name + "/idx"], g[name + "/kps"], g[name + "/score"], g[name + "/count"],
        kps, sc, idx, counts)


@pytest.mark.parametrize("noavg", [False, True])
@pytest.mark.parametrize("shape,cfg", [
    ((2, 480, 640), dict(nms_radius=1, num_pts=4096, use_nms="softnms", thr=1.0, thr_mod="mean")),
    ((1, 896, 1200), dict(nms_radius=2, num_pts=8192, use_nms="softnms", thr=0.3, thr_mod="max")),
    ((1, 300, 400), dict(nms_radius=1, num_pts=5000)),
    ((2, 131, 77), dict(nms_radius=4, num_pts=False, thr=1.2, thr_mod="mean")),
])
def test_detect_ext_vs_oracle(noavg, shape, cfg):
    m = softplus_map(*shape, seed=4321 + shape[1])
    soft = cfg.get("use_nms", True) == "softnms"
    if not noavg and not soft:
        pytest.skip("hard NMS through generate_kpts_single is covered by test_detect_vs_oracle")
    ofn = O.generate_kpts_single_noavg if noavg else O.generate_kpts_single
    okps, osc, oidx, ocnt = ofn(m, return_idx=True, **cfg)
    src = [np.pad(m[b, 0], 1) if noavg else m[b, 0] for b in range(m.shape[0])]
    keys = np.stack([O.detect_keys(x, cfg["nms_radius"], cfg.get("use_nms", True), cfg.get("thr", False),
                                   cfg.get("thr_mod", "mean"))[0].reshape(-1) for x in src])
    kps, sc, idx, counts = (run_gpu_noavg if noavg else run_gpu)(m, **cfg)
    (check_detect_float_key if soft else check_detect)(keys, oidx, okps, osc, ocnt, kps, sc, idx, counts)


def test_detect_ext_errors():
    from posfeat_b200.preprocess_utils import generate_kpts_single, generate_kpts_single_noavg
    m = torch.rand(1, 1, 40, 40, device="cuda")
    with pytest.raises(UnboundLocalError):          # thr_mask undefined in the reference without thr
        generate_kpts_single(m, 1, 200, use_nms="softnms")
    with pytest.raises(UnboundLocalError):          # noavg has no 'abs' mode
        generate_kpts_single_noavg(m, 1, 200, thr=0.5, thr_mod="abs")


def softplus_map(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.softplus(torch.randn(b, 1, h, w, generator=g)).numpy()


@pytest.mark.parametrize("shape,cfg", [
    ((2, 480, 640), dict(nms_radius=1, num_pts=4096, thr=0.9, thr_mod="abs")),      # C1
    ((1, 896, 1200), dict(nms_radius=1, num_pts=8192, thr=0.9, thr_mod="abs")),     # C2
    ((1, 1056, 1600), dict(nms_radius=3, num_pts=20480, thr=0.5, thr_mod="abs")),   # C3 (shipped aachen yaml)
    ((3, 200, 333), dict(nms_radius=2, num_pts=False, thr=0.5, thr_mod="max")),
    ((1, 300, 400), dict(nms_radius=1, num_pts=False)),                             # all survivors
    ((2, 131, 77), dict(nms_radius=4, num_pts=700, use_nms=True)),
    ((1, 240, 320), dict(nms_radius=1, num_pts=30000, use_nms=False, thr=1.0, thr_mod="abs")),
])
def test_detect_vs_oracle(shape, cfg):
    m = softplus_map(*shape, seed=1234 + shape[1])
    okps, osc, oidx, ocnt = O.generate_kpts_single(m, return_idx=True, **cfg)
    keys = np.stack([O.detect_keys(m[b, 0], cfg["nms_radius"], cfg.get("use_nms", True),
                                   cfg.get("thr", False), cfg.get("thr_mod", "mean"))[0].reshape(-1)
                     for b in range(m.shape[0])])
    kps, sc, idx, counts = run_gpu(m, **cfg)
    check_detect(keys, oidx, okps, osc, ocnt, kps, sc, idx, counts)
    # with distinct scores the whole index list is identical
    for b in range(m.shape[0]):
        v = keys[b][oidx[b]]
        if len(np.unique(v)) == len(v):
            np.testing.assert_array_equal(idx[b], oidx[b])


def test_detect_properties_full_size():
    """Size-independent properties at the HPatches shape: sorted, unique, every
    pick is a strict local first-maximum above the threshold."""
    m = softplus_map(2, 896, 1200, seed=77)
    kps, sc, idx, counts = run_gpu(m, nms_radius=1, num_pts=8192, thr=0.9, thr_mod="abs")
    assert kps.shape == (2, 8192, 2)
    for b in range(2):
        inter = m[b, 0, 1:-1, 1:-1]
        v = inter.reshape(-1)[idx[b]]
        assert np.all(np.diff(v) <= 0)
        assert len(np.unique(idx[b])) == 8192
        keep = O.nms_keep_mask(inter, 1) & (inter > np.float32(0.9))
        assert keep.reshape(-1)[idx[b]].all()
        assert counts[b] == keep.sum()
        # nothing better was left behind
        rest = inter.reshape(-1).copy()
        rest[~keep.reshape(-1)] = 0
        rest[idx[b]] = 0
        assert rest.max() <= v[-1]
        assert np.all(np.abs(kps[b]) <= 1.0)


def test_detect_errors():
    import posfeat_b200 as P
    x = torch.ones(1, 1, 32, 32, device="cuda")
    with pytest.raises(NotImplementedError):
        P.generate_kpts_single(x, 1, stable=False)
    k, s = P.generate_kpts_single(x, 1, use_nms="softnms", thr=0.5)       # every pixel passes 0.5 * mean
    assert k.shape == (1, 900, 2)
    with pytest.raises(Exception):
        P.generate_kpts_single(torch.ones(1, 1, 8, 8, device="cuda"), 1)   # 128 > 36 interior pixels
    # cpu tensors are accepted and results come back on the cpu
    k, s = P.generate_kpts_single(torch.rand(1, 1, 40, 40), 1, num_pts=130)
    assert k.device.type == "cpu" and k.shape == (1, 130, 2) and s.shape == (1, 130, 1)


def test_detect_noncontiguous_rows():
    import posfeat_b200 as P
    big = torch.from_numpy(softplus_map(2, 64, 100, seed=5)).cuda()
    view = big[:, :, :, :80]                       # row stride 100, width 80
    a = P.generate_kpts_single(view, 1, num_pts=200, return_idx=True)
    b = P.generate_kpts_single(view.contiguous(), 1, num_pts=200, return_idx=True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)

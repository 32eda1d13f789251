"""GPU parity: the fused pair pipeline, the batched matcher entry and the
Extractor.process / save_desc boundary, against the oracle."""
import os

import numpy as np
import pytest
import torch

from _checks import assert_close_vec, check_mnn_near_tie
from oracle import posfeat_oracle as O

pytestmark = pytest.mark.gpu
CFG = dict(nms_radius=1, num_pts=700, thr=0.9, thr_mod="abs", use_nms=True, stable=True)


def small_pairs(P, seed=3, H=160, W=208, D=128):
    g = torch.Generator().manual_seed(seed)
    score = torch.nn.functional.softplus(torch.randn(2 * P, 1, H, W, generator=g))
    fmap = torch.randn(2 * P, D, H // 4, W // 4, generator=g)
    fmap[1::2] = fmap[0::2] + 0.4 * torch.randn(P, D, H // 4, W // 4, generator=g)
    score[1::2] = score[0::2]
    return score, fmap


@pytest.mark.parametrize("algo", [1, 2])
def test_pair_pipeline_vs_oracle(algo):
    from posfeat_b200.pairs import PairPipeline
    P = 3
    score, fmap = small_pairs(P)
    pipe = PairPipeline(CFG, mnn_algo=algo)
    feats, matches, nm = pipe.run(score.cuda(), fmap.cuda())
    cfg = {k: v for k, v in CFG.items() if k != "stable"}
    okps, osc, oidx, _ = O.generate_kpts_single(score.numpy(), return_idx=True, **cfg)
    np.testing.assert_array_equal(feats["idx"].cpu().numpy(), oidx)          # keypoint indices bit exact
    np.testing.assert_allclose(feats["kps_n"].cpu().numpy(), okps, rtol=1e-5, atol=2e-6)
    odesc = O.sample_feat_by_coord(fmap.numpy(), okps, True)
    assert_close_vec(feats["desc"].cpu().numpy(), odesc, 1e-5)
    np.testing.assert_allclose(feats["kpt"].cpu().numpy(), O.denormalize_coords(okps, 160, 208), rtol=1e-6, atol=1e-4)
    for i in range(P):
        want = O.mnn_matcher(odesc[2 * i], odesc[2 * i + 1], exact=True)
        got = matches[i, :int(nm[i])].cpu().numpy()
        check_mnn_near_tie(odesc[2 * i], odesc[2 * i + 1], got, want)
        assert len(got) > 100


@pytest.mark.parametrize("short", [False, True])
def test_pair_pipeline_streams_equal_plain_path(short):
    """The multi-stream variant (n kept on the device, groups of pairs on separate streams) returns what the
    plain path returns; when an image has fewer survivors than num_pts it falls back to the plain path."""
    from posfeat_b200.pairs import PairPipeline
    score, fmap = small_pairs(5, seed=3)
    cfg = dict(CFG)
    if short:
        cfg["num_pts"] = 6000                      # more than the 160x208 maps can supply -> n < num_pts
    plain = PairPipeline(cfg, streams=1)
    multi = PairPipeline(cfg, streams=2)
    f1, m1, n1 = plain.run(score.cuda(), fmap.cuda())
    f2, m2, n2 = multi.run(score.cuda(), fmap.cuda())
    assert f1["n"] == f2["n"] and (f1["n"] < cfg["num_pts"]) == short
    assert torch.equal(f1["idx"], f2["idx"]) and torch.equal(f1["desc"], f2["desc"]) and torch.equal(n1, n2)
    for i in range(5):
        assert torch.equal(m1[i, :int(n1[i])], m2[i, :int(n2[i])])


def test_fused_sampler_operands_equal_separate_pass():
    """channels-last maps + tensor-core matcher: the sampler writes the matcher's bf16 operands and norms
    itself (posfeat_sample_pairs_f32).  Results must equal the path where the matcher rounds the
    descriptors in its own pass (NCHW map -> plain sampler)."""
    from posfeat_b200 import _lib
    from posfeat_b200.pairs import PairPipeline
    score, fmap = small_pairs(3, seed=11, H=320, W=416)
    cfg = dict(CFG, num_pts=1500)
    pipe = PairPipeline(cfg, mnn_algo=_lib.MNN_TC)
    f_cl = fmap.cuda().contiguous(memory_format=torch.channels_last)
    kps = pipe.extract(score.cuda(), f_cl)["kps_n"].contiguous()
    _, prepared = pipe.sample_for_pairs(f_cl, kps)
    assert prepared
    _, prepared = pipe.sample_for_pairs(fmap.cuda(), kps)
    assert not prepared
    fa, ma, na = pipe.run(score.cuda(), f_cl)
    fb, mb, nb = pipe.run(score.cuda(), fmap.cuda())
    assert torch.equal(fa["idx"], fb["idx"]) and torch.equal(na, nb)
    np.testing.assert_allclose(fa["desc"].cpu().numpy(), fb["desc"].cpu().numpy(), rtol=0, atol=1e-6)
    for i in range(3):
        a_, b_ = fa["desc"][2 * i].cpu().numpy(), fa["desc"][2 * i + 1].cpu().numpy()
        want = O.mnn_matcher(a_, b_, exact=True)
        check_mnn_near_tie(a_, b_, ma[i, :int(na[i])].cpu().numpy(), want)
        assert int(na[i]) > 300


def test_pipeline_host_entry_matches_device_entry():
    from posfeat_b200.pairs import PairPipeline
    score, fmap = small_pairs(2, seed=9)
    pipe = PairPipeline(CFG)
    feats, matches, nm = pipe.run(score.cuda(), fmap.cuda())
    kpt_h, m_h, nm_h = pipe.run_host(score.pin_memory(), fmap.pin_memory())
    assert kpt_h.device.type == "cpu"
    assert torch.equal(kpt_h, feats["kpt"].cpu()) and torch.equal(nm_h, nm.cpu())
    for i in range(2):
        assert torch.equal(m_h[i, :int(nm_h[i])], matches[i, :int(nm[i])].cpu())


@pytest.mark.parametrize("gather", ["stage", "direct", False])
def test_pipeline_host_entry_chunked_equals_unsplit(gather):
    """run_host cut into groups of pairs on separate streams (the default from 8 pairs on) returns exactly what the
    unsplit call returns; when an image has fewer survivors than num_pts (n couples the whole batch through
    min()) the call falls back to the unsplit form."""
    from posfeat_b200.pairs import PairPipeline
    score, fmap = small_pairs(9, seed=13)
    f_cl = fmap.contiguous(memory_format=torch.channels_last).pin_memory()
    sp = score.pin_memory()
    for cfg, short in ((CFG, False), (dict(CFG, num_pts=6000), True)):
        pipe = PairPipeline(cfg)
        one = pipe.run_host(sp, f_cl, gather=gather, chunks=1)
        for chunks in (None, 2, 3):
            got = pipe.run_host(sp, f_cl, gather=gather, chunks=chunks)
            assert all(t.is_pinned() for t in got)
            assert (got[0].shape[1] < cfg["num_pts"]) == short
            for a, b in zip(got, one):
                assert torch.equal(a, b)
        if gather == "stage" and not short:
            assert pipe.staged_pixels() > 0


def test_pipeline_host_gather_equals_full_copy():
    """run_host with a pinned channels-last descriptor map: either exactly the pixels under the keypoints'
    taps are staged on the device (posfeat_fetch_taps_f32, the default) or the sampler reads the taps over the
    host link (posfeat_host_device_pointer).  Same bits as copying the whole map first; pageable memory and
    non-channels-last maps are refused for an explicit gather and fall back to the copy by default."""
    from posfeat_b200 import _lib
    from posfeat_b200._runtime import map_ptr
    from posfeat_b200.pairs import PairPipeline
    from posfeat_b200.preprocess_utils import sample_l2norm
    score, fmap = small_pairs(3, seed=21, H=320, W=416)
    cfg = dict(CFG, num_pts=1500)
    f_cl = fmap.contiguous(memory_format=torch.channels_last).pin_memory()
    assert PairPipeline.host_gather_applies(f_cl) and not PairPipeline.host_gather_applies(fmap.pin_memory())
    assert not PairPipeline.host_gather_applies(fmap.contiguous(memory_format=torch.channels_last))
    for algo in (_lib.MNN_TC, _lib.MNN_SIMT):          # fused sampler+operands, and the plain sampler
        pipe = PairPipeline(cfg, mnn_algo=algo)
        runs = [pipe.run_host(score.pin_memory(), f_cl, gather=g) for g in ("stage", "direct", True, False, None)]
        for r in runs[1:]:
            for x, y in zip(runs[0], r):
                assert torch.equal(x, y)
        assert int(runs[0][2].min()) > 300
    with pytest.raises(ValueError):
        pipe.run_host(score.pin_memory(), fmap.pin_memory(), gather=True)
    with pytest.raises(ValueError):
        pipe.run_host(score.pin_memory(), f_cl, gather="some")
    with pytest.raises(_lib.PosfeatError):
        map_ptr(fmap)                                    # pageable host memory has no device alias
    # the kernel-level sampler takes the pinned map as well (descriptors bit-equal to the device-resident map)
    kps = (torch.rand(6, 777, 2, generator=torch.Generator().manual_seed(3)) * 2 - 1).cuda()
    assert torch.equal(sample_l2norm(f_cl, kps, True), sample_l2norm(f_cl.cuda(), kps, True))


@pytest.mark.parametrize("D,h,w,n", [(128, 40, 52, 300), (64, 9, 7, 50), (256, 17, 33, 1), (384, 5, 70, 999)])
def test_stage_taps_moves_exactly_the_sampled_pixels(D, h, w, n):
    """posfeat_fetch_taps_f32: the device map starts as NaN everywhere; after staging, sampling it equals
    sampling the whole map, every pixel that arrived is one the reference's bilinear taps touch, and the
    kernel's pixel count equals the number of pixels that arrived (coordinates include points outside the
    map, whose taps are partly or wholly dropped by the zeros padding)."""
    from posfeat_b200.pairs import PairPipeline
    from posfeat_b200.preprocess_utils import sample_l2norm
    g = torch.Generator().manual_seed(D + n)
    b = 4
    fmap = torch.randn(b, D, h, w, generator=g).contiguous(memory_format=torch.channels_last).pin_memory()
    kps = (torch.rand(b, n, 2, generator=g) * 2.3 - 1.15)
    kps[0, 0] = torch.tensor([-1.0, -1.0])
    kps[1, 0] = torch.tensor([1.0, 1.0])
    pipe = PairPipeline(CFG)
    f_dev = torch.full_like(fmap, float("nan"), device="cuda")
    assert f_dev.stride() == fmap.stride()
    pipe.stage_taps(fmap, f_dev, kps.cuda())
    moved = pipe.staged_pixels(fmap, f_dev.device)
    got = sample_l2norm(f_dev, kps.cuda(), True)
    want = sample_l2norm(fmap.cuda(), kps.cuda(), True)
    assert torch.equal(got, want)
    arrived = ~torch.isnan(f_dev[:, 0]).cpu()                      # [b,h,w]
    assert int(arrived.sum()) == moved
    assert torch.equal(f_dev.cpu()[arrived.unsqueeze(1).expand(-1, D, -1, -1)],
                       fmap[arrived.unsqueeze(1).expand(-1, D, -1, -1)])
    # expected cover from the reference's unnormalisation (align_corners=False), losses/preprocess_utils.py:48
    ix = ((kps[..., 0] + 1) * w - 1) / 2
    iy = ((kps[..., 1] + 1) * h - 1) / 2
    x0, y0 = torch.floor(ix).long(), torch.floor(iy).long()
    cover = torch.zeros(b, h, w, dtype=torch.bool)
    for dy in (0, 1):
        for dx in (0, 1):
            xx, yy = x0 + dx, y0 + dy
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            bi = torch.arange(b)[:, None].expand(-1, n)
            cover[bi[ok], yy[ok], xx[ok]] = True
    assert torch.equal(arrived, cover)
    with pytest.raises(ValueError):
        pipe.stage_taps(fmap, torch.empty(b, D, h, w, device="cuda"), kps.cuda())     # NCHW strides differ
    # no keypoints: nothing moves and the counter is reset
    f_dev.fill_(float("nan"))
    pipe.stage_taps(fmap, f_dev, kps[:, :0].cuda())
    assert pipe.staged_pixels(fmap, f_dev.device) == 0 and bool(torch.isnan(f_dev).all())


def test_batched_matcher_equals_single_calls():
    import posfeat_b200 as Pb
    from posfeat_b200 import _lib
    from posfeat_b200._runtime import check, lib, stream_ptr, workspace
    L = lib()
    g = torch.Generator().manual_seed(5)
    P, N, M, D = 3, 700, 900, 128
    A = torch.nn.functional.normalize(torch.randn(P, N, D, generator=g), dim=-1).cuda()
    B = torch.nn.functional.normalize(torch.randn(P, M, D, generator=g), dim=-1).cuda()
    for algo in (1, 2):
        nn12 = torch.empty((P, N), dtype=torch.int32, device="cuda")
        nn21 = torch.empty((P, M), dtype=torch.int32, device="cuda")
        matches = torch.empty((P, N, 2), dtype=torch.int64, device="cuda")
        nm = torch.empty(P, dtype=torch.int32, device="cuda")
        ws = workspace("mnn", L.posfeat_mnn_batched_workspace_bytes(P, N, M, D, algo), A.device)
        check(L.posfeat_mnn_batched_f32(A.data_ptr(), A.stride(0), N, A.stride(1), B.data_ptr(), B.stride(0), M,
                                        B.stride(1), D, P, algo, nn12.data_ptr(), nn21.data_ptr(),
                                        matches.data_ptr(), nm.data_ptr(), ws.data_ptr(), ws.numel(),
                                        stream_ptr(A.device)))
        for i in range(P):
            m1, n1, a12, a21 = Pb.mnn_match(A[i], B[i], algo=algo)
            assert int(n1) == int(nm[i])
            assert torch.equal(m1[:int(n1)], matches[i, :int(nm[i])])
            assert torch.equal(a12, nn12[i]) and torch.equal(a21, nn21[i])
            want = O.mnn_matcher(A[i].cpu().numpy(), B[i].cpu().numpy(), exact=True)
            check_mnn_near_tie(A[i].cpu().numpy(), B[i].cpu().numpy(), m1[:int(n1)].cpu().numpy(), want)


def test_extractor_process_and_npz(tmp_path):
    """managers/extractor.py:318-355 + :254-271: dict contract and .npz layout."""
    from posfeat_b200.extractor import FeatureExtractor
    g = torch.Generator().manual_seed(21)
    H, W = 96, 128
    outputs = {"local_point": torch.nn.functional.softplus(torch.randn(1, 1, H, W, generator=g)).cuda(),
               "local_map": torch.randn(1, 128, H // 4, W // 4, generator=g).cuda()}
    inputs = {"im1": torch.zeros(1, 3, H, W), "name1": ["v_seq/1.ppm"]}
    config = {"detector": "generate_kpts_single", "loss_distance": "cos", "postfix": "PoSFeat_test",
              "detector_config": dict(num_pts=300, stable=True, use_nms=True, nms_radius=1, thr=0.9, thr_mod="abs")}
    ex = FeatureExtractor(config, desc_root=str(tmp_path / "desc"))
    pr = ex.process(inputs, outputs)
    assert pr["kpt"].shape == (300, 2) and pr["kpt"].dtype == np.float32
    assert tuple(pr["desc"].shape) == (1, 300, 128) and tuple(pr["kp_score"].shape) == (1, 300, 1)
    okps, osc = O.generate_kpts_single(outputs["local_point"].cpu().numpy(), nms_radius=1, num_pts=300, thr=0.9,
                                       thr_mod="abs")
    np.testing.assert_allclose(pr["kpt"], O.denormalize_coords(okps, H, W)[0], rtol=1e-6, atol=1e-4)
    ex.save_desc(inputs, outputs, pr)
    path = os.path.join(str(tmp_path / "desc"), "v_seq", "1.ppm.PoSFeat_test")
    assert os.path.exists(path)                                  # no '.npz' suffix, like the reference
    z = np.load(path)
    assert set(z.files) == {"keypoints", "scores", "descriptors"}
    assert z["keypoints"].shape == (300, 2) and z["scores"].shape == (300, 1) and z["descriptors"].shape == (300, 128)
    # the evaluation reader's path: np.load -> mnn_matcher on the descriptors
    import posfeat_b200 as Pb
    m = Pb.mnn_matcher(torch.from_numpy(z["descriptors"]), torch.from_numpy(z["descriptors"]))
    np.testing.assert_array_equal(m, np.stack([np.arange(300)] * 2, -1))


def test_async_writer_with_device_tensors(tmp_path):
    """AsyncDescWriter: D2H through pinned staging on a side stream while the GPU keeps working."""
    from posfeat_b200.extractor import AsyncDescWriter
    root = str(tmp_path / "desc")
    want = {}
    with AsyncDescWriter(root, "PoSFeat_t", workers=2) as wr:
        for i in range(8):
            desc = torch.randn(1, 500, 128, device="cuda")
            sc = torch.rand(1, 500, 1, device="cuda")
            kpt = np.random.rand(500, 2).astype(np.float32)
            wr.save({"name1": [f"s/{i}.ppm"]}, {"kpt": kpt, "desc": desc, "kp_score": sc})
            want[i] = (kpt.copy(), sc[0].cpu().numpy(), desc[0].cpu().numpy())
            desc.zero_()                        # later work on the same stream must not leak into the file
            sc.zero_()
    for i, (k, s_, d) in want.items():
        z = np.load(os.path.join(root, "s", f"{i}.ppm.PoSFeat_t"))
        np.testing.assert_array_equal(z["keypoints"], k)
        np.testing.assert_array_equal(z["scores"], s_)
        np.testing.assert_array_equal(z["descriptors"], d)


def test_install_patches_reference_style_module():
    import types
    import posfeat_b200 as Pb
    fake = types.SimpleNamespace(generate_kpts_single=None, sample_feat_by_coord=None, mnn_matcher=None)
    Pb.install(fake)
    det = getattr(fake, "generate_kpts_single")                 # managers/extractor.py:87 style dispatch
    k, s = det(torch.rand(1, 1, 40, 48).cuda() + 0.1, nms_radius=1, num_pts=140)
    assert tuple(k.shape) == (1, 140, 2) and tuple(s.shape) == (1, 140, 1)


def test_matching_driver_on_device(tmp_path):
    """The feature-file driver with the CUDA matcher (its default): descriptors are cached on the device, results
    equal the oracle's per pair and the HPatches score equals the single-process reference loop."""
    import test_matching_driver as T
    from posfeat_b200 import matching_driver as MD
    root = str(tmp_path)
    T._write_features(root)
    res = MD.hpatches_benchmark(T.SEQS, root, T.METHOD, T._homography)
    ri, rv, rtype, rfeats, rmatches = T._reference_loop(root)
    for t in ri:
        assert abs(res[0][t] - ri[t]) < 1e-12 and abs(res[1][t] - rv[t]) < 1e-12
    np.testing.assert_array_equal(res[2][2], rmatches)
    dev = MD.hpatches_benchmark(T.SEQS, root, T.METHOD, T._homography, score_on_device=True)     # matches never leave the device
    for t in ri:
        assert abs(dev[0][t] - ri[t]) < 1e-12 and abs(dev[1][t] - rv[t]) < 1e-12
    np.testing.assert_array_equal(dev[2][2], rmatches)
    pairs = MD.hpatches_pairs(T.SEQS[:2])
    got = MD.match_pairs(pairs, root, T.METHOD)
    for (a, b), m in got.items():
        za, zb = np.load(os.path.join(root, f"{a}.{T.METHOD}")), np.load(os.path.join(root, f"{b}.{T.METHOD}"))
        np.testing.assert_array_equal(m, O.mnn_matcher(za["descriptors"], zb["descriptors"], exact=True))


def test_graphed_pipeline_equals_plain_call():
    """The CUDA-graph form of one pipeline call (GraphedPairPipeline) returns the plain call's results, call after
    call with different inputs, and falls back to the plain path when an image has fewer survivors than num_pts."""
    from posfeat_b200.pairs import GraphedPairPipeline, PairPipeline
    cfg = dict(CFG, num_pts=1500)
    pipe = PairPipeline(cfg)
    score, fmap = small_pairs(2, seed=31, H=320, W=416)
    f_cl = fmap.cuda().contiguous(memory_format=torch.channels_last)
    g = GraphedPairPipeline(pipe, score.shape, f_cl.shape)
    for seed in (31, 32, 33):
        score, fmap = small_pairs(2, seed=seed, H=320, W=416)
        f_cl = fmap.cuda().contiguous(memory_format=torch.channels_last)
        fa, ma, na = pipe.run(score.cuda(), f_cl)
        fb, mb, nb = g(score.cuda(), f_cl)
        assert fb["n"] == fa["n"] == 1500
        assert torch.equal(fa["idx"], fb["idx"]) and torch.equal(fa["desc"], fb["desc"]) and torch.equal(na, nb)
        assert torch.equal(fa["kpt"], fb["kpt"])
        for i in range(2):
            assert torch.equal(ma[i, :int(na[i])], mb[i, :int(nb[i])])
    sparse = torch.full_like(score, 0.1)
    sparse[:, :, 10:300:7, 10:400:9] = 2.0                    # ~1800 isolated maxima in image... fewer than num_pts after NMS in one image
    sparse[1, :, 10:300:7, 10:400:9] = 0.2
    sparse[1, :, 10:100:7, 10:100:9] = 2.0
    fa, ma, na = pipe.run(sparse.cuda(), f_cl)
    fb, mb, nb = g(sparse.cuda(), f_cl)
    assert fa["n"] < 1500 and fb["n"] == fa["n"]
    assert torch.equal(fa["idx"], fb["idx"]) and torch.equal(na, nb)


def test_graphed_matcher_equals_plain_call():
    """GraphedMatcher (the five launches of a matches-only call replayed as one CUDA graph) returns mnn_matcher's
    list, call after call with different descriptors, for device and host inputs; a wrong shape is refused."""
    import posfeat_b200 as P
    gm = P.GraphedMatcher(1500, 1300)
    for seed in (1, 2, 3):
        g = torch.Generator().manual_seed(seed)
        a = torch.nn.functional.normalize(torch.randn(1500, 128, generator=g), dim=1)
        b = torch.nn.functional.normalize(a[torch.randperm(1500, generator=g)[:1300]] + 0.3 * torch.randn(1300, 128, generator=g), dim=1)
        want = O.mnn_matcher(a.numpy(), b.numpy(), exact=True)
        np.testing.assert_array_equal(gm(a.cuda(), b.cuda()), want)
        np.testing.assert_array_equal(gm(a, b), want)                     # host tensors: copied into the static buffers
        np.testing.assert_array_equal(P.mnn_matcher(a.cuda(), b.cuda()), want)
    with pytest.raises(ValueError):
        gm(a[:100].cuda(), b.cuda())


def test_graph_objects_own_their_scratch_memory():
    """A captured graph has the addresses of its scratch memory baked in: every graph object captures inside its own
    workspace scope, so a later, larger call on the shared grow-only cache (or release_workspaces) cannot pull the
    memory from under it, and two live graphs never share scratch."""
    import posfeat_b200 as P
    from posfeat_b200 import _runtime
    g = torch.Generator().manual_seed(9)
    mk = lambda n: torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
    a1, b1, a2, b2 = mk(1500), mk(1300), mk(3000), mk(2500)
    b1[:1000] = torch.nn.functional.normalize(a1[:1000] + 0.2 * torch.randn(1000, 128, generator=g), dim=1)
    b2[:2000] = torch.nn.functional.normalize(a2[:2000] + 0.2 * torch.randn(2000, 128, generator=g), dim=1)
    w1, w2 = O.mnn_matcher(a1.numpy(), b1.numpy(), exact=True), O.mnn_matcher(a2.numpy(), b2.numpy(), exact=True)
    g1 = P.GraphedMatcher(1500, 1300)
    g2 = P.GraphedMatcher(3000, 2500)                      # larger: would have replaced a shared buffer
    np.testing.assert_array_equal(g1(a1.cuda(), b1.cuda()), w1)
    np.testing.assert_array_equal(g2(a2.cuda(), b2.cuda()), w2)
    big = mk(9000)
    P.mnn_matcher(big.cuda(), big.cuda())                  # grows the shared cache
    _runtime.release_workspaces()
    torch.cuda.empty_cache()
    junk = torch.full((64 << 20,), 0x7f, dtype=torch.uint8, device="cuda")     # reuse whatever was freed
    np.testing.assert_array_equal(g1(a1.cuda(), b1.cuda()), w1)
    np.testing.assert_array_equal(g2(a2.cuda(), b2.cuda()), w2)
    keys = [k for k in _runtime._workspaces if k[3] is not None]
    assert len({k[3] for k in keys}) == 2                  # one private scope per graph object
    del g1, junk
    import gc
    gc.collect()
    assert len({k[3] for k in _runtime._workspaces if k[3] is not None}) == 1

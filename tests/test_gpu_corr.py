"""GPU parity: training-side correlation + softmax expectation (subsystem 4)."""
import numpy as np
import pytest
import torch

from _checks import assert_close_vec
import _torch_ref as R

pytestmark = pytest.mark.gpu


def t(x):
    return torch.from_numpy(np.asarray(x)).cuda()


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-30)


def test_dense_expectation_golden(golden):
    """get_expected_correspondence_locs vs the reference's own output (1e-5 relative)."""
    import posfeat_b200.preprocess as PP
    g = golden("corr")
    e, std, kurt, prob = PP.get_expected_correspondence_locs(t(g["f1"]), t(g["fm"]), with_std=True)
    assert rel_err(e.cpu(), g["exp"]) < 1e-5
    assert rel_err(std.cpu(), g["std"]) < 2e-5
    assert rel_err(prob.cpu(), g["prob"]) < 2e-5
    e2 = PP.get_expected_correspondence_locs(t(g["f1"]), t(g["fm"]))
    assert torch.equal(e2, e)


def test_compute_prob_golden(golden):
    """compute_prob (losses/preprocess_utils.py:89-115): probabilities, raw similarities and the gradients of
    a weighted sum of the probabilities w.r.t. both inputs, against the reference's own run (1e-5 relative to
    the largest entry)."""
    import posfeat_b200 as P
    g = golden("prob")
    wgt = t(g["wgt"])
    for name, kw in (("cos", {}), ("cos_scale", dict(with_scale=True)), ("euc", dict(loss_distance="euc"))):
        a, b = t(g["f1"]).requires_grad_(), t(g["f2"]).requires_grad_()
        prob = P.compute_prob(a, b, **kw)
        assert prob.shape == (2, 37, 150) and prob.dtype == torch.float32
        assert rel_err(prob.detach().cpu(), g[f"{name}/prob"]) < 1e-5
        np.testing.assert_allclose(prob.detach().sum(-1).cpu().numpy(), 1.0, rtol=0, atol=2e-6)
        (prob * wgt).sum().backward()
        assert rel_err(a.grad.cpu(), g[f"{name}/g1"]) < 2e-5
        assert rel_err(b.grad.cpu(), g[f"{name}/g2"]) < 2e-5
    prob, sim = P.compute_prob(t(g["f1"]), t(g["f2"]), return_sim=True)
    assert rel_err(prob.cpu(), g["sim/prob"]) < 1e-5
    assert rel_err(sim.cpu(), g["sim/sim"]) < 1e-5
    with pytest.raises(AssertionError):
        P.compute_prob(t(g["f1"]), t(g["f2"]), loss_distance="l1")
    with pytest.raises(AssertionError):
        P.compute_prob(t(g["f1"]), t(g["f2"]), loss_distance="euc", return_sim=True)


def test_compute_prob_vs_oracle_ragged():
    """Sizes that are not multiples of the 64 x 64 tile or the 16-wide K slice, against the numpy oracle."""
    import posfeat_b200 as P
    from oracle import posfeat_oracle as O
    g = torch.Generator().manual_seed(77)
    for (B, m, n, D) in ((1, 1, 1, 1), (3, 65, 130, 33), (2, 200, 9000, 128)):
        f1 = torch.randn(B, m, D, generator=g) * 0.4
        f2 = torch.randn(B, n, D, generator=g) * 0.4
        for kw in ({}, dict(with_scale=True), dict(loss_distance="euc")):
            if kw.get("with_scale") and n > 1000:
                continue                      # sqrt(n)-scaled logits of random data: a one-hot softmax, nothing to compare
            got = P.compute_prob(f1.cuda(), f2.cuda(), **kw).cpu().numpy()
            want = O.compute_prob(f1.numpy(), f2.numpy(), dtype=np.float64, **kw)
            assert rel_err(got, want) < 2e-5, (B, m, n, D, kw)


def test_window_expectation_golden(golden):
    import posfeat_b200.preprocess as PP
    g = golden("corr")
    for fm in (t(g["fmw"]), t(g["fmw"]).contiguous(memory_format=torch.channels_last)):
        e, cg, std, prob = PP.get_expected_correspondence_within_window(t(g["f1w"]), fm, t(g["c2"]), 0.1, with_std=True)
        np.testing.assert_allclose(cg.cpu().numpy(), g["cgw"], rtol=0, atol=1e-7)
        assert rel_err(prob.cpu(), g["probw"]) < 2e-5
        assert rel_err(e.cpu(), g["expw"]) < 1e-5
        assert rel_err(std.cpu(), g["stdw"]) < 1e-4      # sqrt of a small variance amplifies rounding


@pytest.mark.parametrize("B,n,m,D,C,scale", [(2, 150, 170, 128, 4, 60.0), (1, 64, 64, 32, 2, 1.0),
                                              (3, 33, 257, 100, 3, 10.0), (8, 512, 1200, 128, 4, 60.0),
                                              # >= 2^23 logits: backward on the tensor cores as well
                                              (2, 300, 15000, 128, 4, 30.0), (1, 700, 13000, 96, 3, 10.0)])
def test_dense_forward_backward_vs_torch(B, n, m, D, C, scale):
    import posfeat_b200.preprocess as PP
    g = torch.Generator().manual_seed(n + m)
    q = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_()
    k = torch.nn.functional.normalize(torch.randn(B, m, D, generator=g), dim=-1).cuda().requires_grad_()
    v = (torch.rand(B, m, C, generator=g) * 100).cuda()
    go = torch.randn(B, n, C, generator=g).cuda()
    out = PP.corr_expect(q, k, v, scale)
    out.backward(go)
    gq, gk = q.grad.clone(), k.grad.clone()
    q64, k64 = q.detach().double().requires_grad_(), k.detach().double().requires_grad_()
    ref = R.corr_expect_ref(q64, k64, v.double(), scale)
    ref.backward(go.double())
    assert rel_err(out.detach().cpu(), ref.detach().cpu()) < 1e-5
    assert rel_err(gq.cpu(), q64.grad.cpu()) < 1e-4
    assert rel_err(gk.cpu(), k64.grad.cpu()) < 1e-4
    # a shared (unbatched) value table gives the same result
    if B > 1:
        out1 = PP.corr_expect(q.detach(), k.detach(), v[0], scale)
        refb = R.corr_expect_ref(q.detach().double(), k.detach().double(), v[0].double(), scale)
        assert rel_err(out1.cpu(), refb.cpu()) < 1e-5


@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
def test_window_forward_backward_vs_torch(layout):
    import posfeat_b200.preprocess as PP
    g = torch.Generator().manual_seed(17)
    B, n, D, h, w = 2, 90, 128, 30, 40
    fm = (6 * torch.nn.functional.normalize(torch.randn(B, D, h, w, generator=g), dim=1)).cuda()
    if layout == "channels_last":
        fm = fm.contiguous(memory_format=torch.channels_last)
    fm.requires_grad_()
    q = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_()
    c = (torch.rand(B, n, 2, generator=g) * 2.1 - 1.05).cuda()
    off = PP.gen_grid(-0.1, 0.1, -0.1, 0.1, int(0.1 * h), int(0.1 * w)).cuda()
    ge, gs = torch.randn(B, n, 2, generator=g).cuda(), torch.randn(B, n, generator=g).cuda()
    e, std, prob = PP.WindowExpect.apply(q, fm, c, off)
    (e * ge).sum().add((std * gs).sum()).backward()
    gq, gf = q.grad.clone(), fm.grad.clone()
    q64 = q.detach().double().requires_grad_()
    f64 = fm.detach().double().contiguous().requires_grad_()
    er, sr, pr = R.window_ref(q64, f64, c.double(), off.double())
    (er * ge.double()).sum().add((sr * gs.double()).sum()).backward()
    assert rel_err(e.detach().cpu(), er.detach().cpu()) < 1e-5
    assert rel_err(prob.cpu(), pr.detach().cpu()) < 2e-5
    assert rel_err(std.detach().cpu(), sr.detach().cpu()) < 1e-4
    assert rel_err(gq.cpu(), q64.grad.cpu()) < 2e-4
    assert rel_err(gf.cpu(), f64.grad.cpu()) < 2e-4


def test_sample_backward_vs_torch():
    import posfeat_b200 as P
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 128, 20, 24, generator=g).cuda().requires_grad_()
    c = (torch.rand(2, 300, 2, generator=g) * 2.06 - 1.03).cuda()
    go = torch.randn(2, 300, 128, generator=g).cuda()
    out = P.sample_feat_by_coord(x, c, True)
    out.backward(go)
    x64 = x.detach().double().requires_grad_()
    ref = torch.nn.functional.normalize(
        torch.nn.functional.grid_sample(x64, c.double().unsqueeze(2), padding_mode="zeros", align_corners=False).squeeze(-1),
        p=2, dim=1).transpose(1, 2)
    ref.backward(go.double())
    assert_close_vec(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), 1e-5)
    assert rel_err(x.grad.cpu(), x64.grad.cpu()) < 1e-4


def test_preprocess_line2window_golden(golden):
    """The whole Preprocess_Line2Window.forward (+ EpipolarLoss_full gradient) against
    a run of the real reference with the same coordinates and jitter."""
    import posfeat_b200.preprocess as PP
    g = golden("preprocess")
    H, W = int(g["H"]), int(g["W"])
    cfg = dict(kps_generator="generate_kpts_regular_grid_random",
               kps_generator_config=dict(grid_size=16, map_init="identity", keep_spatial=True, random_select="random"),
               window_size=0.1, loss_distance="cos", use_nn_grid=False, use_line_search=True,
               line_search_config=dict(line_step=100, use_nn=True, loc_rand=True),
               temperature_base=60, temperature_max=60)
    P = PP.Preprocess_Line2Window(cfg)
    xf1, xf2 = t(g["xf1"]).requires_grad_(), t(g["xf2"]).requires_grad_()
    B = xf1.shape[0]
    inputs = dict(im1=torch.zeros(B, 3, H, W), im2=torch.zeros(B, 3, H, W), F1=t(g["F1"]), F2=t(g["F2"]))
    outputs = dict(preds1=dict(local_map=xf1, local_point=torch.ones(B, 1, H, W).cuda()),
                   preds2=dict(local_map=xf2, local_point=torch.ones(B, 1, H, W).cuda()), epoch=0)
    pr = P(inputs, outputs, coords=(t(g["coord1_n"]), t(g["coord2_n"])), jitter=(t(g["jitter1"]), t(g["jitter2"])))
    assert set(pr) == {"coord1", "coord2", "feat1g_corloc", "feat2g_corloc", "feat1w_corloc", "feat2w_corloc",
                       "feat1c_corloc_org", "feat2c_corloc_org", "feat1g_std", "feat2g_std", "feat1w_std",
                       "feat2w_std", "temperature", "valid_epi1", "valid_epi2"}
    np.testing.assert_array_equal(pr["valid_epi1"].cpu().numpy(), g["p_valid_epi1"])
    np.testing.assert_array_equal(pr["valid_epi2"].cpu().numpy(), g["p_valid_epi2"])
    scale = float(max(H, W))
    for key, tol in (("coord1", 1e-6), ("feat1g_corloc", 2e-5), ("feat2g_corloc", 2e-5), ("feat1c_corloc_org", 1e-5),
                     ("feat2c_corloc_org", 1e-5), ("feat1w_corloc", 2e-5), ("feat2w_corloc", 2e-5)):
        got, want = pr[key].detach().cpu().numpy(), g["p_" + key]
        denom = scale if np.abs(want).max() > 2 else 1.0
        assert np.max(np.abs(got - want)) / denom < tol, key
    # std = sqrt(E[c^2] - E[c]^2) in float32 (reference :76-81, :749-750): the
    # subtraction cancels to ~1e-7 absolute, so sqrt() turns rounding noise into up to
    # ~5e-4 for sharply peaked windows -- in the reference as much as here.  Compare
    # absolutely, on rows whose epipolar line is valid (the others hold inf/nan-derived
    # garbage in the reference and are masked out of the loss).
    for key, vkey in (("feat1g_std", "valid_epi1"), ("feat2g_std", "valid_epi2"), ("feat1w_std", "valid_epi1"),
                      ("feat2w_std", "valid_epi2")):
        ok = g["p_" + vkey].astype(bool)
        got, want = pr[key].detach().cpu().numpy(), g["p_" + key]
        assert np.max(np.abs(got - want)[ok]) < 1e-3, key
        big = ok & (want > 0.02)
        assert np.max(np.abs(got - want)[big] / want[big]) < 5e-3, key
    # The loss weights are 1/std, detached (epipolarloss.py:25-36): feed the reference's
    # own std values so the comparison of the loss and of its gradient is not dominated
    # by that noise; everything differentiable comes from this implementation.
    pr2 = dict(pr)
    for key in ("feat1g_std", "feat2g_std", "feat1w_std", "feat2w_std"):
        pr2[key] = t(g["p_" + key])
    loss_cfg = dict(grid_cost_thr=0.5, win_cost_thr=0.1, use_std_as_weight=True, weight_grid=0.3, weight_window=1)
    loss = R.epipolar_loss_full(inputs, pr2, loss_cfg)
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-4
    loss.backward()
    assert rel_err(xf1.grad.cpu(), g["gxf1"]) < 2e-3
    assert rel_err(xf2.grad.cpu(), g["gxf2"]) < 2e-3


def test_preprocess_line2window_vs_float64_reference(golden, capsys):
    """How far is this implementation from the TRUTH, and how far is the reference's own float32 run?

    tests/golden/preprocess_f64.npz holds the reference run in float64 (same coordinates; the results of its
    no_grad line search replayed from the float32 run), and the float32 reference's loss / gradients evaluated
    with the float64 run's std values as loss weights.  For every output the deviation of this implementation
    from the float64 run must not exceed 3x the deviation of the reference's own float32 run from it (plus
    1e-5 of the output's scale, the north star's bar).  The measured pairs are printed."""
    import posfeat_b200.preprocess as PP
    g, g64 = golden("preprocess"), golden("preprocess_f64")
    H, W = int(g["H"]), int(g["W"])
    cfg = dict(kps_generator="generate_kpts_regular_grid_random",
               kps_generator_config=dict(grid_size=16, map_init="identity", keep_spatial=True, random_select="random"),
               window_size=0.1, loss_distance="cos", use_nn_grid=False, use_line_search=True,
               line_search_config=dict(line_step=100, use_nn=True, loc_rand=True),
               temperature_base=60, temperature_max=60)
    P = PP.Preprocess_Line2Window(cfg)
    xf1, xf2 = t(g["xf1"]).requires_grad_(), t(g["xf2"]).requires_grad_()
    B = xf1.shape[0]
    inputs = dict(im1=torch.zeros(B, 3, H, W), im2=torch.zeros(B, 3, H, W), F1=t(g["F1"]), F2=t(g["F2"]))
    outputs = dict(preds1=dict(local_map=xf1, local_point=torch.ones(B, 1, H, W).cuda()),
                   preds2=dict(local_map=xf2, local_point=torch.ones(B, 1, H, W).cuda()), epoch=0)
    pr = P(inputs, outputs, coords=(t(g["coord1_n"]), t(g["coord2_n"])), jitter=(t(g["jitter1"]), t(g["jitter2"])))
    rows = []

    def compare(name, ours, ref32, ref64, mask=None, scale=None, factor=3.0, floor=1e-5):
        ours, ref32, ref64 = (np.asarray(x, dtype=np.float64) for x in (ours, ref32, ref64))
        if mask is not None:
            ours, ref32, ref64 = ours[mask], ref32[mask], ref64[mask]
        sc = scale if scale is not None else max(np.abs(ref64).max(), 1e-30)
        e_ours, e_ref = np.abs(ours - ref64).max() / sc, np.abs(ref32 - ref64).max() / sc
        rows.append((name, e_ours, e_ref))
        assert e_ours <= factor * e_ref + floor, f"{name}: ours {e_ours:.3e} vs reference-float32 {e_ref:.3e} (of scale {sc:.3g})"

    v1, v2 = g["p_valid_epi1"].astype(bool), g["p_valid_epi2"].astype(bool)
    # the discrete choice of the line search must be the float32 reference's (it was replayed into the float64 run)
    for key, v in (("feat1c_corloc_org", v1), ("feat2c_corloc_org", v2)):
        got, want = pr[key].detach().cpu().numpy(), g["p_" + key]
        same = np.abs(got - want).max(-1) < 1e-4 * (max(H, W) if np.abs(want).max() > 2 else 1.0)
        assert same[v].mean() > 0.99, key
        v &= same
    for key, v in (("feat1g_corloc", None), ("feat2g_corloc", None), ("feat1w_corloc", v1), ("feat2w_corloc", v2)):
        compare(key, pr[key].detach().cpu().numpy(), g["p_" + key], g64["p_" + key], mask=v, scale=float(max(H, W)))
    # std = sum_xy sqrt(clamp(E[c^2] - E[c]^2)): the subtraction cancels to a few float32 ulps of E[c^2] <= 1, and the
    # square root amplifies that by 1 / (2 std); at the reference's clamps (1e-6 grid stage, 1e-10 window) a
    # variance error of 4 ulp (2.4e-7) per axis becomes up to 2.4e-4 in std.  Which of two float32 evaluations lands
    # closer to the float64 value there is luck, so the floor for std is that bound (5e-4 for the two axes), not 1e-5.
    for key, v in (("feat1g_std", None), ("feat2g_std", None), ("feat1w_std", v1), ("feat2w_std", v2)):
        compare(key, pr[key].detach().cpu().numpy(), g["p_" + key], g64["p_" + key], mask=v, scale=1.0, floor=5e-4)
    # loss and gradients with the float64 run's std values as the (detached) weights, as in the fixture
    pr2 = dict(pr)
    for key in ("feat1g_std", "feat2g_std", "feat1w_std", "feat2w_std"):
        pr2[key] = t(g64["p_" + key])
    loss_cfg = dict(grid_cost_thr=0.5, win_cost_thr=0.1, use_std_as_weight=True, weight_grid=0.3, weight_window=1)
    loss = R.epipolar_loss_full(inputs, pr2, loss_cfg)
    compare("loss", [float(loss)], [float(g64["loss_f32_std64"])], [float(g64["loss"])], floor=1e-6)
    loss.backward()
    compare("grad local_map 1", xf1.grad.cpu().numpy(), g64["gxf1_f32_std64"], g64["gxf1"])
    compare("grad local_map 2", xf2.grad.cpu().numpy(), g64["gxf2_f32_std64"], g64["gxf2"])
    with capsys.disabled():
        print("\n  deviation from the float64 reference run, relative to the output's scale:")
        for name, eo, er in rows:
            print(f"    {name:22s} this implementation {eo:9.2e}   reference's own float32 run {er:9.2e}")


def test_epipolar_line_search_vs_torch():
    """The fused kernel (endpoints + line samples + softmax + reductions) against the reference's sequence of
    tensor ops (losses/preprocess_utils.py:662-719) written with torch: get_endpoints, grid_sample with border
    padding, softmax, nearest-neighbour position, variance."""
    import torch.nn.functional as F
    import posfeat_b200.preprocess as PP
    g = torch.Generator().manual_seed(23)
    B, n, D, h2, w2, H, W, steps, ws = 2, 300, 128, 60, 80, 240, 320, 100, 0.125
    q = F.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda()
    fm = (20 * F.normalize(torch.randn(B, D, h2, w2, generator=g), dim=1)).cuda()
    coord = torch.stack([torch.rand(B, n, generator=g) * (W - 1), torch.rand(B, n, generator=g) * (H - 1)], -1).cuda()
    Fm = torch.randn(B, 3, 3, generator=g).cuda()
    Fm[:, 2, 2] *= 100                      # lines that cross the image for most points
    jit = torch.rand(B, n, 2, generator=g).cuda()
    exp_, org_, valid_, std_, prob_ = PP.epipolar_line_search(coord, Fm, q, fm, H, W, line_step=steps, window_size=ws,
                                                                jitter=jit, visualize=True)
    # torch restatement
    e1, e2, valid = PP.get_endpoints(coord, Fm, H, W)
    t = torch.linspace(0., 1., steps, device="cuda")
    grids = (e2 - e1)[:, :, None, :] * t[None, None, :, None] + e1[:, :, None, :]
    samp = F.grid_sample(fm, grids, padding_mode="border", align_corners=False).permute(0, 2, 3, 1)
    prob = torch.softmax((q.unsqueeze(-2) * samp).sum(-1), -1)
    mask = prob == prob.max(-1, True)[0]
    org = (mask.unsqueeze(-1) * grids).sum(2)
    exp = org + 0.707 * ws * (2 * jit - 1)
    border = (exp[..., 0] >= -1) & (exp[..., 0] <= 1) & (exp[..., 1] >= -1) & (exp[..., 1] <= 1)
    var = torch.sum(grids ** 2 * prob.unsqueeze(-1), dim=2) - exp ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    assert int(valid.sum()) > 0.5 * B * n
    same_valid = (valid_ == (valid & border))
    assert float(same_valid.float().mean()) > 0.995          # a border intersection may sit exactly on the rectangle
    ok = same_valid & valid
    assert rel_err(prob_[ok].cpu(), prob[ok].cpu()) < 2e-4    # logits of magnitude 20: exp amplifies 1e-6
    agree = ok & ((org_ - org).abs().max(-1)[0] < 1e-5)       # near-tied maxima may pick the neighbouring sample
    assert float(agree.float().sum() / ok.float().sum()) > 0.99
    np.testing.assert_allclose(exp_[agree].cpu().numpy(), exp[agree].cpu().numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(std_[agree].cpu().numpy(), std[agree].cpu().numpy(), rtol=1e-3, atol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 128, 30, 40), (3, 128, 17, 23), (1, 96, 5, 7), (2, 32, 1, 1)])
def test_normalize_scale_channels_last(shape):
    """The fused scale * F.normalize(x, dim=1) -> channels_last pass of Preprocess_Line2Window (csrc/normalize.cu)
    against the tensor expression it replaces: values to 2 ulp (the norm is summed in another order), gradient
    against a float64 run of the expression; a zero descriptor (norm below eps) takes the constant-denominator branch."""
    import torch.nn.functional as F
    from posfeat_b200.preprocess import normalize_scale_channels_last
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    x[0, :, 0, 0] = 0.0                                   # |x| < eps
    x[-1, :, -1, -1] *= 1e-3
    T = 60.0
    xc = x.cuda().requires_grad_(True)
    out = normalize_scale_channels_last(xc, T)
    assert out.is_contiguous(memory_format=torch.channels_last) and out.shape == xc.shape
    want = T * F.normalize(x.cuda(), p=2.0, dim=1)
    np.testing.assert_allclose(out.detach().cpu().numpy(), want.cpu().numpy(), rtol=3e-7, atol=1e-30)
    w = torch.randn(*shape, generator=g).cuda()
    (out * w).sum().backward()
    x64 = x.double().cuda().requires_grad_(True)
    ((T * F.normalize(x64, p=2.0, dim=1)) * w.double()).sum().backward()
    ref = x64.grad.cpu().numpy()
    got = xc.grad.cpu().numpy()
    # the zero descriptor: d/dx (x / eps) = 1 / eps, 6e13 here -- compared on its own, relatively
    np.testing.assert_allclose(got[0, :, 0, 0], ref[0, :, 0, 0], rtol=1e-6)
    got[0, :, 0, 0] = ref[0, :, 0, 0] = 0.0
    scale = np.abs(ref).max(axis=1, keepdims=True)      # per pixel: the gradient is a difference of terms of this size
    assert np.abs(got - ref).max() <= 1e-30 + (2e-6 * scale).max()
    assert (np.abs(got - ref) <= 2e-6 * scale + 1e-30).all()
    # non-contiguous input: the tensor expression, same values
    xs = x.cuda()[:, :, ::1, :].transpose(2, 3).contiguous().transpose(2, 3)
    assert not xs.is_contiguous() or shape[2] == 1 or shape[3] == 1
    np.testing.assert_allclose(normalize_scale_channels_last(xs, T).cpu().numpy(), want.cpu().numpy(), rtol=3e-7, atol=1e-30)

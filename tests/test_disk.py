"""DiskLoss dense affinity (losses/kploss.py): oracle vs the reference's outputs (CPU) and the fused
tensor-core path vs both (GPU).  The random point samples of the reference run are replayed."""
import numpy as np
import pytest
import torch

from oracle import posfeat_oracle as O

CFG = dict(grid_size=8, loss_distance="cos", temperature_base=60, temperature_max=60,
           epipolar_reward="constant_reward", reward_config=dict(reward_thr=2, rescale_thr=False),
           cor_detach=True, good_reward=1, bad_reward=-0.25, kp_penalty=-0.001, match_grad=False)
COMPS = ["reinforce", "kp_penalty", "cor minmax", "cor minmean", "cor max", "cor mean", "cor summin", "cor summax",
         "n_kps", "n_pairs"]


def _case(g, tag):
    return {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(tag + "/")}


def _logp_from_maps(kp, coord, acc, gsz=8):
    """log-probability of the replayed draws as a function of the score map (kploss.py:20-38)."""
    from torch.distributions import Bernoulli, Categorical
    u = kp.unfold(2, gsz, gsz).unfold(3, gsz, gsz)
    b, c, h, w, g1, g2 = u.shape
    logits = u.reshape(b, c, h, w, g1 * g2)
    idx = ((coord[..., 1].long() % gsz) * gsz + (coord[..., 0].long() % gsz)).reshape(b, 1, h, w)
    prop = Categorical(logits=logits)
    al = torch.gather(logits, -1, idx[..., None]).squeeze(-1)
    return prop.log_prob(idx) + Bernoulli(logits=al).log_prob(acc.reshape(b, 1, h, w).float())


@pytest.mark.parametrize("tag", ["const", "dyn"])
def test_oracle_disk_vs_reference(golden, tag):
    c = _case(golden("disk"), tag)
    b, _, h, w = c["kp1"].shape
    co1, co2 = c["coord1"].reshape(b, -1, 2), c["coord2"].reshape(b, -1, 2)
    f1 = O.sample_feat_by_coord(c["xf1"], O.normalize_coords(co1, h, w), True)
    f2 = O.sample_feat_by_coord(c["xf2"], O.normalize_coords(co2, h, w), True)
    loss, comp, _, _ = O.disk_loss_dense(f1, f2, co1, co2, c["F1"], c["F2"], c["logp1"].reshape(b, -1),
                                         c["logp2"].reshape(b, -1), c["acc1"].reshape(b, -1), c["acc2"].reshape(b, -1),
                                         T=60.0, dynamic=(tag == "dyn"))
    assert abs(loss - float(c["loss"])) <= 2e-4 * abs(float(c["loss"]))
    for k in COMPS:
        assert abs(comp[k] - float(c["comp/" + k])) <= 2e-4 * abs(float(c["comp/" + k])) + 1e-7, k


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["const", "dyn"])
def test_disk_loss_golden(golden, tag):
    from posfeat_b200.kploss import DiskLoss
    c = _case(golden("disk"), tag)
    cfg = dict(CFG, epipolar_reward="dynamic_reward" if tag == "dyn" else "constant_reward")
    t = lambda x: torch.from_numpy(np.asarray(x)).cuda()
    kp1, kp2 = t(c["kp1"]).requires_grad_(True), t(c["kp2"]).requires_grad_(True)
    co1, co2, a1, a2 = t(c["coord1"]), t(c["coord2"]), t(c["acc1"]), t(c["acc2"])
    lp1, lp2 = _logp_from_maps(kp1, co1, a1), _logp_from_maps(kp2, co2, a2)
    np.testing.assert_allclose(lp1.detach().cpu().numpy(), c["logp1"], rtol=1e-5, atol=1e-6)
    inputs = {"F1": t(c["F1"]), "F2": t(c["F2"])}
    outputs = {"epoch": 0, "preds1": {"local_point": kp1, "local_map": t(c["xf1"])},
               "preds2": {"local_point": kp2, "local_map": t(c["xf2"])}}
    loss, comp = DiskLoss(cfg)(inputs, outputs, None, samples=((co1, lp1, a1), (co2, lp2, a2)))
    assert abs(float(loss) - float(c["loss"])) <= 2e-4 * abs(float(c["loss"]))
    for k in COMPS:
        assert abs(float(comp[k]) - float(c["comp/" + k])) <= 3e-4 * abs(float(c["comp/" + k])) + 1e-7, k
    loss.backward()
    for got, want in ((kp1.grad, c["g_kp1"]), (kp2.grad, c["g_kp2"])):
        scale = np.abs(want).max()
        assert np.abs(got.cpu().numpy() - want).max() <= 3e-4 * scale


@pytest.mark.gpu
def test_disk_loss_vs_oracle_larger():
    """4800-point sides are the real training shape; 1500 x 1300 here keeps the float64 oracle quick."""
    from posfeat_b200.kploss import _rows, _epipolar_lines
    from posfeat_b200.preprocess import corr_expect
    g = torch.Generator().manual_seed(5)
    B, m, n, D, T = 2, 1500, 1300, 128, 60.0
    f1 = torch.nn.functional.normalize(torch.randn(B, m, D, generator=g), dim=-1)
    f2 = torch.nn.functional.normalize(f1[:, torch.randperm(m, generator=g)[:n]] + 0.05 * torch.randn(B, n, D, generator=g), dim=-1)
    c1 = torch.rand(B, m, 2, generator=g) * 300
    c2 = torch.rand(B, n, 2, generator=g) * 300
    F1 = torch.tensor([[0., 0., 0.], [0., 0., -1.], [0., 1., 0.]]).repeat(B, 1, 1)
    F2 = F1.transpose(1, 2).contiguous()
    lp1, lp2 = -torch.rand(B, m, generator=g) * 5, -torch.rand(B, n, generator=g) * 5
    a1, a2 = torch.rand(B, m, generator=g) > 0.3, torch.rand(B, n, generator=g) > 0.3
    loss, comp, g1, g2 = O.disk_loss_dense(f1.numpy(), f2.numpy(), c1.numpy(), c2.numpy(), F1.numpy(), F2.numpy(), lp1.numpy(),
                                           lp2.numpy(), a1.numpy(), a2.numpy(), T=T, kp_penalty=0.0)
    f1c, f2c = f1.cuda(), f2.cuda()
    lse_r = corr_expect(f1c, f2c, torch.ones(n, 1).cuda(), T, want_lse=True)[1] - T
    lse_c = corr_expect(f2c, f1c, torch.ones(m, 1).cuda(), T, want_lse=True)[1] - T
    tab1 = torch.cat([lse_r[..., None], _epipolar_lines(F1.cuda(), c1.cuda()), c1.cuda(), lp1.cuda()[..., None], a1.cuda()[..., None].float()], -1).contiguous()
    tab2 = torch.cat([lse_c[..., None], _epipolar_lines(F2.cuda(), c2.cuda()), c2.cuda(), lp2.cuda()[..., None], a2.cuda()[..., None].float()], -1).contiguous()
    r1 = _rows(f1c, f2c, tab1, tab2, T, 2.0, 2.0, 1.0, -0.25, False).cpu().double().numpy()
    r2 = _rows(f2c, f1c, tab2, tab1, T, 2.0, 2.0, 1.0, -0.25, False).cpu().double().numpy()
    assert abs(-r1[..., 0].sum() - loss) <= 3e-4 * abs(loss)
    assert abs(r1[..., 0].sum() - r2[..., 0].sum()) <= 3e-4 * abs(loss)          # both role assignments see the same pairs
    np.testing.assert_allclose(-r1[..., 1], g1, rtol=0, atol=3e-4 * np.abs(g1).max())
    np.testing.assert_allclose(-r2[..., 1], g2, rtol=0, atol=3e-4 * np.abs(g2).max())
    assert abs(r1[..., 2].sum() / B - comp["n_pairs"]) <= 3e-4 * comp["n_pairs"]
    assert abs(r1[..., 3].max() - comp["cor max"]) <= 3e-4 * comp["cor max"]


OPTS = {"mg": dict(match_grad=True), "mg_cd": dict(match_grad=True, cor_detach=False), "cd": dict(cor_detach=False),
        "rs": dict(reward_config=dict(reward_thr=2, rescale_thr=True)),
        "rs_dyn_mg": dict(reward_config=dict(reward_thr=2, rescale_thr=True), epipolar_reward="dynamic_reward", match_grad=True)}


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(OPTS))
def test_disk_loss_options_golden(golden, tag):
    """match_grad / cor_detach / rescale_thr (losses/kploss.py:52-129, :152-182) against the reference's own run:
    loss, components and the gradients w.r.t. both score maps and both descriptor maps."""
    from posfeat_b200.kploss import DiskLoss
    g = golden("disk_opts")
    c = _case(g, tag)
    cfg = dict(CFG, **OPTS[tag])
    t = lambda x: torch.from_numpy(np.asarray(x)).cuda()
    kp1, kp2 = t(g["kp1"]).requires_grad_(True), t(g["kp2"]).requires_grad_(True)
    xf1, xf2 = t(g["xf1"]).requires_grad_(True), t(g["xf2"]).requires_grad_(True)
    co1, co2, a1, a2 = t(g["coord1"]), t(g["coord2"]), t(g["acc1"]), t(g["acc2"])
    lp1, lp2 = _logp_from_maps(kp1, co1, a1), _logp_from_maps(kp2, co2, a2)
    inputs = {"F1": t(g["F1"]), "F2": t(g["F2"])}
    outputs = {"epoch": 0, "preds1": {"local_point": kp1, "local_map": xf1}, "preds2": {"local_point": kp2, "local_map": xf2}}
    mod = DiskLoss(cfg)
    assert mod.dense
    loss, comp = mod(inputs, outputs, None, samples=((co1, lp1, a1), (co2, lp2, a2)))
    assert abs(float(loss) - float(c["loss"])) <= 2e-4 * abs(float(c["loss"]))
    for k in COMPS + ["scale1", "scale2"]:
        want = np.asarray(c["comp/" + k], dtype=np.float64)
        got = np.asarray(comp[k].detach().cpu().numpy() if torch.is_tensor(comp[k]) else comp[k], dtype=np.float64)
        np.testing.assert_allclose(got.reshape(want.shape), want, rtol=3e-4, atol=1e-7, err_msg=k)
    loss.backward()
    for name, tens in (("g_kp1", kp1), ("g_kp2", kp2), ("g_xf1", xf1), ("g_xf2", xf2)):
        want = c[name]
        got = tens.grad.cpu().numpy() if tens.grad is not None else np.zeros_like(want)
        scale = max(np.abs(want).max(), 1e-12)
        assert np.abs(got - want).max() <= 5e-4 * scale, (name, np.abs(got - want).max(), scale)
    if "mg" in tag:
        assert np.abs(c["g_xf1"]).max() > 0          # the descriptor maps do receive gradient with match_grad

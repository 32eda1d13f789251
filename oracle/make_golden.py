"""Generate tests/golden/*.npz by running the REAL reference from /root/reference.

Run in the authoring container only (the reference checkout does not travel to
the GPU box):   python oracle/make_golden.py
Everything is seeded; fixtures are small (a few hundred KB in total).  The
fixtures pin oracle/posfeat_oracle.py (tests/test_oracle_golden.py) and are
used directly by the GPU parity tests.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("POSFEAT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import losses.preprocess_utils as pu          # noqa: E402
import losses.preprocess as pp                # noqa: E402
from losses.epipolarloss import EpipolarLoss_full  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)


def gen(seed):
    return torch.Generator().manual_seed(seed)


def ref_detect(kp_map, **cfg):
    """Reference detector plus the topk index it used internally."""
    kps, sc = pu.generate_kpts_single(kp_map, **cfg)
    b = kp_map.shape[0]
    inter = kp_map[:, :, 1:-1, 1:-1]
    use_nms = cfg.get("use_nms", True)
    if use_nms == "softnms":
        mask = pu.soft_nms(inter, cfg["nms_radius"])
    elif use_nms:
        mask = pu.nms(inter, cfg["nms_radius"])
    else:
        mask = torch.ones_like(inter)
    nms_only = mask.clone()
    thr = cfg.get("thr", False)
    if thr:
        mod = cfg.get("thr_mod", "mean")
        if mod == "max":
            t = inter.reshape(b, 1, -1).max(2)[0]
        elif mod == "mean":
            t = inter.reshape(b, 1, -1).mean(2)
        else:
            t = torch.tensor(1.).repeat(b)
        mask = (inter > thr * t.view(b, 1, 1, 1)) * mask
    key = (mask * inter).reshape(b, -1)
    n = kps.shape[1]
    _, idx = key.topk(n)
    return dict(kps=kps.numpy(), score=sc.numpy(), idx=idx.numpy(),
                key=key.numpy(), nms_mask=nms_only.numpy().astype(np.float32),
                count=np.asarray(mask.reshape(b, -1).sum(1).numpy()))


def golden_detect():
    cases = {}
    g = gen(101)
    base = F.softplus(torch.randn(2, 1, 48, 64, generator=g))
    cases["r1_abs"] = (base, dict(nms_radius=1, num_pts=300, thr=0.9, thr_mod="abs"))
    cases["r3_abs"] = (base, dict(nms_radius=3, num_pts=False, thr=0.5, thr_mod="abs"))
    cases["r2_max"] = (base, dict(nms_radius=2, num_pts=200, thr=0.3, thr_mod="max"))
    cases["r1_mean"] = (base, dict(nms_radius=1, num_pts=150, thr=1.2, thr_mod="mean"))
    cases["r1_nothr"] = (base, dict(nms_radius=1, num_pts=400))
    cases["nonms_abs"] = (base, dict(nms_radius=1, num_pts=256, use_nms=False, thr=1.5,
                                     thr_mod="abs"))
    # heavy ties: quantised scores create plateaus and equal maxima
    q = torch.round(base * 4) / 4 + 0.25
    cases["ties_r1"] = (q, dict(nms_radius=1, num_pts=200, thr=0.9, thr_mod="abs"))
    cases["ties_r2"] = (q, dict(nms_radius=2, num_pts=False))
    # few survivors -> the num_pts<128 floor pads with masked-out pixels
    cases["few"] = (base, dict(nms_radius=4, num_pts=500, thr=2.5, thr_mod="abs"))
    const = torch.full((1, 1, 20, 24), 0.7)
    cases["const"] = (const, dict(nms_radius=1, num_pts=False))
    # odd sizes (not multiples of 4) and a wide image
    odd = F.softplus(torch.randn(1, 1, 37, 131, generator=g))
    cases["odd_r1"] = (odd, dict(nms_radius=1, num_pts=256, thr=0.9, thr_mod="abs"))
    cases["odd_r5"] = (odd, dict(nms_radius=5, num_pts=False))
    out = {}
    for name, (m, cfg) in cases.items():
        r = ref_detect(m, **cfg)
        out[name + "/map"] = m.numpy()
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        out[name + "/cfg"] = np.array(repr(cfg))
    np.savez_compressed(os.path.join(OUT, "detect.npz"), **out)
    print("detect:", list(cases))


def ref_detect_noavg(kp_map, **cfg):
    kps, sc = pu.generate_kpts_single_noavg(kp_map, **cfg)
    b = kp_map.shape[0]
    use_nms = cfg.get("use_nms", True)
    if use_nms == "softnms":
        mask = pu.soft_nms(kp_map, cfg["nms_radius"])
    elif use_nms:
        mask = pu.nms(kp_map, cfg["nms_radius"])
    else:
        mask = torch.ones_like(kp_map)
    thr = cfg.get("thr", False)
    cnt_mask = mask
    if thr:
        mod = cfg.get("thr_mod", "mean")
        t = kp_map.reshape(b, 1, -1).max(2)[0] if mod == "max" else kp_map.reshape(b, 1, -1).mean(2)
        thr_mask = kp_map > thr * t.view(b, 1, 1, 1)
        mask = thr_mask * mask
        cnt_mask = thr_mask if use_nms == "softnms" else mask
    key = (mask * kp_map).reshape(b, -1)
    _, idx = key.topk(kps.shape[1])
    return dict(kps=kps.numpy(), score=sc.numpy(), idx=idx.numpy(), key=key.numpy(),
                count=np.asarray(cnt_mask.reshape(b, -1).sum(1).numpy()))


def golden_detect_ext():
    """soft NMS (use_nms='softnms') and generate_kpts_single_noavg -- the remaining
    detector_config options of the plugin slot."""
    g = gen(111)
    base = F.softplus(torch.randn(2, 1, 48, 64, generator=g))
    odd = F.softplus(2 * torch.randn(1, 1, 37, 131, generator=g))
    soft = {
        "soft_r1_mean": (base, dict(nms_radius=1, num_pts=300, use_nms="softnms", thr=1.0, thr_mod="mean")),
        "soft_r2_abs": (base, dict(nms_radius=2, num_pts=False, use_nms="softnms", thr=1.2, thr_mod="abs")),
        "soft_r4_max": (odd, dict(nms_radius=4, num_pts=500, use_nms="softnms", thr=0.2, thr_mod="max")),
    }
    noavg = {
        "noavg_r1": (base, dict(nms_radius=1, num_pts=300)),
        "noavg_r2_mean": (base, dict(nms_radius=2, num_pts=False, thr=1.1, thr_mod="mean")),
        "noavg_r5_max": (odd, dict(nms_radius=5, num_pts=200, thr=0.1, thr_mod="max")),
        "noavg_nonms": (base, dict(nms_radius=1, num_pts=256, use_nms=False, thr=1.5, thr_mod="mean")),
        "noavg_soft": (odd, dict(nms_radius=2, num_pts=400, use_nms="softnms", thr=1.0, thr_mod="mean")),
    }
    out = {}
    for name, (m, cfg) in soft.items():
        r = ref_detect(m, **cfg)
        b = m.shape[0]
        inter = m[:, :, 1:-1, 1:-1]
        mod = cfg["thr_mod"]
        t = (inter.reshape(b, 1, -1).max(2)[0] if mod == "max" else inter.reshape(b, 1, -1).mean(2)
             if mod == "mean" else torch.tensor(1.).repeat(b))
        r["count"] = (inter > cfg["thr"] * t.view(b, 1, 1, 1)).reshape(b, -1).sum(1).numpy()   # thr_mask count (:253,:258)
        del r["nms_mask"]
        out[name + "/map"] = m.numpy()
        out[name + "/cfg"] = np.array(repr(cfg))
        for k, v in r.items():
            out[f"{name}/{k}"] = v
    for name, (m, cfg) in noavg.items():
        r = ref_detect_noavg(m, **cfg)
        out[name + "/map"] = m.numpy()
        out[name + "/cfg"] = np.array(repr(cfg))
        for k, v in r.items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "detect_ext.npz"), **out)
    print("detect_ext:", list(soft) + list(noavg))


def golden_sample():
    g = gen(202)
    x = torch.randn(2, 24, 12, 15, generator=g)
    c = torch.rand(2, 200, 2, generator=g) * 2.2 - 1.1       # some fall outside
    c[0, :4] = torch.tensor([[-1., -1.], [1., 1.], [-1., 1.], [0.999, -0.999]])
    out = dict(x=x.numpy(), coord=c.numpy(),
               raw=pu.sample_feat_by_coord(x, c, False).numpy(),
               normed=pu.sample_feat_by_coord(x, c, True).numpy())
    x128 = torch.randn(1, 128, 16, 20, generator=g)
    c128 = torch.rand(1, 300, 2, generator=g) * 2 - 1
    out.update(x128=x128.numpy(), coord128=c128.numpy(),
               normed128=pu.sample_feat_by_coord(x128, c128, True).numpy())
    np.savez_compressed(os.path.join(OUT, "sample.npz"), **out)
    print("sample ok")


def golden_mnn():
    sys.path.insert(0, os.path.join(REF, "evaluations", "aachen"))
    import matchers as am
    g = gen(303)
    a = F.normalize(torch.randn(257, 128, generator=g), dim=1)
    b = F.normalize(a[torch.randperm(257, generator=g)][:201] + 0.4 * torch.randn(201, 128, generator=g), dim=1)
    b = torch.cat([b, F.normalize(torch.randn(100, 128, generator=g), dim=1)])
    out = dict(a=a.numpy(), b=b.numpy(), mnn=pu.mnn_matcher(a, b),
               mutual_nn=am.mutual_nn_matcher(a, b),
               ratio=am.ratio_matcher(a, b, 0.95),
               mutual_ratio=am.mutual_nn_ratio_matcher(a, b, 0.9))
    # duplicates: exact ties must resolve to the first index
    ad = a[:64].clone()
    bd = torch.cat([a[:32], a[:32], a[32:64]])
    out.update(ad=ad.numpy(), bd=bd.numpy(), mnn_dup=pu.mnn_matcher(ad, bd),
               mnn_dup_t=pu.mnn_matcher(bd, ad))
    np.savez_compressed(os.path.join(OUT, "mnn.npz"), **out)
    print("mnn:", out["mnn"].shape, out["mnn_dup"].shape, out["ratio"].shape, out["mutual_ratio"].shape)


def golden_corr():
    g = gen(404)
    f1 = F.normalize(torch.randn(2, 40, 32, generator=g), dim=-1) * 3
    fm = torch.randn(2, 32, 10, 12, generator=g)
    e, std, kurt, prob = pu.get_expected_correspondence_locs(f1, fm, with_std=True)
    out = dict(f1=f1.numpy(), fm=fm.numpy(), exp=e.numpy(), std=std.numpy(), prob=prob.numpy())
    fmw = 6 * F.normalize(torch.randn(2, 32, 40, 60, generator=g), dim=1)
    c2 = torch.rand(2, 40, 2, generator=g) * 2.1 - 1.05
    f1w = F.normalize(torch.randn(2, 40, 32, generator=g), dim=-1)
    ew, cg, stdw, probw = pu.get_expected_correspondence_within_window(f1w, fmw, c2, 0.1, with_std=True)
    out.update(f1w=f1w.numpy(), fmw=fmw.numpy(), c2=c2.numpy(), expw=ew.numpy(), cgw=cg.numpy(),
               stdw=stdw.numpy(), probw=probw.numpy())
    np.savez_compressed(os.path.join(OUT, "corr.npz"), **out)
    print("corr ok", cg.shape)


def golden_prob():
    """compute_prob (losses/preprocess_utils.py:89-115): every option, plus gradients of a scalar of prob."""
    g = gen(505)
    out = {}
    f1 = F.normalize(torch.randn(2, 37, 24, generator=g), dim=-1) * 2.5
    f2 = F.normalize(torch.randn(2, 150, 24, generator=g), dim=-1) * 2.0
    wgt = torch.randn(2, 37, 150, generator=g)
    out.update(f1=f1.numpy(), f2=f2.numpy(), wgt=wgt.numpy())
    for name, kw in (("cos", {}), ("cos_scale", dict(with_scale=True)), ("euc", dict(loss_distance="euc"))):
        a = f1.clone().requires_grad_(True)
        b = f2.clone().requires_grad_(True)
        prob = pu.compute_prob(a, b, **kw)
        (prob * wgt).sum().backward()
        out[f"{name}/prob"] = prob.detach().numpy()
        out[f"{name}/g1"] = a.grad.numpy()
        out[f"{name}/g2"] = b.grad.numpy()
    prob, sim = pu.compute_prob(f1, f2, return_sim=True)
    out["sim/prob"] = prob.numpy()
    out["sim/sim"] = sim.numpy()
    np.savez_compressed(os.path.join(OUT, "prob.npz"), **out)
    print("prob ok", prob.shape)


def random_fundamental(b, h, w, g):
    Fs = []
    for _ in range(b):
        ang = (torch.rand(3, generator=g) - 0.5) * 0.5
        Rx = torch.tensor([[1, 0, 0], [0, torch.cos(ang[0]), -torch.sin(ang[0])], [0, torch.sin(ang[0]), torch.cos(ang[0])]])
        Ry = torch.tensor([[torch.cos(ang[1]), 0, torch.sin(ang[1])], [0, 1, 0], [-torch.sin(ang[1]), 0, torch.cos(ang[1])]])
        Rz = torch.tensor([[torch.cos(ang[2]), -torch.sin(ang[2]), 0], [torch.sin(ang[2]), torch.cos(ang[2]), 0], [0, 0, 1]])
        R = Rz @ Ry @ Rx
        t = F.normalize(torch.randn(3, generator=g), dim=0)
        tx = torch.tensor([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        K = torch.tensor([[float(w), 0, w / 2], [0, float(w), h / 2], [0, 0, 1]])
        Ki = torch.linalg.inv(K)
        Fm = Ki.t() @ tx @ R @ Ki
        Fs.append(Fm / Fm.norm())
    return torch.stack(Fs).float()


def golden_preprocess():
    """Full Preprocess_Line2Window.forward + EpipolarLoss_full with gradients."""
    g = gen(505)
    B, d, H, W = 2, 32, 128, 160
    cfg = dict(kps_generator="generate_kpts_regular_grid_random",
               kps_generator_config=dict(grid_size=16, map_init="identity", keep_spatial=True,
                                         random_select="random"),
               window_size=0.1, loss_distance="cos", use_nn_grid=False, use_line_search=True,
               line_search_config=dict(line_step=100, use_nn=True, loc_rand=True),
               temperature_base=60, temperature_max=60)
    P = pp.Preprocess_Line2Window(cfg)
    xf1 = torch.randn(B, d, H // 4, W // 4, generator=g).requires_grad_()
    xf2 = torch.randn(B, d, H // 4, W // 4, generator=g).requires_grad_()
    inputs = dict(im1=torch.zeros(B, 3, H, W), im2=torch.zeros(B, 3, H, W),
                  F1=random_fundamental(B, H, W, g))
    inputs["F2"] = inputs["F1"].transpose(1, 2).contiguous()
    preds1 = dict(global_map=torch.zeros(B, d, H // 16, W // 16), local_map=xf1,
                  local_point=torch.ones(B, 1, H, W))
    preds2 = dict(global_map=torch.zeros(B, d, H // 16, W // 16), local_map=xf2,
                  local_point=torch.ones(B, 1, H, W))
    outputs = dict(preds1=preds1, preds2=preds2, epoch=0)

    rec = {}
    orig_gen = P.kps_generator

    def gen_wrap(*a, **k):
        r = orig_gen(*a, **k)
        rec["coord1_n"], rec["coord2_n"] = r[0].clone(), r[1].clone()
        rec["all"] = tuple(x.clone() if torch.is_tensor(x) else x for x in r)
        return r
    P.kps_generator = gen_wrap
    rands = []
    orig_rand = torch.rand

    def rand_wrap(*a, **k):
        t = orig_rand(*a, **k)
        rands.append(t.clone())
        return t
    # the line search runs under no_grad and makes a discrete choice (position of the largest probability): its
    # results are recorded so that the float64 run below can replay them instead of re-deciding near ties
    line_rec = []
    orig_line = pp.epipolar_line_search

    def line_wrap(*a, **k):
        r = orig_line(*a, **k)
        line_rec.append(tuple(x.clone() for x in r))
        return r
    torch.manual_seed(7)
    torch.rand = rand_wrap
    pp.epipolar_line_search = line_wrap
    try:
        processed = P(inputs, outputs)
    finally:
        torch.rand = orig_rand
        pp.epipolar_line_search = orig_line
    assert len(line_rec) == 2
    assert len(rands) == 2, len(rands)
    loss_cfg = dict(grid_cost_thr=0.5, win_cost_thr=0.1, use_std_as_weight=True,
                    weight_grid=0.3, weight_window=1)
    L = EpipolarLoss_full(loss_cfg)
    loss, comp = L(inputs, outputs, processed)
    loss.backward()
    out = dict(xf1=xf1.detach().numpy(), xf2=xf2.detach().numpy(),
               F1=inputs["F1"].numpy(), F2=inputs["F2"].numpy(),
               coord1_n=rec["coord1_n"].numpy(), coord2_n=rec["coord2_n"].numpy(),
               jitter1=rands[0].numpy(), jitter2=rands[1].numpy(),
               loss=loss.detach().numpy(), gxf1=xf1.grad.numpy(), gxf2=xf2.grad.numpy(),
               H=np.array(H), W=np.array(W))
    for k, v in processed.items():
        if torch.is_tensor(v):
            out["p_" + k] = v.detach().numpy()
    for k, v in comp.items():
        out["c_" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **out)
    print("preprocess ok; loss", float(loss), "valid", processed["valid_epi1"].float().mean().item())

    # ---- the same run in float64 (coordinates replayed, and the results of the no_grad line search -- a discrete
    # argmax plus jitter -- replayed from the float32 run): the yardstick that tells how much of a float32
    # implementation's deviation is the reference's own rounding noise.  Stored rounded to float32
    # (6e-8 relative, far below the 1e-5 bar) to keep the fixture small.
    P64 = pp.Preprocess_Line2Window(cfg)
    P64.kps_generator = lambda *a, **k: tuple(x.double() if torch.is_tensor(x) and x.dtype.is_floating_point else x
                                              for x in rec["all"])
    x1d = xf1.detach().double().requires_grad_()
    x2d = xf2.detach().double().requires_grad_()
    in64 = dict(im1=inputs["im1"].double(), im2=inputs["im2"].double(), F1=inputs["F1"].double(), F2=inputs["F2"].double())
    out64 = dict(preds1=dict(global_map=preds1["global_map"].double(), local_map=x1d, local_point=preds1["local_point"].double()),
                 preds2=dict(global_map=preds2["global_map"].double(), local_map=x2d, local_point=preds2["local_point"].double()),
                 epoch=0)
    line_replay = [tuple(x.double() if x.dtype.is_floating_point else x for x in r) for r in line_rec]
    pp.epipolar_line_search = lambda *a, **k: line_replay.pop(0)
    try:
        pr64 = P64(in64, out64)
    finally:
        pp.epipolar_line_search = orig_line
    loss64, _ = EpipolarLoss_full(loss_cfg)(in64, out64, pr64)
    loss64.backward()
    # ---- the float32 reference once more with the float64 run's std values as (detached) loss weights: what is
    # left between this and the float64 run is the rounding error of the reference's differentiable float32 path
    # alone (the weights 1/std amplify the float32 cancellation noise of std by orders of magnitude otherwise)
    std_keys = ("feat1g_std", "feat2g_std", "feat1w_std", "feat2w_std")
    xf1.grad = None
    xf2.grad = None
    replay32 = [tuple(x.clone() for x in r) for r in line_rec]
    P.kps_generator = lambda *a, **k: rec["all"]
    pp.epipolar_line_search = lambda *a, **k: replay32.pop(0)
    try:
        pr32 = P(inputs, outputs)
    finally:
        pp.epipolar_line_search = orig_line
    for k in std_keys:
        assert torch.equal(pr32[k], processed[k])           # the replayed float32 run is the recorded one
        pr32[k] = pr64[k].detach().float()
    loss32s, _ = L(inputs, outputs, pr32)
    loss32s.backward()
    o64 = dict(loss=np.array(float(loss64)), gxf1=x1d.grad.numpy().astype(np.float32), gxf2=x2d.grad.numpy().astype(np.float32),
               loss_f32_std64=np.array(float(loss32s)), gxf1_f32_std64=xf1.grad.numpy(), gxf2_f32_std64=xf2.grad.numpy())
    print("float32 reference with float64 weights: loss dev", abs(float(loss32s) - float(loss64)) / float(loss64), "grad dev",
          float((xf1.grad.double() - x1d.grad).abs().max() / x1d.grad.abs().max()))
    for k, v in pr64.items():
        if torch.is_tensor(v):
            o64["p_" + k] = v.detach().numpy().astype(np.float32) if v.dtype.is_floating_point else v.numpy()
    np.savez_compressed(os.path.join(OUT, "preprocess_f64.npz"), **o64)
    print("preprocess float64 ok; loss", float(loss64), "| fp32 run deviates by", abs(float(loss64) - float(loss)) / float(loss64),
          "(loss),", float((xf1.grad.double() - x1d.grad).abs().max() / x1d.grad.abs().max()), "(grad, rel to max)")


def golden_disk():
    """DiskLoss (losses/kploss.py) on small maps: the random point samples are drawn once with the
    reference's own point_sample and replayed, so that every implementation sees the same points."""
    from losses.kploss import DiskLoss
    g = gen(707)
    torch.manual_seed(707)
    b, h, w, d = 2, 64, 96, 32
    cfg = dict(grid_size=8, loss_distance="cos", temperature_base=60, temperature_max=60,
               epipolar_reward="constant_reward", reward_config=dict(reward_thr=2, rescale_thr=False),
               cor_detach=True, good_reward=1, bad_reward=-0.25, kp_penalty=-0.001, match_grad=False)
    out = {}
    for tag, reward in (("const", "constant_reward"), ("dyn", "dynamic_reward")):
        cfg["epipolar_reward"] = reward
        loss_mod = DiskLoss(cfg)
        kp1 = torch.randn(b, 1, h, w, generator=g).requires_grad_(True)
        kp2 = torch.randn(b, 1, h, w, generator=g).requires_grad_(True)
        xf1 = torch.randn(b, d, h // 4, w // 4, generator=g)
        xf2 = xf1 + 0.3 * torch.randn(b, d, h // 4, w // 4, generator=g)
        # a fundamental matrix of a pure horizontal shift: x2 = x1 + 3  (F = [t]_x), so matches are plentiful
        F1 = torch.tensor([[0., 0., 0.], [0., 0., -1.], [0., 1., 0.]]).repeat(b, 1, 1)
        F2 = F1.transpose(1, 2).contiguous()
        s1 = loss_mod.point_sample(kp1)
        s2 = loss_mod.point_sample(kp2)
        draws = iter([s1, s2])
        loss_mod.point_sample = lambda kp_map: next(draws)
        inputs = {"F1": F1, "F2": F2}
        outputs = {"epoch": 0, "preds1": {"local_point": kp1, "local_map": xf1}, "preds2": {"local_point": kp2, "local_map": xf2}}
        loss, comp = loss_mod(inputs, outputs, None)
        loss.backward()
        out.update({f"{tag}/kp1": kp1.detach().numpy(), f"{tag}/kp2": kp2.detach().numpy(), f"{tag}/xf1": xf1.numpy(),
                    f"{tag}/xf2": xf2.numpy(), f"{tag}/F1": F1.numpy(), f"{tag}/F2": F2.numpy(),
                    f"{tag}/coord1": s1[0].detach().numpy(), f"{tag}/logp1": s1[1].detach().numpy(), f"{tag}/acc1": s1[2].numpy(),
                    f"{tag}/coord2": s2[0].detach().numpy(), f"{tag}/logp2": s2[1].detach().numpy(), f"{tag}/acc2": s2[2].numpy(),
                    f"{tag}/loss": loss.detach().numpy(), f"{tag}/g_kp1": kp1.grad.numpy(), f"{tag}/g_kp2": kp2.grad.numpy()})
        for k, v in comp.items():
            out[f"{tag}/comp/{k}"] = np.asarray(v.detach().numpy() if torch.is_tensor(v) else v)
    np.savez_compressed(os.path.join(OUT, "disk.npz"), **out)
    print("disk ok", float(out["const/loss"]), float(out["dyn/loss"]))


def golden_disk_opts():
    """DiskLoss with the settings that put gradient on the affinity or rescale the reward threshold
    (match_grad, cor_detach, rescale_thr; losses/kploss.py:52-129, :152-182): one small shared input, one entry
    per option set with loss, components and the gradients w.r.t. both score maps AND both descriptor maps."""
    from losses.kploss import DiskLoss
    g = gen(808)
    torch.manual_seed(808)
    b, h, w, d = 2, 48, 64, 16
    base = dict(grid_size=8, loss_distance="cos", temperature_base=60, temperature_max=60,
                epipolar_reward="constant_reward", reward_config=dict(reward_thr=2, rescale_thr=False),
                cor_detach=True, good_reward=1, bad_reward=-0.25, kp_penalty=-0.001, match_grad=False)
    kp1_0 = torch.randn(b, 1, h, w, generator=g)
    kp2_0 = torch.randn(b, 1, h, w, generator=g)
    xf1_0 = torch.randn(b, d, h // 4, w // 4, generator=g)
    xf2_0 = xf1_0 + 0.3 * torch.randn(b, d, h // 4, w // 4, generator=g)
    F1 = torch.tensor([[0., 0., 0.], [0., 0., -1.], [0., 1., 0.]]).repeat(b, 1, 1)
    F1 = F1 + 0.01 * torch.randn(b, 3, 3, generator=g)            # not exactly a pure shift: the two distances differ
    F2 = F1.transpose(1, 2).contiguous()
    sampler = DiskLoss(base)
    s1 = sampler.point_sample(kp1_0)
    s2 = sampler.point_sample(kp2_0)
    out = dict(kp1=kp1_0.numpy(), kp2=kp2_0.numpy(), xf1=xf1_0.numpy(), xf2=xf2_0.numpy(), F1=F1.numpy(), F2=F2.numpy(),
               coord1=s1[0].numpy(), acc1=s1[2].numpy(), coord2=s2[0].numpy(), acc2=s2[2].numpy())
    opts = {"mg": dict(match_grad=True), "mg_cd": dict(match_grad=True, cor_detach=False), "cd": dict(cor_detach=False),
            "rs": dict(reward_config=dict(reward_thr=2, rescale_thr=True)),
            "rs_dyn_mg": dict(reward_config=dict(reward_thr=2, rescale_thr=True), epipolar_reward="dynamic_reward", match_grad=True)}
    for tag, o in opts.items():
        cfg = dict(base, **o)
        mod = DiskLoss(cfg)
        kp1, kp2 = kp1_0.clone().requires_grad_(True), kp2_0.clone().requires_grad_(True)
        xf1, xf2 = xf1_0.clone().requires_grad_(True), xf2_0.clone().requires_grad_(True)

        def replay(kp_map, _kp1=kp1, _s=(s1, s2)):
            # the recorded proposals / accept draws, with log-probabilities recomputed from THIS map (autograd)
            coord, _, acc = _s[0] if kp_map is _kp1 else _s[1]
            logits = pu.unfold(kp_map, 8)
            bb, c, hh, ww, _ = logits.shape
            idx = ((coord[..., 1].long() % 8) * 8 + (coord[..., 0].long() % 8)).reshape(bb, 1, hh, ww)
            from torch.distributions import Bernoulli, Categorical
            al = torch.gather(logits, -1, idx[..., None]).squeeze(-1)
            logp = Categorical(logits=logits).log_prob(idx) + Bernoulli(logits=al).log_prob(acc.reshape(bb, 1, hh, ww).float())
            return coord, logp, acc
        mod.point_sample = replay
        outputs = {"epoch": 0, "preds1": {"local_point": kp1, "local_map": xf1}, "preds2": {"local_point": kp2, "local_map": xf2}}
        loss, comp = mod({"F1": F1, "F2": F2}, outputs, None)
        loss.backward()
        out[f"{tag}/loss"] = loss.detach().numpy()
        for name, t in (("g_kp1", kp1), ("g_kp2", kp2), ("g_xf1", xf1), ("g_xf2", xf2)):
            out[f"{tag}/{name}"] = (t.grad if t.grad is not None else torch.zeros_like(t)).numpy()
        for k, v in comp.items():
            out[f"{tag}/comp/{k}"] = np.asarray(v.detach().numpy() if torch.is_tensor(v) else v)
    np.savez_compressed(os.path.join(OUT, "disk_opts.npz"), **out)
    print("disk_opts ok", {t: float(out[f"{t}/loss"]) for t in opts})


if __name__ == "__main__":
    only = set(sys.argv[1:])                      # e.g. `python oracle/make_golden.py detect_ext`
    for fn in (golden_detect, golden_detect_ext, golden_sample, golden_mnn, golden_corr, golden_prob, golden_preprocess, golden_disk, golden_disk_opts):
        if not only or fn.__name__[len("golden_"):] in only:
            fn()
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes:", tot)

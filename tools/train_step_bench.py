"""BASELINE.json configs[3]: the MegaDepth-shaped training step of the post-backbone path.

  B = 8 image pairs of 480x640, fine maps [8,128,120,160], coarse maps [8,128,30,40], n = 512 query points per image;
  forward + backward of  Preprocess_Line2Window (grid<->grid softmax expectations, epipolar line search, 12x16
  window expectation; losses/preprocess.py:24-118)  ->  EpipolarLoss_full (losses/epipolarloss.py:38-101),
  plus the dense coarse (30x40) and fine (120x160) get_expected_correspondence_locs terms BASELINE config 4 names
  (losses/preprocess_utils.py:55-82), with gradients to both descriptor maps;
  then ONE bucketed NCCL all-reduce of the backbone's gradient volume (ResUNet-50: 20.51 M float32 = 82 MB, in
  25 MB buckets) on a side stream, started when the loss backward starts, i.e. overlapped with the correlation
  backward (managers/trainer.py:297-331 is the step this stands for; no SyncBatchNorm).

The conv backbones are out of scope (they stay PyTorch): the maps are synthetic leaf tensors and the gradient
buckets are a synthetic buffer of the right size -- the measurement is the loss path and how much of the
collective it hides.  Reported: step ms without the collective, the collective alone, both overlapped, the hidden
fraction, per-kernel times, and (rank 0, when oracle/_ref is staged) the same step through the reference's own
modules on the same GPU.  Used by  bench.py --workload train_step  (under torchrun for N > 1).
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, H, W, D, N_Q = 8, 480, 640, 128, 512
GRAD_FLOATS = 20_510_000          # ResUNet-50 parameters, 82 MB of float32 gradients
PRE_CFG = dict(kps_generator="generate_kpts_regular_grid_random",
               kps_generator_config=dict(grid_size=16, map_init="identity", keep_spatial=True, random_select="random"),
               window_size=0.1, loss_distance="cos", use_nn_grid=False, use_line_search=True,
               line_search_config=dict(line_step=100, use_nn=True, loc_rand=True),
               temperature_base=60, temperature_max=60)                       # configs/train_desc.yaml:61-80
LOSS_CFG = dict(grid_cost_thr=0.5, win_cost_thr=0.1, use_std_as_weight=True, weight_grid=0, weight_window=1)   # :86-91


def fundamental(b, h, w, g):
    """F of a random relative pose (rotation <= 30 degrees about a random axis, unit translation) between two
    pinhole cameras with focal length max(h, w): x2^T F x1 = 0 in pixel coordinates."""
    import torch
    axis = torch.nn.functional.normalize(torch.randn(b, 3, generator=g), dim=-1)
    ang = (torch.rand(b, generator=g) * 2 - 1) * math.radians(30)
    K = torch.zeros(b, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    eye = torch.eye(3).expand(b, 3, 3)
    R = eye + torch.sin(ang)[:, None, None] * K + (1 - torch.cos(ang))[:, None, None] * (K @ K)
    t = torch.nn.functional.normalize(torch.randn(b, 3, generator=g), dim=-1)
    tx = torch.zeros(b, 3, 3)
    tx[:, 0, 1], tx[:, 0, 2], tx[:, 1, 0] = -t[:, 2], t[:, 1], t[:, 2]
    tx[:, 1, 2], tx[:, 2, 0], tx[:, 2, 1] = -t[:, 0], -t[:, 1], t[:, 0]
    f = float(max(h, w))
    Kin = torch.tensor([[f, 0, (w - 1) / 2], [0, f, (h - 1) / 2], [0, 0, 1.0]])
    Ki = torch.linalg.inv(Kin)
    return Ki.T @ (tx @ R) @ Ki


def make_inputs(dev, seed=4242):
    import torch
    g = torch.Generator().manual_seed(seed)
    xf1 = torch.randn(B, D, H // 4, W // 4, generator=g)
    xf2 = xf1 + 0.5 * torch.randn(B, D, H // 4, W // 4, generator=g)
    xc1 = torch.randn(B, D, H // 16, W // 16, generator=g)
    xc2 = xc1 + 0.5 * torch.randn(B, D, H // 16, W // 16, generator=g)
    F1 = fundamental(B, H, W, g)
    gh, gw = 16, 32                                            # 512 query points on a jittered grid
    def grid():
        yy = (torch.arange(gh).float()[None, :, None] + torch.rand(B, gh, gw, generator=g)) / gh * 2 - 1
        xx = (torch.arange(gw).float()[None, None, :] + torch.rand(B, gh, gw, generator=g)) / gw * 2 - 1
        return torch.stack([xx, yy], -1) * 0.97
    c1, c2 = grid(), grid()
    jit = (torch.rand(B, gh * gw, 2, generator=g), torch.rand(B, gh * gw, 2, generator=g))
    inputs = dict(im1=torch.zeros(B, 3, H, W), im2=torch.zeros(B, 3, H, W), F1=F1.to(dev), F2=F1.transpose(1, 2).contiguous().to(dev))
    maps = [t.to(dev).requires_grad_() for t in (xf1, xf2, xc1, xc2)]
    return inputs, maps, (c1.to(dev), c2.to(dev)), tuple(j.to(dev) for j in jit)


def build_step(dev):
    """Returns (forward() -> loss, maps): one training step of the post-backbone path through this package."""
    import torch
    import posfeat_b200.preprocess as PP
    from posfeat_b200.epipolarloss import EpipolarLoss_full, point_to_line_distance
    from posfeat_b200.preprocess_utils import denormalize_coords
    inputs, maps, coords, jitter = make_inputs(dev)
    xf1, xf2, xc1, xc2 = maps
    pre, crit = PP.Preprocess_Line2Window(PRE_CFG), EpipolarLoss_full(LOSS_CFG)
    ones = torch.ones(B, 1, H, W, device=dev)

    def forward():
        outputs = dict(preds1=dict(local_map=xf1, global_map=xc1, local_point=ones),
                       preds2=dict(local_map=xf2, global_map=xc2, local_point=ones), epoch=0)
        pr = pre(inputs, outputs, coords=coords, jitter=jitter)
        loss, comp = crit(inputs, outputs, pr)
        # dense coarse + fine expectations (BASELINE config 4): queries of image 1 against the whole maps of image 2
        c1n = coords[0].reshape(B, -1, 2)
        q = PP.sample_feat_by_coord_grad(xf1, c1n, True)
        dense = 0.0
        for fm in (xc2, xf2):
            # (the package's one-pass scale * normalize -> channels_last; the keys of the dense expectation are then a
            # view of it, and the gradient comes back in the same layout: no copies either way)
            e = PP.get_expected_correspondence_locs(q, PP.normalize_scale_channels_last(fm, 20.0))
            cost = point_to_line_distance(pr["coord1"], denormalize_coords(e, H, W), inputs["F1"])
            dense = dense + cost.clamp(max=0.5 * H).mean()
        return loss + 0.1 * dense
    return forward, maps


def build_reference_step(dev):
    """The same step through the reference's own modules (oracle/_ref) on cuda tensors, or None."""
    import torch
    from oracle import ref_runner
    if not ref_runner.available():
        return None
    pu, pp, el, _ = ref_runner.modules()
    inputs, maps, coords, jitter = make_inputs(dev)
    xf1, xf2, xc1, xc2 = maps
    pre, crit = pp.Preprocess_Line2Window(PRE_CFG), el.EpipolarLoss_full(LOSS_CFG)
    ones = torch.ones(B, 1, H, W, device=dev)
    c1, c2 = coords
    pre.kps_generator = lambda *a, **k: (c1, c2, torch.ones_like(c1[..., :1]), torch.ones_like(c1[..., :1]))

    def forward():
        outputs = dict(preds1=dict(local_map=xf1, global_map=xc1, local_point=ones),
                       preds2=dict(local_map=xf2, global_map=xc2, local_point=ones), epoch=0)
        pr = pre(inputs, outputs)
        loss, comp = crit(inputs, outputs, pr)
        q = pu.sample_feat_by_coord(xf1, c1.reshape(B, -1, 2), True)
        dense = 0.0
        for fm in (xc2, xf2):
            e = pu.get_expected_correspondence_locs(q, torch.nn.functional.normalize(fm, dim=1) * 20.0)
            cost = crit.epipolar_cost(pr["coord1"], pu.denormalize_coords(e, H, W), inputs["F1"], None)
            dense = dense + cost.clamp(max=0.5 * H).mean()
        return loss + 0.1 * dense
    return forward, maps


def _time(fn, iters, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(dev, world, iters=10):
    import torch
    import torch.distributed as dist
    from posfeat_b200 import _lib
    from posfeat_b200.dist import GradAllReducer
    torch.backends.cuda.matmul.allow_tf32 = False
    forward, maps = build_step(dev)
    grads = [torch.zeros(GRAD_FLOATS // 8, device=dev, requires_grad=True) for _ in range(8)]     # "parameters" of the backbone
    red = GradAllReducer(grads, bucket_mb=25.0)
    for b in red.buckets:
        b.normal_()

    def zero():
        for m in maps:
            m.grad = None

    def step_plain():
        zero()
        forward().backward()

    def step_fwd():
        zero()
        return forward()

    def allreduce_only():
        red.start()
        red.finish()

    def step_overlapped():
        zero()
        loss = forward()
        red.start()               # gradient buckets (of the previous micro-step / the backbone) reduce on the side stream ...
        loss.backward()           # ... while the correlation backward runs
        red.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the step launches ~10 kernels of this library among a few hundred small tensor ops and is bound by the host's
    # launch rate: captured once in a CUDA graph (static shapes, no host round trip anywhere on the path) it replays
    # as one submission
    graph, graph_err = None, None
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                step_plain()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        zero()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = forward()
            static_loss.backward()
        graph.replay()
        torch.cuda.synchronize()
        g_loss = float(static_loss.detach())
        g_grad = float(maps[0].grad.abs().sum())
    except Exception as e:          # report, do not hide
        graph, graph_err = None, repr(e)[:300]
        torch.cuda.synchronize()

    def step_graph():
        graph.replay()

    def step_graph_overlapped():
        red.start()
        graph.replay()
        red.finish()

    barrier()
    t_fwd = _time(step_fwd, iters)
    barrier()
    t_plain = _time(step_plain, iters)
    barrier()
    t_ar = _time(allreduce_only, iters)
    barrier()
    t_ov = _time(step_overlapped, iters)
    t_g = t_gov = float("nan")
    if graph is not None:
        barrier()
        t_g = _time(step_graph, iters)
        barrier()
        t_gov = _time(step_graph_overlapped, iters)
    t = torch.tensor([t_fwd, t_plain, t_ar, t_ov, t_g, t_gov], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_fwd, t_plain, t_ar, t_ov, t_g, t_gov = (float(x) for x in t)
    loss = float(forward().detach())
    _lib.profile_enable(True)
    torch.cuda.synchronize()
    for _ in range(5):
        step_plain()
    torch.cuda.synchronize()
    prof = {k: v[0] / 5 for k, v in _lib.profile_read().items()}
    _lib.profile_enable(False)
    out = {"workload": "train_step_B8_480x640_n512_d128", "n_gpus": world, "pairs_per_step_per_gpu": B, "queries": N_Q,
           "loss": loss, "fwd_ms": t_fwd, "step_ms_no_collective": t_plain, "allreduce_alone_ms": t_ar,
           "step_ms_overlapped": t_ov, "allreduce_bytes": red.nbytes, "buckets": len(red.buckets),
           "allreduce_hidden_fraction": (max(0.0, min(1.0, (t_plain + t_ar - t_ov) / t_ar)) if t_ar > 0.02 else None),
           "allreduce_busbw_GBs": (2 * (world - 1) / world * red.nbytes / (t_ar * 1e-3) / 1e9 if world > 1 else None),
           "pairs_per_s": world * B / (t_ov * 1e-3), "kernel_ms_per_step": {k: round(v, 4) for k, v in prof.items()},
           "timing": "CUDA events, max over ranks"}
    if graph is not None:
        step_plain()
        torch.cuda.synchronize()
        out["cuda_graph"] = {"step_ms_no_collective": t_g, "step_ms_overlapped": t_gov,
                             "allreduce_hidden_fraction": (max(0.0, min(1.0, (t_g + t_ar - t_gov) / t_ar)) if t_ar > 0.02 else None),
                             "pairs_per_s": world * B / (t_gov * 1e-3),
                             "loss_equal_eager": abs(g_loss - loss) <= 1e-5 * abs(loss),
                             "grad_abs_sum_rel_diff_vs_eager": abs(g_grad - float(maps[0].grad.abs().sum())) / max(g_grad, 1e-30)}
    else:
        out["cuda_graph"] = {"error": graph_err}
    if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
        return out
    ref = None
    try:
        ref = build_reference_step(dev)
    except Exception as e:        # the comparison leg must not take the measurement down
        out["reference_on_b200"] = {"error": repr(e)[:200]}
    if ref is not None:
        rfwd, rmaps = ref

        def ref_step():
            for m in rmaps:
                m.grad = None
            rfwd().backward()
        t_ref = _time(ref_step, max(3, iters // 2), warmup=2)
        out["reference_on_b200"] = {"step_ms": t_ref, "loss": float(rfwd().detach()), "speedup": t_ref / t_plain, "speedup_cuda_graph": (t_ref / t_g if graph is not None else None),
                                    "what": "the reference's Preprocess_Line2Window + EpipolarLoss_full + "
                                            "get_expected_correspondence_locs (oracle/_ref) on cuda tensors"}
    return out


if __name__ == "__main__":
    import json
    import torch
    print(json.dumps(run(torch.device("cuda", 0), 1)))

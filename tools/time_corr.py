"""BASELINE.json configs[3] shapes (B=8 pairs of 480x640, n=512 queries, D=128): the
training-side correlation kernels against the same maths in stock PyTorch on this GPU.
Dense coarse 30x40, dense fine 120x160, and the 12x16 window on the fine map; forward and
forward+backward.  Prints one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import _torch_ref as R
from posfeat_b200 import _lib
from posfeat_b200 import preprocess as PP

torch.backends.cuda.matmul.allow_tf32 = False
B, n, D = 8, 512, 128
g = torch.Generator().manual_seed(7)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def timeit_graph(fn, iters=50):
    """Device time of one call: the call captured as a CUDA graph and replayed (no host launch overhead)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    return timeit(gr.replay, iters)


# grid <-> grid expectation of Preprocess_Line2Window (n = m = 512 points per image, temperature 20)
qg = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_(True)
kg = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_(True)
vg = torch.rand(B, n, 4, generator=g).cuda()


def gg_ours():
    out = PP.corr_expect(qg, kg, vg, 20.0)
    out.square().sum().backward()


def gg_ref():
    out = R.corr_expect_ref(qg, kg, vg, 20.0)
    out.square().sum().backward()


with torch.no_grad():
    err = float((PP.corr_expect(qg, kg, vg, 20.0) - R.corr_expect_ref(qg, kg, vg, 20.0)).abs().max())
print(json.dumps({"case": "grid_grid_512x512_T20", "B": B, "n": n, "max_abs_err_vs_torch": err,
                  "ours_fwd_bwd_us": timeit(gg_ours), "torch_fwd_bwd_us": timeit(gg_ref)}), flush=True)

for name, (h, w) in (("dense_coarse_30x40", (30, 40)), ("dense_fine_120x160", (120, 160))):
    q = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_(True)
    fmap = torch.nn.functional.normalize(torch.randn(B, D, h, w, generator=g), dim=1).cuda().requires_grad_(True)
    grid = PP.gen_grid(-1, 1, -1, 1, h, w).cuda()
    table = torch.cat([grid, grid ** 2], -1)

    def ours_fwd():
        with torch.no_grad():
            return PP.get_expected_correspondence_locs(q, fmap)

    def ours_fb():
        out = PP.get_expected_correspondence_locs(q, fmap)
        out.square().sum().backward()

    def ref_fwd():
        with torch.no_grad():
            return R.corr_expect_ref(q, fmap.reshape(B, D, h * w).transpose(1, 2), grid, 1.0)

    def ref_fb():
        out = R.corr_expect_ref(q, fmap.reshape(B, D, h * w).transpose(1, 2), grid, 1.0)
        out.square().sum().backward()

    err = float((ours_fwd() - ref_fwd()).abs().max())
    flops = 2.0 * B * n * h * w * D
    r = {"case": name, "B": B, "n": n, "hw": h * w, "max_abs_err_vs_torch": err,
         "ours_fwd_us": timeit(ours_fwd), "torch_fwd_us": timeit(ref_fwd),
         "ours_fwd_bwd_us": timeit(ours_fb), "torch_fwd_bwd_us": timeit(ref_fb)}
    r["ours_fwd_graph_us"], r["torch_fwd_graph_us"] = timeit_graph(ours_fwd), timeit_graph(ref_fwd)
    r["ours_fwd_tflops_fp32"] = flops / (r["ours_fwd_graph_us"] * 1e-6) / 1e12
    print(json.dumps(r), flush=True)

# window on the fine map
h, w = 120, 160
q = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=-1).cuda().requires_grad_(True)
fmap = torch.nn.functional.normalize(torch.randn(B, D, h, w, generator=g), dim=1).cuda().requires_grad_(True)
centre = (torch.rand(B, n, 2, generator=g) * 1.8 - 0.9).cuda()
ws = 0.1
offsets = PP.gen_grid(-ws, ws, -ws, ws, int(ws * h), int(ws * w)).cuda()


def w_ours_fwd():
    with torch.no_grad():
        return PP.get_expected_correspondence_within_window(q, fmap, centre, ws, with_std=True)


def w_ours_fb():
    e, _, s, _ = PP.get_expected_correspondence_within_window(q, fmap, centre, ws, with_std=True)
    (e.square().sum() + s.sum()).backward()


def w_ref_fwd():
    with torch.no_grad():
        return R.window_ref(q, fmap, centre, offsets)


def w_ref_fb():
    e, s, _ = R.window_ref(q, fmap, centre, offsets)
    (e.square().sum() + s.sum()).backward()


err = float((w_ours_fwd()[0] - w_ref_fwd()[0]).abs().max())
r = {"case": f"window_{offsets.shape[0]}taps_fine_120x160", "B": B, "n": n, "max_abs_err_vs_torch": err,
     "ours_fwd_us": timeit(w_ours_fwd), "torch_fwd_us": timeit(w_ref_fwd),
     "ours_fwd_bwd_us": timeit(w_ours_fb), "torch_fwd_bwd_us": timeit(w_ref_fb)}
print(json.dumps(r), flush=True)


# epipolar line search (mode 1: 100 samples on a line, border padding), random fundamental matrices
Fm = torch.randn(B, 3, 3, generator=g).cuda()
coord = torch.stack([torch.rand(B, n, generator=g) * 639, torch.rand(B, n, generator=g) * 479], -1).cuda()
jit = torch.rand(B, n, 2, generator=g).cuda()
qn = q.detach()
fm = fmap.detach()


def line_ours():
    return PP.epipolar_line_search(coord, Fm, qn, fm, 480, 640, jitter=jit)


r = {"case": "epipolar_line_100steps_fine_120x160", "B": B, "n": n, "ours_us": timeit(line_ours)}
print(json.dumps(r), flush=True)

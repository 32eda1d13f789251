"""Training-side correlation + softmax expectation (subsystem 4) behind the
reference's names: drop-ins for ``losses/preprocess.py`` (Preprocess_Line2Window)
and for the correlation helpers of ``losses/preprocess_utils.py``
(get_expected_correspondence_locs :55-82, get_expected_correspondence_within_window
:721-758, epipolar_line_search :662-694, get_endpoints :697-719).

The kernels never materialise the reference's [B,n,m] probability tensors of the
dense stage nor the [B,n,m,D] gathered windows; torch.autograd.Function wrappers
provide the backward passes (also CUDA kernels).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ._runtime import check, lib, ptr, require_cuda, stream_ptr, workspace
from .preprocess_utils import denormalize_coords, normalize_coords, sample_l2norm


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


# ----------------------------------------------------------------- gen_grid
def gen_grid(h_min, h_max, w_min, w_max, len_h, len_w):
    """losses/preprocess_utils.py:84-87: row-major [(x, y)] nodes, [len_h*len_w, 2]."""
    xs = torch.linspace(w_min, w_max, len_w)
    ys = torch.linspace(h_min, h_max, len_h)
    return torch.stack((xs[None, :].expand(len_h, len_w), ys[:, None].expand(len_h, len_w)), -1).reshape(-1, 2).float()


_grid_cache = {}


def _device_grid(h, w, device):
    """[-1, 1] node grid of an h x w map on `device`, built once (the reference rebuilds it on the host
    and copies it on every call, losses/preprocess_utils.py:58-60)."""
    key = (h, w, str(device))
    g = _grid_cache.get(key)
    if g is None:
        g = gen_grid(-1, 1, -1, 1, h, w).to(device)
        g = _grid_cache[key] = (g, torch.cat([g, g ** 2], -1).contiguous())
    return g


def homogenize(coord):
    return torch.cat((coord, torch.ones_like(coord[..., [0]])), -1)


# ------------------------------------------------------ sampler with backward
class SampleFeat(torch.autograd.Function):
    """Bilinear gather (zeros padding, align_corners=False) with a gradient to
    the feature map; the L2 normalisation is applied outside in plain torch so
    autograd differentiates it."""

    @staticmethod
    def forward(ctx, x, coord_n):
        require_cuda()
        coord = _f32c(coord_n)
        ctx.save_for_backward(coord)
        ctx.xmeta = (tuple(x.shape), tuple(x.stride()), x.dtype)
        return sample_l2norm(x.detach(), coord, False)

    @staticmethod
    def backward(ctx, g):
        (coord,) = ctx.saved_tensors
        shape, stride, dtype = ctx.xmeta
        b, c, h, w = shape
        gx = torch.zeros(shape, dtype=torch.float32, device=g.device)
        if tuple(gx.stride()) != stride and stride[1] == 1:
            gx = gx.contiguous(memory_format=torch.channels_last)
        g = _f32c(g)
        with torch.cuda.device(g.device):
            check(lib().posfeat_sample_bwd_f32(g.data_ptr(), b, c, h, w, gx.stride(0), gx.stride(1), gx.stride(2),
                                               gx.stride(3), coord.data_ptr(), coord.shape[1], gx.data_ptr(),
                                               stream_ptr(g.device)))
        return gx.to(dtype), None


def sample_feat_by_coord_grad(x, coord_n, norm=False):
    """Differentiable sample_feat_by_coord (losses/preprocess_utils.py:40-53)."""
    raw = SampleFeat.apply(x, coord_n)
    return F.normalize(raw, p=2, dim=2) if norm else raw


# ------------------------------------------------- scale * normalize, channels-last
class NormalizeScale(torch.autograd.Function):
    """scale * F.normalize(x, p=2, dim=1) of a [B,D,h,w] map, returned channels-last, in one pass each way
    (csrc/normalize.cu) instead of ~5 tensor operations forwards and ~10 backwards over the whole map."""

    @staticmethod
    def forward(ctx, x, scale, eps):
        require_cuda()
        xd = x.detach()
        B, D, h, w = xd.shape
        out = torch.empty((B, D, h, w), dtype=torch.float32, device=xd.device, memory_format=torch.channels_last)
        norm = torch.empty((B, h * w), dtype=torch.float32, device=xd.device)
        with torch.cuda.device(xd.device):
            check(lib().posfeat_normalize_scale_fwd_f32(xd.data_ptr(), B, D, h * w, xd.stride(0), xd.stride(1), float(scale),
                                                        float(eps), out.data_ptr(), norm.data_ptr(), stream_ptr(xd.device)))
        ctx.save_for_backward(xd, norm)
        ctx.scale, ctx.eps = float(scale), float(eps)
        return out

    @staticmethod
    def backward(ctx, g):
        x, norm = ctx.saved_tensors
        B, D, h, w = x.shape
        g = g.detach().to(torch.float32).contiguous(memory_format=torch.channels_last)     # a no-op after our own kernels
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(lib().posfeat_normalize_scale_bwd_f32(g.data_ptr(), x.data_ptr(), norm.data_ptr(), B, D, h * w, x.stride(0),
                                                        x.stride(1), ctx.scale, ctx.eps, gx.data_ptr(), stream_ptr(x.device)))
        return gx, None, None


class SampleAndNormalize(torch.autograd.Function):
    """(sample_feat_by_coord(x, coord, norm=False), scale * F.normalize(x, dim=1) channels-last) of one map as ONE
    autograd node: the backward pass writes the dense gradient of the normalisation and scatters the sampling
    gradient into the same buffer -- no zero-filled map, no map-sized addition to merge the two paths."""

    @staticmethod
    def forward(ctx, x, coord_n, scale, eps):
        require_cuda()
        xd = x.detach()
        coord = _f32c(coord_n)
        B, D, h, w = xd.shape
        raw = sample_l2norm(xd, coord, False)
        out = torch.empty((B, D, h, w), dtype=torch.float32, device=xd.device, memory_format=torch.channels_last)
        norm = torch.empty((B, h * w), dtype=torch.float32, device=xd.device)
        with torch.cuda.device(xd.device):
            check(lib().posfeat_normalize_scale_fwd_f32(xd.data_ptr(), B, D, h * w, xd.stride(0), xd.stride(1), float(scale),
                                                        float(eps), out.data_ptr(), norm.data_ptr(), stream_ptr(xd.device)))
        ctx.save_for_backward(xd, norm, coord)
        ctx.scale, ctx.eps = float(scale), float(eps)
        return raw, out

    @staticmethod
    def backward(ctx, g_raw, g_out):
        x, norm, coord = ctx.saved_tensors
        B, D, h, w = x.shape
        L = lib()
        with torch.cuda.device(x.device):
            if g_out is not None:
                g = g_out.detach().to(torch.float32).contiguous(memory_format=torch.channels_last)
                gx = torch.empty_like(x)
                check(L.posfeat_normalize_scale_bwd_f32(g.data_ptr(), x.data_ptr(), norm.data_ptr(), B, D, h * w, x.stride(0),
                                                        x.stride(1), ctx.scale, ctx.eps, gx.data_ptr(), stream_ptr(x.device)))
            else:
                gx = torch.zeros_like(x)
            if g_raw is not None:
                gr = _f32c(g_raw)
                check(L.posfeat_sample_bwd_f32(gr.data_ptr(), B, D, h, w, gx.stride(0), gx.stride(1), gx.stride(2), gx.stride(3),
                                               coord.data_ptr(), coord.shape[1], gx.data_ptr(), stream_ptr(x.device)))
        return gx, None, None, None


def _fused_map_ok(x):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous() and x.shape[1] <= 512
            and x.shape[0] <= 65535 and x.numel() > 0)


def normalize_scale_channels_last(x, scale=1.0, eps=1e-12):
    """``(scale * F.normalize(x, p=2, dim=1, eps=eps)).contiguous(memory_format=torch.channels_last)`` -- what
    Preprocess_Line2Window hands to the line-search and window kernels.  Contiguous float32 NCHW maps on the device
    take the fused kernels; anything else the tensor expression."""
    if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous() and x.shape[1] <= 1024
            and x.shape[0] <= 65535 and x.numel() > 0):
        return NormalizeScale.apply(x, scale, eps)
    return (scale * F.normalize(x, p=2.0, dim=1, eps=eps)).contiguous(memory_format=torch.channels_last)


# ----------------------------------------------------------- dense expectation
class CorrExpect(torch.autograd.Function):
    """out[b,i,:] = sum_j softmax_j(scale * <q_i,k_j>) v[j,:]   (C <= 4, D <= 128)."""

    @staticmethod
    def forward(ctx, q, k, v, scale):
        require_cuda()
        qd, kd, vd = _f32c(q), _f32c(k), _f32c(v)
        B, n, D = qd.shape
        m = kd.shape[1]
        v_batched = int(vd.dim() == 3)
        C = vd.shape[-1]
        out = torch.empty((B, n, C), dtype=torch.float32, device=qd.device)
        lse = torch.empty((B, n), dtype=torch.float32, device=qd.device)
        with torch.cuda.device(qd.device):
            nbytes = lib().posfeat_corr_expect_workspace_bytes(B, n, m, D, C)
            ws = workspace("corr", nbytes, qd.device) if nbytes else None
            check(lib().posfeat_corr_expect_fwd_f32(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), v_batched, B, n, m, D,
                                                    C, float(scale), out.data_ptr(), lse.data_ptr(), ptr(ws),
                                                    ws.numel() if ws is not None else 0, stream_ptr(qd.device)))
        ctx.save_for_backward(qd, kd, vd, out, lse)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        qd, kd, vd, out, lse = ctx.saved_tensors
        B, n, D = qd.shape
        m = kd.shape[1]
        C = vd.shape[-1]
        g = _f32c(g_out)
        gq = torch.empty_like(qd) if ctx.needs_input_grad[0] else None
        gk = torch.empty_like(kd) if ctx.needs_input_grad[1] else None
        if gq is not None or gk is not None:
            with torch.cuda.device(qd.device):
                nbytes = lib().posfeat_corr_expect_bwd_workspace_bytes(B, n, m, D, C)
                ws = workspace("corr_bwd", nbytes, qd.device) if nbytes else None
                check(lib().posfeat_corr_expect_bwd_f32(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), int(vd.dim() == 3),
                                                        B, n, m, D, C, ctx.scale, out.data_ptr(), lse.data_ptr(),
                                                        g.data_ptr(), ptr(gq), ptr(gk), ptr(ws),
                                                        ws.numel() if ws is not None else 0, stream_ptr(qd.device)))
        return gq, gk, None, None


def corr_expect(q, k, v, scale=1.0, want_lse=False):
    """out[b,i,:] = sum_j softmax_j(scale <q_i,k_j>) v_j.  ``want_lse`` (no autograd): also the per-row
    log-sum-exp of the scaled logits, returned as (out, lse)."""
    if not want_lse:
        return CorrExpect.apply(q, k, v, scale)
    require_cuda()
    qd, kd, vd = _f32c(q), _f32c(k), _f32c(v)
    B, n, D = qd.shape
    m = kd.shape[1]
    C = vd.shape[-1]
    out = torch.empty((B, n, C), dtype=torch.float32, device=qd.device)
    lse = torch.empty((B, n), dtype=torch.float32, device=qd.device)
    with torch.cuda.device(qd.device):
        nbytes = lib().posfeat_corr_expect_workspace_bytes(B, n, m, D, C)
        ws = workspace("corr", nbytes, qd.device) if nbytes else None
        check(lib().posfeat_corr_expect_fwd_f32(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), int(vd.dim() == 3), B, n, m, D,
                                                C, float(scale), out.data_ptr(), lse.data_ptr(), ptr(ws),
                                                ws.numel() if ws is not None else 0, stream_ptr(qd.device)))
    return out, lse


class ComputeProb(torch.autograd.Function):
    """softmax over the last axis of the (scaled) similarity / negative squared distance, materialised by
    posfeat_compute_prob_f32; the backward pass is the softmax Jacobian followed by two batched products on
    the saved probabilities (plain tensor ops: nothing on the training path differentiates through this
    stand-alone form, the fused expectations have their own backward kernels)."""

    @staticmethod
    def forward(ctx, feat1, feat2, mode, scale, want_sim):
        require_cuda()
        a, b = _f32c(feat1), _f32c(feat2)
        B, m, D = a.shape
        n = b.shape[1]
        prob = torch.empty((B, m, n), dtype=torch.float32, device=a.device)
        sim = torch.empty((B, m, n), dtype=torch.float32, device=a.device) if want_sim else None
        with torch.cuda.device(a.device):
            check(lib().posfeat_compute_prob_f32(a.data_ptr(), b.data_ptr(), B, m, n, D, int(mode), float(scale),
                                                 prob.data_ptr(), ptr(sim), stream_ptr(a.device)))
        ctx.save_for_backward(a, b, prob)
        ctx.mode, ctx.scale = int(mode), float(scale)
        return prob, sim

    @staticmethod
    def backward(ctx, g_prob, g_sim):
        a, b, prob = ctx.saved_tensors
        g = g_prob.to(torch.float32)
        gl = prob * (g - (g * prob).sum(-1, keepdim=True))            # d loss / d logits
        if g_sim is not None:                                          # return_sim: sim = raw <a,b> also carries gradient
            gl_sim = g_sim.to(torch.float32)
        else:
            gl_sim = None
        if ctx.mode == 0:
            gs = ctx.scale * gl if gl_sim is None else ctx.scale * gl + gl_sim
            ga = gs @ b if ctx.needs_input_grad[0] else None
            gb = gs.transpose(1, 2) @ a if ctx.needs_input_grad[1] else None
        else:   # logits = -(|a|^2 + |b|^2 - 2 a.b)
            ga = (2 * (gl @ b) - 2 * a * gl.sum(2, keepdim=True)) if ctx.needs_input_grad[0] else None
            gb = (2 * (gl.transpose(1, 2) @ a) - 2 * b * gl.sum(1).unsqueeze(-1)) if ctx.needs_input_grad[1] else None
        return ga, gb, None, None, None


def compute_prob(feat1, feat2, loss_distance='cos', with_scale=False, return_sim=False):
    """losses/preprocess_utils.py:89-115: feat1 [B,m,d], feat2 [B,n,d] -> prob [B,m,n]
    (softmax over n of scale*<f1,f2> for 'cos', of -|f1-f2|^2 for 'euc'); ``return_sim`` also returns
    the raw similarities ('cos' only).  Same assertion behaviour as the reference."""
    assert loss_distance in ['cos', 'euc']
    if return_sim:
        assert loss_distance == 'cos'
    if feat1.dim() != 3 or feat2.dim() != 3 or feat1.shape[0] != feat2.shape[0] or feat1.shape[2] != feat2.shape[2]:
        raise RuntimeError(f"compute_prob: expected [B,m,d] and [B,n,d], got {tuple(feat1.shape)} {tuple(feat2.shape)}")
    mode = 0 if loss_distance == 'cos' else 1
    # the reference builds the scale as a float32 tensor: sqrt is taken in float32
    scale = float(torch.tensor(float(feat2.shape[1]), dtype=torch.float32).sqrt()) if (with_scale and mode == 0) else 1.0
    prob, sim = ComputeProb.apply(feat1, feat2, mode, scale, bool(return_sim))
    out_dtype = feat1.dtype if feat1.dtype.is_floating_point else torch.float32
    if return_sim:
        return prob.to(out_dtype), sim.to(out_dtype)
    return prob.to(out_dtype)


def get_expected_correspondence_locs(feat1, featmap2, with_std=False):
    """losses/preprocess_utils.py:55-82.  feat1 [B,n,d], featmap2 [B,d,h,w] ->
    expected normalised xy [B,n,2]; with_std: (xy, std [B,n], kurtosis [B,n], prob [B,n,hw])."""
    B, d, h2, w2 = featmap2.shape
    grid, table = _device_grid(h2, w2, featmap2.device)                   # [hw, 2]; x, y, x^2, y^2
    keys = featmap2.reshape(B, d, h2 * w2).transpose(1, 2)                # [B, hw, d]
    out = corr_expect(feat1, keys, table, 1.0)
    exp_xy = out[..., :2]
    if not with_std:
        return exp_xy
    var = out[..., 2:] - exp_xy ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    # prob / kurtosis are auxiliary outputs nothing on the training path consumes; prob is materialised by
    # the stand-alone compute_prob kernel only because the reference returns it
    prob = compute_prob(feat1, keys)
    kurt = torch.pow(grid[None, None] - exp_xy.unsqueeze(-2), 4).mean(-2) / torch.pow(var, 2)
    return exp_xy, std, (kurt / 10.).clamp(0, 1).mean(-1), prob


# ---------------------------------------------------------- window expectation
class WindowExpect(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, fmap, centre, offsets):
        require_cuda()
        qd, cd, od = _f32c(q), _f32c(centre), _f32c(offsets)
        fd = fmap.detach()
        if fd.dtype != torch.float32:
            fd = fd.float()
        # The gather kernels read one 512-byte descriptor per tap when channels are innermost; an NCHW
        # map would cost one 32-byte sector per channel.  One layout pass over the map (a no-op when the
        # backbone already runs channels_last) is far cheaper than that; the gradient comes back in the
        # same strides.
        if fd.shape[1] >= 32:
            fd = fd.contiguous(memory_format=torch.channels_last)
        elif not fd.is_contiguous():
            fd = fd.contiguous()
        B, D, h, w = fd.shape
        n, m = qd.shape[1], od.shape[0]
        dev = qd.device
        exp_xy = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
        std = torch.empty((B, n), dtype=torch.float32, device=dev)
        prob = torch.empty((B, n, m), dtype=torch.float32, device=dev)
        lse = torch.empty((B, n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().posfeat_window_expect_fwd_f32(fd.data_ptr(), B, D, h, w, fd.stride(0), fd.stride(1), fd.stride(2),
                                                      fd.stride(3), qd.data_ptr(), cd.data_ptr(), n, od.data_ptr(), m, 0,
                                                      exp_xy.data_ptr(), std.data_ptr(), prob.data_ptr(), lse.data_ptr(),
                                                      stream_ptr(dev)))
        ctx.save_for_backward(qd, fd, cd, od, exp_xy, prob)
        ctx.mark_non_differentiable(prob)
        return exp_xy, std, prob

    @staticmethod
    def backward(ctx, g_exp, g_std, _g_prob):
        qd, fd, cd, od, exp_xy, prob = ctx.saved_tensors
        B, D, h, w = fd.shape
        n, m = qd.shape[1], od.shape[0]
        dev = qd.device
        g_exp = _f32c(g_exp) if g_exp is not None else torch.zeros_like(exp_xy)
        g_std = _f32c(g_std) if g_std is not None else torch.zeros((B, n), dtype=torch.float32, device=dev)
        gq = torch.empty_like(qd)
        gf = torch.zeros_like(fd)          # preserves the memory format of fmap
        with torch.cuda.device(dev):
            check(lib().posfeat_window_expect_bwd_f32(fd.data_ptr(), B, D, h, w, gf.stride(0), gf.stride(1), gf.stride(2),
                                                      gf.stride(3), qd.data_ptr(), cd.data_ptr(), n, od.data_ptr(), m,
                                                      exp_xy.data_ptr(), prob.data_ptr(), g_exp.data_ptr(),
                                                      g_std.data_ptr(), gq.data_ptr(), gf.data_ptr(), stream_ptr(dev)))
        return gq, gf, None, None


def get_expected_correspondence_within_window(feat1, featmap2, coord2_n, window_size, with_std=False, with_sim=False):
    """losses/preprocess_utils.py:721-758.  Returns (expected xy [B,n,2], window
    coords [B,n,m,2][, std [B,n], prob [B,n,m]]) like the reference."""
    if with_sim:
        raise NotImplementedError("with_sim=True (visualisation only) is not supported")
    B, d, h2, w2 = featmap2.shape
    okey = ("win", float(window_size), h2, w2, str(coord2_n.device), coord2_n.dtype)
    offsets = _grid_cache.get(okey)
    if offsets is None:       # built once: a per-call host->device copy would also forbid CUDA-graph capture of the step
        offsets = _grid_cache[okey] = gen_grid(-window_size, window_size, -window_size, window_size,
                                               int(window_size * h2), int(window_size * w2)).to(coord2_n)
    exp_xy, std, prob = WindowExpect.apply(feat1, featmap2, coord2_n, offsets)
    coord_grid = coord2_n.unsqueeze(-2) + offsets[None, None]
    if with_std:
        return exp_xy, coord_grid, std, prob
    return exp_xy, coord_grid


# ---------------------------------------------------------- epipolar line search
@torch.no_grad()
def get_endpoints(coords, Fmat, h, w):
    """losses/preprocess_utils.py:697-719: clip each epipolar line F x~ to the image
    rectangle; returns two normalised endpoints [B,n,2] and a validity mask [B,n]
    (exactly two of the four border intersections inside the image)."""
    B, n, _ = coords.shape
    line = Fmat.bmm(homogenize(coords).transpose(1, 2))
    a, b, c = line[:, 0, :], line[:, 1, :], line[:, 2, :]
    zero, wm, hm = torch.zeros_like(a), (w - 1) * torch.ones_like(a), (h - 1) * torch.ones_like(a)
    pts = torch.stack([torch.stack([zero, -c / b], -1), torch.stack([wm, -(a * (w - 1) + c) / b], -1),
                       torch.stack([-(b * (h - 1) + c) / a, hm], -1), torch.stack([-c / a, zero], -1)], 2)  # B,n,4,2
    inside = (pts[..., 0] >= 0) & (pts[..., 0] <= w - 1) & (pts[..., 1] >= 0) & (pts[..., 1] <= h - 1)
    valid = inside.sum(-1) == 2
    inside[~valid] = torch.tensor([True, True, False, False], device=inside.device)
    two = pts[inside].reshape(B, n, 2, 2)
    return normalize_coords(two[:, :, 0, :], h, w), normalize_coords(two[:, :, 1, :], h, w), valid


@torch.no_grad()
def epipolar_line_search(coord, Fmat, feat1, featmap2, h, w, line_step=100, use_nn=True, loc_rand=True,
                         window_size=0.125, visualize=False, jitter=None):
    """losses/preprocess_utils.py:662-694.  The [B,n,line_step,D] gathered samples
    are never materialised (kernel mode 1: line between the two endpoints, border
    padding).  ``jitter`` ([B,n,2] uniform in [0,1)) replaces the internal
    torch.rand draw so runs can be reproduced."""
    B, d, h2, w2 = featmap2.shape
    n = coord.shape[1]
    qd = _f32c(feat1)
    cd = _f32c(coord)
    Fd = _f32c(Fmat).reshape(B, 9)
    fd = featmap2.detach().float()
    if d >= 32:
        fd = fd.contiguous(memory_format=torch.channels_last)     # one 512-byte read per tap (see WindowExpect)
    elif not fd.is_contiguous():
        fd = fd.contiguous()
    dev = qd.device
    ends = torch.empty((B, n, 4), dtype=torch.float32, device=dev)
    valid_u8 = torch.empty((B, n), dtype=torch.uint8, device=dev)
    exp_soft = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
    nn_xy = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
    m2 = torch.empty((B, n, 2), dtype=torch.float32, device=dev)
    prob = torch.empty((B, n, line_step), dtype=torch.float32, device=dev) if visualize else None
    # one launch: endpoints (get_endpoints), the line_step samples, softmax, and the reductions the
    # reference takes over the [B,n,line_step] tensors afterwards
    with torch.cuda.device(dev):
        check(lib().posfeat_line_search_f32(fd.data_ptr(), B, d, h2, w2, fd.stride(0), fd.stride(1), fd.stride(2),
                                            fd.stride(3), qd.data_ptr(), cd.data_ptr(), Fd.data_ptr(), n, int(h), int(w),
                                            int(line_step), ends.data_ptr(), valid_u8.data_ptr(), exp_soft.data_ptr(),
                                            nn_xy.data_ptr(), m2.data_ptr(), ptr(prob), stream_ptr(dev)))
    valid = valid_u8.bool()
    expected = nn_xy if use_nn else exp_soft          # nn: sum of the positions whose probability is the maximum
    expected_org = expected
    if loc_rand:
        u = jitter if jitter is not None else torch.rand(expected.shape)
        expected = expected + 0.707 * window_size * (2 * u.type_as(expected) - 1)
    border = (expected[:, :, 0] >= -1) & (expected[:, :, 0] <= 1) & (expected[:, :, 1] >= -1) & (expected[:, :, 1] <= 1)
    valid = valid & border
    var = m2 - expected ** 2                            # m2 = sum_p prob_p * grid_p^2 (from the kernel)
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    if visualize:
        return expected, expected_org, valid, std, prob
    return expected, expected_org, valid, std


# ------------------------------------------------------- keypoint generator (input side)
@torch.no_grad()
def generate_kpts_regular_grid_random(inputs, outputs, grid_size, map_init="identity", keep_spatial=False,
                                      random_select="random"):
    """losses/preprocess_utils.py:598-659, the configuration the shipped yaml uses
    (map_init 'identity', random_select 'random'): one uniformly random pixel of
    every grid_size x grid_size cell of the full-resolution map.  It only produces
    the query coordinates for the correlation stage."""
    if map_init != "identity" or random_select != "random":
        raise NotImplementedError("only map_init='identity', random_select='random' (configs/train_desc.yaml:66-69)")

    def one(kp_map):
        b, _, h, w = kp_map.shape
        gh, gw = h // grid_size, w // grid_size
        pick = torch.randint(0, grid_size * grid_size, (b, gh, gw), device=kp_map.device)
        yy = torch.arange(gh, device=kp_map.device)[None, :, None] * grid_size + pick // grid_size
        xx = torch.arange(gw, device=kp_map.device)[None, None, :] * grid_size + pick % grid_size
        xs = torch.linspace(-1, 1, w, device=kp_map.device)
        ys = torch.linspace(-1, 1, h, device=kp_map.device)
        kps = torch.stack([xs[xx], ys[yy]], -1)                                            # b,gh,gw,2
        return kps, torch.ones(b, gh, gw, 1, device=kp_map.device)
    k1, s1 = one(outputs["preds1"]["local_point"])
    k2, s2 = one(outputs["preds2"]["local_point"])
    if not keep_spatial:
        b = k1.shape[0]
        return k1.reshape(b, -1, 2), k2.reshape(b, -1, 2), s1.reshape(b, -1, 1), s2.reshape(b, -1, 1)
    return k1, k2, s1, s2


# ------------------------------------------------------------- the nn.Module
class Preprocess_Line2Window(nn.Module):
    """Drop-in for losses/preprocess.py:7-118 (same config keys, same 15-key output
    dict).  ``coords`` / ``jitter`` keyword arguments inject the random draws."""

    def __init__(self, configs, device=None, vis=False):
        super().__init__()
        self.__lossname__ = "Preprocess_Line2Window"
        self.config = configs
        if configs["kps_generator"] != "generate_kpts_regular_grid_random":
            raise NotImplementedError(f"kps_generator {configs['kps_generator']!r}")
        self.kps_generator = generate_kpts_regular_grid_random
        self.t_base = configs["temperature_base"]
        self.t_max = configs["temperature_max"]
        if device is not None:
            self.device = device

    def name(self):
        return self.__lossname__

    def forward(self, inputs, outputs, coords=None, jitter=None):
        if self.config.get("use_nn_grid"):
            raise NotImplementedError("use_nn_grid=True references an undefined name in the reference (preprocess.py:68)")
        preds1, preds2 = outputs["preds1"], outputs["preds2"]
        xf1, xf2 = preds1["local_map"], preds2["local_map"]
        h1i, w1i = inputs["im1"].size()[2:]
        h2i, w2i = inputs["im2"].size()[2:]
        b = xf1.shape[0]
        T = min(self.t_base + outputs["epoch"], self.t_max)
        cos = self.config["loss_distance"] == "cos"

        if coords is None:
            c1n, c2n, _, _ = self.kps_generator(inputs, outputs, **self.config["kps_generator_config"])
        else:
            c1n, c2n = coords
        c1n = c1n.reshape(b, -1, 2).to(xf1.device).float()
        c2n = c2n.reshape(b, -1, 2).to(xf1.device).float()
        coord1 = denormalize_coords(c1n, h1i, w1i)
        coord2 = denormalize_coords(c2n, h2i, w2i)

        # :56-57 and the normalised maps of :84-106, one autograd node per map when the fused kernels apply
        fused = _fused_map_ok(xf1) and _fused_map_ok(xf2)
        if fused:
            r1, m1 = SampleAndNormalize.apply(xf1, c1n, T, 1e-12)
            r2, m2 = SampleAndNormalize.apply(xf2, c2n, T, 1e-12)
            f1 = F.normalize(r1, p=2, dim=2) if cos else r1
            f2 = F.normalize(r2, p=2, dim=2) if cos else r2
        else:
            f1 = sample_feat_by_coord_grad(xf1, c1n, cos)
            f2 = sample_feat_by_coord_grad(xf2, c2n, cos)

        # grid <-> grid softmax expectations (:59-63, :76-81) without the [b,m,n] tensors
        o1 = corr_expect(f1, f2, torch.cat([coord2, c2n ** 2], -1), T)      # rows: softmax over image-2 points
        o2 = corr_expect(f2, f1, torch.cat([coord1, c1n ** 2], -1), T)      # cols: softmax over image-1 points
        feat1g_corloc, feat2g_corloc = o1[..., :2], o2[..., :2]
        with torch.no_grad():
            feat1g_corloc_n = normalize_coords(feat1g_corloc, h2i, w2i)
            feat2g_corloc_n = normalize_coords(feat2g_corloc, h1i, w1i)
        feat1g_std = (o1[..., 2:] - feat1g_corloc_n ** 2).clamp(min=1e-6).sqrt().sum(-1)
        feat2g_std = (o2[..., 2:] - feat2g_corloc_n ** 2).clamp(min=1e-6).sqrt().sum(-1)

        # channels innermost: the line and window kernels then read whole descriptors (converted once here)
        if not fused:
            m2 = normalize_scale_channels_last(xf2, T)
            m1 = normalize_scale_channels_last(xf1, T)
        if self.config["use_line_search"]:
            j1, j2 = jitter if jitter is not None else (None, None)
            ws = self.config["window_size"]
            c1_, c1_org, valid1, _ = epipolar_line_search(coord1, inputs["F1"], f1, m2, h2i, w2i, window_size=ws,
                                                          jitter=j1, **self.config["line_search_config"])
            c2_, c2_org, valid2, _ = epipolar_line_search(coord2, inputs["F2"], f2, m1, h1i, w1i, window_size=ws,
                                                          jitter=j2, **self.config["line_search_config"])
            feat1c_corloc_org = denormalize_coords(c1_org, h2i, w2i)
            feat2c_corloc_org = c2_org            # the reference returns the normalised tensor here (:113)
        else:
            c1_, c2_ = feat1g_corloc_n.detach(), feat2g_corloc_n.detach()
            feat1c_corloc_org, feat2c_corloc_org = c1_, c2_
            valid1 = torch.ones_like(feat1g_std).bool()
            valid2 = torch.ones_like(feat2g_std).bool()

        w1n, _, feat1w_std, _ = get_expected_correspondence_within_window(f1, m2, c1_, self.config["window_size"], True)
        w2n, _, feat2w_std, _ = get_expected_correspondence_within_window(f2, m1, c2_, self.config["window_size"], True)
        return {"coord1": coord1, "coord2": coord2,
                "feat1g_corloc": feat1g_corloc, "feat2g_corloc": feat2g_corloc,
                "feat1w_corloc": denormalize_coords(w1n, h2i, w2i), "feat2w_corloc": denormalize_coords(w2n, h1i, w1i),
                "feat1c_corloc_org": feat1c_corloc_org, "feat2c_corloc_org": feat2c_corloc_org,
                "feat1g_std": feat1g_std, "feat2g_std": feat2g_std,
                "feat1w_std": feat1w_std, "feat2w_std": feat2w_std,
                "temperature": T, "valid_epi1": valid1, "valid_epi2": valid2}


class Preprocess_Skip(nn.Module):
    """losses/preprocess.py:120-129."""

    def __init__(self, **kargs):
        super().__init__()
        self.__lossname__ = "Preprocess_Skip"

    def forward(self, inputs, outputs):
        return None

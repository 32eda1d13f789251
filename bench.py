#!/usr/bin/env python
"""bench.py -- matched image pairs/s through the post-backbone feature pipeline.

One "step" = `--passes` passes of the hot path (NMS + threshold + top-k keypoints, descriptor sampling +
L2 norm, mutual-NN matching) over a resident batch of `--pairs` synthetic HPatches-shaped image pairs
(BASELINE.json configs[1]: 896x1200 score maps, 128x224x300 descriptor maps, 8192 keypoints, D=128), so
that the timed region of `--steps 20` lasts about a second instead of 54 ms.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  torchrun ... bench.py --gpus N ...      (one rank per GPU; pairs are sharded, no collective)

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same metric through the
public host-buffer call (PairPipeline.run_host) with the H2D/D2H copies inside the timed region,
`roofline` the tcgen05 matcher kernel against the measured bf16 peak, `cpu_baseline` the reference's own
torch-CPU functions (staged under oracle/_ref by oracle/build_ref.py; the numpy port when they are absent)
on the host cores.  `--impl reference` times that CPU path only.

Other workloads (`--workload`; the default line also carries a short run of each under `workloads`):
  c1         BASELINE configs[0]: one 480x640 pair, 4096 keypoints (latency of the single-pair call)
  aachen     configs[2]: 1056x1600 maps, r=3, thr 0.5 (configs/extract_aachen.yaml:27-40), one query
             against 20 retrieved database images, extraction amortised 1:20
  mnn_sweep  configs[4]: matcher N = M in {1k, 4k, 16k, 64k} against the tensor-pipe roofline
  train_step configs[3]: B=8 640x480 pairs, 512 queries, Preprocess_Line2Window + EpipolarLoss_full
             forward+backward and the bucketed NCCL gradient all-reduce (82 MB) overlapped with it
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, D, KPTS = 896, 1200, 128, 8192
DET_CFG = dict(nms_radius=1, num_pts=KPTS, thr=0.9, thr_mod="abs", use_nms=True, stable=True)
WORKLOAD = "hpatches_pairs_896x1200_k8192_d128"
C1_CFG = dict(nms_radius=1, num_pts=4096, thr=0.9, thr_mod="abs", use_nms=True, stable=True)      # configs/extract_hpatches.yaml:26-33 at k=4096
AACHEN_CFG = dict(nms_radius=3, num_pts=20480, thr=0.5, thr_mod="abs", use_nms=True, stable=True)  # configs/extract_aachen.yaml:27-40


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "bf16": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


def ncu_traffic(kernel):
    """DRAM read+write bytes per pair of `kernel` from the committed ncu --set full summary
    (profiles/ncu_traffic.json, written by tools/collect_profiles.py); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        return t[kernel]
    except Exception:
        return None


def synth_pairs(P, seed, device, layout="channels_last", h=H, w=W):
    """Synthetic post-backbone maps: score = softplus(randn),
    descriptor map = randn; image 2i+1 is a noisy copy of image 2i so that the
    pairs have plenty of true matches.  The descriptor map has the reference's
    logical shape [2P, D, H/4, W/4]; `layout` picks its memory format:
    channels_last (what a channels_last backbone emits; the sampler then reads
    128-bit vectors) or nchw (the reference's default contiguous layout)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    score = torch.empty(2 * P, 1, h, w)
    fmap = torch.empty(2 * P, D, h // 4, w // 4)
    for i in range(P):
        s = torch.randn(1, h, w, generator=g)
        f = torch.randn(D, h // 4, w // 4, generator=g)
        score[2 * i] = torch.nn.functional.softplus(s)
        score[2 * i + 1] = torch.nn.functional.softplus(s + 0.05 * torch.randn(1, h, w, generator=g))
        fmap[2 * i] = f
        fmap[2 * i + 1] = f + 0.3 * torch.randn(D, h // 4, w // 4, generator=g)
    fmap = fmap.to(device)
    if layout == "channels_last":
        fmap = fmap.contiguous(memory_format=torch.channels_last)
    return score.to(device), fmap


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- the reference on the host cores
def ref_available():
    from oracle import ref_runner
    return ref_runner.available()


def cpu_reference_pair(score2, fmap2, cfg=None, need_idx=True):
    """ONE pair through the reference's own torch-CPU functions (oracle/_ref) -- or, when the staged
    reference is absent, through the numpy port (oracle/posfeat_oracle.py).  Returns (kind, idx [2,n],
    matches (K,2))."""
    cfg = cfg or DET_CFG
    if ref_available():
        from oracle import ref_runner
        if not need_idx:                       # timed calls: the stock code path, nothing observed
            kps, desc, m = ref_runner.run_pair(score2, fmap2, dict(cfg))
            return "reference", None, m
        kps, desc, m, idx = ref_runner.run_pair(score2, fmap2, dict(cfg), want_idx=True)
        return "reference", idx.cpu().numpy(), m
    from oracle import posfeat_oracle as O
    sn, fn = score2.numpy(), fmap2.numpy()
    kps, sc, idx, _ = O.generate_kpts_single(sn, return_idx=True, **{k: v for k, v in cfg.items() if k != "stable"})
    desc = O.sample_feat_by_coord(fn, kps, True)
    return "port", idx, O.mnn_matcher(desc[0], desc[1])


def pixel_matches(idx2, m):
    """Match list as a set of (interior pixel in image a, interior pixel in image b): independent of how
    equal-score keypoints are ordered (torch.topk leaves that open)."""
    return {(int(idx2[0][i]), int(idx2[1][j])) for i, j in m}


def time_cpu_reference(score, fmap, budget_s=15.0, max_pairs=64):
    """Bounded sample of the same workload on the host cores: the first pairs of the batch, one at a time,
    as the reference's scripts run them (batch_size 1)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    score, fmap = score[:8].cpu(), fmap[:8].cpu().contiguous()       # the reference's maps are NCHW-contiguous
    P = score.shape[0] // 2
    kind, _, _ = cpu_reference_pair(score[:2], fmap[:2])             # warm-up (thread pool, allocator)
    done, t0, first = 0, time.perf_counter(), None
    while done < max_pairs and (done == 0 or time.perf_counter() - t0 < budget_s):
        i = done % P
        cpu_reference_pair(score[2 * i:2 * i + 2], fmap[2 * i:2 * i + 2], need_idx=False)
        done += 1
    dt = time.perf_counter() - t0
    first = cpu_reference_pair(score[0:2], fmap[0:2])          # untimed: pair 0 with the keypoint indices, for the parity check
    return kind, done / dt, done, dt, first


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    score, fmap = synth_pairs(2, 1234, "cpu", layout="nchw")
    for _ in range(max(args.warmup, 1)):
        kind, _, _ = cpu_reference_pair(score[0:2], fmap[0:2])
    steps = max(1, min(args.steps, 100))        # each step = 1 pair (a bounded sample of the 64-pair workload)
    t0 = time.perf_counter()
    for s in range(steps):
        i = s % 2
        cpu_reference_pair(score[2 * i:2 * i + 2], fmap[2 * i:2 * i + 2], need_idx=False)
    dt = time.perf_counter() - t0
    v = steps / dt
    what = ("the reference's generate_kpts_single / sample_feat_by_coord / mnn_matcher (oracle/_ref) on CPU tensors"
            if kind == "reference" else "numpy port of the reference (oracle/posfeat_oracle.py; oracle/_ref not staged)")
    out = {"impl": "reference", "metric": "matched image pairs/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
           "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": 1, "keypoints": KPTS, "descriptor_dim": D,
                      "detector": dict(DET_CFG), "note": "each step is one pair of the workload (bounded sample)"},
           "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": os.cpu_count(), "threads": torch.get_num_threads(),
                            "kind": kind, "sample": f"{steps} pairs, {what}"},
           "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- helpers
def cuda_time(fn, iters, warmup=2):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, r


def kernel_times(fn, iters):
    """Per-kernel CUDA-event durations (library slots), ms per call of fn."""
    import torch
    from posfeat_b200 import _lib
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    return {k: v[0] / iters for k, v in prof.items()}


E2E_PATHS = {
    "stage": "score maps copied whole; descriptor-map pixels under the keypoints' taps fetched once each from pinned host "
             "memory (posfeat_fetch_taps_f32; bytes = pixels moved, counted by the kernel, x 512 B)",
    "direct": "score maps copied whole; the sampler reads its taps from pinned host memory (bytes requested = 4 taps x 512 B "
              "per keypoint)",
    False: "score and descriptor maps copied whole",
}


# --------------------------------------------------------------------------- other workloads
def workload_c1(dev, iters=200):
    """BASELINE configs[0]: ONE 480x640 pair, 4096 keypoints -- the latency of the single-pair call."""
    import torch
    from posfeat_b200.pairs import PairPipeline
    score, fmap = synth_pairs(1, 4321, dev, h=480, w=640)
    pipe = PairPipeline(C1_CFG)
    ms, (feats, matches, nm) = cuda_time(lambda: pipe.run(score, fmap), iters, warmup=5)
    ker = kernel_times(lambda: pipe.run(score, fmap), 20)
    out = {"workload": "c1_pair_480x640_k4096", "ms_per_pair": ms, "pairs_per_s": 1e3 / ms, "keypoints": int(feats["n"]),
           "matches": int(nm[0]), "kernel_ms": {k: round(v, 5) for k, v in ker.items()}}
    try:        # the same call captured in a CUDA graph (one submission instead of ~12 launches and a mid-call host wait)
        from posfeat_b200.pairs import GraphedPairPipeline
        gp = GraphedPairPipeline(pipe, score.shape, fmap.shape)
        ms_g, (gf, gm, gnm) = cuda_time(lambda: gp(score, fmap), iters, warmup=5)
        out["cuda_graph"] = {"ms_per_pair": ms_g, "pairs_per_s": 1e3 / ms_g,
                             "equal_plain_call": bool(torch.equal(gf["idx"], feats["idx"]) and torch.equal(gnm, nm) and
                                                      torch.equal(gm[0, :int(nm[0])], matches[0, :int(nm[0])]))}
    except Exception as e:
        out["cuda_graph"] = {"error": repr(e)[:200]}
    try:
        kind, idx_r, m_r = cpu_reference_pair(score.cpu(), fmap.cpu().contiguous(), C1_CFG)
        k0 = int(nm[0])
        ours = pixel_matches(feats["idx"].cpu().numpy(), matches[0, :k0].cpu().numpy())
        out["parity_vs_" + kind] = {"keypoint_sets_equal": bool(all(set(a.tolist()) == set(b.tolist()) for a, b in
                                                                    zip(feats["idx"].cpu().numpy(), idx_r))),
                                    "matches_differing": len(ours ^ pixel_matches(idx_r, m_r))}
    except Exception as e:      # the checker must not take the measurement down
        out["parity_error"] = repr(e)[:200]
    return out


def workload_aachen(dev, n_img=32, n_ret=20, iters=3, num_pts=None):
    """BASELINE configs[2]: 1056x1600 maps, detector of configs/extract_aachen.yaml:27-40 (r=3, thr 0.5 abs), every
    image extracted once and matched against its 20 retrieved neighbours (image q vs images q+1..q+20 of the ring)."""
    import torch
    from posfeat_b200.pairs import PairPipeline
    h, w = 1056, 1600
    cfg = dict(AACHEN_CFG)
    if num_pts:
        cfg["num_pts"] = int(num_pts)
    score, fmap = synth_pairs(n_img // 2, 777, dev, h=h, w=w)
    pipe = PairPipeline(cfg)

    def step():
        f = pipe.extract(score, fmap)
        desc = f["desc"]
        n = desc.shape[1]
        tot = 0
        outs = []
        for q in range(n_img):
            lo, hi = q + 1, q + 1 + n_ret
            segs = [(lo, min(hi, n_img))] + ([(0, hi - n_img)] if hi > n_img else [])
            for a, b in segs:
                if b > a:
                    outs.append(pipe.match_views(desc[q].unsqueeze(0).expand(b - a, n, desc.shape[2]), desc[a:b]))
                    tot += b - a
        return f, outs, tot

    ms, (f, outs, tot) = cuda_time(step, iters, warmup=3)
    ker = kernel_times(step, 2)
    n = int(f["n"])
    nms_ms = ker.get("nms_candidates")
    pk = peaks()
    out = {"workload": f"aachen_1056x1600_r3_thr0.5_k{cfg['num_pts']}_ret{n_ret}", "images_per_step": n_img,
           "pairs_per_step": tot, "keypoints": n, "ms_per_step": ms, "pairs_per_s": tot / (ms * 1e-3),
           "images_per_s": n_img / (ms * 1e-3), "mean_matches": float(torch.cat([o[1] for o in outs]).float().mean()),
           "kernel_ms": {k: round(v, 4) for k, v in ker.items()}}
    if nms_ms:
        nb = 4.0 * n_img * h * w
        out["nms_hbm_frac"] = nb / (nms_ms * 1e-3) / 1e9 / pk["hbm"]
        sel = ker.get("select_topk", 0.0) + ker.get("keypoint_outputs", 0.0)
        out["nms_topk_hbm_frac"] = (nb + 20.0 * n * n_img) / ((nms_ms + sel) * 1e-3) / 1e9 / pk["hbm"]
    return out


def workload_mnn_sweep(dev, sizes=(1024, 4096, 16384, 65536)):
    """BASELINE configs[4]: single-pair matcher calls against the tensor-pipe roofline."""
    import torch
    from posfeat_b200 import _lib
    from posfeat_b200.preprocess_utils import mnn_match
    pk = peaks()
    rows = []
    for n in sizes:
        g = torch.Generator().manual_seed(n)
        a = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
        b = torch.nn.functional.normalize(a[torch.randperm(n, generator=g)] + 0.06 * torch.randn(n, 128, generator=g), dim=1)
        a, b = a.to(dev), b.to(dev)
        iters = 50 if n <= 4096 else (20 if n <= 16384 else 5)
        # (the 64k call touches gigabytes of fresh workspace: its first three calls are 10x / 8 % / 2 % slower)
        ms, (m, nm, _, _) = cuda_time(lambda: mnn_match(a, b, algo=2, want_nn21=False), iters, warmup=4)
        ker = kernel_times(lambda: mnn_match(a, b, algo=2, want_nn21=False), iters)
        flops = 2.0 * n * n * 128
        tc = ker.get("mnn_tc")
        graph_us = None
        if n <= 16384:                         # launch-bound sizes: the same call replayed as one CUDA graph
            from posfeat_b200.matchers import GraphedMatcher
            gm = GraphedMatcher(n, n, 128, algo=2, device=dev)
            gm.a.copy_(a); gm.b.copy_(b)
            gms, _ = cuda_time(gm.graph.replay, iters, warmup=2)
            assert torch.equal(gm.matches[:int(gm.nm.item())], m[:int(nm.item())])
            graph_us = 1e3 * gms
            del gm
        rows.append({"N": n, "us_per_call": 1e3 * ms, "cuda_graph_us_per_call": graph_us, "frac_of_bf16_peak_whole_call": flops / (ms * 1e-3) / 1e12 / pk["bf16"],
                     "tc_kernel_us": 1e3 * tc if tc else None,
                     "tc_kernel_frac_of_bf16_peak": flops / (tc * 1e-3) / 1e12 / pk["bf16"] if tc else None,
                     "kernel_us": {k: round(1e3 * v, 1) for k, v in ker.items()}, "matches": int(nm.item())})
        del a, b
    return {"workload": "mnn_sweep_d128_matches_only", "rows": rows}


def workload_train_step(dev, world, iters=10):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_step_bench
    return train_step_bench.run(dev, world, iters)


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hpatches", choices=["hpatches", "c1", "aachen", "mnn_sweep", "train_step"])
    ap.add_argument("--pairs", type=int, default=64, help="pairs of the resident batch, per GPU")
    ap.add_argument("--passes", type=int, default=20, help="passes over the resident batch per step (device-resident leg)")
    ap.add_argument("--e2e-passes", type=int, default=1, help="run_host calls per step of the e2e leg")
    ap.add_argument("--streams", type=int, default=1, help="CUDA streams the pair batch is split over (1 = plain path)")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="run_host groups (0 = the call's default)")
    ap.add_argument("--no-host-gather", action="store_true",
                    help="e2e leg: copy the whole descriptor map to the device instead of gathering taps over the host link")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-on-B200 comparison legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other workloads")
    ap.add_argument("--fmap-layout", default="channels_last", choices=["channels_last", "nchw"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from posfeat_b200 import _lib
    from posfeat_b200.pairs import PairPipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: posfeat_b200 has no CPU fallback")
    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")             # at these levels NCCL prints a version banner to stdout;
                                                     # this script's stdout is exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    if args.workload != "hpatches":
        fn = {"c1": lambda: workload_c1(dev), "aachen": lambda: workload_aachen(dev),
              "mnn_sweep": lambda: workload_mnn_sweep(dev), "train_step": lambda: workload_train_step(dev, world)}[args.workload]
        out = fn()
        if rank == 0:
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    P = args.pairs
    score, fmap = synth_pairs(P, 1234 + 1000 * rank, dev, args.fmap_layout)   # each rank owns its shard of the pair list
    pipe = PairPipeline(DET_CFG, streams=args.streams)
    n_streams = pipe.streams

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass():
        return pipe.run(score, fmap)

    def step():
        for _ in range(args.passes):
            r = one_pass()
        return r

    for _ in range(warmup):
        feats, matches, nm = step()
    torch.cuda.synchronize()
    n_kp = feats["n"]
    mean_matches = float(nm.float().mean().item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident throughput -----------------------------------
    l0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    # ---- end to end through the host-buffer call ------------------------
    score_h, fmap_h = score.cpu().pin_memory(), fmap.cpu().pin_memory()   # .cpu() preserves the memory format

    def time_host(gather, steps):
        for _ in range(2):
            pipe.run_host(score_h, fmap_h, gather=gather, chunks=args.e2e_chunks or None)
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(steps * args.e2e_passes):
            res = pipe.run_host(score_h, fmap_h, gather=gather, chunks=args.e2e_chunks or None)
        g1.record()
        torch.cuda.synchronize()
        return max(g0.elapsed_time(g1), 1e3 * (time.perf_counter() - t0)), res   # host is in the loop: take wall clock

    # the call's default: once the keypoints are known, exactly the descriptor-map pixels under their taps are
    # fetched from the pinned map, each once ("stage"); beside it (fewer steps: they only document the alternatives)
    # the sampler reading its taps from the pinned map directly ("direct") and the plain form that copies the whole map
    modes = (["stage", "direct"] if PairPipeline.host_gather_applies(fmap_h) and not args.no_host_gather else []) + [False]
    e2e_runs = {}
    for mi, mode in enumerate(modes):
        st = args.steps if mi == 0 else max(2, args.steps // 5)
        ms_m, res = time_host(mode, st)
        e2e_runs[mode] = [ms_m, res, PairPipeline.h2d_bytes(score, fmap), st]
        if mode == "stage":
            e2e_runs[mode][2] = score.numel() * 4 + pipe.staged_pixels(fmap_h, dev) * fmap.shape[1] * 4
        elif mode == "direct":
            e2e_runs[mode][2] = PairPipeline.h2d_bytes(score, fmap, int(n_kp))
    head = modes[0]
    kpt_h, matches_h, nm_h = e2e_runs[head][1]
    e2e_same = all(torch.equal(a, b) for m in modes for a, b in zip(e2e_runs[m][1], e2e_runs[head][1]))
    e2e_same = bool(e2e_same and torch.equal(nm_h, nm.cpu()) and torch.equal(kpt_h, feats["kpt"].cpu())
                    and torch.equal(matches_h[0, :int(nm_h[0])], matches[0, :int(nm_h[0])].cpu()))
    for m in modes:
        e2e_runs[m][1] = None
    # ---- per-kernel durations (CUDA events on the launching stream) -----
    pipe.streams = 1                      # per-kernel durations are taken with the kernels running alone
    prof_iters = max(args.steps, 10)
    _lib.profile_enable(True)
    for _ in range(prof_iters):
        one_pass()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    pipe.streams = n_streams
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms] + [e2e_runs[m][0] for m in modes], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    for i, m in enumerate(modes):
        e2e_runs[m][0] = float(t[1 + i])
    total_pairs = P * args.passes * args.steps * world
    value = total_pairs / (ms / 1e3)

    def e2e_rate(m):
        return P * args.e2e_passes * e2e_runs[m][3] * world / (e2e_runs[m][0] / 1e3)

    if rank == 0:
        pk = peaks()
        kern = {k: {"ms_per_launch": v[0] / v[1], "launches_per_pass": v[1] / prof_iters} for k, v in prof.items()}
        pass_ms_prof = sum(v[0] for v in prof.values()) / prof_iters
        for k, v in prof.items():
            kern[k]["share_of_kernel_time"] = (v[0] / prof_iters) / pass_ms_prof
        # dominant kernel: the tcgen05 matcher.  Algorithmic work = 2*N*M*D flops per pair.
        tc_launches = kern.get("mnn_tc", {}).get("launches_per_pass", 1)
        pairs_per_launch = P / max(tc_launches, 1e-9)            # the batched matcher serves all pairs in one launch
        flops = 2.0 * n_kp * n_kp * D * pairs_per_launch
        tc_ms = kern.get("mnn_tc", {}).get("ms_per_launch")
        roof = None
        if tc_ms:
            ach = flops / (tc_ms * 1e-3) / 1e12
            tr = ncu_traffic("mnn_tc_kernel")
            roof = {"bound": "tensor", "kernel": "mnn_tc_kernel", "achieved": ach, "peak": pk["bf16"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16"],
                    # dram read+write of mnn_tc_kernel per launch: the ncu --set full capture committed under profiles/
                    # (bytes per pair there) scaled to this launch's pair count
                    "traffic": tr["dram_bytes_per_pair"] * pairs_per_launch if tr else None,
                    "traffic_source": tr.get("source") if tr else None,
                    "peak_source": pk["src"] + " (burst bf16)", "algorithmic_flops_per_launch": flops,
                    "frac_of_sustained_peak": ach / pk["bf16_sustained"]}
        nms_ms = kern.get("nms_candidates", {}).get("ms_per_launch")
        sel_ms = kern.get("select_topk", {}).get("ms_per_launch")
        if sel_ms:
            sel_ms += kern.get("keypoint_outputs", {}).get("ms_per_launch", 0.0)
        extra = {}
        if nms_ms:
            nbytes = 4.0 * 2 * P * H * W
            gbs = nbytes / (nms_ms * 1e-3) / 1e9
            extra["nms_hbm"] = {"bound": "hbm", "kernel": "nms_quad_r1_kernel", "achieved": gbs, "peak": pk["hbm"],
                                "unit": "GB/s", "frac": gbs / pk["hbm"], "algorithmic_bytes_per_launch": nbytes}
            if sel_ms:
                nb2 = nbytes + 20.0 * n_kp * 2 * P
                g2 = nb2 / ((nms_ms + sel_ms) * 1e-3) / 1e9
                extra["nms_topk_hbm"] = {"bound": "hbm", "kernels": "NMS + selection + keypoint outputs", "achieved": g2,
                                         "peak": pk["hbm"], "unit": "GB/s", "frac": g2 / pk["hbm"],
                                         "algorithmic_bytes_per_launch": nb2, "ms": nms_ms + sel_ms}
        mnn_keys = ("mnn_prep", "mnn_tc", "mnn_rescore", "mnn_scan", "mnn_verify", "mnn_compact")
        mnn_ms = sum(kern.get(k, {}).get("ms_per_launch", 0) * kern.get(k, {}).get("launches_per_pass", 0)
                     for k in mnn_keys) / max(P, 1)
        if mnn_ms:
            f1 = 2.0 * n_kp * n_kp * D
            extra["mnn_total"] = {"ms_per_pair": mnn_ms, "achieved_tflops": f1 / (mnn_ms * 1e-3) / 1e12,
                                  "frac_of_bf16_peak": f1 / (mnn_ms * 1e-3) / 1e12 / pk["bf16"],
                                  "kernels": [k for k in mnn_keys if k in kern]}
        cpu = None
        if world == 1 and not args.no_cpu:
            import numpy as np
            kind, v, n_done, dt, first = time_cpu_reference(score, fmap)
            cpu = {"value": v, "unit": "pairs/s", "cores": os.cpu_count(), "threads": torch.get_num_threads(), "kind": kind,
                   "sample": f"{n_done} pairs of the same workload in {dt:.1f} s, " +
                             ("the reference's own functions on CPU tensors (oracle/_ref)" if kind == "reference"
                              else "numpy port (oracle/_ref not staged)")}
            # parity of pair 0 against the checker (done here because its output is at hand)
            _, idx_o, m_o = first
            idx_g = feats["idx"][:2].cpu().numpy()
            k0 = int(nm[0].item())
            gs, os_ = pixel_matches(idx_g, matches[0, :k0].cpu().numpy()), pixel_matches(idx_o, m_o)
            cpu["parity_pair0"] = {"keypoint_idx_equal": bool(np.array_equal(idx_g, idx_o)),
                                   "keypoint_sets_equal": bool(all(set(a.tolist()) == set(b.tolist()) for a, b in zip(idx_g, idx_o))),
                                   "matches_gpu": len(gs), "matches_checker": len(os_), "matches_differing": len(gs ^ os_)}
        # ---- the reference's functions, and the same pipeline in stock PyTorch ops, on THIS GPU (bounded sample) ----
        eager = ref_b200 = None
        if not args.no_eager:
            n_eager = min(P, 8)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import torch_eager_pipeline as tep
                ev, (eidx, ematch) = tep.time_pairs(score, fmap, DET_CFG, n_eager)
                k0 = int(nm[0].item())
                eager = {"value": ev, "unit": "pairs/s",
                         "sample": f"{n_eager} pairs of the same workload, device-resident, fp32 matmul (restatement in stock ops)",
                         "keypoint_idx_equal_pair0": bool(torch.equal(eidx, feats["idx"][:2, :eidx.shape[1]])),
                         "matches_differing_pair0": len({tuple(x) for x in ematch.tolist()} ^
                                                        {tuple(x) for x in matches[0, :k0].tolist()})}
            except Exception as e:
                eager = {"error": repr(e)[:200]}
            if ref_available():
                try:
                    from oracle import ref_runner
                    fm_nchw = fmap[:2 * n_eager].contiguous()            # the reference's layout
                    prev = torch.backends.cuda.matmul.allow_tf32
                    torch.backends.cuda.matmul.allow_tf32 = False
                    def ref_pass():
                        out = None
                        for i in range(n_eager):
                            out = ref_runner.run_pair(score[2 * i:2 * i + 2], fm_nchw[2 * i:2 * i + 2], dict(DET_CFG))
                        return out
                    ms_r, _ = cuda_time(ref_pass, 3, warmup=2)
                    kps_r, _, m_r, idx_r = ref_runner.run_pair(score[0:2], fm_nchw[0:2], dict(DET_CFG), want_idx=True)
                    torch.backends.cuda.matmul.allow_tf32 = prev
                    idx_r = idx_r.cpu().numpy()
                    k0 = int(nm[0].item())
                    gs = pixel_matches(feats["idx"][:2].cpu().numpy(), matches[0, :k0].cpu().numpy())
                    ref_b200 = {"value": n_eager / (ms_r * 1e-3), "unit": "pairs/s",
                                "sample": f"{n_eager} pairs, the reference's own functions (oracle/_ref) on cuda tensors, one pair "
                                          "per call as its scripts do, host-built grid and syncs included",
                                "speedup_device_resident": value / world / (n_eager / (ms_r * 1e-3)),
                                "matches_differing_pair0": len(gs ^ pixel_matches(idx_r, m_r))}
                except Exception as e:
                    ref_b200 = {"error": repr(e)[:200]}
        others = None
        if world == 1 and not args.no_extras:
            others = {}
            for name, fn in (("c1", lambda: workload_c1(dev, 100)), ("aachen_k20480", lambda: workload_aachen(dev, iters=2)),
                             ("aachen_k8192", lambda: workload_aachen(dev, iters=2, num_pts=8192)),
                             ("mnn_sweep", lambda: workload_mnn_sweep(dev)),
                             ("train_step", lambda: workload_train_step(dev, 1, 5))):
                try:
                    del_cache = torch.cuda.empty_cache
                    others[name] = fn()
                    del_cache()
                except Exception as e:
                    others[name] = {"error": repr(e)[:300]}
        head_rate = e2e_rate(head)
        out = {"metric": "matched image pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
               "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16 MMA + f64 rescoring (matcher), f32 (detect/sample)",
               "data": "synthetic",
               "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": P * args.passes, "resident_batch_pairs": P,
                          "passes_per_step": args.passes, "streams": n_streams, "keypoints": int(n_kp),
                          "descriptor_dim": D, "fmap_layout": args.fmap_layout, "detector": {k: v for k, v in DET_CFG.items()},
                          "mean_matches_per_pair": mean_matches,
                          "l2": f"every pass re-reads {((score.numel() + fmap.numel()) * 4) >> 20} MiB of inputs > 126 MB L2 (no flush needed)",
                          "timed_region_ms": ms,
                          "parallelism": f"pairs sharded over {world} rank(s), no collective"},
               "e2e": {"value": head_rate, "unit": "pairs/s", "ms_per_step": e2e_runs[head][0] / e2e_runs[head][3],
                       "pairs_per_step_per_gpu": P * args.e2e_passes,
                       "h2d_bytes_per_step": int(e2e_runs[head][2]) * args.e2e_passes,
                       "d2h_bytes_per_step": PairPipeline.d2h_bytes(2 * P, int(n_kp), P) * args.e2e_passes,
                       "input_path": E2E_PATHS[head],
                       "other_input_paths": {str(m): {"value": e2e_rate(m), "ms_per_step": e2e_runs[m][0] / e2e_runs[m][3],
                                                      "h2d_bytes_per_step": int(e2e_runs[m][2]) * args.e2e_passes,
                                                      "input_path": E2E_PATHS[m]}
                                             for m in modes[1:]},
                       "results_equal_device_path": e2e_same},
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
               "reference_on_b200": ref_b200, "torch_eager_b200": eager, "kernels": kern, "extra_rooflines": extra,
               "workloads": others}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- matched image pairs/s through the post-backbone feature pipeline.

One "step" = one pass of the hot path (NMS + threshold + top-k keypoints,
descriptor sampling + L2 norm, mutual-NN matching) over a batch of P synthetic
HPatches-shaped image pairs (BASELINE.json configs[1]: 896x1200 score maps,
128x224x300 descriptor maps, 8192 keypoints, D=128).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...      (one rank per GPU; pairs are sharded, no collective)

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e`
the same metric through the public host-buffer call (PairPipeline.run_host) with
the H2D/D2H copies inside the timed region, `roofline` the tcgen05 matcher
kernel against the measured bf16 peak, `cpu_baseline` the oracle port on the
host cores.  `--impl reference` times the reference algorithm's CPU port only.
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


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "bf16": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


def synth_pairs(P, seed, device, layout="channels_last"):
    """Synthetic post-backbone maps: score = softplus(randn),
    descriptor map = randn; image 2i+1 is a noisy copy of image 2i so that the
    pairs have plenty of true matches.  The descriptor map has the reference's
    logical shape [2P, D, H/4, W/4]; `layout` picks its memory format:
    channels_last (what a channels_last backbone emits; the sampler then reads
    128-bit vectors) or nchw (the reference's default contiguous layout)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    score = torch.empty(2 * P, 1, H, W)
    fmap = torch.empty(2 * P, D, H // 4, W // 4)
    for i in range(P):
        s = torch.randn(1, H, W, generator=g)
        f = torch.randn(D, H // 4, W // 4, generator=g)
        score[2 * i] = torch.nn.functional.softplus(s)
        score[2 * i + 1] = torch.nn.functional.softplus(s + 0.05 * torch.randn(1, H, W, generator=g))
        fmap[2 * i] = f
        fmap[2 * i + 1] = f + 0.3 * torch.randn(D, H // 4, W // 4, generator=g)
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


def cpu_port_pair(score2, fmap2):
    """The oracle port of the reference path for ONE pair (numpy, host cores)."""
    from oracle import posfeat_oracle as O
    kps, sc, idx, _ = O.generate_kpts_single(score2, return_idx=True, **{k: v for k, v in DET_CFG.items()
                                                                        if k not in ("stable",)})
    desc = O.sample_feat_by_coord(fmap2, kps, True)
    m = O.mnn_matcher(desc[0], desc[1])
    return idx, kps, desc, m


def time_cpu_port(score, fmap, budget_s=12.0, max_pairs=16):
    score, fmap = score[:8], fmap[:8]          # a bounded sample: at most the first 4 pairs are cycled through
    """Oracle on the host cores over a bounded sample of the same workload."""
    sn, fn = score.cpu().numpy(), fmap.cpu().numpy()
    P = sn.shape[0] // 2
    done, t0 = 0, time.perf_counter()
    first = None
    while done < max_pairs and (done == 0 or time.perf_counter() - t0 < budget_s):
        i = done % P
        r = cpu_port_pair(sn[2 * i:2 * i + 2], fn[2 * i:2 * i + 2])
        if first is None:
            first = r
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, done, dt, first


def run_reference_arm(args, rank, world):
    """--impl reference: the reference algorithm's CPU port (oracle/) on the host
    cores; the Python reference itself cannot travel to the GPU box."""
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    score, fmap = synth_pairs(2, 1234, "cpu")
    sn, fn = score.numpy(), fmap.numpy()
    for _ in range(min(args.warmup, 1)):
        cpu_port_pair(sn[0:2], fn[0:2])
    steps = max(1, min(args.steps, 60))  # each step = 1 pair (bounded sample, ~1.3 s on 16 cores): K steps as asked, capped so
                                         # that the whole run stays within a couple of minutes
    t0 = time.perf_counter()
    for s in range(steps):
        i = s % 2
        cpu_port_pair(sn[2 * i:2 * i + 2], fn[2 * i:2 * i + 2])
    dt = time.perf_counter() - t0
    v = steps / dt
    out = {"impl": "reference", "metric": "matched image pairs/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
           "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "pairs_per_step": 1, "keypoints": KPTS, "descriptor_dim": D},
           "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"{steps} pairs, numpy oracle (BLAS threads = all cores)"},
           "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


E2E_PATHS = {
    "stage": "score maps copied whole; descriptor-map pixels under the keypoints' taps fetched once each from pinned host "
             "memory (posfeat_fetch_taps_f32; bytes = pixels moved, counted by the kernel, x 512 B)",
    "direct": "score maps copied whole; the sampler reads its taps from pinned host memory (bytes requested = 4 taps x 512 B "
              "per keypoint)",
    False: "score and descriptor maps copied whole",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="pairs per step per GPU")
    ap.add_argument("--streams", type=int, default=1, help="CUDA streams the pair batch is split over (1 = plain path)")
    ap.add_argument("--no-host-gather", action="store_true",
                    help="e2e leg: copy the whole descriptor map to the device instead of gathering taps over the host link")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager-on-B200 comparison leg")
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

    P = args.pairs
    score, fmap = synth_pairs(P, 1234 + 1000 * rank, dev, args.fmap_layout)   # each rank owns its shard of the pair list
    pipe = PairPipeline(DET_CFG, streams=args.streams)
    n_streams = pipe.streams

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return pipe.run(score, fmap)

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
    def time_host(gather):
        for _ in range(2):
            pipe.run_host(score_h, fmap_h, gather=gather)
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            res = pipe.run_host(score_h, fmap_h, gather=gather)
        g1.record()
        torch.cuda.synchronize()
        return max(g0.elapsed_time(g1), 1e3 * (time.perf_counter() - t0)), res   # host is in the loop: take wall clock

    # the call's default: once the keypoints are known, exactly the descriptor-map pixels under their taps are
    # fetched from the pinned map, each once ("stage"); beside it the sampler reading its taps from the pinned
    # map directly ("direct") and the plain form that copies the whole dense map first
    modes = (["stage", "direct"] if PairPipeline.host_gather_applies(fmap_h) and not args.no_host_gather else []) + [False]
    e2e_runs = {}
    for mode in modes:
        ms_m, res = time_host(mode)
        e2e_runs[mode] = [ms_m, res, PairPipeline.h2d_bytes(score, fmap)]
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
    _lib.profile_enable(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    pipe.streams = n_streams
    clocks = sampler.stop() if rank == 0 else None
    # ---- the same pipeline in stock PyTorch ops on this GPU (rank 0, bounded sample; not part of `value`) ----
    eager = None
    if rank == 0 and not args.no_eager:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import torch_eager_pipeline as tep
        n_eager = min(P, 8)
        ev, (eidx, ematch) = tep.time_pairs(score, fmap, DET_CFG, n_eager)
        k0 = int(nm[0].item())
        eager = {"value": ev, "unit": "pairs/s", "sample": f"{n_eager} pairs of the same workload, device-resident, fp32 matmul",
                 "keypoint_idx_equal_pair0": bool(torch.equal(eidx, feats["idx"][:2, :eidx.shape[1]])),
                 "matches_differing_pair0": len({tuple(x) for x in ematch.tolist()} ^
                                                {tuple(x) for x in matches[0, :k0].tolist()})}

    t = torch.tensor([ms] + [e2e_runs[m][0] for m in modes], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    for i, m in enumerate(modes):
        e2e_runs[m][0] = float(t[1 + i])
    ms_e2e = e2e_runs[head][0]
    total_pairs = P * args.steps * world
    value = total_pairs / (ms / 1e3)
    e2e_value = total_pairs / (ms_e2e / 1e3)

    if rank == 0:
        pk = peaks()
        kern = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
        step_ms_prof = sum(v[0] for v in prof.values()) / args.steps
        for k, v in prof.items():
            kern[k]["share_of_kernel_time"] = (v[0] / args.steps) / step_ms_prof
        # dominant kernel: the tcgen05 matcher.  Algorithmic work = 2*N*M*D flops per pair.
        tc_launches = kern.get("mnn_tc", {}).get("launches_per_step", P)
        pairs_per_launch = P / max(tc_launches, 1e-9)            # the batched matcher serves all pairs in one launch
        flops = 2.0 * n_kp * n_kp * D * pairs_per_launch
        tc_ms = kern.get("mnn_tc", {}).get("ms_per_launch")
        roof = None
        if tc_ms:
            ach = flops / (tc_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "mnn_tc_kernel", "achieved": ach, "peak": pk["bf16"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16"],
                    # dram read+write of mnn_tc_kernel from the ncu --set full capture (113.8 MB per 8-pair launch,
                    # profiles/r01_ncu_full_mnn_tc_kernel_P8.csv), scaled to this launch's pair count
                    "traffic": 14.23e6 * pairs_per_launch, "peak_source": pk["src"] + " (burst bf16)",
                    "algorithmic_flops_per_launch": flops}
        nms_ms = kern.get("nms_candidates", {}).get("ms_per_launch")
        extra = {}
        if nms_ms:
            nbytes = 4.0 * 2 * P * H * W
            gbs = nbytes / (nms_ms * 1e-3) / 1e9
            extra["nms_hbm"] = {"bound": "hbm", "kernel": "nms_quad_r1_kernel", "achieved": gbs, "peak": pk["hbm"],
                                "unit": "GB/s", "frac": gbs / pk["hbm"], "algorithmic_bytes_per_launch": nbytes}
        mnn_ms = sum(kern.get(k, {}).get("ms_per_launch", 0) * kern.get(k, {}).get("launches_per_step", 0)
                     for k in ("mnn_prep", "mnn_tc", "mnn_rescore", "mnn_scan", "mnn_verify", "mnn_compact")) / max(P, 1)
        if mnn_ms:
            f1 = 2.0 * n_kp * n_kp * D
            extra["mnn_total"] = {"ms_per_pair": mnn_ms, "achieved_tflops": f1 / (mnn_ms * 1e-3) / 1e12,
                                  "frac_of_bf16_peak": f1 / (mnn_ms * 1e-3) / 1e12 / pk["bf16"]}
        cpu = None
        if world == 1 and not args.no_cpu:
            import numpy as np
            torch.set_num_threads(os.cpu_count() or 1)
            v, n_done, dt, first = time_cpu_port(score, fmap)
            cpu = {"value": v, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{n_done} pairs of the same workload in {dt:.1f} s, numpy oracle (BLAS uses all cores)"}
            # parity of pair 0 against the oracle (checked here because the oracle output is at hand)
            idx_o, kps_o, desc_o, m_o = first
            same_idx = bool(np.array_equal(feats["idx"][:2].cpu().numpy(), idx_o))
            k0 = int(nm[0].item())
            m_g = matches[0, :k0].cpu().numpy()
            gs, os_ = {tuple(x) for x in m_g.tolist()}, {tuple(x) for x in m_o.tolist()}
            cpu["parity_pair0"] = {"keypoint_idx_equal": same_idx, "matches_gpu": len(gs), "matches_oracle": len(os_),
                                   "matches_differing": len(gs ^ os_)}
        out = {"metric": "matched image pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world,
               "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16 MMA + f64 rescoring (matcher), f32 (detect/sample)",
               "data": "synthetic",
               "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": P, "streams": n_streams, "keypoints": int(n_kp),
                          "descriptor_dim": D, "fmap_layout": args.fmap_layout, "detector": {k: v for k, v in DET_CFG.items()},
                          "mean_matches_per_pair": mean_matches,
                          "l2": f"inputs {((score.numel() + fmap.numel()) * 4) >> 20} MiB per step > 126 MB L2 (no flush needed)",
                          "parallelism": f"pairs sharded over {world} rank(s), no collective"},
               "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                       "h2d_bytes_per_step": int(e2e_runs[head][2]),
                       "d2h_bytes_per_step": PairPipeline.d2h_bytes(2 * P, int(n_kp), P),
                       "input_path": E2E_PATHS[head],
                       "other_input_paths": {str(m): {"value": total_pairs / (e2e_runs[m][0] / 1e3),
                                                      "ms_per_step": e2e_runs[m][0] / args.steps,
                                                      "h2d_bytes_per_step": int(e2e_runs[m][2]), "input_path": E2E_PATHS[m]}
                                             for m in modes[1:]},
                       "results_equal_device_path": e2e_same},
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
               "torch_eager_b200": eager, "kernels": kern, "extra_rooflines": extra}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host-link ceiling for the e2e leg: plain pinned host->device (and device->host) copies issued by all ranks at the
same time, 1 GiB each, CUDA-event timed.  `torchrun --nproc-per-node N tools/h2d_ceiling.py` (or python for N = 1)
prints one JSON line: per-rank and aggregate GB/s.  PairPipeline.run_host can not move its bytes faster than this."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 28                                          # float32 elements: 1 GiB
host = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
devb = torch.empty(n, dtype=torch.float32, device=dev)


def timed(fn, iters=5):
    fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 4.0 * n * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


h2d = timed(lambda: devb.copy_(host, non_blocking=True))
d2h = timed(lambda: host.copy_(devb, non_blocking=True))
t = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
if world > 1:
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
else:
    allv = [t]
if rank == 0:
    topo = ""
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        pass
    numa = [ln for ln in topo.splitlines() if ln.startswith("GPU")]
    print(json.dumps({"n_gpus": world, "h2d_GBs_per_rank": [round(float(v[0]), 2) for v in allv],
                      "d2h_GBs_per_rank": [round(float(v[1]), 2) for v in allv],
                      "h2d_GBs_aggregate": round(sum(float(v[0]) for v in allv), 1),
                      "d2h_GBs_aggregate": round(sum(float(v[1]) for v in allv), 1),
                      "host_cpus": os.cpu_count(), "topo_rows": numa[:8]}), flush=True)
if world > 1:
    dist.destroy_process_group()

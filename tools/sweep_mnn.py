"""BASELINE.json configs[4]: MNN matcher sweep N = M in {1k, 4k, 16k, 64k}, D = 128, against the
tensor-pipe roofline.  Prints one JSON object per size: time of the tcgen05 kernel alone and of
the whole matcher (library CUDA events on the launching stream), the resulting TFLOP/s
(2*N*M*D per pair) and, for sizes where it is affordable, the torch fp32 matmul+max matcher."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from posfeat_b200 import _lib
from posfeat_b200.preprocess_utils import mnn_match

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch_eager_pipeline as tep

peak = bench.peaks()
sizes = [int(x) for x in sys.argv[1:]] or [1024, 4096, 8192, 16384, 65536]
for n in sizes:
    g = torch.Generator().manual_seed(n)
    a = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=1)
    b = torch.nn.functional.normalize(a[torch.randperm(n, generator=g)] + 0.06 * torch.randn(n, 128, generator=g), dim=1)
    a, b = a.cuda(), b.cuda()
    iters = 20 if n <= 16384 else 3
    out = {"N": n, "M": n, "D": 128}
    for want_nn21, tag in ((False, "matches_only"), (True, "with_nn21")):
        for _ in range(2):
            m, nm, _, _ = mnn_match(a, b, algo=2, want_nn21=want_nn21)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m, nm, _, _ = mnn_match(a, b, algo=2, want_nn21=want_nn21)
        e1.record()
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        ker = {k: 1e3 * v[0] / iters for k, v in pr.items() if k.startswith("mnn")}
        flops = 2.0 * n * n * 128 * (2 if want_nn21 else 1)       # both directions are contracted when nn21 is wanted
        tc_us = ker.get("mnn_tc", 0.0)
        out[tag] = {"us_per_call_stream": 1e3 * e0.elapsed_time(e1) / iters, "kernel_us": {k: round(v, 1) for k, v in ker.items()},
                    "tc_tflops": flops / (tc_us * 1e-6) / 1e12 if tc_us else None,
                    "tc_frac_of_peak": flops / (tc_us * 1e-6) / 1e12 / peak["bf16"] if tc_us else None,
                    "matcher_tflops_useful": 2.0 * n * n * 128 / (sum(ker.values()) * 1e-6) / 1e12,
                    "matches": int(nm.item())}
    if n <= 16384:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        for _ in range(2):
            r = tep.mnn(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = tep.mnn(a, b)
        e1.record()
        torch.cuda.synchronize()
        torch.backends.cuda.matmul.allow_tf32 = prev
        out["torch_fp32_us"] = 1e3 * e0.elapsed_time(e1) / 5
        got = m[:int(nm.item())]
        out["matches_equal_torch"] = bool(got.shape == r.shape and torch.equal(got, r))
    out["peak_bf16_tflops"] = peak["bf16"]
    print(json.dumps(out), flush=True)
    del a, b
    torch.cuda.empty_cache()

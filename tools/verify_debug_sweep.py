"""Bring-up timing of the verification kernel under POSFEAT_TC_DEBUG bits (results are garbage by design): 0x10000
return after the threshold + group-entry scan + competitor list, 0x40000 additionally collect the members, 0x20000
everything but the evaluation of non-member rows, 0x80000 threshold and first block of group entries only.
The rescoring and tensor-kernel times of the same calls are printed beside (POSFEAT_MNN_RESCORE_WARP=1: warp-per-row rescoring).  Two data sets: the bench-like matched pairs and unrelated pairs."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posfeat_b200 import _lib
from posfeat_b200.pairs import PairPipeline

P, n = 64, 8192
g = torch.Generator().manual_seed(0)
pipe = PairPipeline(dict(nms_radius=1, num_pts=n), mnn_algo=_lib.MNN_TC)
for name, noise in (("matched_0.06", 0.06), ("matched_0.3", 0.3), ("unrelated", None)):
    a = torch.nn.functional.normalize(torch.randn(P, n, 128, generator=g), dim=-1)
    b = torch.randn(P, n, 128, generator=g) if noise is None else a[:, torch.randperm(n, generator=g)] + noise * torch.randn(P, n, 128, generator=g)
    b = torch.nn.functional.normalize(b, dim=-1)
    desc = torch.stack([a, b], 1).reshape(2 * P, n, 128).cuda()
    os.environ["POSFEAT_MNN_DEBUG"] = "1"
    os.environ.pop("POSFEAT_TC_DEBUG", None)
    pipe.match(desc)
    del os.environ["POSFEAT_MNN_DEBUG"]
    for dbg in (0, 0x20000, 0x40000, 0x10000, 0x80000):
        os.environ["POSFEAT_TC_DEBUG"] = str(dbg)
        for _ in range(2):
            pipe.match(desc)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(5):
            pipe.match(desc)
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        print(f"{name} debug={dbg:#07x}: " + " ".join(f"{k} {1e3 * pr[k][0] / pr[k][1]:7.1f}" for k in ("mnn_tc", "mnn_rescore", "mnn_verify", "mnn_compact")), flush=True)

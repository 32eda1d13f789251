"""Bring-up timing of the rescoring kernel under POSFEAT_TC_DEBUG bits (results are garbage by design): 0x2000 return
after the rescoring kernel, 0x4000 every candidate re-evaluates chunk 0 (loads come out of L1: what is left is the
instruction cost), 0x8000 no dot products at all (list walk and outputs only).  POSFEAT_MNN_RESCORE_WARP=1 selects the
warp-per-row kernel."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posfeat_b200 import _lib
from posfeat_b200.pairs import PairPipeline

P, n = 64, 8192
g = torch.Generator().manual_seed(0)
a = torch.nn.functional.normalize(torch.randn(P, n, 128, generator=g), dim=-1)
b = torch.nn.functional.normalize(a[:, torch.randperm(n, generator=g)] + 0.3 * torch.randn(P, n, 128, generator=g), dim=-1)
desc = torch.stack([a, b], 1).reshape(2 * P, n, 128).cuda()
pipe = PairPipeline(dict(nms_radius=1, num_pts=n), mnn_algo=_lib.MNN_TC)
os.environ["POSFEAT_MNN_DEBUG"] = "1"
pipe.match(desc)
del os.environ["POSFEAT_MNN_DEBUG"]
for dbg in (0, 0x4000, 0x8000):
    os.environ["POSFEAT_TC_DEBUG"] = str(dbg | 0x2000)
    for _ in range(2):
        pipe.match(desc)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(5):
        pipe.match(desc)
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    _lib.profile_enable(False)
    print(f"debug={dbg:#06x}: mnn_rescore {1e3 * pr['mnn_rescore'][0] / pr['mnn_rescore'][1]:8.1f} us per {P} pairs", flush=True)

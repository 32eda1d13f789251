"""Where the verification kernel's time goes on UNRELATED descriptor sets (the Aachen-shaped workload: a query against
retrieved images that mostly do not match): per-kernel times and the POSFEAT_MNN_DEBUG counters, list form vs table form."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posfeat_b200 import _lib
from posfeat_b200.pairs import PairPipeline

P, n = 20, 8192
g = torch.Generator().manual_seed(0)
for name, noise in (("unrelated", None), ("matched_0.3", 0.3)):
    a = torch.nn.functional.normalize(torch.randn(P, n, 128, generator=g), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(P, n, 128, generator=g) if noise is None else a + noise * torch.randn(P, n, 128, generator=g), dim=-1)
    desc = torch.stack([a, b], 1).reshape(2 * P, n, 128).cuda()
    pipe = PairPipeline(dict(nms_radius=1, num_pts=n), mnn_algo=_lib.MNN_TC)
    for table in (0, 1):
        os.environ["POSFEAT_MNN_TABLE"] = str(table)
        os.environ.pop("POSFEAT_MNN_DEBUG", None)
        for _ in range(2):
            pipe.match(desc)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(3):
            m, nm = pipe.match(desc)
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        print(name, "table" if table else "lists", {k: round(1e3 * v[0] / v[1] / P, 1) for k, v in pr.items()}, "us per pair; matches", float(nm.float().mean()), flush=True)
        if not table:
            os.environ["POSFEAT_MNN_DEBUG"] = "1"
            pipe.match(desc)
            torch.cuda.synchronize()

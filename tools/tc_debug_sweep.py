"""Bring-up timing of mnn_tc_kernel alone under POSFEAT_TC_DEBUG bit combinations (results are garbage by design):
4 no MMA, 16 MMA ignores the epilogue handshake, 64 no table / group-entry stores, 128 no group-entry reduction,
256 no row lists, 0x1000 return after the tensor kernel."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posfeat_b200 import _lib
from posfeat_b200.pairs import PairPipeline

P, n = 64, 8192
g = torch.Generator().manual_seed(0)
a = torch.nn.functional.normalize(torch.randn(P, n, 128, generator=g), dim=-1)
b = torch.nn.functional.normalize(a + 0.3 * torch.randn(P, n, 128, generator=g), dim=-1)
desc = torch.stack([a, b], 1).reshape(2 * P, n, 128).cuda()
pipe = PairPipeline(dict(nms_radius=1, num_pts=n), mnn_algo=_lib.MNN_TC)
combos = [int(x, 0) for x in sys.argv[1:]] or [0, 64, 128, 256, 128 | 256, 4, 4 | 128 | 256, 16, 16 | 128 | 256]
for table in (0, 1):
    os.environ["POSFEAT_MNN_TABLE"] = str(table)
    for dbg in combos:
        if table and (dbg & (128 | 256)):
            continue
        os.environ["POSFEAT_TC_DEBUG"] = str(dbg | 0x1000)
        for _ in range(2):
            pipe.match(desc)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(5):
            pipe.match(desc)
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        print(f"table={table} debug={dbg:#06x}: mnn_tc {1e3 * pr['mnn_tc'][0] / pr['mnn_tc'][1]:8.1f} us per {P} pairs", flush=True)

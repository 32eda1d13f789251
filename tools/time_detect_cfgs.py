"""Detector timing (library event slots) for the three extraction configs of BASELINE.json:
C1 480x640 k4096 r1, C2 896x1200 k8192 r1, C3 1056x1600 k20480 r3 (configs/extract_aachen.yaml)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from posfeat_b200 import _lib
from posfeat_b200.preprocess_utils import detect_topk

for name, B, H, W, r, k, thr in (("C1", 256, 480, 640, 1, 4096, 0.9), ("C2", 128, 896, 1200, 1, 8192, 0.9),
                                 ("C3", 64, 1056, 1600, 3, 20480, 0.5)):
    torch.manual_seed(0)
    score = torch.nn.functional.softplus(torch.randn(B, 1, H, W, device="cuda"))
    for _ in range(3):
        res = detect_topk(score, r, k, True, thr, "abs", sync=False)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(10):
        res = detect_topk(score, r, k, True, thr, "abs", sync=False)
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    nms = prof["nms_candidates"][0] / prof["nms_candidates"][1]
    sel = prof["select_topk"][0] / prof["select_topk"][1]
    gbs = 4.0 * B * H * W / (nms * 1e-3) / 1e9
    print(f"{name}: {B} maps {H}x{W} r={r} k={k}: nms {nms*1e3:7.1f} us ({gbs:6.0f} GB/s, {nms*1e3/B:5.2f} us/map)  "
          f"select {sel*1e3:7.1f} us ({sel*1e3/B:5.2f} us/map)  survivors/map {float(res['counts'].float().mean()):8.0f} n={int(res['n'])}")
    del score

"""Device time of crdr_conv_wgrad on the training step's shapes (CUDA events, 20 launches after 3 warm-ups).
    CRDR_WGRAD_HALO=0|1 python tools/wgrad_bench.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crdr_b200 import backward as bw  # noqa: E402
from crdr_b200.engine import Act  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
CASES = [  # (name, hs, ws, ca (dY channels), cb (X channels), k)
    ("ChARM c1 5x5 480->224 @16", 16, 16, 224, 480, 5),
    ("ChARM c2 5x5 224->128 @16", 16, 16, 128, 224, 5),
    ("ChARM c3 3x3 128->32  @16", 16, 16, 32, 128, 3),
    ("g_a 3x3 96->96   @128", 128, 128, 96, 96, 3),
    ("g_s 3x3 128->128 @128", 128, 128, 128, 128, 3),
    ("g_s 3x3 128->128 @32", 32, 32, 128, 128, 3),
    ("NLAM 3x3 160->160 @16", 16, 16, 160, 160, 3),
]
print(f"batch {B}, CRDR_WGRAD_HALO={os.environ.get('CRDR_WGRAD_HALO', '1')}")
for name, hs, ws, ca, cb, k in CASES:
    dy = Act(torch.randn(B, hs, ws, ca, device="cuda").half(), None)
    x = Act(torch.randn(B, hs, ws, cb, device="cuda").half(), None)
    out = torch.zeros(ca, cb, k, k, device="cuda")
    taps = [(i - k // 2, j - k // 2) for i in range(k) for j in range(k)]
    run = lambda: bw.wgrad(dy, 0, ca, x, 0, cb, taps, 1, out, cb * k * k, k * k, 1)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * B * hs * ws * ca * cb * k * k
    print(f"  {name:28s} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")

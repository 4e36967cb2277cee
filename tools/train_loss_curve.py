"""Loss per step on a fixed batch with fixed noise: python tools/train_loss_curve.py <config> <graphs 0|1> [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fixtures  # noqa: E402
from crdr_b200.train import CodecTrainer  # noqa: E402

DEV = "cuda:0"
config, graphs = sys.argv[1], sys.argv[2] == "1"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
model, _ = fixtures.build_model(seed=6, calibrated=False, config=config)
tr = CodecTrainer(model, device=DEV, lr=1e-4, clip_max_norm=1.0)
tr.use_graphs = graphs
n, h, w = 2, 128, 128
x = fixtures.image(n, h, w, seed=22).to(DEV).contiguous()
g = torch.Generator(device=DEV).manual_seed(5)
mk = lambda c, a, b: torch.rand((n, c, a, b), dtype=torch.float32, device=DEV, generator=g) - 0.5
noise = {"z": mk(192, h // 64, w // 64), "y": mk(320, h // 16, w // 16)}
beta = 2.56 if config == "crdr.yaml" else None
for it in range(steps):
    ld = tr.train_step(x, q=2.0, noise=noise, beta=beta)
    print(it, f"rate {float(ld['rate']):.5f} dist {float(ld['distortion']):.5f} bpp {float(ld['bpp']):.4f} w {float(ld['rate_weight']):.4f} aux {float(ld['aux']):.2f}")

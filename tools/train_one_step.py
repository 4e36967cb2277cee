"""Two eager training steps on 8 crops of 256 x 256 (ncu launch lists: the second step is the steady state)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fixtures  # noqa: E402
from crdr_b200.train import CodecTrainer  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model, _ = fixtures.build_model(seed=0, calibrated=False, device=DEV, config="crdr_stage_2.yaml")
tr = CodecTrainer(model, device=DEV, lr=1e-4, clip_max_norm=1.0)
tr.use_graphs = False
x = fixtures.image(B, 256, 256, seed=3).to(DEV)
gen = torch.Generator(device=DEV).manual_seed(0)
for i in range(2):
    torch.cuda.nvtx.range_push(f"step{i}")
    tr.train_step(x, q=2.0, generator=gen)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done")

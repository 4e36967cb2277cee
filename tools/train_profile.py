"""Where a training step spends its time: host enqueue vs device, per phase (synchronising between phases)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fixtures  # noqa: E402
from crdr_b200 import native as nv  # noqa: E402
from crdr_b200.train import CodecTrainer  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model, _ = fixtures.build_model(seed=0, calibrated=False, device=DEV, config="crdr_stage_2.yaml")
tr = CodecTrainer(model, device=DEV, lr=1e-4, clip_max_norm=1.0)
tr.use_graphs = False
x = fixtures.image(B, 256, 256, seed=3).to(DEV)
gen = torch.Generator(device=DEV).manual_seed(0)
mk = lambda c, a, b: torch.rand((B, c, a, b), dtype=torch.float32, device=DEV, generator=gen) - 0.5
noise = {"z": mk(192, 4, 4), "y": mk(320, 16, 16)}
for _ in range(3):
    tr.train_step(x, q=2.0, noise=noise)
torch.cuda.synchronize()


def phase(name, fn, acc):
    torch.cuda.synchronize()
    l0 = nv.LAUNCH_COUNT[0]
    t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0, 0])
    a[0] += t1 - t0
    a[1] += t2 - t0
    a[2] += nv.LAUNCH_COUNT[0] - l0
    return r


acc = {}
N = 5
for _ in range(N):
    out = phase("forward", lambda: tr.forward(x, 2.0, noise), acc)
    ld = phase("losses", lambda: tr.losses(x, out, 2.0), acc)
    phase("backward", lambda: tr.backward(x, out, ld["rate_weight"]), acc)
    phase("aux", lambda: tr.aux_step(), acc)
    phase("optimizer", lambda: tr.optimizer_step(), acc)
print(f"batch {B} x 256x256; per phase: host enqueue ms | enqueue + device ms | C-ABI calls")
for k, (a, b, c) in acc.items():
    print(f"{k:10s} {a / N * 1e3:8.2f} {b / N * 1e3:8.2f} {c // N:6d}")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    tr.train_step(x, q=2.0, noise=noise)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"whole step: enqueue {1e3 * (t1 - t0) / N:.2f} ms, total {1e3 * (t2 - t0) / N:.2f} ms")
tr.use_graphs = True
for _ in range(3):
    tr.train_step(x, q=2.0, noise=noise)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    tr.train_step(x, q=2.0, noise=noise)
torch.cuda.synchronize()
print(f"graph replay: {1e3 * (time.perf_counter() - t0) / 10:.2f} ms per step")

"""Data-parallel training step vs the same global batch on one GPU (run under torchrun with 2+ ranks):
    torchrun --nproc-per-node 2 tools/dp_check.py
Every rank takes its slice of a fixed batch (and of the noise); after forward + rate switch + backward + gradient all-reduce
the flat gradient must equal the single-process gradient of the whole batch (rank 0 computes that with a 1-rank group)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fixtures  # noqa: E402
from crdr_b200.train import CodecTrainer  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
solo = dist.new_group([0])                      # collectives of the reference trainer are no-ops (world size 1)
per, h, w, q = 2, 128, 128, 2.0
n = per * world
x = fixtures.image(n, h, w, seed=40)
g = torch.Generator().manual_seed(41)
noise = {"z": torch.rand(n, 192, h // 64, w // 64, generator=g) - 0.5, "y": torch.rand(n, 320, h // 16, w // 16, generator=g) - 0.5}


def grads(tr, xs, ns):
    tr.loss_scale = 2.0 ** 11                   # the same scale in both runs (the default follows the local pixel count)
    ld = tr._core_forward(xs.to(dev).contiguous(), q, {k: v.to(dev).contiguous() for k, v in ns.items()})
    tr._decide_rate()
    tr._decide_skip()
    tr._core_backward(xs.to(dev).contiguous())
    tr.all_reduce_grads()
    torch.cuda.synchronize()
    return tr.ctx.flat_g.clone(), float(tr._rate_w), float(tr._qbpp)


model, _ = fixtures.build_model(seed=6, calibrated=True, device=dev, config="crdr_stage_2.yaml")
sl = slice(rank * per, (rank + 1) * per)
g_dp, w_dp, qbpp_dp = grads(CodecTrainer(model, device=dev), x[sl], {k: v[sl] for k, v in noise.items()})
if rank == 0:
    model2, _ = fixtures.build_model(seed=6, calibrated=True, device=dev, config="crdr_stage_2.yaml")
    g_one, w_one, qbpp_one = grads(CodecTrainer(model2, device=dev, process_group=solo), x, noise)
    rel = float((g_dp - g_one).norm() / g_one.norm())
    print(f"world {world}: rate weight {w_dp} vs {w_one}, qbpp {qbpp_dp:.6f} vs {qbpp_one:.6f}, "
          f"|g_dp - g_single| / |g_single| = {rel:.3e}, max |diff| {float((g_dp - g_one).abs().max()):.3e}")
    assert w_dp == w_one and abs(qbpp_dp - qbpp_one) < 1e-5 and rel < 2e-3, rel
    print("dp_check ok")
dist.barrier()
dist.destroy_process_group()

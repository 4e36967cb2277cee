"""Kernel timeline of one graph-replayed training step from CUPTI (torch.profiler): device busy time per stream, the
union busy time (any stream) and the idle time of the step.  Attribution only, never a bench number."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fixtures  # noqa: E402
from crdr_b200.train import CodecTrainer  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model, _ = fixtures.build_model(seed=0, calibrated=False, device=DEV, config="crdr_stage_2.yaml")
tr = CodecTrainer(model, device=DEV, lr=1e-4, clip_max_norm=1.0)
x = fixtures.image(B, 256, 256, seed=3).to(DEV)
gen = torch.Generator(device=DEV).manual_seed(0)
mk = lambda c, a, b: torch.rand((B, c, a, b), dtype=torch.float32, device=DEV, generator=gen) - 0.5
noise = {"z": mk(192, 4, 4), "y": mk(320, 16, 16)}
for _ in range(4):
    tr.train_step(x, q=2.0, noise=noise)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(x, q=2.0, noise=noise)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
short = lambda n: n.replace("void ", "").replace("crdr::", "").split("(")[0][:48]
ks = sorted(((e.time_range.start, e.time_range.end, short(e.name)) for e in evs), key=lambda t: t[0])
span = ks[-1][1] - ks[0][0]
busy = sum(e - s for s, e, _ in ks)
# union of busy intervals
union, cur_s, cur_e = 0.0, None, None
for s, e, _ in ks:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            union += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
union += cur_e - cur_s
print(f"batch {B}: {len(ks)} device activities, span {span / 1e3:.2f} ms, sum of kernel times {busy / 1e3:.2f} ms, "
      f"any-kernel-running {union / 1e3:.2f} ms, idle {100 * (span - union) / span:.1f} %")
agg = collections.OrderedDict()
for s, e, n in ks:
    t = agg.setdefault(n, [0, 0.0])
    t[0] += 1
    t[1] += e - s
print("device time by kernel:")
for n, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"  {cnt:4d} {tot / 1e3:9.3f} ms  avg {tot / cnt:7.1f} us  {n}")
# phases: forward ends at the first epi_bwd_kernel / mse_bwd, optimiser starts at sumsq
t0 = ks[0][0]
first = lambda name: next((s for s, e, n in ks if name in n), None)
for name in ("mse_bwd_kernel", "sumsq_kernel", "pack_weights_multi_kernel"):
    f = first(name)
    if f is not None:
        print(f"first {name} at {(f - t0) / 1e3:.2f} ms")

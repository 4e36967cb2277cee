"""Kernel timeline of one encode+decode step from CUPTI (torch.profiler): where the step's non-kernel time sits.
    python tools/timeline_gaps.py [batch] [H W]  -> kernel time, idle time, idle time by (previous kernel -> next kernel)
Never a timing source for bench numbers (the profiler adds host overhead); it only attributes the idle time.
"""
import collections
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT + "/tests")
import fixtures  # noqa: E402


def short(name):
    name = name.replace("void ", "").replace("crdr::", "")
    return name.split("(")[0][:60]


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 768)
    model, _ = fixtures.build_model(seed=0, calibrated=True)
    eng = model.engine()
    x = fixtures.image(b, h, w).cuda()

    def step():
        a = eng.analysis(x, 1.5)
        eng.decode_device(a["z_sym"], a["y_sym"], 1.5, 3.84, (h, w))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, short(e.name)) for e in evs), key=lambda t: t[0])
    if not ks:
        print("no CUDA activity records")
        return
    span = ks[-1][1] - ks[0][0]
    busy = sum(e - s for s, e, _ in ks)
    gaps = collections.OrderedDict()
    idle = 0.0
    big = []
    for (s0, e0, n0), (s1, e1, n1) in zip(ks, ks[1:]):
        g = max(0.0, s1 - e0)
        idle += g
        t = gaps.setdefault((n0, n1), [0, 0.0])
        t[0] += 1
        t[1] += g
        big.append((g, n0, n1))
    print(f"batch {b} {h}x{w}: {len(ks)} device activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {idle / 1e3:.2f} ms "
          f"({100 * idle / span:.1f} %), mean gap {idle / max(len(ks) - 1, 1):.1f} us")
    print("idle time by kernel pair (count, total us, mean us):")
    for (n0, n1), (cnt, tot) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:15]:
        print(f"  {cnt:4d} {tot:9.1f} {tot / cnt:7.1f}   {n0}  ->  {n1}")
    print("largest single gaps (us):")
    for g, n0, n1 in sorted(big, reverse=True)[:10]:
        print(f"  {g:8.1f}   {n0}  ->  {n1}")
    agg = collections.OrderedDict()
    for s, e, n in ks:
        t = agg.setdefault(n, [0, 0.0])
        t[0] += 1
        t[1] += e - s
    print("device time by kernel:")
    for n, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"  {cnt:4d} {tot / 1e3:9.3f} ms  {n}")


if __name__ == "__main__":
    main()

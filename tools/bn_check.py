"""Bring-up check of the fused bottleneck tail (crdr_bottleneck_bc) against the three-launch form; prints timings.
    python tools/bn_check.py [n h w C mid]"""
import sys
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT)
from crdr_b200 import codec, native as nv  # noqa: E402
from crdr_b200.engine import Act  # noqa: E402

n, h, w, c, mid = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (2, 40, 56, 256, 128)
g = torch.Generator().manual_seed(0)
sd = {"a.weight": torch.randn(mid, c, 1, 1, generator=g) / c ** 0.5, "a.bias": torch.randn(mid, generator=g) * 0.1,
      "b.weight": torch.randn(mid, mid, 3, 3, generator=g) / (9 * mid) ** 0.5, "b.bias": torch.randn(mid, generator=g) * 0.1,
      "c.weight": torch.randn(c, mid, 1, 1, generator=g) / mid ** 0.5, "c.bias": torch.randn(c, generator=g) * 0.1}
blk = codec.Bottleneck(sd, ["a", "b", "c"], codec.NetCfg("cuda", 1))
print("fused eligible:", blk.fused)
x = Act.from_nchw(torch.randn(n, c, h, w, generator=g).cuda(), two=False)
add = tuple((torch.randn(k, generator=g) * 0.3).cuda() for k in (mid, mid, c))
scale, shift = (torch.rand(c, generator=g) + 0.5).cuda(), (torch.randn(c, generator=g) * 0.2).cuda()
if "novec" in sys.argv:   # experiment: no per-channel epilogue vectors at all (how much do their shared-memory loads cost?)
    add, scale, shift = (None, None, None), None, None
    for cv in (blk.c1, blk.c2, blk.c3):
        cv.op.bias = None
        for ph in cv.op.phases:
            ph.tmpl.bias = None
    if blk.fused:
        blk.tmpl.bias2 = blk.tmpl.bias3 = None
res = {}
for fused in (False, True):
    codec.FUSE_BC[0] = fused
    for _ in range(2):
        out = blk(x, add=add, scale=scale, shift=shift)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = blk(x, add=add, scale=scale, shift=shift)
    e1.record()
    torch.cuda.synchronize()
    nv.status_check()
    res[fused] = out.hi.clone()
    print(f"fused={fused}: {e0.elapsed_time(e1) / 5:.3f} ms per bottleneck ({n}x{h}x{w}, C={c}, mid={mid})")
d = (res[True].float() - res[False].float()).abs()
print("bit-identical:", torch.equal(res[True], res[False]), "max |diff|", d.max().item(), "mismatching elements", int((d > 0).sum()))
if not torch.equal(res[True], res[False]):
    bad = (d > 0).nonzero()
    print("first mismatches (n, h, w, c):", bad[:8].tolist())
    print("rows with mismatch per image:", [(int(i), int((d[i] > 0).any(-1).sum())) for i in range(n)])
    sys.exit(1)

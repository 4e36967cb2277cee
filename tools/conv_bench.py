"""Time one conv shape on the tcgen05 engine: python tools/conv_bench.py n cin cout h w k stride transposed prec [tile_n]"""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from crdr_b200 import native as nv
from crdr_b200.engine import Act, ConvOp


def bench(n, cin, cout, h, w, k, stride, tr, prec, tile_n=None, epi=0, iters=10):
    g = torch.Generator().manual_seed(0)
    x = Act.from_nchw(torch.randn(n, cin, h, w, generator=g).cuda(), two=True)
    wt = torch.randn(cin, cout, k, k, generator=g) if tr else torch.randn(cout, cin, k, k, generator=g)
    op = ConvOp(wt / (cin * k * k) ** 0.5, torch.zeros(cout), transposed=bool(tr), stride=stride, padding=k // 2,
                output_padding=stride - 1 if tr else 0)
    kw = {}
    if epi:  # residual + per-channel gain, like the last 1x1 of a bottleneck block
        ho, wo = op.out_hw(h, w)
        r = torch.randn(n, cout, ho, wo, generator=g).cuda()
        if epi in (1, 3):
            kw.update(mode=nv.EPI_RESIDUAL, res=Act.from_nchw(r, two=(prec == 0)))
        if epi == 4:
            kw.update(mode=nv.EPI_RESIDUAL, res=r.permute(0, 2, 3, 1).contiguous())
        if epi == 5:
            kw.update(mode=nv.EPI_RESIDUAL, res=Act.from_nchw(r, two=False))
        if epi in (1, 2):
            kw.update(scale=torch.rand(cout).cuda() + 0.5, shift=torch.randn(cout).cuda())
    tile_n = tile_n or None
    out = op(x, precision=prec, tile_n=tile_n, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        op(x, out=out, precision=prec, tile_n=tile_n, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ho, wo = op.out_hw(h, w)
    fl = 2.0 * n * ho * wo * cout * cin * k * k / (stride * stride if tr else 1)
    mult = 3 if prec == 0 else 1
    print(f"n={n} {cin}->{cout} {h}x{w} k{k} s{stride} tr{tr} prec{prec} tile={tile_n} epi={epi}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s alg, {mult*fl/ms/1e9:.1f} MMA")
    nv.status_check()
    import os, ctypes
    if os.environ.get("CRDR_CONV_TRACE"):  # counters exist only in a CRDR_BUILD_TRACE=1 build of the library
        c = (ctypes.c_ulonglong * 10)()
        nv.lib().crdr_debug_counters(c)
        tot, full, d0, acc, patch, kbs, tmma, tcommit, ptot, pwait = [int(v) for v in c]
        print(f"   MMA thread (CTA0): {tot} cyc total, {kbs} k-blocks -> {tot/max(kbs,1):.0f} cyc/kb; waits: full {full/tot:.0%} d0_empty {d0/tot:.0%} acc_empty {acc/tot:.0%} patch {patch/tot:.0%}; issue mma {tmma/tot:.0%} commit {tcommit/tot:.0%}; B producer {ptot} cyc, empty wait {pwait/max(ptot,1):.0%}")


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    bench(*a)

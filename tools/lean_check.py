"""LEAN (TMA-in / TMA-out) epilogue vs the staged epilogue: bit-exact comparison on small / ragged shapes and timing on
the bottleneck shapes of the batch-24 Kodak workload.

    python tools/lean_check.py [check] [bench]
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from crdr_b200 import native as nv
from crdr_b200.engine import Act, ConvOp

X3, X1 = nv.PREC_F16X3, nv.PREC_F16X1


def make(n, cin, cout, h, w, k, prec, epi, seed=0):
    g = torch.Generator().manual_seed(seed)
    two = prec == X3
    x = Act.from_nchw(torch.randn(n, cin, h, w, generator=g).cuda(), two=two)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    op = ConvOp(wt, torch.randn(cout, generator=g), padding=k // 2)
    kw = {}
    if "res" in epi:
        kw.update(mode=nv.EPI_RESIDUAL, res=Act.from_nchw(torch.randn(n, cout, h, w, generator=g).cuda(), two=two))
    if "relu" in epi:
        kw.update(relu=True)
    if "add" in epi:
        kw.update(add_vec=torch.randn(cout, generator=g).cuda())
    if "aff" in epi:
        kw.update(scale=torch.rand(cout, generator=g).cuda() + 0.5, shift=torch.randn(cout, generator=g).cuda())
    return op, x, kw


def run(op, x, kw, prec, lean, swz=1, tile_n=None):
    nv.set_conv_epilogue(lean, swz)
    out = op(x, precision=prec, tile_n=tile_n, **kw)
    torch.cuda.synchronize()
    nv.status_check()
    return out


def check():
    cases = [
        # n, cin, cout, h, w, k, prec, epi, tile_n
        (1, 128, 256, 16, 24, 1, X1, "res", None),
        (3, 128, 256, 16, 24, 1, X1, "res aff add", 256),
        (2, 128, 256, 20, 30, 1, X1, "res aff", 128),
        (1, 256, 128, 16, 8, 1, X1, "relu add", None),
        (2, 128, 128, 19, 13, 3, X1, "relu add", None),
        (1, 96, 192, 16, 24, 1, X3, "res", None),
        (3, 96, 192, 17, 23, 1, X3, "res aff", 96),
        (2, 192, 96, 33, 9, 1, X3, "relu", None),
        (1, 128, 192, 32, 32, 1, X3, "aff", 64),
        (5, 160, 320, 8, 12, 1, X1, "res", 160),
        (1, 64, 64, 1, 1, 1, X3, "", None),
        (7, 96, 256, 24, 40, 1, X3, "res aff", 128),
        (9, 64, 256, 48, 24, 1, X1, "res", 256),
        (4, 128, 128, 64, 64, 3, X1, "res relu", 128),
        (1, 64, 32, 40, 40, 1, X1, "res relu", 32),
    ]
    bad = 0
    for (n, cin, cout, h, w, k, prec, epi, tile_n) in cases:
        op, x, kw = make(n, cin, cout, h, w, k, prec, epi)
        ref = run(op, x, kw, prec, 0, tile_n=tile_n)
        for swz in (1, 0):
            got = run(op, x, kw, prec, 1, swz, tile_n=tile_n)
            same = torch.equal(ref.hi, got.hi) and (ref.lo is None or torch.equal(ref.lo, got.lo))
            if not same:
                bad += 1
                dh = (ref.hi.float() - got.hi.float()).abs()
                print(f"FAIL n={n} {cin}->{cout} {h}x{w} k{k} prec{prec} [{epi}] tile={tile_n} swz={swz}: "
                      f"{int((dh > 0).sum())} of {dh.numel()} hi elements differ, max {dh.max().item():.3e}", flush=True)
            else:
                print(f"PASS n={n} {cin}->{cout} {h}x{w} k{k} prec{prec} [{epi}] tile={tile_n} swz={swz}", flush=True)
    print(f"lean_check: {'ALL PASS' if not bad else str(bad) + ' FAILED'}", flush=True)
    return bad


def bench():
    shapes = [
        ("g_s c3 128->256+skip x1", 24, 128, 256, 256, 384, 1, X1, "res aff"),
        ("g_s c1 256->128 relu x1", 24, 256, 128, 256, 384, 1, X1, "relu add"),
        ("g_s c2 3x3 128->128 x1", 24, 128, 128, 256, 384, 3, X1, "relu add"),
        ("g_a c3 96->192+skip x3", 24, 96, 192, 256, 384, 1, X3, "res aff"),
        ("g_a c1 192->96 relu x3", 24, 192, 96, 256, 384, 1, X3, "relu"),
        ("g_s c3 128->256+skip x1 (M/4)", 24, 128, 256, 128, 192, 1, X1, "res"),
        ("g_a c3 96->192+skip x3 (M/4)", 24, 96, 192, 128, 192, 1, X3, "res"),
    ]
    for (name, n, cin, cout, h, w, k, prec, epi) in shapes:
        op, x, kw = make(n, cin, cout, h, w, k, prec, epi)
        line = f"{name:34s}"
        for lean in (0, 1):
            nv.set_conv_epilogue(lean, 1)
            out = op(x, precision=prec, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                op(x, out=out, precision=prec, **kw)
            e1.record()
            torch.cuda.synchronize()
            nv.status_check()
            line += f"  {'lean' if lean else 'staged'} {e0.elapsed_time(e1) / 10:.3f} ms"
        print(line, flush=True)
        del op, x, kw, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1:] or ["check", "bench"]
    rc = 0
    if "check" in what:
        rc = check()
    if "bench" in what and not rc:
        bench()
    sys.exit(1 if rc else 0)

"""GPU bring-up check for crdr_conv2d: every engine against a float64 torch convolution.

    python tools/conv_check.py [simt|notma|tma] [--big]
Each engine should be run in its own process (a device trap poisons the CUDA context).
"""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from crdr_b200 import native as nv
from crdr_b200.engine import Act, ConvOp

ENG = {"simt": nv.ENGINE_SIMT, "notma": nv.ENGINE_TCGEN05_NOTMA, "tma": nv.ENGINE_TCGEN05}


def ref_conv(x, w, b, transposed, stride, pad, opad):
    x, w, b = x.double(), w.double(), b.double()
    if transposed:
        return F.conv_transpose2d(x, w, b, stride=stride, padding=pad, output_padding=opad)
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def run_case(name, eng, n, cin, cout, h, w, k, stride, transposed, prec, epi="none", seed=0, tile_n=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    pad = k // 2
    opad = stride - 1 if transposed else 0
    x = torch.randn(n, cin, h, w, generator=g).cuda()
    wt = (torch.randn(cin, cout, k, k, generator=g) if transposed else torch.randn(cout, cin, k, k, generator=g))
    wt = (wt / (cin * k * k) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    op = ConvOp(wt, b, transposed=transposed, stride=stride, padding=pad, output_padding=opad)
    xa = Act.from_nchw(x, two=True)
    xq = xa.to_nchw() if prec == nv.PREC_F16X3 else xa.hi.float().permute(0, 3, 1, 2)
    if prec == nv.PREC_F16X3:
        wq = wt  # 22-bit operands: compare against the exact weights
    else:
        wq = wt.half().float()
    ref = ref_conv(xq, wq, b, transposed, stride, pad, opad)
    kw = {}
    ho, wo = ref.shape[2:]
    if epi == "relu_affine":
        sc, sh = torch.rand(cout, generator=g).cuda() + 0.5, torch.randn(cout, generator=g).cuda()
        av = torch.randn(cout, generator=g).cuda()
        kw = dict(relu=True, add_vec=av, scale=sc, shift=sh)
        ref = (torch.relu(ref) + av.double().view(1, -1, 1, 1)) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    elif epi in ("residual", "gate", "half_tanh"):
        r = torch.randn(n, cout, ho, wo, generator=g).cuda()
        ra = Act.from_nchw(r, two=True)
        rq = ra.to_nchw().double()
        if epi == "residual":
            kw = dict(mode=nv.EPI_RESIDUAL, res=ra)
            ref = ref + rq
        elif epi == "gate":
            t = torch.randn(n, cout, ho, wo, generator=g).cuda()
            ta = Act.from_nchw(t, two=True)
            kw = dict(mode=nv.EPI_GATE, res=ra, trunk=ta)
            ref = rq + ta.to_nchw().double() * torch.sigmoid(ref)
        else:
            rf = r.permute(0, 2, 3, 1).contiguous()
            kw = dict(mode=nv.EPI_HALF_TANH, res=rf)
            ref = r.double() + 0.5 * torch.tanh(ref)
    out_f32 = torch.full((n, ho, wo, cout), float("nan"), device="cuda")
    nv.status_reset()
    torch.cuda.synchronize()
    t0 = time.time()
    out = op(xa, precision=prec, engine=eng, out_f32=out_f32, tile_n=tile_n, **kw)
    torch.cuda.synchronize()
    dt = time.time() - t0
    nv.status_check()
    got = out_f32.permute(0, 3, 1, 2).double()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    perr = (out.to_nchw().double() - ref).abs().max().item() / scale
    tol = 2e-5 if prec == nv.PREC_F16X3 else 2e-3
    ptol = tol if prec == nv.PREC_F16X3 else 4e-3
    ok = err < tol and perr < ptol and bool(torch.isfinite(got).all())
    print(f"{'PASS' if ok else 'FAIL'} {name:34s} err_f32={err:.2e} err_planes={perr:.2e} ({dt*1e3:.1f} ms)", flush=True)
    return ok


def main():
    eng_name = sys.argv[1] if len(sys.argv) > 1 else "simt"
    eng = ENG[eng_name]
    big = "--big" in sys.argv
    X3, X1 = nv.PREC_F16X3, nv.PREC_F16X1
    cases = [
        # name, n, cin, cout, h, w, k, stride, transposed, prec, epi
        ("1x1 64->64 x3", 1, 64, 64, 16, 16, 1, 1, False, X3, "none"),
        ("1x1 64->64 x1", 1, 64, 64, 16, 16, 1, 1, False, X1, "none"),
        ("3x3 64->64 x3", 1, 64, 64, 16, 16, 3, 1, False, X3, "none"),
        ("3x3 96->96 x3 relu_affine", 2, 96, 96, 12, 20, 3, 1, False, X3, "relu_affine"),
        ("5x5s2 192->192 x3", 1, 192, 192, 32, 48, 5, 2, False, X3, "none"),
        ("5x5s2 8->192 x3 (odd dims)", 1, 8, 192, 30, 42, 5, 2, False, X3, "none"),
        ("1x1 96->192 x3 residual", 1, 96, 192, 16, 24, 1, 1, False, X3, "residual"),
        ("1x1 160->320 x3 gate", 1, 160, 320, 8, 12, 1, 1, False, X3, "gate"),
        ("3x3 128->32 x3 half_tanh", 1, 128, 32, 8, 12, 3, 1, False, X3, "half_tanh"),
        ("5x5 352->224 x3", 1, 352, 224, 8, 12, 5, 1, False, X3, "none"),
        ("deconv5x5s2 192->256 x3", 1, 192, 256, 8, 12, 5, 2, True, X3, "none"),
        ("deconv3x3s1 256->320 x3", 1, 256, 320, 8, 12, 3, 1, True, X3, "none"),
        ("deconv5x5s2 256->3 x1", 1, 256, 3, 16, 24, 5, 2, True, X1, "none"),
        ("3x3 128->128 x1 relu_affine", 2, 128, 128, 16, 24, 3, 1, False, X1, "relu_affine"),
        ("1x1 128->256 x1 residual", 1, 128, 256, 16, 24, 1, 1, False, X1, "residual"),
    ]
    if big:
        cases += [
            ("5x5 480->224 x3 (charm, n=4)", 4, 480, 224, 32, 48, 5, 1, False, X3, "none"),
            ("5x5s2 192->192 x3 (g_a conv2)", 1, 192, 192, 256, 384, 5, 2, False, X3, "none"),
            ("deconv5x5s2 256->256 x1 (g_s)", 1, 256, 256, 64, 96, 5, 2, True, X1, "none"),
        ]
    bad = 0
    for c in cases:
        try:
            bad += not run_case(c[0], eng, *c[1:])
        except Exception as e:  # keep going unless the context is dead
            bad += 1
            print(f"ERROR {c[0]}: {type(e).__name__}: {e}", flush=True)
            if "CUDA" in str(e) or "cuda" in str(e) or "status" in str(e):
                break
    print(f"[{eng_name}] {len(cases) - bad}/{len(cases)} cases passed", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

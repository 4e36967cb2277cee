"""GPU parity of crdr_conv2d (tcgen05 engine, TMA weights) against float64 torch convolutions, and of the
tcgen05 engine against the scalar cross-check engine.  Tolerances: F16X3 2e-5 of max|out| (fp32-class),
F16X1 2e-3 (fp16 operands)."""
import os
import sys

import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu

X3, X1 = 0, 1
CASES = [
    ("1x1 64->64 x3", 1, 64, 64, 16, 16, 1, 1, False, X3, "none"),
    ("1x1 64->64 x1", 1, 64, 64, 16, 16, 1, 1, False, X1, "none"),
    ("3x3 96->96 x3 relu_affine", 2, 96, 96, 12, 20, 3, 1, False, X3, "relu_affine"),
    ("5x5s2 192->192 x3", 1, 192, 192, 32, 48, 5, 2, False, X3, "none"),
    ("5x5s2 8->192 x3 odd dims", 1, 8, 192, 30, 42, 5, 2, False, X3, "none"),
    ("5x5s2 192->192 x3 odd dims n=3", 3, 192, 192, 31, 45, 5, 2, False, X3, "relu_affine"),
    ("5x5s2 320->256 x3 (h_a)", 2, 320, 256, 32, 48, 5, 2, False, X3, "none"),
    ("5x5s2 256->192 x3 tiny (h_a)", 1, 256, 192, 8, 12, 5, 2, False, X3, "none"),
    ("3x3s2 64->64 x1", 1, 64, 64, 20, 20, 3, 2, False, X1, "none"),
    ("1x1 96->192 x3 residual", 1, 96, 192, 16, 24, 1, 1, False, X3, "residual"),
    ("1x1 160->320 x3 gate", 1, 160, 320, 8, 12, 1, 1, False, X3, "gate"),
    ("3x3 128->32 x3 half_tanh", 1, 128, 32, 8, 12, 3, 1, False, X3, "half_tanh"),
    ("5x5 352->224 x3", 1, 352, 224, 8, 12, 5, 1, False, X3, "none"),
    ("deconv5x5s2 192->256 x3", 1, 192, 256, 8, 12, 5, 2, True, X3, "none"),
    ("deconv3x3s1 256->320 x3", 1, 256, 320, 8, 12, 3, 1, True, X3, "none"),
    ("deconv5x5s2 256->3 x1", 1, 256, 3, 16, 24, 5, 2, True, X1, "none"),
    ("3x3 128->128 x1 relu_affine", 2, 128, 128, 16, 24, 3, 1, False, X1, "relu_affine"),
    ("1x1 128->256 x1 residual", 1, 128, 256, 16, 24, 1, 1, False, X1, "residual"),
    ("5x5 480->224 x3 charm n=4", 4, 480, 224, 32, 48, 5, 1, False, X3, "none"),
    ("single pixel 1x1", 1, 64, 32, 1, 1, 1, 1, False, X3, "none"),
    ("M=129 crosses a tile", 1, 64, 64, 3, 43, 3, 1, False, X3, "none"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("engine", ["tma", "notma"])
def test_conv_matches_float64(case, engine):
    import conv_check
    assert conv_check.run_case(case[0], conv_check.ENG[engine], *case[1:])


def test_simt_cross_check_engine():
    import conv_check
    for case in CASES[:6]:
        assert conv_check.run_case(case[0], conv_check.ENG["simt"], *case[1:])


def test_channel_segments_and_in_place_output():
    """Two input channel ranges of one wide tensor, output written at a channel offset (ChARM support tensor)."""
    import torch.nn.functional as F
    from crdr_b200 import native as nv
    from crdr_b200.engine import Act, ConvOp
    g = torch.Generator().manual_seed(1)
    wide = torch.randn(2, 160, 10, 14, generator=g).cuda()
    T = Act.from_nchw(wide)
    w = (torch.randn(32, 96, 5, 5, generator=g) / 50).cuda()
    b = torch.randn(32, generator=g).cuda()
    op = ConvOp(w, b, padding=2, seg_lens=[64, 32])
    out = Act.zeros(2, 10, 14, 64)
    op(T, segs=[(8, 64), (120, 32)], out=out, out_coff=32)
    torch.cuda.synchronize()
    xin = torch.cat([T.to_nchw()[:, 8:72], T.to_nchw()[:, 120:152]], dim=1).double()
    ref = F.conv2d(xin, w.double(), b.double(), padding=2)
    got = out.to_nchw().double()
    assert (got[:, 32:] - ref).abs().max() / ref.abs().max() < 2e-5
    assert got[:, :32].abs().max() == 0
    nv.status_check()


def test_bad_descriptors_raise():
    from crdr_b200 import native as nv
    from crdr_b200.engine import Act, ConvOp
    op = ConvOp(torch.randn(64, 64, 1, 1), torch.zeros(64))
    x = Act.zeros(1, 4, 4, 64)
    with pytest.raises(nv.NativeError, match="tile_n"):
        op(x, tile_n=24)
    with pytest.raises(nv.NativeError, match="multiples of 8"):
        op(x, segs=[(4, 64)])
    with pytest.raises(nv.NativeError, match="residual"):
        op(x, mode=nv.EPI_RESIDUAL)


def test_fp16_overflow_is_flagged_not_silent():
    from crdr_b200 import native as nv
    from crdr_b200.engine import Act, ConvOp
    w = torch.full((64, 64, 1, 1), 100.0)
    op = ConvOp(w, torch.zeros(64))
    x = Act.from_nchw(torch.full((1, 64, 4, 4), 50.0).cuda())
    nv.status_reset()
    op(x)
    with pytest.raises(nv.NativeError, match="overflow"):
        nv.status_check()
    nv.status_reset()


BN_CASES = [  # (name, n, h, w, C, mid, beta adds, gain)
    ("g_s block 256/128 ragged beta gain", 3, 37, 50, 256, 128, True, True),
    ("g_s block 256/128 single tile", 1, 16, 8, 256, 128, False, False),
    ("g_s block 256/128 many tiles", 5, 64, 96, 256, 128, True, False),
    ("192/96 partial channel block", 2, 33, 24, 192, 96, True, True),
    ("128/64", 1, 20, 20, 128, 64, False, True),
]


@pytest.mark.parametrize("case", BN_CASES, ids=[c[0] for c in BN_CASES])
def test_fused_bottleneck_tail_is_bit_identical(case):
    """crdr_bottleneck_bc (3x3 -> ReLU -> 1x1 + skip in one launch, mid tensor in shared memory) against the same block
    as three crdr_conv2d launches: identical bits, and both within fp16-operand tolerance of a float64 reference."""
    import torch.nn.functional as F
    from crdr_b200 import codec, native as nv
    from crdr_b200.engine import Act
    _, n, h, w, c, mid, beta, gain = case
    g = torch.Generator().manual_seed(n * 1000 + h)
    sd = {"a.weight": torch.randn(mid, c, 1, 1, generator=g) / c ** 0.5, "a.bias": torch.randn(mid, generator=g) * 0.1,
          "b.weight": torch.randn(mid, mid, 3, 3, generator=g) / (9 * mid) ** 0.5, "b.bias": torch.randn(mid, generator=g) * 0.1,
          "c.weight": torch.randn(c, mid, 1, 1, generator=g) / mid ** 0.5, "c.bias": torch.randn(c, generator=g) * 0.1}
    blk = codec.Bottleneck(sd, ["a", "b", "c"], codec.NetCfg("cuda", X1))
    assert blk.fused
    x32 = torch.randn(n, c, h, w, generator=g)
    x = Act.from_nchw(x32.cuda(), two=False)
    vec = lambda k, s=0.3: (torch.randn(k, generator=g) * s).cuda()
    add = (vec(mid), vec(mid), vec(c)) if beta else (None, None, None)
    scale, shift = (vec(c).abs() + 0.5, vec(c)) if gain else (None, None)
    nv.status_reset()
    outs = []
    for fused in (True, False):
        codec.FUSE_BC[0] = fused
        try:
            outs.append(blk(x, add=add, scale=scale, shift=shift).hi.clone())
        finally:
            codec.FUSE_BC[0] = True
    nv.status_check()
    assert torch.equal(outs[0], outs[1]), f"fused != unfused: {(outs[0].float() - outs[1].float()).abs().max().item():.3e}"
    xd = x.to_nchw().double().cpu()
    cw = lambda k: sd[k].double()
    ad = [a.double().cpu().reshape(1, -1, 1, 1) if a is not None else 0.0 for a in add]
    t = torch.relu(F.conv2d(xd, cw("a.weight"), cw("a.bias"))) + ad[0]
    t = torch.relu(F.conv2d(t, cw("b.weight"), cw("b.bias"), padding=1)) + ad[1]
    ref = F.conv2d(t, cw("c.weight"), cw("c.bias")) + ad[2] + xd
    if gain:
        ref = ref * scale.double().cpu().reshape(1, -1, 1, 1) + shift.double().cpu().reshape(1, -1, 1, 1)
    got = outs[0].float().permute(0, 3, 1, 2).double().cpu()
    assert (got - ref).abs().max() / ref.abs().max() < 4e-3

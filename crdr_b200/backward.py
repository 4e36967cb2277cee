"""Backward primitives of the training step above the C ABI (``include/crdr_b200.h``, "Training step").

The reference's backward is ``torch.autograd`` behind ``l_total.backward()`` (rate_distortion_trainer.py:84).  Here the
two contractions autograd would dispatch per convolution are native:

  * wgrad  -- ``crdr_conv_wgrad`` (tcgen05, MN-major operands straight from the NHWC planes, deterministic split-K);
  * dgrad  -- ``crdr_conv_dgrad`` = the forward convolution kernel fed with transposed / flipped weight matrices.

Weight matrices are (re)packed on the device by ``crdr_pack_weights`` through index maps, so a parameter updated by the
optimiser reaches the tensor-core layout with one gather per matrix (forward, dgrad, phase-packed forms alike).
"""
import ctypes as C

import torch

from . import native as nv
from .engine import ConvOp

_WORKSPACE = {}


def workspace(nbytes, device):
    """Growable per-device scratch (split-K partial sums, per-block column sums)."""
    buf = _WORKSPACE.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = _WORKSPACE[device] = torch.empty(max(int(nbytes), 1 << 24), dtype=torch.uint8, device=device)
    return buf


def wgrad(s, s_coff, ca, b, b_coff, cb, taps, stride, out, sa, sb, st, scale=1.0, accumulate=False, ws=None):
    """out[t*st + a*sa + b*sb] (+)= scale * sum_p S[p][a] * B[p*stride + tap_t][b].
    s, b: Act (hi planes are read); out: fp32 device tensor (the parameter gradient, in the parameter's layout).
    ws: scratch tensor of the calling stream (default: the shared per-device workspace); a smaller one than the kernel
    would like only lowers the split-K count."""
    assert s.n == b.n
    d = nv.WgradDesc()
    d.s = nv.Planes(s.hi.data_ptr(), None, s.c, s_coff)
    d.ca = ca
    d.b = nv.Planes(b.hi.data_ptr(), None, b.c, b_coff)
    d.cb = cb
    d.n, d.hs, d.ws, d.hb, d.wb, d.stride = s.n, s.h, s.w, b.h, b.w, stride
    d.ntaps = len(taps)
    for i, (a_, b_) in enumerate(taps):
        d.dh[i], d.dw[i] = a_, b_
    d.out = out.data_ptr()
    d.sa, d.sb, d.st = sa, sb, st
    d.scale, d.accumulate = scale, 1 if accumulate else 0
    L = nv.lib()
    need = L.crdr_conv_wgrad_workspace(C.byref(d))
    if need == 0:
        raise nv.NativeError(f"crdr_conv_wgrad_workspace: {L.crdr_last_error().decode()}")
    if ws is None:
        ws = workspace(need, s.hi.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), min(need, ws.numel() * ws.element_size())
    nv.check(L.crdr_conv_wgrad(C.byref(d), nv.stream_handle()))


def index_weight(shape):
    """A weight-shaped fp32 tensor whose values are 1 + the flat element index (0 = structural zero): pushing it through
    the engines' host-side packing code yields, for every packed matrix element, the master element it comes from."""
    n = 1
    for s_ in shape:
        n *= s_
    assert n < (1 << 24), "fp32 index trick needs < 2^24 elements per parameter"
    return (torch.arange(n, dtype=torch.float32) + 1.0).reshape(shape)


class PackedConv:
    """A ConvOp whose matrices are gathered on the device from a master fp32 parameter (``repack``)."""

    def __init__(self, master, weight_of_index, two_planes, bias=None, **conv_kwargs):
        """master: fp32 device parameter (any shape).  weight_of_index(idx) -> the tensor ConvOp takes as `weight`,
        computed from the index tensor by pure indexing ops (permute / flip / slice / zero padding).  bias: the live
        fp32 device vector the launches read (or None)."""
        self.master = master
        w = weight_of_index(index_weight(tuple(master.shape)))
        self.op = ConvOp(w, bias, device=master.device, index_mode=True, two_planes=two_planes, **conv_kwargs)

    def repack(self):
        L, st = nv.lib(), nv.stream_handle()
        for phs in self.op.phases:
            nv.check(L.crdr_pack_weights(self.master.data_ptr(), phs.map.data_ptr(), phs.map.numel(), phs.w_hi.data_ptr(),
                                         nv.ptr(phs.w_lo), st))


class PackTable:
    """Device-resident job table for crdr_pack_weights_multi: every matrix of a set of PackedConvs in one launch."""

    def __init__(self, packed_convs, device):
        import numpy as np
        jobs = [(pc.master.data_ptr(), phs.map.data_ptr(), phs.map.numel(), phs.w_hi.data_ptr(),
                 phs.w_lo.data_ptr() if phs.w_lo is not None else 0) for pc in packed_convs for phs in pc.op.phases]
        self.count = len(jobs)
        arr = (nv.PackJob * self.count)()
        for i, (m, mp, cnt, hi, lo) in enumerate(jobs):
            arr[i].master, arr[i].map, arr[i].count, arr[i].hi, arr[i].lo = m, mp, cnt, hi, lo or None
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.table = torch.from_numpy(raw).to(device)
        self.keep = list(packed_convs)

    def run(self):
        if self.count:
            nv.check(nv.lib().crdr_pack_weights_multi(self.table.data_ptr(), self.count, nv.stream_handle()))


def dgrad_spec(transposed, stride, padding, kh):
    """(weight_of_index, ConvOp kwargs) of the convolution that maps dL/d(out) to dL/d(in) for a forward
    nn.Conv2d [co, ci, kh, kw] / nn.ConvTranspose2d [ci, co, kh, kw] of the path (5x5 s2 p2 (op 1), 3x3 / 5x5 / 1x1 s1)."""
    if not transposed and stride == 1:
        return (lambda w: w.permute(1, 0, 2, 3).flip(2, 3).contiguous()), dict(padding=kh - 1 - padding)
    if not transposed:     # strided conv: its adjoint is the transposed conv with the same weight tensor
        return (lambda w: w), dict(transposed=True, stride=stride, padding=padding, output_padding=stride - 1)
    if stride == 1:        # ConvTranspose2d stride 1: adjoint = plain conv with [Cout'=ci][Cin'=co], no flip
        return (lambda w: w), dict(padding=padding)
    return (lambda w: w), dict(stride=stride, padding=padding)


def _pad_axis(w, axis, pad):
    if pad == 0:
        return w
    shape = list(w.shape)
    shape[axis] = pad
    return torch.cat([w, torch.zeros(shape, dtype=w.dtype)], dim=axis)


MAX_DGRAD_COUT = 320   # per-CTA epilogue parameter cache of the convolution kernel (kMaxCout)


class DgradSet:
    """dL/d(in) of one forward convolution: one PackedConv per (input channel range, <= 320-channel chunk), each
    accumulating in place (residual epilogue) into the matching channel range of the input's gradient tensor."""

    def __init__(self, master, transposed, stride, padding, kh, segs, cin_real=None, pre=None):
        """segs: [(channel offset in the input tensor, length)] in the forward's logical input-channel order.
        pre: optional transform parameter -> the forward ConvOp's weight (pure indexing ops), applied before slicing."""
        fn0, kw = dgrad_spec(transposed, stride, padding, kh)
        fn = fn0 if pre is None else None
        if pre is not None:
            assert not transposed, "transformed weights are plain convolutions on this path"
        self.parts = []
        start = 0
        for off, ln in segs:
            real = ln if cin_real is None else min(ln, max(0, cin_real - start))
            for c0 in range(0, real, MAX_DGRAD_COUT):
                cnt = min(MAX_DGRAD_COUT, real - c0)
                lo, hi = start + c0, start + c0 + cnt
                pad = (-cnt) % 8      # e.g. the 3 image channels of the discriminator's first layer live in 8-channel planes
                if transposed:   # master [ci, co, kh, kw]
                    sl = (lambda w, lo=lo, hi=hi, pad=pad: _pad_axis(fn(w[lo:hi]), 0, pad))
                else:            # master [co, ci, kh, kw]; the adjoint's weight has the input channels on axis 0 (plain
                    #              convolution, stride 1) or on axis 1 (transposed convolution of a strided forward)
                    if pre is None:
                        sl = (lambda w, lo=lo, hi=hi, pad=pad, ax=(0 if stride == 1 else 1): _pad_axis(fn(w[:, lo:hi]), ax, pad))
                    else:
                        sl = (lambda w, lo=lo, hi=hi, pad=pad, ax=(0 if stride == 1 else 1): _pad_axis(fn0(pre(w)[:, lo:hi]), ax, pad))
                self.parts.append((PackedConv(master, sl, two_planes=False, **kw), off + c0, cnt))
            start += ln

    def repack(self):
        for pc, _, _ in self.parts:
            pc.repack()

    def packed(self):
        return [pc for pc, _, _ in self.parts]

    def run(self, dv, grad, accumulate=True, dv_coff=0, dv_c=None):
        """dv: Act (fp16 gradient of the convolution result; channels [dv_coff, dv_coff + dv_c) when it is a range of a
        wider tensor); grad: Act the input gradient is accumulated into."""
        segs = None if dv_c is None else [(dv_coff, dv_c)]
        for pc, coff, cnt in self.parts:
            if accumulate:
                pc.op(dv, segs=segs, out=grad, out_coff=coff, precision=nv.PREC_F16X1, mode=nv.EPI_RESIDUAL, res=grad, res_coff=coff)
            else:
                pc.op(dv, segs=segs, out=grad, out_coff=coff, precision=nv.PREC_F16X1)

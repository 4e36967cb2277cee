"""Host-side layer engine: activation planes, weight packing and convolution launches.

Everything here is plumbing above the C ABI (``include/crdr_b200.h``): tensors are allocated with
torch, the arithmetic happens in ``libcrdr_sm100.so``.
"""
import ctypes as C
import math
import os

import torch

from . import native as nv

LO_SCALE = 2048.0
LAUNCH_COUNT = nv.LAUNCH_COUNT
PROFILE_ON = [False]  # bench.py: bracket every conv launch with CUDA events on the launching stream
PROFILE = []          # (flops, start_event, end_event)


def round_up(x, m):
    return (x + m - 1) // m * m


class Act:
    """NHWC activation stored as fp16 planes: value ~= hi + lo / 2048 (lo is None for single-term tensors)."""

    __slots__ = ("hi", "lo", "n", "h", "w", "c")

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo
        self.n, self.h, self.w, self.c = hi.shape

    @staticmethod
    def empty(n, h, w, c, two=True, device="cuda"):
        hi = torch.empty((n, h, w, c), dtype=torch.float16, device=device)
        lo = torch.empty((n, h, w, c), dtype=torch.float16, device=device) if two else None
        return Act(hi, lo)

    @staticmethod
    def zeros(n, h, w, c, two=True, device="cuda"):
        hi = torch.zeros((n, h, w, c), dtype=torch.float16, device=device)
        lo = torch.zeros((n, h, w, c), dtype=torch.float16, device=device) if two else None
        return Act(hi, lo)

    @staticmethod
    def from_nchw(x, two=True):
        """Test helper (torch ops, not a hot-path kernel): fp32 NCHW -> planes."""
        xh = x.permute(0, 2, 3, 1).contiguous().float()
        hi = xh.half()
        lo = ((xh - hi.float()) * LO_SCALE).half() if two else None
        return Act(hi, lo)

    def to_nchw(self):
        """Test helper: planes -> fp32 NCHW."""
        v = self.hi.float()
        if self.lo is not None:
            v = v + self.lo.float() / LO_SCALE
        return v.permute(0, 3, 1, 2).contiguous()

    def planes(self, coff=0):
        return nv.Planes(self.hi.data_ptr(), self.lo.data_ptr() if self.lo is not None else None, self.c, coff)

    @property
    def pixels(self):
        return self.n * self.h * self.w


NULL_PLANES = nv.Planes(None, None, 0, 0)


def split_weight(w2d, device):
    """[rows, k] fp32 -> (hi, lo) fp16 with lo = (w - hi) * 2^11.  Packed on the host (one-off, at engine build) and
    uploaded with plain copies, so building an engine launches no device kernels of its own."""
    w2d = w2d.detach().to(device="cpu", dtype=torch.float32).contiguous()
    hi = w2d.half()
    lo = ((w2d - hi.float()) * LO_SCALE).half()
    return hi.contiguous().to(device), lo.contiguous().to(device)


class Phase:
    """One GEMM of a (possibly transposed) convolution: taps, packed weights and output phase."""

    __slots__ = ("dh", "dw", "w_hi", "w_lo", "k_pad", "out_ph", "out_pw", "tmpl", "map")


_TILE_N_CACHE = {}


def pick_tile_n(cout_pad, m_tiles, three):
    """Largest N tile (multiple of 16, divides cout_pad, <= 256 / 128) that still yields >= 148 CTAs; tiles below
    64 columns (shared-memory-bound MMAs) are used only when the layer itself is that narrow."""
    key = (cout_pad, min(m_tiles, 4096), three)
    hit = _TILE_N_CACHE.get(key)
    if hit is None:
        hit = _TILE_N_CACHE[key] = _pick_tile_n(cout_pad, m_tiles, three)
    return hit


MIN_TILE_N = int(os.environ.get("CRDR_MIN_TILE_N", "64"))


def _pick_tile_n(cout_pad, m_tiles, three):
    cap = 128 if three else 256  # F16X3 keeps two D0 buffers, D1 and the fp32 total in 512 TMEM columns
    cands = [t for t in range(16, min(cout_pad, cap) + 1, 16) if cout_pad % t == 0]
    good = [t for t in cands if t >= min(MIN_TILE_N, cands[-1])]
    best = good[0]
    for t in good:
        if m_tiles * (cout_pad // t) >= 148:
            best = t
    return best


class ConvOp:
    """Conv2d / ConvTranspose2d of the reference lowered to implicit-GEMM launches.

    weight: Conv2d  [Cout, Cin, kh, kw]   (nn.Conv2d, e.g. elic_autoencoder.py:42)
            Deconv  [Cin, Cout, kh, kw]   (nn.ConvTranspose2d, elic_layers.py:21)
    """

    def __init__(self, weight, bias, transposed=False, stride=1, padding=0, output_padding=0, cin_pad=None,
                 device="cuda", seg_lens=None, algo=None, index_mode=False, two_planes=True):
        """index_mode (training, backward.PackedConv): `weight` holds 1 + master element indices instead of values; the
        phases keep an int32 gather map and uninitialised matrices that crdr_pack_weights fills from the live parameter,
        and `bias` (if any) is the live fp32 device parameter itself."""
        w = weight.detach().to(torch.float32).cpu()
        self.transposed, self.stride, self.padding, self.output_padding = transposed, stride, padding, output_padding
        if transposed:
            cin, cout, kh, kw = w.shape
        else:
            cout, cin, kh, kw = w.shape
        self.cin_real, self.cout, self.kh, self.kw = cin, cout, kh, kw
        self.cin = cin_pad or cin
        assert self.cin % 8 == 0, "input channels must be padded to a multiple of 8"
        self.cout_pad = round_up(cout, 16)
        if index_mode:
            self.bias = bias
        else:
            self.bias = None if bias is None else bias.detach().to(device="cpu", dtype=torch.float32).contiguous().to(device)
        self.device = device
        self.index_mode = index_mode
        # channel ranges the input is read from at launch (ChARM: hyper ++ slices); decides the 64-channel blocking
        self.seg_lens = list(seg_lens) if seg_lens else [self.cin]
        assert sum(self.seg_lens) == self.cin
        # "patch": halo-patch engine (input stride 1, >= 64 channels); "gather": per-tap im2col gather (any stride)
        # A stride-2 convolution runs on the patch engine as four parity classes of stride-1 taps (each class reads the
        # input sub-sampled by a strided TMA box), so only tiny-Cin layers still need the gather engine.
        in_stride = 1 if transposed else stride
        default = "patch" if (in_stride in (1, 2) and min(self.seg_lens) >= 8 and self.cin >= 64) else "gather"
        self.algo = algo or os.environ.get("CRDR_CONV_ALGO") or default
        if in_stride not in (1, 2) or self.cin < 64 or (in_stride == 2 and os.environ.get("CRDR_CONV_S2_PATCH", "1") == "0"):
            self.algo = "gather"
        self.phases = []
        s, p = stride, padding
        phase_list = [(0, 0)] if not transposed else [(a, b) for a in range(s) for b in range(s)]
        for ph, pw in phase_list:
            taps, cols = [], []
            for i in range(kh):
                for j in range(kw):
                    if transposed:
                        if (ph + p - i) % s or (pw + p - j) % s:
                            continue
                        taps.append(((ph + p - i) // s, (pw + p - j) // s))
                        blk = w[:, :, i, j].t()  # [Cout, Cin]
                    else:
                        taps.append((i - p, j - p))
                        blk = w[:, :, i, j]  # [Cout, Cin]
                    if self.cin != cin:
                        blk = torch.nn.functional.pad(blk, (0, self.cin - cin))
                    cols.append(blk)
            assert 0 < len(taps) <= nv.MAX_TAPS
            if self.algo == "patch" and in_stride > 1:
                # group the taps by parity class (dh mod s, dw mod s), ascending: the K order the strided patch engine walks
                order = sorted(range(len(taps)), key=lambda t: ((taps[t][0] % in_stride) * in_stride + taps[t][1] % in_stride, t))
                taps, cols = [taps[t] for t in order], [cols[t] for t in order]
            phs = Phase()
            phs.dh = [t[0] for t in taps]
            phs.dw = [t[1] for t in taps]
            phs.out_ph, phs.out_pw = ph, pw
            w2d = self._pack(cols)
            phs.k_pad = w2d.shape[1]
            if index_mode:
                phs.map = (w2d.round().to(torch.int32) - 1).contiguous().to(device)
                phs.w_hi = torch.empty(w2d.shape, dtype=torch.float16, device=device)
                phs.w_lo = torch.empty(w2d.shape, dtype=torch.float16, device=device) if two_planes else None
            else:
                phs.w_hi, phs.w_lo = split_weight(w2d, device)
            # launch-descriptor template: everything that does not depend on the call (taps, packed weights, bias).
            # A launch copies it and fills in the tensors (the per-field ctypes stores were most of the enqueue cost).
            t = nv.ConvDesc()
            t.ntaps = len(taps)
            for i, (a, b) in enumerate(taps):
                t.dh[i], t.dw[i] = a, b
            t.w_hi, t.w_lo, t.k_pad = phs.w_hi.data_ptr(), (phs.w_lo.data_ptr() if phs.w_lo is not None else None), phs.k_pad
            t.cout_pad, t.cout = self.cout_pad, self.cout
            t.out_ph, t.out_pw = ph, pw
            t.bias = self.bias.data_ptr() if self.bias is not None else None
            t.k_order = 1 if self.algo == "patch" else 0
            if transposed:
                t.in_stride, t.out_stride = 1, stride
            else:
                t.in_stride, t.out_stride = stride, 1
            phs.tmpl = t
            self.phases.append(phs)

    def _pack(self, cols):
        """cols: per-tap [Cout, Cin] fp32 blocks -> [cout_pad, k_pad] in the K order of the selected engine."""
        ntaps = len(cols)
        if self.algo == "gather":
            k_real = ntaps * self.cin
            w2d = torch.zeros((self.cout_pad, round_up(k_real, 64)), dtype=torch.float32)
            w2d[:self.cout, :k_real] = torch.cat(cols, dim=1)
            return w2d
        blocks = []  # (first logical channel, count) of every 64-channel block, per input range
        start = 0
        for ln in self.seg_lens:
            for b in range(0, ln, 64):
                blocks.append((start + b, min(64, ln - b)))
            start += ln
        w2d = torch.zeros((self.cout_pad, len(blocks) * ntaps * 64), dtype=torch.float32)
        for cb, (c0, cnt) in enumerate(blocks):
            for t in range(ntaps):
                k0 = (cb * ntaps + t) * 64
                w2d[:self.cout, k0:k0 + cnt] = cols[t][:, c0:c0 + cnt]
        return w2d

    def out_hw(self, h, w):
        if self.transposed:
            f = lambda x, k: (x - 1) * self.stride - 2 * self.padding + k + self.output_padding
        else:
            f = lambda x, k: (x + 2 * self.padding - k) // self.stride + 1
        return f(h, self.kh), f(w, self.kw)

    def __call__(self, x, out=None, out_coff=0, segs=None, relu=False, add_vec=None, mode=nv.EPI_NONE, res=None,
                 res_coff=0, trunk=None, scale=None, shift=None, out_f32=None, out_f32_coff=0,
                 precision=nv.PREC_F16X3, engine=nv.ENGINE_TCGEN05, two_out=None, tile_n=None, want_planes=True):
        """x: Act.  segs: [(channel_offset, length), ...] (<= 2) selecting the input channels (default: all).
        res: Act or fp32 NHWC tensor.  Returns the output Act (or None when want_planes is False)."""
        hout, wout = self.out_hw(x.h, x.w)
        if segs is None:
            segs = [(0, self.cin)]
        assert [l for _, l in segs] == self.seg_lens, (segs, self.seg_lens)
        if two_out is None:
            two_out = precision == nv.PREC_F16X3
        if out is None and want_planes:
            out = Act.empty(x.n, hout, wout, self.cout, two=two_out, device=x.hi.device)
        engine_ = nv.ENGINE_TCGEN05 if (self.algo == "patch" and engine == nv.ENGINE_TCGEN05_NOTMA) else engine  # no cp.async patch variant
        st = nv.stream_handle()
        L = nv.lib()
        three = precision == nv.PREC_F16X3
        seg1 = segs[1] if len(segs) > 1 else (0, 0)
        for phs in self.phases:
            d = nv.ConvDesc.from_buffer_copy(phs.tmpl)
            d.inp = x.planes(0)
            d.n, d.hin, d.win = x.n, x.h, x.w
            d.seg0_off, d.seg0_len = segs[0]
            d.seg1_off, d.seg1_len = seg1
            d.hout, d.wout = hout, wout
            if out is not None:
                d.out = out.planes(out_coff)
            if out_f32 is not None:
                d.out_f32, d.out_f32_cs, d.out_f32_coff = out_f32.data_ptr(), out_f32.shape[-1], out_f32_coff
            if relu:
                d.relu = 1
            if add_vec is not None:
                d.add_vec = add_vec.data_ptr()
            if mode:
                d.mode = mode
                if isinstance(res, Act):
                    d.res = res.planes(res_coff)
                elif res is not None:
                    d.res_f32, d.res_f32_cs, d.res_f32_coff = res.data_ptr(), res.shape[-1], res_coff
                if trunk is not None:
                    d.trunk = trunk.planes(0)
            if scale is not None:
                d.scale = scale.data_ptr()
            if shift is not None:
                d.shift = shift.data_ptr()
            d.precision, d.engine = precision, engine_
            if self.transposed:
                hb = -(-(hout - phs.out_ph) // self.stride)
                wb = -(-(wout - phs.out_pw) // self.stride)
            else:
                hb, wb = hout, wout
            d.hb, d.wb = hb, wb
            m_tiles = x.n * (-(-hb // 16)) * (-(-wb // 8)) if self.algo == "patch" else -(-(x.n * hb * wb) // 128)
            d.tile_n = tile_n or pick_tile_n(self.cout_pad, m_tiles, three)
            if PROFILE_ON[0]:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nv.check(L.crdr_conv2d(C.byref(d), st))
                e1.record()
                PROFILE.append((2.0 * x.n * hb * wb * self.cout * len(phs.dh) * self.cin_real, e0, e1,
                                dict(m=x.n * hb * wb, n=self.cout, k=len(phs.dh) * self.cin, taps=len(phs.dh), tile_n=d.tile_n,
                                     prec=precision, transposed=self.transposed, stride=self.stride)))
            else:
                nv.check(L.crdr_conv2d(C.byref(d), st))
        return out

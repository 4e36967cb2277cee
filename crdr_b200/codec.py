"""Lowered CRDR codec: the reference's subnets expressed as sequences of C-ABI kernel launches.

Each ``*Engine`` is built from the corresponding slice of the model ``state_dict`` (reference layout,
fp32 OIHW / IOHW) and owns the packed (hi, lo) fp16 weight matrices.  Activations stay NHWC fp16
planes between launches; element-wise work (bias, ReLU, InterpChAtt gain, beta bias, skip, NLAM gate,
LRP tanh) is fused into the producing convolution's epilogue.

Reference call stacks this replaces: SURVEY.md section 3.1-3.3; per-engine citations below.
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F

import os

from . import engine as engine_mod
from . import native as nv
from .engine import Act, ConvOp, NULL_PLANES

FUSE_BC = [os.environ.get("CRDR_FUSE_BC", "1") != "0"]   # fused bottleneck tail (tests flip it to compare the two forms)

X3, X1 = nv.PREC_F16X3, nv.PREC_F16X1


_SIDE_STREAMS = {}
STREAMS_ON = [os.environ.get("CRDR_CHARM_STREAMS", "1") != "0"]


def run_concurrent(jobs, enabled=True):
    """Enqueue independent launch chains (callables) on side streams, fork / join with events on the current stream.
    Every chain's arithmetic is untouched, so results are bit-identical to running them one after the other."""
    if not (enabled and STREAMS_ON[0]) or len(jobs) < 2:
        for job in jobs:
            job()
        return
    main = torch.cuda.current_stream()
    pool = _SIDE_STREAMS.setdefault(main.device, [])
    while len(pool) < min(len(jobs), 10):
        pool.append(torch.cuda.Stream(device=main.device))
    fork = torch.cuda.Event()
    fork.record(main)
    for i, job in enumerate(jobs):
        side = pool[i % len(pool)]
        side.wait_event(fork)
        with torch.cuda.stream(side):
            job()
        done = torch.cuda.Event()
        done.record(side)
        main.wait_event(done)


def _sub(sd, prefix):
    """state_dict entries below ``prefix.`` with the prefix stripped."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


class Conv:
    """ConvOp + the launch-time defaults (precision / engine) shared by a whole sub-network.

    ``transform`` / ``bias_transform`` map the checkpoint tensors to the tensors the ConvOp takes (g_a conv1 as a 1x1
    convolution over im2col patches, the phase-packed last up-convolution); they are pure indexing ops, so in training
    mode (``cfg.train``: a train.TrainContext) the same function applied to an index tensor yields the gather map that
    re-packs the live parameter on the device after every optimiser step (backward.PackedConv)."""

    def __init__(self, sd, name, cfg, transform=None, bias_transform=None, **kw):
        self.cfg, self.name, self.kw = cfg, name, kw
        w, b = sd[name + ".weight"], sd.get(name + ".bias")
        if cfg.train is None:
            self.op = ConvOp(transform(w) if transform else w, bias_transform(b) if bias_transform else b, device=cfg.device, **kw)
        else:
            cfg.train.adopt(self, w, b, transform, bias_transform, two_planes=cfg.precision == X3, **kw)

    def __call__(self, x, bwd_res=None, no_input_grad=False, **kw):
        kw.setdefault("precision", self.cfg.precision)
        kw.setdefault("engine", self.cfg.engine)
        out = self.op(x, **kw)
        if self.cfg.train is not None and self.cfg.train.tape is not None:
            self.cfg.train.tape.append(("conv", self, x, kw, out, bwd_res, no_input_grad))
        return out


class NetCfg:
    def __init__(self, device, precision, engine=nv.ENGINE_TCGEN05, train=None):
        self.device, self.precision, self.engine, self.train = device, precision, engine, train


class Bottleneck:
    """1x1 -> ReLU -> 3x3 -> ReLU -> 1x1 (+ skip): BaseBlock (elic_layers.py:23-36), NLAMResBlock
    (cheng_nlam.py:32-47) and BetaCondBaseBlock (elic_interpca_beta_cond_autoencoder.py:42-66).

    F16X1 blocks whose accumulators fit TMEM (mid <= 128, C <= 256: every g_s block except the 320-channel NLAM) run the
    3x3 and the last 1x1 as ONE launch (crdr_bottleneck_bc: the mid tensor stays in shared memory); the others as
    three crdr_conv2d launches.  Both forms produce identical bits (CRDR_FUSE_BC=0 selects the unfused form)."""

    def __init__(self, sd, names, cfg):
        self.c1 = Conv(sd, names[0], cfg)
        self.c2 = Conv(sd, names[1], cfg, padding=1)
        self.c3 = Conv(sd, names[2], cfg)
        self.cfg = cfg
        mid, cout = self.c2.op.cout, self.c3.op.cout
        self.fused = (cfg.train is None and cfg.precision == X1 and cfg.engine == nv.ENGINE_TCGEN05 and self.c2.op.algo == "patch"
                      and self.c3.op.algo == "patch" and mid % 32 == 0 and 32 <= mid <= 128 and cout % 32 == 0 and cout <= 256
                      and self.c2.op.cin == mid and self.c3.op.cin == mid)
        if self.fused:
            d = nv.BottleneckDesc()
            p2, p3 = self.c2.op.phases[0], self.c3.op.phases[0]
            d.mid, d.cout = mid, cout
            d.w2, d.k2_pad, d.mid_pad = p2.w_hi.data_ptr(), p2.k_pad, self.c2.op.cout_pad
            d.w3, d.k3_pad, d.cout_pad = p3.w_hi.data_ptr(), p3.k_pad, self.c3.op.cout_pad
            d.bias2 = self.c2.op.bias.data_ptr() if self.c2.op.bias is not None else None
            d.bias3 = self.c3.op.bias.data_ptr() if self.c3.op.bias is not None else None
            d.precision = X1
            self.tmpl = d

    def __call__(self, x, add=(None, None, None), scale=None, shift=None):
        t = self.c1(x, relu=True, add_vec=add[0])
        if self.fused and FUSE_BC[0]:
            return self._tail(t, x, add[1], add[2], scale, shift)
        t = self.c2(t, relu=True, add_vec=add[1])
        return self.c3(t, add_vec=add[2], mode=nv.EPI_RESIDUAL, res=x, scale=scale, shift=shift)

    def _tail(self, t, x, add2, add3, scale, shift):
        out = Act.empty(x.n, x.h, x.w, self.c3.op.cout, two=False, device=x.hi.device)
        d = nv.BottleneckDesc.from_buffer_copy(self.tmpl)
        d.inp, d.res, d.out = t.planes(0), x.planes(0), out.planes(0)
        d.n, d.h, d.w = x.n, x.h, x.w
        d.add2 = add2.data_ptr() if add2 is not None else None
        d.add3 = add3.data_ptr() if add3 is not None else None
        d.scale = scale.data_ptr() if scale is not None else None
        d.shift = shift.data_ptr() if shift is not None else None
        st = nv.stream_handle()
        if engine_mod.PROFILE_ON[0]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nv.check(nv.lib().crdr_bottleneck_bc(C.byref(d), st))
            e1.record()
            m, mid, cout = x.n * x.h * x.w, self.c2.op.cout, self.c3.op.cout
            engine_mod.PROFILE.append((2.0 * m * (9 * mid * mid + mid * cout), e0, e1,
                                       dict(m=m, n=cout, k=9 * mid + mid, taps=10, tile_n=cout, prec=X1, transposed=False, stride=1)))
        else:
            nv.check(nv.lib().crdr_bottleneck_bc(C.byref(d), st))
        return out


class Nlam:
    """ChengNLAM (cheng_nlam.py:5-29): x + trunk(x) * sigmoid(conv1x1(attn(x)))."""

    def __init__(self, sd, cfg):
        rb = lambda p: Bottleneck(sd, [p + ".c1", p + ".c2", p + ".c3"], cfg)
        self.trunk = [rb(f"trunk_block.{i}") for i in range(3)]
        self.attn = [rb(f"attention_block.{i}") for i in range(3)]
        self.conv = Conv(sd, "conv", cfg)

    def __call__(self, x, scale=None, shift=None, out_f32=None):
        res = {}

        def chain(name, blocks):
            t = x
            for b in blocks:
                t = b(t)
            res[name] = t
        # the two branches are independent three-bottleneck chains: side by side (their kernels overlap in the tails)
        run_concurrent([lambda: chain("t", self.trunk), lambda: chain("a", self.attn)])
        train = self.conv.cfg.train
        if train is None:
            return self.conv(res["a"], mode=nv.EPI_GATE, res=x, trunk=res["t"], scale=scale, shift=shift, out_f32=out_f32)
        # training: the gate is its own step so that the logits survive for the backward (crdr_gate_forward / _backward)
        a = self.conv(res["a"])
        out = Act.empty(x.n, x.h, x.w, x.c, two=x.lo is not None, device=x.hi.device)
        d = nv.GateDesc()
        d.x, d.t, d.a, d.out = x.planes(0), res["t"].planes(0), a.planes(0), out.planes(0)
        d.m, d.c, d.scale, d.shift = x.pixels, x.c, nv.ptr(scale), nv.ptr(shift)
        d.blocks = max(1, min(1024, x.pixels // 64))
        if out_f32 is not None:
            d.out_f32, d.out_f32_cs, d.out_f32_coff = out_f32.data_ptr(), out_f32.shape[-1], 0
        nv.check(nv.lib().crdr_gate_forward(C.byref(d), nv.stream_handle()))
        if train.tape is not None:
            train.tape.append(("gate", x, res["t"], a, out, scale, shift))
        return out


class InterpGain:
    """InterpChAtt (interp_channel_attention.py:39-73) reduced to its two per-channel vectors for a
    given quality index: scale = softplus(lerp(weight)), shift = lerp(bias)."""

    def __init__(self, weight, bias, device):
        # the lerp / softplus run on the host (two C-vectors per quality index, cached) and are uploaded once
        self.w = weight.detach().to(device="cpu", dtype=torch.float32)  # (L, 1, C, 1, 1)
        self.b = bias.detach().to(device="cpu", dtype=torch.float32)
        self.device = device
        self.levels = self.w.shape[0]
        self._cache = {}

    def vectors_cpu(self, q):
        """(scale, shift) as CPU fp32 vectors for quality index q."""
        q = float(q)
        if not (0.0 <= q <= self.levels - 1):
            raise AssertionError(f"rate_ind = {q} should be in [0, {self.levels - 1}]")
        ind = torch.tensor(q, dtype=torch.float32)
        lo = torch.floor(ind)
        hi = torch.minimum(lo + 1.0, torch.tensor(float(self.levels - 1)))
        alpha = hi - ind
        l, r = int(lo.item()), int(hi.item())
        w = self.w[l] * alpha + self.w[r] * (1 - alpha)
        b = self.b[l] * alpha + self.b[r] * (1 - alpha)
        return F.softplus(w).reshape(-1).contiguous(), b.reshape(-1).contiguous()

    def vectors(self, q):
        q = float(q)
        if q not in self._cache:
            if not (0.0 <= q <= self.levels - 1):
                raise AssertionError(f"rate_ind = {q} should be in [0, {self.levels - 1}]")
            ind = torch.tensor(q, dtype=torch.float32)
            lo = torch.floor(ind)
            hi = torch.minimum(lo + 1.0, torch.tensor(float(self.levels - 1)))
            alpha = hi - ind
            l, r = int(lo.item()), int(hi.item())
            w = self.w[l] * alpha + self.w[r] * (1 - alpha)
            b = self.b[l] * alpha + self.b[r] * (1 - alpha)
            self._cache[q] = (F.softplus(w).reshape(-1).contiguous().to(self.device),
                              b.reshape(-1).contiguous().to(self.device))
        return self._cache[q]


def _gains(sd, count, device):
    return [InterpGain(sd[f"interp_ca_list.{i}.weight"], sd[f"interp_ca_list.{i}.bias"], device) for i in range(count)]


class StaticVectors:
    """Per-channel epilogue vectors that depend on a call parameter (quality index -> InterpChAtt gains, beta ->
    conditioning biases) behind FIXED device addresses: the kernels always read one flat device buffer, and selecting
    another parameter value is one device-to-device copy from that value's cached flat tensor.  Launch sequences (and
    CUDA graphs of them) are therefore independent of q and beta.  One parameter value is live per engine at a time
    (stream ordered): a caller that interleaves different q / beta on several streams needs one engine per stream."""

    def __init__(self, sizes, device):
        self.sizes, self.device = list(sizes), device
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=device)
        self.views, o = [], 0
        for n in self.sizes:
            self.views.append(self.flat[o:o + n])
            o += n
        self._cache, self._live = {}, None

    def load(self, key, build):
        """build() -> list of CPU fp32 vectors (one per view) for `key`; cached on the device per key."""
        if key != self._live:
            src = self._cache.get(key)
            if src is None:
                src = self._cache[key] = torch.cat([v.reshape(-1).float() for v in build()]).contiguous().to(self.device)
                assert src.numel() == self.flat.numel()
            self.flat.copy_(src)
            self._live = key
        return self.views


class AnalysisEngine:
    """g_a: ElicInterpCaEncoder.forward (elic_interpca_autoencoder.py:22-56; layers elic_autoencoder.py:42-56).
    Every InterpChAtt that follows a layer is folded into that layer's last epilogue."""
    PATCH_CH = 80   # channels of the im2col'd image (crdr_image_to_patches): 75 + zero padding to a multiple of 16

    def __init__(self, sd, device, precision=X3, engine=nv.ENGINE_TCGEN05, train=None):
        cfg = NetCfg(device, precision, engine, train)
        self.cfg = cfg
        # conv1 (5x5, stride 2, 3 -> C; elic_autoencoder.py:42) runs as a 1x1 convolution over the im2col'd image written
        # by crdr_image_to_patches: patch channel (kh*5+kw)*3+c, 75 real + 5 zero = 80 channels = one full K block and one
        # K step of a second (the TMA box of that block zero-fills past channel 80, so the tensor stays 80 wide)
        self.conv1 = Conv(sd, "conv1", cfg, transform=self._patch_weight)
        self.conv2 = Conv(sd, "conv2", cfg, stride=2, padding=2)
        self.conv3 = Conv(sd, "conv3", cfg, stride=2, padding=2)
        self.conv4 = Conv(sd, "conv4", cfg, stride=2, padding=2)
        blk = lambda p: [Bottleneck(_sub(sd, f"{p}.block{i}"), ["conv.0", "conv.2", "conv.4"], cfg) for i in range(3)]
        self.block1, self.block2, self.block3 = blk("block1"), blk("block2"), blk("block3")
        self.attn2 = Nlam(_sub(sd, "attn2"), cfg)
        self.attn4 = Nlam(_sub(sd, "attn4"), cfg)
        # ElicEncoder (stage 1, elic_autoencoder.py:33-72) has no InterpChAtt layers: no gains, q is ignored
        self.gains = _gains(sd, 9, device) if "interp_ca_list.0.weight" in sd else None
        self.gain_vecs = StaticVectors([n for gn in self.gains for n in (gn.w.shape[2],) * 2], device) if self.gains else None
        self.out_ch = sd["conv4.weight"].shape[0]

    @classmethod
    def _patch_weight(cls, w1):
        """(C, 3, 5, 5) -> (C, 80, 1, 1) over the patch channels (kh*5+kw)*3+c."""
        wp = torch.zeros(w1.shape[0], cls.PATCH_CH, 1, 1, dtype=w1.dtype)
        wp[:, :75, 0, 0] = w1.permute(0, 2, 3, 1).reshape(w1.shape[0], 75)
        return wp

    def gain_pairs(self, q):
        """[(scale, shift)] * 9 at fixed device addresses, holding the vectors of quality index q."""
        if not self.gains:
            return [(None, None)] * 9
        if self.cfg.train is not None:
            v = self.cfg.train.live_gains(self, "encoder", q)
        else:
            v = self.gain_vecs.load(float(q), lambda: [t for gn in self.gains for t in gn.vectors_cpu(q)])
        return [(v[2 * i], v[2 * i + 1]) for i in range(len(self.gains))]

    @staticmethod
    def _blocks(blocks, x, g):
        for i, b in enumerate(blocks):
            x = b(x, scale=g[0], shift=g[1]) if i == len(blocks) - 1 else b(x)
        return x

    def run(self, img, q):
        """img: Act (n, H/2, W/2, 128) image patches (crdr_image_to_patches) -> (y planes Act, y fp32 NHWC tensor)."""
        g = self.gain_pairs(q)
        x = self.conv1(img, scale=g[0][0], shift=g[0][1], no_input_grad=True)
        x = self._blocks(self.block1, x, g[1])
        x = self.conv2(x, scale=g[2][0], shift=g[2][1])
        x = self._blocks(self.block2, x, g[3])
        x = self.attn2(x, scale=g[4][0], shift=g[4][1])
        x = self.conv3(x, scale=g[5][0], shift=g[5][1])
        x = self._blocks(self.block3, x, g[6])
        x = self.conv4(x, scale=g[7][0], shift=g[7][1])
        y32 = torch.empty((x.n, x.h, x.w, self.out_ch), dtype=torch.float32, device=x.hi.device)
        y = self.attn4(x, scale=g[8][0], shift=g[8][1], out_f32=y32)
        return y, y32


class HyperAnalysisEngine:
    """h_a: Minnen20HyperEncoder.forward (minnen20_hyperprior.py:9-27)."""

    def __init__(self, sd, device, precision=X3, engine=nv.ENGINE_TCGEN05, train=None):
        cfg = NetCfg(device, precision, engine, train)
        self.conv1 = Conv(sd, "conv1", cfg, padding=1)
        self.conv2 = Conv(sd, "conv2", cfg, stride=2, padding=2)
        self.conv3 = Conv(sd, "conv3", cfg, stride=2, padding=2)
        self.out_ch = sd["conv3.weight"].shape[0]

    def run(self, y):
        t = self.conv1(y, relu=True)
        t = self.conv2(t, relu=True)
        hz, wz = self.conv3.op.out_hw(t.h, t.w)
        z32 = torch.empty((t.n, hz, wz, self.out_ch), dtype=torch.float32, device=t.hi.device)
        self.conv3(t, out_f32=z32, want_planes=False)
        return z32


class HyperSynthesisEngine:
    """h_s: Minnen20HyperDecoder.forward (minnen20_hyperprior.py:30-58); the two branches write their
    320 channels straight into the ChARM support tensor (no torch.cat)."""

    def __init__(self, sd, device, precision=X3, engine=nv.ENGINE_TCGEN05, train=None):
        cfg = NetCfg(device, precision, engine, train)
        mk = lambda p: [Conv(sd, f"{p}.conv1", cfg, transposed=True, stride=2, padding=2, output_padding=1),
                        Conv(sd, f"{p}.conv2", cfg, transposed=True, stride=2, padding=2, output_padding=1),
                        Conv(sd, f"{p}.conv3", cfg, transposed=True, stride=1, padding=1)]
        self.mu, self.std = mk("hd_mu"), mk("hd_std")
        self.out_ch = sd["hd_mu.conv3.weight"].shape[1]

    def run(self, zhat, support, mean_coff, scale_coff, hyper32=None):
        """zhat: Act.  support: Act the outputs are written into.  hyper32: optional fp32 NHWC [.., 2*C]
        receiving cat[mu, std] (the reference's hyper_out) for API / test use."""
        def chain(branch, coff, f32off):
            t = branch[0](zhat, relu=True)
            t = branch[1](t, relu=True)
            branch[2](t, out=support, out_coff=coff, out_f32=hyper32, out_f32_coff=f32off)
        run_concurrent([lambda: chain(self.mu, mean_coff, 0), lambda: chain(self.std, scale_coff, self.out_ch)])


class SliceNet:
    """SliceTransform (minnen20_charm_context_model.py:26-38): conv5x5 -> ReLU -> conv5x5 -> ReLU -> conv3x3."""

    def __init__(self, sd, cfg, seg_lens):
        self.c1 = Conv(sd, "model.0", cfg, padding=2, seg_lens=seg_lens)
        self.c2 = Conv(sd, "model.2", cfg, padding=2)
        self.c3 = Conv(sd, "model.4", cfg, padding=1)

    def __call__(self, x, segs, bwd_res=None, **last):
        t = self.c1(x, segs=segs, relu=True)
        t = self.c2(t, relu=True)
        return self.c3(t, bwd_res=bwd_res, **last)


class GaussianParams:
    """Constants of the GaussianConditional the kernels need (scale table on device, bound)."""

    def __init__(self, scale_table, scale_bound, device):
        self.table = scale_table.detach().to(device=device, dtype=torch.float32).contiguous()
        self.bound = float(scale_bound)


class CharmEngine:
    """Minnen20CharmContextModel.forward / forward_compress / forward_decompress
    (minnen20_charm_context_model.py:88-240).

    Support tensor layout (one NHWC Act, `cs` channels):  [hyper_scale | hyper_mean | y_hat slices | scratch slice]
    so that the mean / LRP nets read ``hyper_mean ++ y_hat[:k]`` as ONE contiguous channel range and the
    scale net reads two ranges; nothing is concatenated or copied.
    """

    def __init__(self, sd, num_slices, slice_ch, hyper_ch, max_support, device, precision=X3,
                 engine=nv.ENGINE_TCGEN05, train=None):
        cfg = NetCfg(device, precision, engine, train)
        self.cfg = cfg
        self.S, self.sc, self.hc = num_slices, slice_ch, hyper_ch
        self.max_support = max_support
        self.off_scale, self.off_mean = 0, hyper_ch
        self.off_y = 2 * hyper_ch
        self.off_tmp = self.off_y + num_slices * slice_ch
        self.cs = self.off_tmp + slice_ch * num_slices  # one scratch slot per slice (slices of a group run together)
        self.yc = num_slices * slice_ch
        lens = lambda segs: [l for _, l in segs]
        self.mean = [SliceNet(_sub(sd, f"mean_slice_transforms.{i}"), cfg, lens(self._segs_mean(i))) for i in range(num_slices)]
        self.scale = [SliceNet(_sub(sd, f"scale_slice_transforms.{i}"), cfg, lens(self._segs_scale(i))) for i in range(num_slices)]
        self.lrp = [SliceNet(_sub(sd, f"lrp_slice_transforms.{i}"), cfg, lens(self._segs_lrp(i))) for i in range(num_slices)]

    def n_support(self, s):
        return s if self.max_support < 0 else min(s, self.max_support)

    # -- small batches: independent slice networks side by side ---------------------------------------------------
    # One slice network of a small batch fills a fraction of the GPU (Kodak batch 1: M = 1536 = 6 CTA pairs of 74 per N
    # tile), and its K loop cannot be split without changing the summation order that the decoder must reproduce.  What
    # CAN run concurrently are the networks that do not depend on each other: (mean, scale) of a slice, and all networks of
    # the slices 5..9 group.  They are enqueued on side streams (fork / join with events); every network's arithmetic is
    # untouched, so the results are bit-identical to the serial order (batch-size invariance is a decoder requirement).
    # Side streams are used when one launch occupies at most this many CTA pairs per N tile (CRDR_CHARM_STREAMS_MAX).
    # Measured on B200 (ms per encode + decode step, serial -> concurrent): batch 1 13.7 -> 10.8, batch 8 29.9 -> 26.6,
    # batch 24 69.9 -> 69.4 (the persistent kernels of a full-GPU launch only overlap in their last wave), so: always.
    CONCURRENT_MAX_PAIR_TILES = int(os.environ.get("CRDR_CHARM_STREAMS_MAX", str(1 << 30)))

    def _concurrent(self, T):
        pair_tiles = (T.n * (-(-T.h // 16)) * (-(-T.w // 8)) + 1) // 2
        return pair_tiles <= self.CONCURRENT_MAX_PAIR_TILES

    def _run_jobs(self, jobs, concurrent):
        run_concurrent(jobs, concurrent)

    def new_support(self, n, h, w, device):
        # zero-filled: the 64-channel blocks of the patch engine may read channels that are written later
        # (their weights are zero, but NaN bit patterns of uninitialised memory would still poison the sum)
        return Act.zeros(n, h, w, self.cs, two=True, device=device)

    def groups(self):
        """Slices whose (mu, sigma) depend only on already-finished slices can be processed together."""
        if self.max_support < 0:
            return [[s] for s in range(self.S)]
        head = [[s] for s in range(min(self.max_support, self.S))]
        tail = list(range(self.max_support, self.S))
        return head + ([tail] if tail else [])

    def _segs_mean(self, s):
        return [(self.off_mean, self.hc + self.sc * self.n_support(s))]

    def _segs_scale(self, s):
        k = self.n_support(s)
        return [(self.off_scale, self.hc)] + ([(self.off_y, self.sc * k)] if k else [])

    def _segs_lrp(self, s):
        return [(self.off_mean, self.hc + self.sc * self.n_support(s)), (self.off_tmp + s * self.sc, self.sc)]

    def params(self, T, s, ms):
        """mu_s, sigma_s -> fp32 NHWC scratch `ms` laid out [mu of all slices | sigma of all slices], so that the
        element-wise kernels cover a whole dependency group (consecutive slices) in one launch."""
        for job in self.param_jobs(T, s, ms):
            job()

    def param_jobs(self, T, s, ms):
        return [lambda: self.mean[s](T, self._segs_mean(s), out_f32=ms, out_f32_coff=s * self.sc, want_planes=False),
                lambda: self.scale[s](T, self._segs_scale(s), out_f32=ms, out_f32_coff=self.yc + s * self.sc, want_planes=False)]

    def refine(self, T, s, yq32, yhat32):
        """LRP: y_hat_s = yq_s + 0.5 tanh(lrp(...)) -> support tensor (planes) and yhat32 (fp32 NHWC)."""
        self.lrp[s](T, self._segs_lrp(s), mode=nv.EPI_HALF_TANH, res=yq32, res_coff=s * self.sc, out=T,
                    out_coff=self.off_y + s * self.sc, out_f32=yhat32, out_f32_coff=s * self.sc,
                    bwd_res=(T, self.off_tmp + s * self.sc))   # backward: the quantised slice's planes live in the scratch slot

    def gauss_desc(self, gp, T, s, cnt, n, hw, ms, y32=None, yq32=None, sym=None, idx=None, lik=None, sym16=None,
                   idx8=None, noise=None, lik_noisy=None):
        """Descriptor covering `cnt` consecutive slices starting at slice s."""
        d = nv.GaussDesc()
        if y32 is not None:
            d.y, d.y_cs, d.y_coff = y32.data_ptr(), self.yc, s * self.sc
        d.mu, d.sigma = ms.data_ptr(), ms.data_ptr()
        d.ms_cs, d.mu_coff, d.sigma_coff = ms.shape[-1], s * self.sc, self.yc + s * self.sc
        d.n, d.hw, d.c = n, hw, self.sc * cnt
        d.scale_bound, d.scale_table, d.ntable = gp.bound, gp.table.data_ptr(), gp.table.numel()
        d.yq_planes = T.planes(self.off_tmp + s * self.sc)
        if yq32 is not None:
            d.yq_f32, d.yq_f32_cs, d.yq_f32_coff = yq32.data_ptr(), self.yc, s * self.sc
        d.symbols = sym.data_ptr() if sym is not None else None
        d.indexes = idx.data_ptr() if idx is not None else None
        d.likelihood = lik.data_ptr() if lik is not None else None
        d.symbols16 = sym16.data_ptr() if sym16 is not None else None
        d.indexes8 = idx8.data_ptr() if idx8 is not None else None
        d.noise = noise.data_ptr() if noise is not None else None
        d.likelihood_noisy = lik_noisy.data_ptr() if lik_noisy is not None else None
        d.c_total, d.nchw_coff = self.yc, s * self.sc
        return d

    def encode(self, T, y32, gp, compact=False, noise=None):
        """Encoder-side pass.  Returns y_hat fp32 NHWC, symbols / indexes int32 NCHW, likelihood fp32 NCHW and, with
        ``compact``, the int16 symbols / uint8 indexes copies the host range coder reads.  ``noise`` (NCHW fp32, uniform
        in [-1/2, 1/2)) selects the training-mode forward (minnen20_charm_context_model.py:88-141 with is_train=True): the
        likelihood of y + noise is appended to the result; y_hat / likelihood are then the straight-through forward values
        and the quantised likelihoods (q_likelihoods)."""
        n, h, w = T.n, T.h, T.w
        dev = T.hi.device
        ms = torch.empty((n, h, w, 2 * self.yc), dtype=torch.float32, device=dev)
        yq32 = torch.empty((n, h, w, self.yc), dtype=torch.float32, device=dev)
        yhat32 = torch.empty((n, h, w, self.yc), dtype=torch.float32, device=dev)
        sym = torch.empty((n, self.yc, h, w), dtype=torch.int32, device=dev)
        idx = torch.empty((n, self.yc, h, w), dtype=torch.int32, device=dev)
        lik = torch.empty((n, self.yc, h, w), dtype=torch.float32, device=dev)
        sym16 = torch.empty((n, self.yc, h, w), dtype=torch.int16, device=dev) if compact else None
        idx8 = torch.empty((n, self.yc, h, w), dtype=torch.uint8, device=dev) if compact else None
        lik_noisy = torch.empty((n, self.yc, h, w), dtype=torch.float32, device=dev) if noise is not None else None
        L = nv.lib()
        conc = self._concurrent(T)
        for grp in self.groups():
            self._run_jobs([j for s in grp for j in self.param_jobs(T, s, ms)], conc)
            d = self.gauss_desc(gp, T, grp[0], len(grp), n, h * w, ms, y32=y32, yq32=yq32, sym=sym, idx=idx, lik=lik,
                                sym16=sym16, idx8=idx8, noise=noise, lik_noisy=lik_noisy)
            nv.check(L.crdr_gauss_quantize(C.byref(d), nv.stream_handle()))
            if self.cfg.train is not None and self.cfg.train.tape is not None:
                self.cfg.train.tape.append(("gauss", grp[0], len(grp), T, y32, ms, noise))
            self._run_jobs([(lambda s=s: self.refine(T, s, yq32, yhat32)) for s in grp], conc)
        if noise is not None:
            return yhat32, sym, idx, lik, lik_noisy
        if compact:
            return yhat32, sym, idx, lik, sym16, idx8
        return yhat32, sym, idx, lik

    def decode_steps(self, T, gp, compact=False):
        """Decoder-side pass as a generator: yields ``(first_slice, count, idx_nchw)`` whenever the symbols of a slice group
        are needed (``idx_nchw`` holds the CDF indexes of that group: int32, or uint8 with ``compact``) and expects the
        int32 NCHW symbols (device tensor, full [n, yc, h, w] buffer with that slice range filled) to be sent back.
        Returns y_hat fp32 NHWC.  The generator form lets the caller interleave several independent batches on one stream."""
        n, h, w = T.n, T.h, T.w
        dev = T.hi.device
        ms = torch.empty((n, h, w, 2 * self.yc), dtype=torch.float32, device=dev)
        yq32 = torch.empty((n, h, w, self.yc), dtype=torch.float32, device=dev)
        yhat32 = torch.empty((n, h, w, self.yc), dtype=torch.float32, device=dev)
        idx = torch.empty((n, self.yc, h, w), dtype=torch.uint8 if compact else torch.int32, device=dev)
        L = nv.lib()
        conc = self._concurrent(T)
        for grp in self.groups():
            self._run_jobs([j for s in grp for j in self.param_jobs(T, s, ms)], conc)
            d = self.gauss_desc(gp, T, grp[0], len(grp), n, h * w, ms, **({"idx8": idx} if compact else {"idx": idx}))
            nv.check(L.crdr_gauss_indexes(C.byref(d), nv.stream_handle()))
            sym = yield grp[0], len(grp), idx
            d = self.gauss_desc(gp, T, grp[0], len(grp), n, h * w, ms, yq32=yq32, sym=sym)
            nv.check(L.crdr_gauss_dequantize(C.byref(d), nv.stream_handle()))
            self._run_jobs([(lambda s=s: self.refine(T, s, yq32, yhat32)) for s in grp], conc)
        return yhat32

    def decode(self, T, gp, symbol_source):
        """Decoder-side pass.  ``symbol_source(first_slice, count, idx_nchw)`` returns the int32 NCHW symbols
        (device tensor, full [n, yc, h, w] buffer with that slice range filled) for the group.
        Returns y_hat fp32 NHWC."""
        steps = self.decode_steps(T, gp)
        try:
            req = next(steps)
            while True:
                req = steps.send(symbol_source(*req))
        except StopIteration as done:
            return done.value


class SynthesisEngine:
    """g_s: ElicInterpCaBetaCondDecoder.forward (elic_interpca_beta_cond_autoencoder.py:88-162).
    InterpChAtt precedes each layer here, so gain i is folded into the epilogue of layer i-1; gain 0 is
    applied while converting y_hat to planes.  The beta embedding (fourier_cond.py:21-37), its MLP and
    the 27 projections are evaluated once per beta and enter as per-channel epilogue vectors."""

    def __init__(self, sd, max_beta=0.0, L=0, use_pi=False, include_x=False, use_tanh=False, device="cuda", precision=X1,
                 engine=nv.ENGINE_TCGEN05, train=None):
        if use_tanh:
            raise NotImplementedError("use_tanh=True decoders are not lowered (crdr.yaml uses use_tanh: False)")
        cfg = NetCfg(device, precision, engine, train)
        self.cfg, self.device = cfg, device
        self.two = precision == X3
        up = lambda name: Conv(sd, name, cfg, transposed=True, stride=2, padding=2, output_padding=1)
        self.attn1 = Nlam(_sub(sd, "attn1"), cfg)
        self.conv1, self.conv2, self.conv3 = up("conv1"), up("conv2"), up("conv3")
        self.attn2 = Nlam(_sub(sd, "attn2"), cfg)
        self.conv4p = Conv(sd, "conv4", cfg, transform=self._phase_weight, bias_transform=self._phase_bias, padding=1)
        self.blocks, self.proj = {}, {}
        # ElicDecoder / ElicInterpCaDecoder (stages 1 / 2: elic_autoencoder.py:75-119, elic_interpca_autoencoder.py:59-97)
        # have no beta conditioning; ElicDecoder has no InterpChAtt gains either
        self.has_cond = "mlp.0.weight" in sd
        for b in ("block1", "block2", "block3"):
            self.blocks[b] = [Bottleneck(_sub(sd, f"{b}.block{i}"), ["conv.0", "conv.2", "conv.4"], cfg) for i in range(3)]
            if self.has_cond:
                # beta embedding MLP and the 27 projections are GEMVs evaluated once per beta: host side, results uploaded
                self.proj[b] = [[(sd[f"{b}.block{i}.proj_{k}.weight"].detach().to(device="cpu", dtype=torch.float32).flatten(1),
                                  sd[f"{b}.block{i}.proj_{k}.bias"].detach().to(device="cpu", dtype=torch.float32))
                                 for k in (1, 2, 3)] for i in range(3)]
        if self.has_cond:
            self.mlp = [(sd[f"mlp.{i}.weight"].detach().to(device="cpu", dtype=torch.float32),
                         sd[f"mlp.{i}.bias"].detach().to(device="cpu", dtype=torch.float32)) for i in (0, 2)]
        self.gains = _gains(sd, 9, device) if "interp_ca_list.0.weight" in sd else None
        self.gain_vecs = StaticVectors([n for gn in self.gains for n in (gn.w.shape[2],) * 2], device) if self.gains else None
        if self.has_cond:
            self.cond_vecs = StaticVectors([w.shape[0] for b in ("block1", "block2", "block3") for blk in self.proj[b] for (w, _) in blk],
                                           device)
        self.max_beta, self.L, self.include_x = float(max_beta), int(L), bool(include_x)
        self.freq = torch.pow(torch.Tensor([2]), torch.arange(L))
        if use_pi:
            self.freq = self.freq * math.pi
        self._beta_cache = {}
        self.in_ch = sd["conv1.weight"].shape[0]

    @staticmethod
    def _phase_weight(w):
        """The last up-convolution (ConvTranspose2d 5x5, stride 2, padding 2, output_padding 1, Cout = 3) as ONE
        stride-1 3x3 convolution with 4 phases x 3 = 12 (padded to 16) output channels: out[2a+ph, 2b+pw, c] uses the
        taps kh = ph + 2 - 2*dh, kw = pw + 2 - 2*dw (dh, dw in -1..1).  2.8x fewer (tiny-N) MMAs than four phase
        launches; exact same products, so results are bit-identical to the phase form up to summation order."""
        cin, cout, kh, kw = w.shape
        assert (kh, kw) == (5, 5) and cout * 4 <= 16
        wc = torch.zeros(16, cin, 3, 3, dtype=torch.float32)
        wf = w.detach().float().cpu()
        for ph in range(2):
            for pw in range(2):
                o = (ph * 2 + pw) * cout
                for dh in (-1, 0, 1):
                    for dw in (-1, 0, 1):
                        i, j = ph + 2 - 2 * dh, pw + 2 - 2 * dw
                        if 0 <= i < 5 and 0 <= j < 5:
                            wc[o:o + cout, :, dh + 1, dw + 1] = wf[:, :, i, j].t()
        return wc

    @staticmethod
    def _phase_bias(b):
        bc = torch.zeros(16, dtype=torch.float32, device=b.device)
        bc[:4 * b.numel()] = b.detach().float().repeat(4)
        return bc

    def _cond_cpu(self, beta):
        """The 27 conditioning bias vectors (block, bottleneck, proj_1..3) for beta, CPU fp32 (fourier_cond.py:21-37,
        elic_interpca_beta_cond_autoencoder.py:142-152,52-65)."""
        if not (0 <= beta <= self.max_beta):
            raise AssertionError(f"beta = {beta} should be in [0, {self.max_beta}]")
        b = torch.Tensor([beta]).float()
        nb = (b / self.max_beta - 0.5) * 2
        emb = torch.cat([torch.sin(nb * self.freq), torch.cos(nb * self.freq)], dim=0)
        if self.include_x:
            emb = torch.cat([nb, emb], dim=0)
        c = emb.unsqueeze(0)
        c = F.linear(torch.relu(F.linear(c, *self.mlp[0])), *self.mlp[1])  # [1, cond_ch]
        return [F.linear(c, w, bb).reshape(-1) for b_ in ("block1", "block2", "block3") for blk in self.proj[b_] for (w, bb) in blk]

    def cond_vectors(self, beta):
        """{block: [(add1, add2, add3)] * 3} at fixed device addresses, holding the vectors of `beta`."""
        if not self.has_cond:
            return {b: [(None, None, None)] * 3 for b in self.blocks}
        if self.cfg.train is not None:
            v = self.cfg.train.live_cond(self, beta)
        else:
            beta = float(beta)
            v = self.cond_vecs.load(beta, lambda: self._cond_cpu(beta))
        return {b_: [tuple(v[(bi * 3 + i) * 3 + k] for k in range(3)) for i in range(3)]
                for bi, b_ in enumerate(("block1", "block2", "block3"))}

    def gain_pairs(self, q):
        if not self.gains:
            return [(None, None)] * 9
        if self.cfg.train is not None:
            v = self.cfg.train.live_gains(self, "decoder", q)
        else:
            v = self.gain_vecs.load(float(q), lambda: [t for gn in self.gains for t in gn.vectors_cpu(q)])
        return [(v[2 * i], v[2 * i + 1]) for i in range(len(self.gains))]

    def _blocks(self, name, x, cond, g):
        blks = self.blocks[name]
        for i, b in enumerate(blks):
            last = i == len(blks) - 1
            x = b(x, add=cond[name][i], scale=g[0] if last else None, shift=g[1] if last else None)
        return x

    def run(self, yhat32, q, beta):
        """yhat32: fp32 NHWC [n, h, w, C] -> phase-packed fp32 image [n, 8h, 8w, 16] (see _phase_packed)."""
        g = self.gain_pairs(q)
        cond = self.cond_vectors(beta)
        n, h, w, c = yhat32.shape
        x = Act.empty(n, h, w, c, two=self.two, device=yhat32.device)
        nv.check(nv.lib().crdr_affine_to_planes(yhat32.data_ptr(), c, 0, n * h * w, c, nv.ptr(g[0][0]), nv.ptr(g[0][1]),
                                                x.planes(0), nv.stream_handle()))
        if self.cfg.train is not None and self.cfg.train.tape is not None:
            self.cfg.train.tape.append(("affine_in", x, g[0][0], g[0][1], yhat32))
        x = self.attn1(x, scale=g[1][0], shift=g[1][1])
        x = self.conv1(x, scale=g[2][0], shift=g[2][1])
        x = self._blocks("block1", x, cond, g[3])
        x = self.conv2(x, scale=g[4][0], shift=g[4][1])
        x = self.attn2(x, scale=g[5][0], shift=g[5][1])
        x = self._blocks("block2", x, cond, g[6])
        x = self.conv3(x, scale=g[7][0], shift=g[7][1])
        x = self._blocks("block3", x, cond, g[8])
        # phase-packed final up-convolution: [n, h/2, w/2, 16] with channel = (phase_y*2 + phase_x)*3 + c
        img = torch.empty((n, x.h, x.w, 16), dtype=torch.float32, device=yhat32.device)
        self.conv4p(x, out_f32=img, want_planes=False)
        return img


class DiscriminatorEngine:
    """CLIC21GVAEDiscriminator.forward (clic21_gvae_discriminator.py:27-50, norm_type none): 3x3 convolutions with
    LeakyReLU(0.2) between them on single-plane fp16 tensors.  The image enters as 8-channel NHWC planes
    (crdr_image_to_planes: 3 channels + zero padding); layers wider than 256 output channels run as two launches over
    output-channel halves (the kernel's per-CTA parameter cache holds 320); the 1-channel head is padded to 8 channels
    (channel 0 is the logit map)."""
    SLOPE = 0.2

    def __init__(self, sd, strides, device, precision=X1, engine=nv.ENGINE_TCGEN05, train=None):
        cfg = NetCfg(device, precision, engine, train)
        self.cfg = cfg
        self.layers = []
        names = sorted((k[:-len(".weight")] for k in sd if k.endswith(".weight")), key=lambda k: int(k.split(".")[1]))
        assert len(names) == len(strides)
        for i, (name, stride) in enumerate(zip(names, strides)):
            cout, cin = sd[name + ".weight"].shape[:2]
            last = i == len(names) - 1
            kw = dict(stride=stride, padding=1, cin_pad=8 if cin < 8 else None)
            if last:
                parts = [(Conv(sd, name, cfg, transform=self._pad_head, bias_transform=self._pad_head_bias, **kw), 0, 8)]
                width = 8
            elif cout > 256:
                half = cout // 2
                parts = [(Conv(sd, name, cfg, transform=(lambda w, a=a: w[a:a + half]), bias_transform=(lambda b, a=a: b[a:a + half]), **kw),
                          a, half) for a in (0, half)]
                width = cout
            else:
                parts = [(Conv(sd, name, cfg, **kw), 0, cout)]
                width = cout
            self.layers.append((parts, width, stride, not last))

    @staticmethod
    def _pad_head(w):
        out = torch.zeros((8,) + tuple(w.shape[1:]), dtype=torch.float32)
        out[:w.shape[0]] = w.detach().float().cpu()
        return out

    @staticmethod
    def _pad_head_bias(b):
        out = torch.zeros(8, dtype=torch.float32, device=b.device)
        out[:b.numel()] = b.detach().float()
        return out

    def run(self, x, no_input_grad=False):
        """x: Act [n, h, w, 8] (single plane).  Returns the logits as an Act [n, h/16, w/16, 8] (channel 0)."""
        train = self.cfg.train
        for li, (parts, width, stride, act) in enumerate(self.layers):
            ho, wo = parts[0][0].op.out_hw(x.h, x.w)
            out = Act.empty(x.n, ho, wo, width, two=False, device=x.hi.device)
            for conv, coff, cnt in parts:
                conv(x, out=out, out_coff=coff, no_input_grad=(no_input_grad and li == 0))
            if act:
                nv.check(nv.lib().crdr_leaky_relu(out.planes(0), out.pixels, width, self.SLOPE, nv.stream_handle()))
                if train is not None and train.tape is not None:
                    train.tape.append(("lrelu", out, self.SLOPE))
            x = out
        return x

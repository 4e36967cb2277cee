"""Model API of the reference (`src/models/comp_model/*`), driven by the lowered CUDA engines.

`BetaCondInterpCaHyperpriorCharmModel` keeps the reference's public surface -- ``compress`` /
``decompress`` / ``forward`` / ``run_model`` / ``codec_setup`` / ``load_learned_weight`` /
``load_state_dict`` / ``separate_aux_parameters`` / ``aux_loss`` / ``validation``
(beta_cond_interpca_hyperprior_charm_model.py:13-149, beta_cond_interpca_hyperprior_model.py:28-64,
hyperprior_model.py:60-136, base_model.py:35-167) -- and the 595-key checkpoint layout.  Additions:
``compress_batch`` / ``decompress_batch`` (the reference asserts N == 1) and the device-only
``encode_device`` / ``decode_device`` used for device-timed throughput.

There is no PyTorch/CPU fallback: the arithmetic runs in libcrdr_sm100.so or not at all.
"""
import ctypes as C
import os
from collections import OrderedDict
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from . import native as nv
from . import rans
from .codec import GaussianParams
from .codec_utils import HeaderHandler, MultiRateHeaderHandler
from .engine import Act
from .entropy import EntropyBottleneck, GaussianMeanScaleConditional, get_scale_table
from .logger import get_root_logger
from .registry import (CONTEXTMODEL_REGISTRY, DECODER_REGISTRY, ENCODER_REGISTRY, ENTROPYMODEL_REGISTRY,
                       HYPERDECODER_REGISTRY, HYPERENCODER_REGISTRY, MODEL_REGISTRY)
from . import subnets  # noqa: F401  (registers the sub-network classes)

_REGISTRY_OF = {
    "encoder": ENCODER_REGISTRY, "decoder": DECODER_REGISTRY, "hyperencoder": HYPERENCODER_REGISTRY,
    "hyperdecoder": HYPERDECODER_REGISTRY, "context_model": CONTEXTMODEL_REGISTRY,
    "entropy_model": ENTROPYMODEL_REGISTRY,
}


def build_subnet(subnet_opt, subnet_type):
    """`type:` selects the class, the remaining keys are its kwargs (subnet/__init__.py:16-43)."""
    kw = dict(deepcopy(subnet_opt))
    cls = _REGISTRY_OF[subnet_type].get(kw.pop("type"))
    return cls(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})


def build_comp_model(opt):
    """models/__init__.py:21-30."""
    opt = deepcopy(opt)
    return MODEL_REGISTRY.get(opt["model_type"])(opt)


class CodecEngine:
    """All lowered sub-networks of one model on one device + the launch sequences of the codec."""

    def __init__(self, model, device, precision_main=nv.PREC_F16X3, precision_synthesis=nv.PREC_F16X1,
                 conv_engine=nv.ENGINE_TCGEN05):
        kw = dict(engine=conv_engine)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.ga = model.encoder.lower(device, precision=precision_main, **kw)
            self.ha = model.hyperencoder.lower(device, precision=precision_main, **kw)
            self.hs = model.hyperdecoder.lower(device, precision=precision_main, **kw)
            self.charm = model.context_model.lower(device, precision=precision_main, **kw)
            self.gs = model.decoder.lower(device, precision=precision_synthesis, **kw)
            self.eb_params, self.eb_medians = model.entropy_model_z.kernel_params(device)
            table = model.entropy_model_y.scale_table
            if table.numel() == 0:
                table = get_scale_table()
            self.gp = GaussianParams(table, float(model.entropy_model_y.scale_bound.item()), device)
        self.zc = model.entropy_model_z.channels
        self.stride = 64
        self._graphs = OrderedDict()   # CUDA graphs of the device-only launch sequences of small calls (LRU)
        self.decode_graph_sets = OrderedDict()   # model._DecodeGraphs per decompress chunk shape (LRU)
        self.decode_graph_seen = OrderedDict()   # eager calls per chunk shape that has no set yet

    # ------------------------------------------------------------------ device-side stages
    def padded(self, h, w):
        s = self.stride
        return -(-h // s) * s, -(-w // s) * s

    def analysis(self, images, q, compact=False, noise=None):
        """images: NCHW on the device, fp32 in [-1,1] or uint8 RGB (normalised on the fly like the reference's
        ToTensor + Normalize).  Returns a dict of device tensors (SURVEY 3.1 encode); ``compact`` adds the int16 symbol /
        uint8 index copies for the host coder.  ``noise`` = {"y": NCHW fp32, "z": NCHW fp32} (uniform in [-1/2, 1/2))
        selects the training-mode forward: the likelihoods of y + u / z + u are added as y_lik_noisy / z_lik_noisy."""
        L, st = nv.lib(), nv.stream_handle()
        n, _, h, w = images.shape
        if not images.is_contiguous():
            raise ValueError("analysis() reads the image through a raw pointer: pass a contiguous NCHW tensor")
        hp, wp = self.padded(h, w)
        dev = images.device
        img = Act.empty(n, hp // 2, wp // 2, self.ga.PATCH_CH, two=True, device=dev)   # im2col of g_a conv1 fused with the reflect pad
        if images.dtype == torch.uint8:
            nv.check(L.crdr_image_u8_to_patches(images.data_ptr(), n, h, w, hp, wp, img.planes(0), st))
        else:
            nv.check(L.crdr_image_to_patches(images.data_ptr(), n, h, w, hp, wp, img.planes(0), st))
        y_act, y32 = self.ga.run(img, q)
        z32 = self.ha.run(y_act)
        hz, wz = z32.shape[1:3]
        zhat = Act.empty(n, hz, wz, self.zc, two=True, device=dev)
        z_sym = torch.empty((n, self.zc, hz, wz), dtype=torch.int32, device=dev)
        z_hat = torch.empty((n, self.zc, hz, wz), dtype=torch.float32, device=dev)
        z_lik = torch.empty((n, self.zc, hz, wz), dtype=torch.float32, device=dev)
        d = nv.EbDesc()
        d.z, d.z_cs, d.n, d.hw, d.c = z32.data_ptr(), self.zc, n, hz * wz, self.zc
        d.params, d.medians = self.eb_params.data_ptr(), self.eb_medians.data_ptr()
        d.zhat_planes = zhat.planes(0)
        d.symbols, d.zhat_nchw, d.likelihood = z_sym.data_ptr(), z_hat.data_ptr(), z_lik.data_ptr()
        z_lik_noisy = None
        if noise is not None:
            nz, ny = noise["z"], noise["y"]
            if (tuple(nz.shape) != (n, self.zc, hz, wz) or tuple(ny.shape) != (n, self.charm.yc, y_act.h, y_act.w)
                    or nz.dtype != torch.float32 or ny.dtype != torch.float32 or not nz.is_contiguous() or not ny.is_contiguous()):
                raise ValueError("noise must be contiguous fp32 NCHW tensors shaped like z and y")
            z_lik_noisy = torch.empty((n, self.zc, hz, wz), dtype=torch.float32, device=dev)
            d.noise, d.likelihood_noisy = nz.data_ptr(), z_lik_noisy.data_ptr()
        nv.check(L.crdr_eb_quantize(C.byref(d), st))
        T = self.charm.new_support(n, y_act.h, y_act.w, dev)
        self.hs.run(zhat, T, self.charm.off_mean, self.charm.off_scale)
        enc = self.charm.encode(T, y32, self.gp, compact=compact and noise is None, noise=None if noise is None else noise["y"])
        yhat32, y_sym, y_idx, y_lik = enc[:4]
        out = dict(y32=y32, z32=z32, z_sym=z_sym, z_hat=z_hat, z_lik=z_lik, yhat32=yhat32, y_sym=y_sym, y_idx=y_idx,
                   y_lik=y_lik, size=(h, w))
        if noise is not None:
            out.update(y_lik_noisy=enc[4], z_lik_noisy=z_lik_noisy)
        elif compact:
            out.update(y_sym16=enc[4], y_idx8=enc[5])
        return out

    def hyper_from_symbols(self, z_sym):
        """z symbols (int32 NCHW, device) -> (support tensor with h_s output, z_hat NCHW)."""
        L, st = nv.lib(), nv.stream_handle()
        n, c, hz, wz = z_sym.shape
        dev = z_sym.device
        zhat = Act.empty(n, hz, wz, c, two=True, device=dev)
        z_hat = torch.empty((n, c, hz, wz), dtype=torch.float32, device=dev)
        d = nv.EbDesc()
        d.n, d.hw, d.c = n, hz * wz, c
        d.params, d.medians = self.eb_params.data_ptr(), self.eb_medians.data_ptr()
        d.zhat_planes = zhat.planes(0)
        d.symbols, d.zhat_nchw = z_sym.data_ptr(), z_hat.data_ptr()
        nv.check(L.crdr_eb_dequantize(C.byref(d), st))
        T = self.charm.new_support(n, hz * 4, wz * 4, dev)
        self.hs.run(zhat, T, self.charm.off_mean, self.charm.off_scale)
        return T, z_hat

    def synthesis(self, yhat32, q, beta, size, out_uint8=False, clamp=True):
        """y_hat fp32 NHWC -> cropped, clamped NCHW image: fp32 in [-1,1], or uint8 through the reference's PNG
        conversion (img_utils.py:30-42, truncation).  ``clamp=False``: the raw reconstruction of the training-mode forward."""
        img = self.gs.run(yhat32, q, beta)  # phase-packed: [n, hp/2, wp/2, 16]
        n, hb, wb, cs = img.shape
        h, w = size
        if out_uint8:
            out = torch.empty((n, 3, h, w), dtype=torch.uint8, device=img.device)
            nv.check(nv.lib().crdr_phases_to_image_u8(img.data_ptr(), cs, n, hb, wb, h, w, out.data_ptr(), nv.stream_handle()))
            return out
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=img.device)
        nv.check(nv.lib().crdr_phases_to_image_ex(img.data_ptr(), cs, n, hb, wb, h, w, out.data_ptr(), 1 if clamp else 0,
                                                  nv.stream_handle()))
        return out

    def decode_device(self, z_sym, y_sym, q, beta, size):
        """Decoder arithmetic with the symbols already on the device (the device-timed decode span)."""
        T, z_hat = self.hyper_from_symbols(z_sym)
        yhat32 = self.charm.decode(T, self.gp, lambda s0, cnt, idx: y_sym)
        return self.synthesis(yhat32, q, beta, size), yhat32, z_hat

    # ------------------------------------------------------------------ CUDA graphs for launch-bound (small) calls
    # One Kodak image is ~190 launches per direction of 10-30 us each: the Python enqueue (5 ms) costs more than the
    # kernels.  Small calls therefore replay a captured graph of the same launch sequence (side-stream forks included);
    # the arithmetic and its order are untouched, so results are bit-identical to the eager path.  Keys carry everything
    # a launch sequence depends on (shapes, dtype, quality index, beta: the per-q / per-beta vectors are baked in as
    # pointers to their cached device copies).  Outputs live in the graph's private pool and are overwritten by the next
    # replay of the same key: callers consume them on the same stream (or clone what they hand out).
    GRAPH_MAX_PIXELS = int(os.environ.get("CRDR_GRAPH_MAX_PIXELS", str(4 * 512 * 768)))
    GRAPH_CACHE = 12
    graphs_enabled = os.environ.get("CRDR_GRAPHS", "1") != "0"   # (instance attribute when toggled: per-launch event timing needs eager launches)

    def _graphed(self, key, statics, fn):
        """statics: input tensors to copy into the graph's static buffers; fn(*static_inputs) -> outputs."""
        hit = self._graphs.get(key)
        if hit is None:
            bufs = [torch.empty_like(t) for t in statics]
            for b, t in zip(bufs, statics):
                b.copy_(t)
            fn(*bufs)                                   # eager warm-up: fills vector / tensor-map caches, sets kernel attributes
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = nv.LAUNCH_COUNT[0]
            with torch.cuda.graph(g):
                outs = fn(*bufs)
            hit = self._graphs[key] = (g, bufs, outs, nv.LAUNCH_COUNT[0] - l0)
            while len(self._graphs) > self.GRAPH_CACHE:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        g, bufs, outs, launches = hit
        for b, t in zip(bufs, statics):
            b.copy_(t, non_blocking=True)
        g.replay()
        nv.LAUNCH_COUNT[0] += launches                  # kernels launched by the replay (bench.py reports the count)
        return outs

    def _small(self, n, h, w):
        hp, wp = self.padded(h, w)
        return self.graphs_enabled and n * hp * wp <= self.GRAPH_MAX_PIXELS

    def analysis_fast(self, images, q, compact=False, tag=""):
        """analysis() through a CUDA graph when the call is small enough to be launch bound (same results).  ``tag``
        separates callers whose outputs must coexist (the pipelined chunks of one compress_batch call)."""
        n, _, h, w = images.shape
        if not self._small(n, h, w):
            return self.analysis(images, q, compact=compact)
        key = ("analysis", tag, n, h, w, images.dtype, bool(compact))
        self.prepare(q=q)   # the gain vectors live at fixed addresses: the captured sequence is independent of q
        return self._graphed(key, [images.contiguous()], lambda x: self.analysis(x, q, compact=compact))

    def decode_device_fast(self, z_sym, y_sym, q, beta, size):
        n = z_sym.shape[0]
        if not self._small(n, *size):
            return self.decode_device(z_sym, y_sym, q, beta, size)
        key = ("decode", n, tuple(size))
        self.prepare(q=q, beta=beta)
        return self._graphed(key, [z_sym, y_sym], lambda zs, ys: self.decode_device(zs, ys, q, beta, size))

    def prepare(self, q=None, beta=None):
        """Fill the per-quality / per-beta vector caches on the current stream (before work fans out to side streams)."""
        if q is not None:
            self.ga.gain_pairs(q)
            self.gs.gain_pairs(q)
        if beta is not None:
            self.gs.cond_vectors(beta)

    # ------------------------------------------------------------------ small reductions / layout
    def bits(self, lik):
        n = lik.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=lik.device)
        nv.check(nv.lib().crdr_bits_from_likelihood(lik.data_ptr(), n, lik[0].numel(), out.data_ptr(), nv.stream_handle()))
        return out

    def max_abs(self, x):
        """Per-image max |x| in one launch."""
        n = x.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=x.device)
        nv.check(nv.lib().crdr_max_abs_batch(x.data_ptr(), n, x[0].numel(), out.data_ptr(), nv.stream_handle()))
        return out

    def to_nchw(self, x32):
        n, h, w, c = x32.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x32.device)
        nv.check(nv.lib().crdr_nhwc_to_nchw(x32.data_ptr(), c, 0, n, h * w, c, out.data_ptr(), nv.stream_handle()))
        return out


class _PinnedPool:
    """Reusable page-locked staging buffers for the device<->host boundary of the coder (symbols, CDF indexes).
    One growable flat byte buffer per tag (views are taken for each shape), so a dataset with many image sizes or a
    long-running service does not keep pinning new memory."""

    def __init__(self):
        self._bufs = {}

    def get(self, tag, shape, dtype):
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        buf = self._bufs.get(tag)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 64), dtype=torch.uint8, pin_memory=True)
            self._bufs[tag] = buf
        return buf[:nbytes].view(dtype).view(tuple(shape))

    def fetch(self, tag, dev_tensor):
        """Asynchronous device -> pinned host copy on the current stream (caller synchronises)."""
        buf = self.get(tag, dev_tensor.shape, dev_tensor.dtype)
        buf.copy_(dev_tensor, non_blocking=True)
        return buf


class _DecodeGraphs:
    """CUDA graphs of the device segments of ONE pipelined decompress chunk (n images of one size).

    The decoder alternates between the host coder and the device, group of slices after group: with eager
    launches the Python enqueue of a segment (~12 launches on forked side streams, 0.8 ms) and the host decode (0.8 ms) add
    up on one thread while the device waits (profiles/e2e_timeline_r02.txt: the scheduler never waited for the device).  Here
    every segment -- [host -> device copy of the decoded symbols, dequantise, LRP refinement, (mean, scale) networks of the
    next group, CDF indexes, device -> host copy of the indexes] -- is one graph replay; the first segment also holds h_s and
    the last one g_s.  The pinned staging buffers are owned by the set (fixed addresses inside the memcpy nodes), all segments
    share one memory pool and are replayed in capture order, and the per-quality / per-beta vectors stay behind their fixed
    device addresses (StaticVectors), so one set serves every q and beta.  Same launch sequence, same arithmetic: the
    output is bit-identical to the eager path."""

    def __init__(self, model, eng, n, h, w, out_uint8, q, beta):
        self.n = n
        dev = eng.device
        hp, wp = eng.padded(h, w)
        hz, wz = hp // model.model_stride, wp // model.model_stride
        sc, yC = eng.charm.sc, model.yC
        self.groups = eng.charm.groups()
        pin = lambda shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True)
        self.z_host = pin((n, model.zC, hz, wz), torch.int32)
        hy, wy = 4 * hz, 4 * wz
        self.ix_host = [pin((n, len(g) * sc, hy, wy), torch.uint8) for g in self.groups]
        self.sym_host = [pin((n, len(g) * sc, hy, wy), torch.int32) for g in self.groups]
        # the coder side: persistent decoder objects and marshalled calls on the pinned buffers
        zi = _channel_indexes(model.zC, hz, wz)
        self.z_dec = [rans.Decoder() for _ in range(n)]
        self.y_dec = [rans.Decoder() for _ in range(n)]
        zv = self.z_host.numpy().reshape(n, -1)
        self.z_plan = rans.DecodePlan(self.z_dec, [zi] * n, [zv[i] for i in range(n)])
        self.y_plans = []
        for ix, sy in zip(self.ix_host, self.sym_host):
            iv, sv = ix.numpy().reshape(n, -1), sy.numpy().reshape(n, -1)
            self.y_plans.append(rans.DecodePlan(self.y_dec, [iv[i] for i in range(n)], [sv[i] for i in range(n)]))
        # capture (nothing executes): the generator of the slice loop runs across the captures
        eng.prepare(q=q, beta=beta)
        torch.cuda.current_stream().synchronize()
        self.pool = torch.cuda.graph_pool_handle()
        self.segments = []
        st = {}

        def capture(fn):
            g = torch.cuda.CUDAGraph()
            l0 = nv.LAUNCH_COUNT[0]
            # thread_local: other host threads (PNG workers of compress.py, NCCL's watchdog) keep making CUDA calls
            with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                fn()
            self.segments.append((g, nv.LAUNCH_COUNT[0] - l0))

        def fetch_indexes(k, req):
            s0, cnt, idx = req
            assert s0 == self.groups[k][0] and cnt == len(self.groups[k])
            self.ix_host[k].copy_(idx[:, s0 * sc:(s0 + cnt) * sc].contiguous(), non_blocking=True)

        def first():
            st["z_sym"] = self.z_host.to(dev, non_blocking=True)
            st["T"], st["z_hat"] = eng.hyper_from_symbols(st["z_sym"])
            st["y_sym"] = torch.empty((n, yC, st["T"].h, st["T"].w), dtype=torch.int32, device=dev)
            st["steps"] = eng.charm.decode_steps(st["T"], eng.gp, compact=True)
            fetch_indexes(0, next(st["steps"]))

        def middle(k):   # symbols of group k-1 in, indexes of group k out
            g = self.groups[k - 1]
            st["y_sym"][:, g[0] * sc:(g[-1] + 1) * sc].copy_(self.sym_host[k - 1], non_blocking=True)
            fetch_indexes(k, st["steps"].send(st["y_sym"]))

        def last():
            g = self.groups[-1]
            st["y_sym"][:, g[0] * sc:(g[-1] + 1) * sc].copy_(self.sym_host[-1], non_blocking=True)
            try:
                st["steps"].send(st["y_sym"])
                raise RuntimeError("decode_steps yielded past the last slice group")
            except StopIteration as done:
                yhat32 = done.value
            st["img"] = eng.synthesis(yhat32, q, beta, (h, w), out_uint8=out_uint8)
            st["y_hat"] = eng.to_nchw(yhat32)

        capture(first)
        for k in range(1, len(self.groups)):
            capture(lambda k=k: middle(k))
        capture(last)
        self.img, self.z_hat, self.y_hat = st["img"], st["z_hat"], st["y_hat"]
        self._keep = st   # tensors of the shared pool that later segments read

    def replay(self, k):
        g, launches = self.segments[k]
        g.replay()
        nv.LAUNCH_COUNT[0] += launches


def _channel_indexes(c, h, w):
    return np.ascontiguousarray(np.broadcast_to(np.arange(c, dtype=np.int32)[:, None, None], (c, h, w))).reshape(-1)


class _CodecModelBase(nn.Module):
    """Shared plumbing: BaseModel (base_model.py:16-170) + HyperpriorModel bit accounting."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.device = opt.device
        self.convert_img_range = opt.get("convert_img_range_to_01", False)
        if self.convert_img_range:
            raise NotImplementedError("convert_img_range_to_01 is not used by the CRDR configs")
        self._build_subnets()
        self.stride = 64
        self._engine = None
        self._engine_opts = {}
        self._pinned = _PinnedPool()

    # -- engine lifecycle -------------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            dev = torch.device(self.device)
            if dev.type != "cuda":
                raise nv.NativeError(
                    f"device '{self.device}': the codec hot path exists only as sm_100a CUDA kernels; there is no CPU "
                    "implementation in this package (use the oracle under oracle/ for CPU checks)")
            nv.lib()  # fail loudly if the extension is missing
            self._engine = CodecEngine(self, dev, **self._engine_opts)
        return self._engine

    def set_engine_options(self, **kw):
        """precision_main / precision_synthesis / conv_engine (see crdr_b200.native)."""
        self._engine_opts = kw
        self._engine = None

    def invalidate_engine(self):
        self._engine = None

    # -- checkpoints ---------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        self.entropy_model_z._resize_buffers_from(state_dict, "entropy_model_z.", ["_quantized_cdf", "_offset", "_cdf_length"])
        self.entropy_model_y._resize_buffers_from(state_dict, "entropy_model_y.",
                                                  ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"])
        self.invalidate_engine()
        return super().load_state_dict(state_dict, strict=strict)

    def load_learned_weight(self, ckpt_path):
        get_root_logger().info(f"load checkpoint: {ckpt_path}")
        ckpt = torch.load(ckpt_path, map_location="cpu")
        incoming = OrderedDict((k[7:] if "module." in k else k, v) for k, v in ckpt["comp_model"].items())
        merged = self.state_dict()
        merged.update({k: v for k, v in incoming.items() if k in merged})
        self.load_state_dict(merged)
        self.entropy_model_z.update(force=False)

    def separate_aux_parameters(self):
        named = {n: p for n, p in self.named_parameters() if p.requires_grad}
        aux = {n: p for n, p in sorted(named.items()) if n.endswith(".quantiles")}
        main = {n: p for n, p in sorted(named.items()) if not n.endswith(".quantiles")}
        assert not (main.keys() & aux.keys()) and len(main) + len(aux) == len(named)
        return main, aux

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    # -- reference helpers ---------------------------------------------------------------------
    @staticmethod
    def likelihood_to_bit(likelihood, num_pixel):
        dims = tuple(range(1, likelihood.ndim))
        bits = -(torch.log(likelihood).sum(dim=dims)) / np.log(2)
        return bits, bits / num_pixel

    def codec_setup(self):
        self.header_handler = (MultiRateHeaderHandler if self.uses_rate else HeaderHandler)(use_non_zero_ind=False)
        self.entropy_model_z.update(force=True)
        self.entropy_model_y.update_scale_table(get_scale_table(), force=True)
        self.yC, self.zC = self.encoder.latent_ch, self.hyperencoder.latent_ch
        self.y_stride = 2 ** self.encoder.num_downscale
        self.model_stride = self.y_stride * 2 ** self.hyperencoder.num_downscale
        self.invalidate_engine()


class _CharmModelCore(_CodecModelBase):
    """The three ChARM models of the reference share everything but their conditioning inputs:
    HyperpriorCharmModel (stage 1: none), InterpCaHyperpriorCharmModel (stage 2: quality index q) and
    BetaCondInterpCaHyperpriorCharmModel (stage 3 / crdr.yaml: q and the realism weight beta).  The core takes
    (rate_ind, beta) everywhere; the registered classes below restore the reference's public signatures."""
    uses_rate = True
    uses_beta = True

    def _build_subnets(self):
        sn = self.opt.subnet
        if self.uses_rate:
            self.rate_level = sn.encoder.rate_level
            assert sn.encoder.rate_level == sn.decoder.rate_level
            if self.opt.get("batch_rate_ind_sample", False):
                raise NotImplementedError("batch_rate_ind_sample is not supported yet.")   # interpca_hyperprior_model.py:25-26
        if self.uses_beta:
            self.max_beta = float(sn.decoder.max_beta)
        self.encoder = build_subnet(sn.encoder, "encoder")
        self.decoder = build_subnet(sn.decoder, "decoder")
        self.hyperencoder = build_subnet(sn.hyperencoder, "hyperencoder")
        self.hyperdecoder = build_subnet(sn.hyperdecoder, "hyperdecoder")
        self.entropy_model_z = build_subnet(sn.entropy_model_z, "entropy_model")
        self.entropy_model_y = build_subnet(sn.entropy_model_y, "entropy_model")
        self.context_model = build_subnet(sn.context_model, "context_model")
        if not isinstance(self.entropy_model_y, GaussianMeanScaleConditional):
            raise NotImplementedError("entropy_model_y must be a mean-scale Gaussian conditional")

    # -- sampling helpers of the training API ------------------------------------------------
    def sample_rate_ind(self, num_sample=1):
        return torch.randint(self.rate_level, (num_sample,))

    def sample_beta(self):
        return self.max_beta * (float(np.random.randint(0, 101)) / 100.0)

    @staticmethod
    def _q(rate_ind):
        if rate_ind is None:
            return None
        if isinstance(rate_ind, torch.Tensor):
            assert rate_ind.numel() == 1, "one quality index per batch (batch_rate_ind_sample is unsupported upstream)"
            return float(rate_ind.reshape(-1)[0].item())
        return float(rate_ind)

    def _to_device(self, images):
        """fp32 [-1,1] (the reference's API) or uint8 RGB (normalised on the device) -> contiguous device tensor; the copy
        is asynchronous when the source is page-locked."""
        dt = torch.uint8 if images.dtype == torch.uint8 else torch.float32
        nb = images.device.type == "cpu" and images.is_pinned()
        return images.to(device=self.device, dtype=dt, non_blocking=nb).contiguous()

    # -- forward / run_model ---------------------------------------------------------------------
    def draw_noise(self, n, h, w, generator=None):
        """Uniform noise in [-1/2, 1/2) for the training-mode likelihoods, shaped like z and y of n padded h x w images
        (the reference draws it inside CompressAI's quantize(mode="noise"); here it is an explicit input so a CPU oracle
        can replay the exact values)."""
        dev = self.engine().device
        hy, wy, hz, wz = h // 16, w // 16, h // 64, w // 64
        mk = lambda c, a, b: torch.rand((n, c, a, b), dtype=torch.float32, device=dev, generator=generator) - 0.5
        return {"z": mk(self.hyperencoder.latent_ch, hz, wz), "y": mk(self.encoder.latent_ch, hy, wy)}

    @torch.no_grad()
    def forward(self, real_images, rate_ind, beta, is_train=True, noise=None):
        """Forward VALUES of the reference's forward (beta_cond_interpca_hyperprior_charm_model.py:34-78).  Training mode
        (is_train=True): likelihoods of the noise-perturbed latents (``noise`` = draw_noise(...) or the caller's tensors),
        straight-through rounded codes, quantised q_likelihoods, unclamped reconstruction.  No autograd graph is built
        here: the optimisation step (this forward on taped engines + the backward kernels + Adam) is train.CodecTrainer /
        trainers.RateDistortionTrainer, which keep their own device copy of the parameters."""
        eng = self.engine()
        q = self._q(rate_ind)
        x = self._to_device(real_images)
        n, _, h, w = x.shape
        if h % self.stride or w % self.stride:
            raise ValueError("forward() expects images padded to a multiple of 64 (use run_model)")
        if is_train and noise is None:
            noise = self.draw_noise(n, h, w)
        a = eng.analysis(x, q, noise=noise if is_train else None)
        fake = eng.synthesis(a["yhat32"], q, beta, (h, w), clamp=not is_train)
        y, z = eng.to_nchw(a["y32"]), eng.to_nchw(a["z32"])
        y_hat = eng.to_nchw(a["yhat32"])
        nv.status_check()  # a clamped fp16 overflow must surface as an error, never as a silent wrong value
        lik = {"y": a["y_lik_noisy"], "z": a["z_lik_noisy"]} if is_train else {"y": a["y_lik"], "z": a["z_lik"]}
        return {
            "fake_images": fake,
            "likelihoods": lik,
            "latent_code": {"y": y, "z": z},
            "quantized_code": {"y": y_hat, "z": a["z_hat"]},
            "q_likelihoods": {"y": a["y_lik"], "z": a["z_lik"]},
        }

    @torch.no_grad()
    def run_model(self, real_images, rate_ind=None, beta=None, is_train=True, noise=None):
        if is_train:
            # beta_cond_interpca_hyperprior_model.py:28-64: one quality index and one beta per batch when not given;
            # training crops are multiples of 64, so data_preprocess / data_postprocess are identities (base_model.py:35-57)
            if rate_ind is None and self.uses_rate:
                rate_ind = self.sample_rate_ind(1)
            if beta is None and self.uses_beta:
                beta = self.sample_beta()
            n, _, h, w = real_images.shape
            out = _CharmModelCore.forward(self, real_images, rate_ind, beta, is_train=True, noise=noise)
            num_pixel = h * w
            eng = self.engine()
            bpp = (eng.bits(out["likelihoods"]["y"]) + eng.bits(out["likelihoods"]["z"])) / num_pixel
            qbpp = (eng.bits(out["q_likelihoods"]["y"]) + eng.bits(out["q_likelihoods"]["z"])) / num_pixel
            return dict(real_images=self._to_device(real_images), fake_images=out["fake_images"],
                        y_hat=out["quantized_code"]["y"], z_hat=out["quantized_code"]["z"], rate_ind=rate_ind, beta=beta,
                        y_likelihood=out["likelihoods"]["y"], z_likelihood=out["likelihoods"]["z"], bpp=bpp,
                        y_q_likelihood=out["q_likelihoods"]["y"], z_q_likelihood=out["q_likelihoods"]["z"], qbpp=qbpp)
        if rate_ind is None:
            raise ValueError('"rate_ind" must be specified if is_train=False')
        if beta is None:
            raise ValueError('"beta" must be specified if is_train=False')
        eng = self.engine()
        q = self._q(rate_ind)
        x = self._to_device(real_images)
        n, _, h, w = x.shape
        a = eng.analysis(x, q)
        fake = eng.synthesis(a["yhat32"], q, beta, (h, w))
        num_pixel = h * w
        y_bpp, z_bpp = eng.bits(a["y_lik"]) / num_pixel, eng.bits(a["z_lik"]) / num_pixel
        nv.status_check()
        return dict(real_images=x.clamp(-1, 1), fake_images=fake, y_hat=eng.to_nchw(a["yhat32"]), z_hat=a["z_hat"],
                    rate_ind=rate_ind, beta=beta, y_likelihood=a["y_lik"], z_likelihood=a["z_lik"],
                    bpp=y_bpp + z_bpp, y_q_likelihood=a["y_lik"], z_q_likelihood=a["z_lik"], qbpp=y_bpp + z_bpp)

    # -- codec -------------------------------------------------------------------------------------
    # Host entropy coding and device arithmetic of one call overlap through a software pipeline on ONE stream and ONE host
    # thread: the batch is cut into `pipeline_chunks` chunks whose device segments are enqueued in order, and whenever a
    # chunk needs the host coder (after an event on its device -> host copy) the device already holds the other
    # chunks' next segments.  (Two host threads on two streams were tried first: the chunks' full-GPU persistent kernels
    # interleave one for one, both chunks reach the host coder at the same moment and nothing overlaps.)
    # Results are per image and independent of the chunking (deterministic kernels, one rANS stream per image).
    pipeline_chunks = int(os.environ.get("CRDR_PIPELINE_CHUNKS", "2"))
    pipeline_min_images = 8
    pipeline_weights = (tuple(float(v) for v in os.environ["CRDR_PIPELINE_WEIGHTS"].split(","))
                        if os.environ.get("CRDR_PIPELINE_WEIGHTS") else (2.0, 1.0))
    # relative chunk sizes of decompress_batch: a smaller LAST chunk shortens the pipeline drain (the synthesis transform of
    # the last chunk runs after the last host decode); measured 88.4 vs 89.2 and 91.2 vs 95.9 ms per 24-image step on two boxes

    pipeline_chunks_compress = int(os.environ.get("CRDR_PIPELINE_CHUNKS_COMPRESS", "0"))   # 0: pipeline_chunks

    pipeline_weights_compress = (tuple(float(v) for v in os.environ["CRDR_PIPELINE_WEIGHTS_COMPRESS"].split(","))
                                 if os.environ.get("CRDR_PIPELINE_WEIGHTS_COMPRESS") else (3.0, 1.0))
    # compress_batch: what stays exposed is the host encode of the LAST chunk, and one image's stream is one sequential coder
    # pass whatever the chunk size -- a short last chunk only shortens the device span that precedes it.  Measured
    # (tools/e2e_pipeline_sweep.py, 24 x 512x768, ms per call): one chunk 48.6, 1:1 46.9, 2:1 46.0, 3:1 44.4, 1:1:1 47.6

    # Few coder threads (one process per GPU on a slice of the host: 4 cores per rank at N = 8): the host coder, not the
    # device, paces the pipeline, so the chunks are cut for the coder -- several rounds of streams per chunk hide behind the
    # next chunk's device span.  Measured on one GPU confined to 4 cores (`CRDR_CODER_THREADS=4 taskset -c 0-3 python
    # tools/e2e_pipeline_sweep.py`, ms per 24-image call): compress 3:1 50.1, 2:1 45.5, 1:1 48.7, 3:2:1 45.5, 1:1:1 48.3;
    # decompress 2:1 42.7, 1:1 38.2, 1:1:1 36.2, 1:1:1:1 41.1.  Applies when neither the environment nor the instance
    # sets a pipeline shape.
    pipeline_few_threads = int(os.environ.get("CRDR_PIPELINE_FEW_THREADS", "12"))
    pipeline_weights_few = (1.0, 1.0, 1.0)
    pipeline_weights_compress_few = (2.0, 1.0)

    def _pipeline_shape(self, compress):
        """(chunk count, weights) of compress_batch / decompress_batch."""
        names = ("pipeline_chunks", "pipeline_weights", "pipeline_chunks_compress", "pipeline_weights_compress")
        envs = ("CRDR_PIPELINE_CHUNKS", "CRDR_PIPELINE_WEIGHTS", "CRDR_PIPELINE_CHUNKS_COMPRESS", "CRDR_PIPELINE_WEIGHTS_COMPRESS")
        explicit = any(k in self.__dict__ for k in names) or any(os.environ.get(e) for e in envs)
        if not explicit and rans.pool_info()[0] < self.pipeline_few_threads:
            w = self.pipeline_weights_compress_few if compress else self.pipeline_weights_few
            return len(w), w
        if compress:
            return (self.pipeline_chunks_compress or self.pipeline_chunks), self.pipeline_weights_compress
        return self.pipeline_chunks, self.pipeline_weights

    def _chunks(self, n, chunks=None, weights=None):
        k = (chunks or self.pipeline_chunks) if n >= self.pipeline_min_images else 1
        k = max(1, min(k, n))
        weights = weights or self.pipeline_weights
        w = list(weights) if weights and len(weights) == k else [1] * k
        tot, acc, edges = float(sum(w)), 0.0, [0]
        for wi in w:
            acc += wi
            edges.append(int(round(acc * n / tot)))
        return [(edges[i], edges[i + 1]) for i in range(k) if edges[i + 1] > edges[i]]

    @staticmethod
    def _drive(gens):
        """Round-robin scheduler of chunk generators.  A generator yields a recorded CUDA event when it needs the host
        (its device -> host copy is complete once the event is); it is resumed after the event, runs its host stage,
        enqueues its next device segment and yields again.  Returns the generators' return values in order."""
        from collections import deque
        results = [None] * len(gens)
        queue = deque()
        for i, g in enumerate(gens):          # prime: every chunk's first device segment is enqueued, in order
            try:
                queue.append((i, g, next(g)))
            except StopIteration as done:
                results[i] = done.value
        while queue:
            i, g, ev = queue.popleft()
            ev.synchronize()
            try:
                queue.append((i, g, g.send(None)))
            except StopIteration as done:
                results[i] = done.value
        return results

    @staticmethod
    def _event():
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def _compress_gen(self, tag, real_images, rate_ind, return_tensors, coder_threads):
        eng = self.engine()
        q = self._q(rate_ind)
        x = self._to_device(real_images)
        n, _, h, w = x.shape
        a = eng.analysis_fast(x, q, compact=True, tag=tag)
        y_bits, z_bits = eng.bits(a["y_lik"]), eng.bits(a["z_lik"])
        y_max = eng.max_abs(a["yhat32"])
        # ---- device -> host boundary (the reference moves y, z here; we move int16 symbols and uint8 table indexes)
        pp = self._pinned
        z_sym, y_sym, y_idx = (pp.fetch(tag + "z_sym", a["z_sym"]), pp.fetch(tag + "y_sym16", a["y_sym16"]),
                               pp.fetch(tag + "y_idx8", a["y_idx8"]))
        y_bits, z_bits, y_max = (pp.fetch(tag + "y_bits", y_bits), pp.fetch(tag + "z_bits", z_bits),
                                 pp.fetch(tag + "y_max", y_max))
        flags = pp.get(tag + "flags", (1,), torch.int32)
        nv.check(nv.lib().crdr_status_peek_async(flags.data_ptr(), nv.stream_handle()), counts=False)
        if return_tensors:
            y_hat = eng.to_nchw(a["yhat32"])
        yield self._event()  # the pinned buffers are valid once this event is
        z_sym, y_sym, y_idx = z_sym.numpy(), y_sym.numpy(), y_idx.numpy()
        if int(flags[0]) & nv.FLAG_SYM_RANGE:
            # a symbol outside int16 (never seen with sane weights): take the int32 tensors instead, clear the condition
            y_sym = a["y_sym"].cpu().numpy()
            y_idx = a["y_idx"].cpu().numpy()
            nv.check(nv.lib().crdr_status_clear_bits(nv.FLAG_SYM_RANGE, nv.stream_handle()), counts=False)
        y_bits, z_bits, y_max = y_bits.numpy().copy(), z_bits.numpy().copy(), y_max.numpy().copy()
        zc, hz, wz = z_sym.shape[1:]
        zi = _channel_indexes(zc, hz, wz)
        z_strs = rans.encode_batch([z_sym[i] for i in range(n)], [zi] * n, self.entropy_model_z.coder_tables(), coder_threads)
        y_strs = rans.encode_batch([y_sym[i] for i in range(n)], [y_idx[i] for i in range(n)],
                                   self.entropy_model_y.coder_tables(), coder_threads)
        out = []
        for i in range(n):
            header = self.header_handler.encode((h, w), rate_ind=q, max_abs=float(y_max[i]))   # q is None for single-rate models
            r = {"string_list": [header, z_strs[i], y_strs[i]],
                 "pred_y_bit": float(y_bits[i]), "pred_y_bpp": float(y_bits[i]) / (h * w),
                 "pred_z_bit": float(z_bits[i]), "pred_z_bpp": float(z_bits[i]) / (h * w)}
            if return_tensors:   # clones: the tensors may live in a CUDA graph's pool that the next call overwrites
                r.update(z_hat=a["z_hat"][i:i + 1].clone(), y_hat=y_hat[i:i + 1], z_likelihood=a["z_lik"][i:i + 1].clone(),
                         y_likelihood=a["y_lik"][i:i + 1].clone())
            out.append(r)
        return out

    @torch.no_grad()
    def compress_batch(self, real_images, rate_ind, return_tensors=False, coder_threads=0):
        """N images of one shape -> list of N result dicts (same keys as ``compress``).  ``real_images``: fp32 NCHW in
        [-1,1] like the reference, or uint8 NCHW RGB (normalised on the device exactly like ToTensor + Normalize)."""
        if not hasattr(self, "header_handler"):
            raise RuntimeError("call codec_setup() before compress()")
        n = real_images.shape[0]
        with torch.cuda.device(self.engine().device):
            nv.status_reset()
            res = self._drive([self._compress_gen(f"c{k}_", real_images[lo:hi], rate_ind, return_tensors, coder_threads)
                               for k, (lo, hi) in enumerate(self._chunks(n, *self._pipeline_shape(compress=True)))])
            nv.status_check()
        return [r for chunk in res for r in chunk]

    @torch.no_grad()
    def compress(self, real_images, rate_ind):
        n = real_images.shape[0]
        assert n == 1, f"In compress mode, batchsize must be 1, but {n}"
        return self.compress_batch(real_images, rate_ind, return_tensors=True)[0]

    def _decompress_gen(self, tag, string_lists, h, w, q, beta, coder_threads, out_uint8=False):
        eng = self.engine()
        n = len(string_lists)
        hp, wp = eng.padded(h, w)
        hz, wz = hp // self.model_stride, wp // self.model_stride
        dev = eng.device
        zi = _channel_indexes(self.zC, hz, wz)
        zt, yt = self.entropy_model_z.coder_tables(), self.entropy_model_y.coder_tables()
        z_dec = [rans.Decoder(sl[1]) for sl in string_lists]
        z_host = self._pinned.get(tag + "z_dec", (n, self.zC, hz, wz), torch.int32)
        zv = z_host.numpy().reshape(n, -1)
        rans.decode_batch(z_dec, [zi] * n, zt, coder_threads, outs=[zv[i] for i in range(n)])
        z_sym = z_host.to(dev, non_blocking=True)
        T, z_hat = eng.hyper_from_symbols(z_sym)
        y_dec = [rans.Decoder(sl[2]) for sl in string_lists]
        hy, wy = T.h, T.w
        y_sym = torch.empty((n, self.yC, hy, wy), dtype=torch.int32, device=dev)
        sc = eng.charm.sc
        steps = eng.charm.decode_steps(T, eng.gp, compact=True)
        try:
            req = next(steps)
            while True:
                s0, cnt, idx = req
                c0, c1 = s0 * sc, (s0 + cnt) * sc
                # device -> host: uint8 table indexes of this group (pinned); the event also covers the previous group's
                # host -> device symbol copy, so the pinned symbol buffer below may be reused
                ix_host = self._pinned.fetch(f"{tag}y_idx_{cnt}", idx[:, c0:c1].contiguous())
                yield self._event()
                ix = ix_host.numpy()
                sym_host = self._pinned.get(f"{tag}y_sym_{cnt}", (n, c1 - c0, hy, wy), torch.int32)
                sv = sym_host.numpy().reshape(n, -1)
                rans.decode_batch(y_dec, [ix[i] for i in range(n)], yt, coder_threads, outs=[sv[i] for i in range(n)])
                y_sym[:, c0:c1].copy_(sym_host, non_blocking=True)  # host -> device: decoded symbols
                req = steps.send(y_sym)
        except StopIteration as done:
            yhat32 = done.value
        img = eng.synthesis(yhat32, q, beta, (h, w), out_uint8=out_uint8)
        y_hat = eng.to_nchw(yhat32)
        return img, z_hat, y_hat

    # CUDA graphs of the decode segments (class _DecodeGraphs): a chunk shape is coded eagerly the first
    # DECODE_GRAPH_AFTER times it is seen (those calls also load every kernel the sequence uses) and captured at the end of
    # the last of them; later calls replay.  A data set whose image sizes never repeat therefore never pays for a capture.
    DECODE_GRAPH_MAX_PIXELS = int(os.environ.get("CRDR_DECODE_GRAPH_MAX_PIXELS", str(32 * 512 * 768)))
    DECODE_GRAPH_AFTER = int(os.environ.get("CRDR_DECODE_GRAPH_AFTER", "2"))
    DECODE_GRAPH_CACHE = 6
    decode_graphs_enabled = os.environ.get("CRDR_DECODE_GRAPHS", "1") != "0"

    def _decode_graph_key(self, tag, n, h, w, out_uint8):
        eng = self.engine()
        hp, wp = eng.padded(h, w)
        if not (self.decode_graphs_enabled and eng.graphs_enabled and n * hp * wp <= self.DECODE_GRAPH_MAX_PIXELS):
            return None
        return (tag, n, h, w, bool(out_uint8))

    def _decompress_gen_graphed(self, gs, string_lists, q, beta, coder_threads):
        """_decompress_gen with every device segment replayed from the chunk's graph set (same results)."""
        eng = self.engine()
        zt, yt = self.entropy_model_z.coder_tables(), self.entropy_model_y.coder_tables()
        eng.prepare(q=q, beta=beta)
        for d, sl in zip(gs.z_dec, string_lists):
            d.set_stream(sl[1])
        gs.z_plan.run(zt, coder_threads)
        gs.replay(0)
        for d, sl in zip(gs.y_dec, string_lists):
            d.set_stream(sl[2])
        for k, plan in enumerate(gs.y_plans):
            yield self._event()      # the indexes of group k are in pinned memory once this event is
            plan.run(yt, coder_threads)
            gs.replay(k + 1)
        return gs.img, gs.z_hat, gs.y_hat

    @torch.no_grad()
    def decompress_batch(self, string_lists, beta=0.0, coder_threads=0, out_uint8=False):
        """Streams of N images with identical size and quality -> (images [N,3,H,W], z_hat, y_hat).  ``out_uint8``: the
        images come back as uint8 RGB through the reference's PNG conversion (truncation) instead of fp32 [-1,1]."""
        if not hasattr(self, "header_handler"):
            raise RuntimeError("call codec_setup() before decompress()")
        for sl in string_lists:
            assert len(sl) == 3, f"String list length should be 3 (header, z, and y), but got {len(sl)}"
        heads = [self.header_handler.decode(sl[0]) for sl in string_lists]
        if any(hd["img_size"] != heads[0]["img_size"] or hd["rate_ind"] != heads[0]["rate_ind"] for hd in heads):
            raise ValueError("decompress_batch needs streams of one image size and one quality index")
        h, w = heads[0]["img_size"]
        q = heads[0]["rate_ind"]
        n = len(string_lists)
        with torch.cuda.device(self.engine().device):
            nv.status_reset()
            sets = self.engine().decode_graph_sets   # owned by the engine: they hold its weight pointers
            gens, fresh = [], []
            for k, (lo, hi) in enumerate(self._chunks(n, *self._pipeline_shape(compress=False))):
                key = self._decode_graph_key(f"d{k}_", hi - lo, h, w, out_uint8)
                gs = sets.get(key) if key is not None else None
                if gs is not None:
                    sets.move_to_end(key)
                    gens.append(self._decompress_gen_graphed(gs, string_lists[lo:hi], q, beta, coder_threads))
                else:
                    if key is not None:
                        seen = self.engine().decode_graph_seen
                        seen[key] = seen.pop(key, 0) + 1
                        while len(seen) > 64:
                            seen.popitem(last=False)
                        if seen[key] >= self.DECODE_GRAPH_AFTER:
                            fresh.append((key, hi - lo))
                    gens.append(self._decompress_gen(f"d{k}_", string_lists[lo:hi], h, w, q, beta, coder_threads, out_uint8))
            res = self._drive(gens)
            nv.status_check()
            # graph-owned outputs are overwritten by the next call on the same chunk shape: hand out copies
            out = tuple(torch.cat([r[j] for r in res], dim=0) if len(res) > 1 else res[0][j].clone() for j in range(3))
            for key, cnt in fresh:
                self.engine().decode_graph_seen.pop(key, None)
                sets[key] = _DecodeGraphs(self, self.engine(), cnt, h, w, out_uint8, q, beta)
                while len(sets) > self.DECODE_GRAPH_CACHE:
                    sets.popitem(last=False)
            return out

    @torch.no_grad()
    def decompress(self, string_list, beta=0.0):
        return self.decompress_batch([string_list], beta=beta)

    # -- validation (beta_cond_interpca_hyperprior_model.py:137-208) -----------------------------------
    @torch.no_grad()
    def validation(self, dataloader, max_sample_size, beta=None, save_img=False, save_dir="", use_tqdm=False):
        import pandas as pd
        from .img_utils import calc_psnr
        beta = self.max_beta / 2.0 if beta is None else beta
        rows = []
        for idx, data in enumerate(dataloader):
            row = {"idx": idx + 1}
            for q in range(self.rate_level):
                out = _CharmModelCore.run_model(self, **data, rate_ind=float(q), beta=beta, is_train=False)
                row[f"bpp_{q + 1}"] = out["bpp"].mean().item()
                row[f"psnr_{q + 1}"] = calc_psnr(out["real_images"], out["fake_images"], 255)
            rows.append(row)
            if idx + 1 >= min(len(dataloader), max_sample_size):
                break
        return pd.json_normalize(rows)


@MODEL_REGISTRY.register()
class BetaCondInterpCaHyperpriorCharmModel(_CharmModelCore):
    """config/crdr.yaml, config/crdr_stage_3.yaml (beta_cond_interpca_hyperprior_charm_model.py:13-149): the core's
    signatures are this model's."""


@MODEL_REGISTRY.register()
class InterpCaHyperpriorCharmModel(_CharmModelCore):
    """config/crdr_stage_2.yaml: variable rate, no realism conditioning (interpca_hyperprior_charm_model.py:22-146,
    interpca_hyperprior_model.py:31-62)."""
    uses_beta = False

    def forward(self, real_images, rate_ind, is_train=True, noise=None):
        return super().forward(real_images, rate_ind, None, is_train=is_train, noise=noise)

    def run_model(self, real_images, rate_ind=None, is_train=True, noise=None):
        out = super().run_model(real_images, rate_ind=rate_ind, beta=0.0, is_train=is_train, noise=noise)
        out.pop("beta", None)
        return out

    def decompress_batch(self, string_lists, coder_threads=0, out_uint8=False):
        return super().decompress_batch(string_lists, beta=None, coder_threads=coder_threads, out_uint8=out_uint8)

    def decompress(self, string_list):
        return self.decompress_batch([string_list])

    def validation(self, dataloader, max_sample_size, save_img=False, save_dir="", use_tqdm=False):
        return super().validation(dataloader, max_sample_size, beta=0.0)


@MODEL_REGISTRY.register()
class HyperpriorCharmModel(_CharmModelCore):
    """config/crdr_stage_1.yaml: single rate (hyperprior_charm_model.py:21-147, hyperprior_model.py:38-58); the
    5-byte header carries no quality index."""
    uses_rate = False
    uses_beta = False

    def forward(self, real_images, is_train=True, noise=None):
        return super().forward(real_images, None, None, is_train=is_train, noise=noise)

    def run_model(self, real_images, is_train=True, noise=None):
        out = super().run_model(real_images, rate_ind=0.0, beta=0.0, is_train=is_train, noise=noise)
        out.pop("beta", None)
        out.pop("rate_ind", None)
        return out

    def compress_batch(self, real_images, return_tensors=False, coder_threads=0):
        return super().compress_batch(real_images, None, return_tensors=return_tensors, coder_threads=coder_threads)

    def compress(self, real_images):
        n = real_images.shape[0]
        assert n == 1, f"In compress mode, batchsize must be 1, but {n}"
        return self.compress_batch(real_images, return_tensors=True)[0]

    def decompress_batch(self, string_lists, coder_threads=0, out_uint8=False):
        return super().decompress_batch(string_lists, beta=None, coder_threads=coder_threads, out_uint8=out_uint8)

    def decompress(self, string_list):
        return self.decompress_batch([string_list])

    def validation(self, dataloader, max_sample_size, save_img=False, save_dir="", use_tqdm=False):
        import pandas as pd
        from .img_utils import calc_psnr
        rows = []
        for idx, data in enumerate(dataloader):
            out = self.run_model(**data, is_train=False)
            rows.append({"idx": idx + 1, "bpp": out["bpp"].mean().item(), "psnr": calc_psnr(out["real_images"], out["fake_images"], 255)})
            if idx + 1 >= min(len(dataloader), max_sample_size):
                break
        return pd.json_normalize(rows)

"""Training step of the ChARM codec models on the CUDA engines: forward (training mode) + backward + optimiser.

What it replaces in the reference: ``RateDistortionTrainer.optimize_parameters`` (rate_distortion_trainer.py:57-101):
``run_model(is_train=True)`` -> rate + distortion losses -> ``l_total.backward()`` -> ``clip_grad_norm_`` ->
``g_optimizer.step()`` -> aux optimiser.  The reference delegates the backward to torch.autograd; here it is an explicit
reverse sweep over the tape of launches the forward recorded:

  conv record   -> crdr_epilogue_backward (ReLU mask / gain / skip / LRP tanh, per-channel sums), crdr_conv_wgrad (tcgen05),
                   crdr_conv_dgrad (the forward kernel over device-packed adjoint matrices, accumulating in place)
  gate record   -> crdr_gate_backward                 (ChengNLAM, cheng_nlam.py:23-29)
  gauss record  -> crdr_gauss_backward                (rate term + straight-through path, ste_gaussian_conditional.py:20-27)
  loss roots    -> crdr_mse_backward (distortion_loss.py:40-46) and the rate coefficient (rate_loss.py:84-110)

Activation gradients are fp16 planes under a power-of-two loss scale; parameter gradients, Adam moments and the master
weights are fp32 in one flat buffer each (one fused Adam launch, one NCCL all-reduce per bucket for data parallelism).
The factorised prior over z (18 k elements per Kodak image, 58 parameters per channel) is differentiated with torch
autograd on the device: it is 0.003 % of the step's arithmetic and not worth a kernel.

Scope: every ChARM model of the reference (stage 1, stage 2 = BASELINE config 5, and the beta-conditioned crdr.yaml model)
with the rate + MSE losses (+ the weight-free stand-in of the LPIPS term: pretrained AlexNet weights are not available
offline); GanCodecTrainer below adds the stage-3 discriminators and the beta * (perceptual + adversarial) terms of
multirate_hr_rgan_beta_cond_rate_distortion_trainer.py:62-66.
"""
import ctypes as C
import math
import os

import torch
import torch.nn.functional as F

from . import backward as bw
from . import native as nv
from .engine import Act

AUX_SUFFIX = "quantiles"


class TrainContext:
    """Owns the flat fp32 parameter / gradient buffers and everything the engines need in training mode."""

    def __init__(self, model, device):
        self.device = torch.device(device)
        self.tape = None
        self.convs = []
        named = list(model.named_parameters())
        main = [(k, p) for k, p in named if not k.endswith(AUX_SUFFIX)]
        aux = [(k, p) for k, p in named if k.endswith(AUX_SUFFIX)]
        self.names = [k for k, _ in main + aux]
        sizes = [p.numel() for _, p in main + aux]
        self.n_main = sum(p.numel() for _, p in main)
        total = sum(sizes)
        # every view starts on a 16-byte boundary (vector loads of biases / gains)
        self.offsets, o = [], 0
        for s in sizes:
            self.offsets.append(o)
            o += (s + 3) // 4 * 4
        self.total = o
        self.n_main_padded = self.offsets[len(main)] if aux else o
        self.flat_p = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.params, self.grads = {}, {}
        for (k, p), off in zip(main + aux, self.offsets):
            v = self.flat_p[off:off + p.numel()].view(p.shape)
            v.copy_(p.detach().to(self.device, torch.float32))
            self.params[k] = v
            self.grads[k] = self.flat_g[off:off + p.numel()].view(p.shape)
        self.buffers = {k: b.detach().to(self.device) for k, b in model.named_buffers()}
        self._grad_by_ptr = {self.params[k].data_ptr(): self.grads[k] for k in self.names}
        self._pack_table = None
        self.param_grads = True
        self.live = {}          # sub-network name -> live state dict (views of flat_p)
        self.gain_state = {}    # "encoder" / "decoder" -> (engine, l, r, alpha, lerped weights)
        self.gain_grad = {}
        self._gain_by_ptr = {}
        self.cond_grad, self.cond_state = None, None

    def live_sd(self, prefix):
        p = prefix + "."
        sd = {k[len(p):]: v for k, v in self.params.items() if k.startswith(p)}
        sd.update({k[len(p):]: v for k, v in self.buffers.items() if k.startswith(p)})
        self.live[prefix] = sd
        return sd

    def grad_view(self, master):
        return self._grad_by_ptr[master.data_ptr()]

    # ---- engines call these ------------------------------------------------------------------------------------
    def adopt(self, conv, w, b, transform, bias_transform, two_planes, **kw):
        conv.master, conv.bias_master, conv.bias_transform, conv.packed_transform = w, b, bias_transform, transform
        conv.bias_buf = bias_transform(b) if (bias_transform and b is not None) else None
        bias = conv.bias_buf if conv.bias_buf is not None else b
        conv.packed = bw.PackedConv(w, transform or (lambda t: t), two_planes, bias=bias, **kw)
        conv.op = conv.packed.op
        conv.wmap = None
        if transform is not None:
            conv.wmap = (transform(bw.index_weight(tuple(w.shape))).round().to(torch.int64) - 1).reshape(-1)
            conv.wmap_src = torch.nonzero(conv.wmap >= 0).reshape(-1).to(self.device)      # positions in the ConvOp's layout
            conv.wmap_dst = conv.wmap[conv.wmap >= 0].to(self.device)                      # ... and in the parameter
            conv.wmap = conv.wmap.to(self.device)
        if conv.bias_buf is not None:
            bidx = bias_transform(torch.arange(b.numel(), dtype=torch.float32) + 1.0).round().to(torch.int64).reshape(-1) - 1
            conv.bmap_src = torch.nonzero(bidx >= 0).reshape(-1).to(self.device)
            conv.bmap_dst = bidx[bidx >= 0].to(self.device)
        conv.dgrad = None
        self.convs.append(conv)

    def repack(self):
        """After an optimiser step: re-pack every forward / adjoint matrix (one launch over a device-resident job table,
        rebuilt whenever a backward pass has created new adjoint matrices) and refresh the derived bias vectors."""
        packed = []
        for conv in self.convs:
            packed.append(conv.packed)
            if conv.dgrad is not None:
                packed += conv.dgrad.packed()
        if self._pack_table is None or self._pack_table.count != sum(len(pc.op.phases) for pc in packed):
            self._pack_table = bw.PackTable(packed, self.device)
        self._pack_table.run()
        for conv in self.convs:
            if conv.bias_buf is not None:
                conv.bias_buf.copy_(conv.bias_transform(conv.bias_master))

    def live_gains(self, engine, which, q):
        """InterpChAtt vectors (interp_channel_attention.py:54-73) from the live parameters, written behind the fixed
        addresses the launches read."""
        sd = self.live[which]
        levels = engine.gains[0].levels
        q = float(q)
        if not (0.0 <= q <= levels - 1):
            raise AssertionError(f"rate_ind = {q} should be in [0, {levels - 1}]")
        l = int(math.floor(q))
        r = min(l + 1, levels - 1)
        alpha = float(r) - q
        vecs, lerped = [], []
        for i in range(len(engine.gains)):
            W, B = sd[f"interp_ca_list.{i}.weight"], sd[f"interp_ca_list.{i}.bias"]
            w = (W[l] * alpha + W[r] * (1 - alpha)).reshape(-1)
            b = (B[l] * alpha + B[r] * (1 - alpha)).reshape(-1)
            lerped.append(w)
            vecs += [F.softplus(w), b]
        flat = engine.gain_vecs.flat
        flat.copy_(torch.cat(vecs))
        engine.gain_vecs._live = None
        views = engine.gain_vecs.views
        if which not in self.gain_grad:
            self.gain_grad[which] = torch.zeros_like(flat)
            o = 0
            for i, n in enumerate(engine.gain_vecs.sizes):
                self._gain_by_ptr[views[i].data_ptr()] = (which, o, n)
                o += n
        self.gain_state[which] = (engine, l, r, alpha, lerped)
        return views

    def live_cond(self, engine, beta):
        """The 27 conditioning bias vectors (fourier_cond.py:21-37 embedding -> MLP -> proj_1..3 of every BetaCondBaseBlock,
        elic_interpca_beta_cond_autoencoder.py:52-65,142-152) from the live parameters, under autograd: the forward copies
        their values behind the fixed addresses the launches read, the backward feeds the per-channel gradient sums the
        epilogue-backward kernels produce into torch.autograd.grad (a handful of GEMVs).  beta: device scalar tensor."""
        sd = self.live["decoder"]
        names = [k for k in sd if k.startswith("mlp.") or ".proj_" in k]
        leaves = {k: sd[k].detach().requires_grad_(True) for k in names}
        with torch.enable_grad():
            nb = (beta.reshape(1) / engine.max_beta - 0.5) * 2
            if getattr(engine, "_freq_dev", None) is None:
                engine._freq_dev = engine.freq.to(self.device)     # first (eager) step: no host copies inside a capture
            freq = engine._freq_dev
            emb = torch.cat([torch.sin(nb * freq), torch.cos(nb * freq)], dim=0)
            if engine.include_x:
                emb = torch.cat([nb, emb], dim=0)
            c = F.linear(torch.relu(F.linear(emb.unsqueeze(0), leaves["mlp.0.weight"], leaves["mlp.0.bias"])),
                         leaves["mlp.2.weight"], leaves["mlp.2.bias"])
            vecs = [F.linear(c, leaves[f"{b}.block{i}.proj_{k}.weight"].flatten(1), leaves[f"{b}.block{i}.proj_{k}.bias"]).reshape(-1)
                    for b in ("block1", "block2", "block3") for i in range(3) for k in (1, 2, 3)]
            flat_vec = torch.cat(vecs)
        flat = engine.cond_vecs.flat
        flat.copy_(flat_vec.detach())
        engine.cond_vecs._live = None
        views = engine.cond_vecs.views
        if self.cond_grad is None:
            self.cond_grad = torch.zeros_like(flat)
            o = 0
            for i, n in enumerate(engine.cond_vecs.sizes):
                self._gain_by_ptr[views[i].data_ptr()] = ("cond", o, n)
                o += n
            self.gain_grad["cond"] = self.cond_grad
        self.cond_state = (names, leaves, flat_vec)
        return views

    def finish_cond_grads(self, inv_scale):
        if self.cond_state is None:
            return
        names, leaves, flat_vec = self.cond_state
        grads = torch.autograd.grad(flat_vec, [leaves[k] for k in names], grad_outputs=self.cond_grad * inv_scale, allow_unused=True)
        for k, g in zip(names, grads):
            if g is not None:
                self.grads["decoder." + k].add_(g)
        self.cond_grad.zero_()
        self.cond_state = None

    def gain_grad_slot(self, vec):
        which, o, n = self._gain_by_ptr[vec.data_ptr()]
        return self.gain_grad[which][o:o + n]

    def finish_gain_grads(self, inv_scale, which_sets=("encoder", "decoder")):
        """d(scale), d(shift) sums -> interp_ca_list.{i}.weight / .bias gradients (softplus and the level lerp)."""
        for which, (engine, l, r, alpha, lerped) in self.gain_state.items():
            if which not in which_sets:
                continue
            gg = self.gain_grad[which]
            o = 0
            for i in range(len(engine.gains)):
                n = engine.gain_vecs.sizes[2 * i]
                dscale, dshift = gg[o:o + n] * inv_scale, gg[o + n:o + 2 * n] * inv_scale
                o += 2 * n
                dw = dscale * torch.sigmoid(lerped[i])
                gW = self.grads[f"{which}.interp_ca_list.{i}.weight"]
                gB = self.grads[f"{which}.interp_ca_list.{i}.bias"]
                shape = gW[l].shape
                gW[l] += alpha * dw.view(shape)
                gB[l] += alpha * dshift.view(shape)
                if r != l or alpha != 1.0:
                    gW[r] += (1 - alpha) * dw.view(shape)
                    gB[r] += (1 - alpha) * dshift.view(shape)
            gg.zero_()


def _eb_likelihood(sd, v):
    """CompressAI EntropyBottleneck._likelihood on points v [C, 1, K] with the live parameters (autograd on the device)."""
    def logits(x):
        h = x
        for i in range(5):
            h = torch.matmul(F.softplus(sd[f"_matrix{i}"]), h) + sd[f"_bias{i}"]
            if i < 4:
                h = h + torch.tanh(sd[f"_factor{i}"]) * torch.tanh(h)
        return h
    lower, upper = logits(v - 0.5), logits(v + 0.5)
    sign = (-torch.sign(lower + upper)).detach()
    return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))


class CodecTrainer:
    """One data-parallel replica of the rate-distortion training step (stage 1 / stage 2 models)."""

    def __init__(self, model, device="cuda:0", lr=1e-4, betas=(0.9, 0.999), eps=1e-8, clip_max_norm=None,
                 lambda_mse=150.0, rate_lambda_a=(3.6, 1.8, 0.8, 0.4, 0.1), rate_lambda_b=2.0 ** -6,
                 target_rate=(0.08, 0.16, 0.36, 0.72, 1.2), aux_lr=1e-3, loss_scale=None, process_group=None,
                 perceptual_weight=0.0):
        self.model, self.device = model, torch.device(device)
        # programmatic dependent launch between consecutive convolutions: the next launch's prologue overlaps the previous
        # one's tail; neutral for the full-GPU inference launches, -3.7 % on the small launches of a training step
        os.environ.setdefault("CRDR_CONV_PDL", "1")
        self.ctx = ctx = TrainContext(model, device)
        X3, X1 = nv.PREC_F16X3, nv.PREC_F16X1
        with torch.cuda.device(self.device):
            self.ga = model.encoder.lower(device, sd=ctx.live_sd("encoder"), precision=X3, train=ctx)
            self.ha = model.hyperencoder.lower(device, sd=ctx.live_sd("hyperencoder"), precision=X3, train=ctx)
            self.hs = model.hyperdecoder.lower(device, sd=ctx.live_sd("hyperdecoder"), precision=X3, train=ctx)
            self.charm = model.context_model.lower(device, sd=ctx.live_sd("context_model"), precision=X3, train=ctx)
            self.gs = model.decoder.lower(device, sd=ctx.live_sd("decoder"), precision=X1, train=ctx)
        self.uses_beta = bool(self.gs.has_cond)
        self._beta = torch.zeros(1, dtype=torch.float32, device=self.device)   # device-resident (graph replays read it)
        self.eb_sd = ctx.live_sd("entropy_model_z")
        from .codec import GaussianParams
        from .entropy import get_scale_table
        table = model.entropy_model_y.scale_table
        self.gp = GaussianParams(table if table.numel() else get_scale_table(), float(model.entropy_model_y.scale_bound.item()), device)
        self.zc = model.entropy_model_z.channels
        self.uses_rate = bool(getattr(model, "uses_rate", False))
        self.lr, self.betas, self.eps, self.clip = lr, betas, eps, clip_max_norm
        self.aux_lr = aux_lr
        self.lambda_mse = float(lambda_mse)
        # LPIPSLoss (perceptual_loss.py:11-31) needs pretrained AlexNet weights that are not available offline.  What can be
        # lowered is the weight-free stand-in the oracle stack uses in its place (oracle/shims/lpips.py: the per-image mean
        # squared difference on the [-1, 1] scale), i.e. a second squared-error term of this weight (0: term absent).
        self.perceptual_weight = float(perceptual_weight)
        self._pfac = torch.ones(1, dtype=torch.float32, device=torch.device(device))   # stage 3: beta (device-resident)
        as_cfg = lambda v: list(v) if isinstance(v, (list, tuple)) else float(v)     # per quality level, or one value
        self.lambda_a, self.lambda_b, self.target = as_cfg(rate_lambda_a), as_cfg(rate_lambda_b), as_cfg(target_rate)
        self.loss_scale = None if loss_scale is None else float(loss_scale)   # None: from the batch's pixel count
        self.m = torch.zeros_like(ctx.flat_p)
        self.v = torch.zeros_like(ctx.flat_p)
        self.step_count = 0
        self.pg = process_group
        self._sumsq = torch.zeros(1024 + 1, dtype=torch.float32, device=self.device)
        # everything a captured step reads that changes from step to step lives in device memory: the HiFiC rate weight,
        # the learning rates and Adam's bias corrections
        f32 = dict(dtype=torch.float32, device=self.device)
        self._rate_w = torch.zeros(1, **f32)
        self._qbpp = torch.zeros(1, **f32)
        self._level = 0
        self._loss_vals = torch.zeros(2, **f32)          # mean bpp | weighted distortion (+ perceptual) of the current step
        self._skip = torch.zeros(1, **f32)               # 1: loss is nan / inf / > 10000 -> no update (base_trainer.py:228-238)
        self._lr = torch.tensor([lr, aux_lr], **f32)
        self._step_dev = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._hyper = torch.zeros(2, 4, **f32)           # rows: main / aux; columns: lr, 1 - b1^t, sqrt(1 - b2^t), unused
        self._clip_coef = torch.ones(1, **f32)
        ctx.repack()
        self._grads = {}     # activation storage pointer -> gradient Act of the current step
        self._keep = []
        self._warm, self._graphs, self._pool = set(), {}, None
        # flat-buffer ranges whose gradients are final after phase 1 of the backward (decoder.*, context_model.*) and the rest
        early = [i for i, k in enumerate(ctx.names) if k.startswith(("decoder.", "context_model."))]
        ends = ctx.offsets[1:] + [ctx.total]
        self._early_ranges, self._late_ranges, pos = [], [], 0
        for i in range(len(ctx.names)):
            target = self._early_ranges if i in set(early) else self._late_ranges
            if target and target[-1][1] == ctx.offsets[i]:
                target[-1] = (target[-1][0], ends[i])
            else:
                target.append((ctx.offsets[i], ends[i]))
        self.use_graphs = True
        bw.workspace(256 << 20, self.device)   # split-K partial sums / column sums: sized once (no growth inside a capture)
        # parameter-gradient work (wgrad + split-K reduction, bias / gain column sums) hangs off the critical chain
        # epilogue-backward -> dgrad -> next layer: it runs on side streams with their own scratch and joins at the end
        self.side_streams = int(os.environ.get("CRDR_TRAIN_SIDE_STREAMS", "4"))
        self._sides = [(torch.cuda.Stream(device=self.device), torch.empty(64 << 20, dtype=torch.uint8, device=self.device))
                       for _ in range(self.side_streams)]
        self._side_next, self._side_used = 0, set()

    # ------------------------------------------------------------------ forward (training mode, taped)
    def _eb_kernel_params(self):
        sd, c = self.eb_sd, self.zc
        cols = []
        for i in range(5):
            cols.append(F.softplus(sd[f"_matrix{i}"]).reshape(c, -1))
            cols.append(sd[f"_bias{i}"].reshape(c, -1))
            if i < 4:
                cols.append(torch.tanh(sd[f"_factor{i}"]).reshape(c, -1))
        return torch.cat(cols, dim=1).contiguous(), sd["quantiles"][:, 0, 1].contiguous()

    def forward(self, images, q, noise, beta=None):
        """images: [n, 3, h, w] fp32 device tensor (h, w multiples of 64); noise = {"y", "z"} uniform in [-1/2, 1/2);
        beta: realism weight of the beta-conditioned decoder (float: stored in the device scalar; None: keep it).
        Returns the dict of device tensors of the training-mode forward and leaves the tape for backward()."""
        ctx, L, st = self.ctx, nv.lib(), nv.stream_handle()
        if beta is not None:
            self._beta.fill_(float(beta))
        ctx.tape = []
        self._grads, self._keep = {}, []
        n, _, h, w = images.shape
        assert h % 64 == 0 and w % 64 == 0 and images.is_contiguous() and images.dtype == torch.float32
        dev = images.device
        img = Act.empty(n, h // 2, w // 2, self.ga.PATCH_CH, two=True, device=dev)
        nv.check(L.crdr_image_to_patches(images.data_ptr(), n, h, w, h, w, img.planes(0), st))
        y_act, y32 = self.ga.run(img, q)
        z32 = self.ha.run(y_act)
        hz, wz = z32.shape[1:3]
        zhat = Act.empty(n, hz, wz, self.zc, two=True, device=dev)
        z_sym = torch.empty((n, self.zc, hz, wz), dtype=torch.int32, device=dev)
        z_hat = torch.empty((n, self.zc, hz, wz), dtype=torch.float32, device=dev)
        z_lik = torch.empty((n, self.zc, hz, wz), dtype=torch.float32, device=dev)
        z_lik_noisy = torch.empty((n, self.zc, hz, wz), dtype=torch.float32, device=dev)
        eb_params, eb_medians = self._eb_kernel_params()
        d = nv.EbDesc()
        d.z, d.z_cs, d.n, d.hw, d.c = z32.data_ptr(), self.zc, n, hz * wz, self.zc
        d.params, d.medians = eb_params.data_ptr(), eb_medians.data_ptr()
        d.zhat_planes = zhat.planes(0)
        d.symbols, d.zhat_nchw, d.likelihood = z_sym.data_ptr(), z_hat.data_ptr(), z_lik.data_ptr()
        d.noise, d.likelihood_noisy = noise["z"].data_ptr(), z_lik_noisy.data_ptr()
        nv.check(L.crdr_eb_quantize(C.byref(d), st))
        self._keep += [eb_params, eb_medians]
        ctx.tape.append(("eb", z32, zhat, noise["z"]))
        T = self.charm.new_support(n, y_act.h, y_act.w, dev)
        self.hs.run(zhat, T, self.charm.off_mean, self.charm.off_scale)
        self._y_act, self._T, self._zhat = y_act, T, zhat
        self._out_y_sym = None
        yhat32, y_sym, y_idx, y_lik, y_lik_noisy = self.charm.encode(T, y32, self.gp, noise=noise["y"])
        self._out_y_sym = y_sym      # the integer rounding decisions of this forward (tests replay them in the oracle)
        fake_packed = self.gs.run(yhat32, q, self._beta)
        fake = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
        nv.check(L.crdr_phases_to_image_ex(fake_packed.data_ptr(), fake_packed.shape[-1], n, h // 2, w // 2, h, w, fake.data_ptr(), 0, st))
        return dict(fake_images=fake, fake_packed=fake_packed, y32=y32, z32=z32, yhat32=yhat32, z_hat=z_hat, y_lik=y_lik, y_sym=y_sym,
                    z_lik=z_lik, y_lik_noisy=y_lik_noisy, z_lik_noisy=z_lik_noisy, T=T, y_act=y_act, size=(h, w))

    # ------------------------------------------------------------------ losses (values, for logging and the rate switch)
    def _bits(self, lik):
        n = lik.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=lik.device)
        nv.check(nv.lib().crdr_bits_from_likelihood(lik.data_ptr(), n, lik[0].numel(), out.data_ptr(), nv.stream_handle()))
        return out

    def _losses_device(self, images, out, q, decide=True):
        """rate_distortion_trainer.py:70-76 with HificVariableRateLoss (rate_loss.py:84-175) and MSELoss (0_1 scale), all
        on the device: the rate weight (lambda_A if the quantised bpp exceeds the target else lambda_B) is decided by a
        torch.where and left in self._rate_w for the backward kernels, so a step needs no host round trip."""
        h, w = out["size"]
        bpp = (self._bits(out["y_lik_noisy"]) + self._bits(out["z_lik_noisy"])) / (h * w)
        qbpp = (self._bits(out["y_lik"]) + self._bits(out["z_lik"])) / (h * w)
        self._qbpp.copy_(qbpp.mean().reshape(1))
        self._level = int(q) if self.uses_rate else 0
        if decide:
            self._decide_rate()
        mse = torch.mean(((images + 1) / 2 - (out["fake_images"] + 1) / 2) ** 2)
        ld = dict(bpp_mean=bpp.mean(), distortion=self.lambda_mse * mse, bpp=bpp.mean(), qbpp=qbpp.mean())
        if self.perceptual_weight:
            ld["perceptual"] = (self._pfac * (self.perceptual_weight * 4.0) * mse).reshape(())   # mean (x - y)^2 on [-1, 1] = 4 x the 0..1 MSE
        self._loss_vals.copy_(torch.stack([bpp.mean(), ld["distortion"] + ld.get("perceptual", 0.0)]))
        if decide:
            self._decide_skip()
        return ld

    def _decide_rate(self):
        """HiFiC rate switch on the mean quantised bpp of the GLOBAL batch (all ranks), on the device."""
        from .sharding import allreduce_mean_scalar
        allreduce_mean_scalar(self._qbpp, self.pg)
        pick = lambda v: float(v[self._level]) if isinstance(v, (list, tuple)) else float(v)
        lam_a, lam_b, tgt = pick(self.lambda_a), pick(self.lambda_b), pick(self.target)
        self._rate_w.copy_(torch.where(self._qbpp > tgt, torch.full_like(self._rate_w, lam_a), torch.full_like(self._rate_w, lam_b)))

    def _decide_skip(self):
        """check_loss_nan_inf (base_trainer.py:228-238) on the device, after the rate weight is known: a nan / inf / > 10000
        total loss makes the optimiser kernels of this step no-ops (on every rank: the flag is max-reduced)."""
        total = self._rate_w * self._loss_vals[0] + self._loss_vals[1]
        bad = (~torch.isfinite(total)) | (total > 10000.0)
        self._skip.copy_(bad.to(torch.float32))
        if self._distributed():
            import torch.distributed as dist
            dist.all_reduce(self._skip, op=dist.ReduceOp.MAX, group=self.pg)

    def _distributed(self):
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1

    def _finish_losses(self, ld):
        ld["skipped"] = self._skip.reshape(())
        ld["rate_weight"] = self._rate_w.reshape(())
        ld["rate"] = ld["rate_weight"] * ld.pop("bpp_mean")
        return ld

    def losses(self, images, out, q):
        ld = self._finish_losses(self._losses_device(images, out, q))
        ld["rate_weight"] = float(ld["rate_weight"].item())
        return ld

    # ------------------------------------------------------------------ backward
    def _grad(self, key, like=None, shape=None):
        """Gradient Act (fp16, zero-initialised on first use) of the tensor whose storage starts at `key`."""
        g = self._grads.get(key)
        if g is None:
            n, h, w, c = shape if shape is not None else (like.n, like.h, like.w, like.c)
            g = self._grads[key] = Act.zeros(n, h, w, c, two=False, device=self.device)
        return g

    @staticmethod
    def _initial_scale(n, h, w):
        """The loss roots carry 1 / (n h w): a power of two proportional to the pixel count keeps the activation gradients
        in fp16's normal range whatever the batch.  Measured (tools/activation_range.py, profiles/activation_range_r02.txt):
        with n h w / 2 the largest gradient of a step is 2^10 (default-init weights) to 2^13 (calibrated stage-2 weights) and
        saturates for the calibrated beta-conditioned model, hence n h w / 16; an overflow lowers the scale further."""
        return 2.0 ** math.floor(math.log2(max(2.0, n * h * w / 16.0)))

    def _on_overflow(self):
        """fp16 overflow in a backward kernel (gradients were clamped, the device flag is already cleared): lower the loss
        scale and drop the captured graphs, which have the old scale baked into their launch parameters."""
        self.loss_scale = max(1.0, self.loss_scale / 8.0)
        self._graphs.clear()
        self._warm.clear()

    def _partial(self, blocks, nsums, c):
        """Per-record scratch for the per-block column sums, alive until the step ends (side streams read it later)."""
        t = torch.empty(blocks * nsums * c, dtype=torch.float32, device=self.device)
        self._keep.append(t)
        return t

    class _Side:
        """with trainer._side() as ws: ... -- enqueue on the next side stream (forked from the current stream here)."""

        def __init__(self, tr):
            self.tr = tr

        def __enter__(self):
            tr = self.tr
            if not tr._sides:
                return None
            self.main = torch.cuda.current_stream()
            i = tr._side_next
            tr._side_next = (i + 1) % len(tr._sides)
            stream, ws = tr._sides[i]
            ev = torch.cuda.Event()
            ev.record(self.main)
            stream.wait_event(ev)
            tr._side_used.add(i)
            self.ctx = torch.cuda.stream(stream)
            self.ctx.__enter__()
            return ws

        def __exit__(self, *exc):
            if self.tr._sides:
                self.ctx.__exit__(*exc)
            return False

    def _side(self):
        return CodecTrainer._Side(self)

    def _join_sides(self):
        main = torch.cuda.current_stream()
        for i in sorted(self._side_used):
            ev = torch.cuda.Event()
            ev.record(self._sides[i][0])
            main.wait_event(ev)
        self._side_used.clear()

    def backward(self, images, out, rate_weight=None, image_grad=None, image_grad_scale=1.0, phase=None):
        """Reverse sweep over the tape; fills ctx.flat_g (unscaled fp32 gradients of mean-reduced losses).  rate_weight:
        a float overrides the device-resident weight the last losses() call decided.  image_grad: an extra loss-scaled
        gradient w.r.t. the reconstruction as 8-channel NHWC planes (the adversarial term's path through a discriminator).
        phase: None = the whole sweep; 1 = loss roots, g_s and ChARM (afterwards the gradients of `decoder.*` and
        `context_model.*`, 90 % of the parameters, are final: their all-reduce can start); 2 = h_s, the factorised prior, h_a, g_a."""
        if phase == 2:
            return self._sweep(self._pending, ("encoder",), 1.0 / self.loss_scale)
        if rate_weight is not None:
            self._rate_w.fill_(float(rate_weight))
        ctx, L, st = self.ctx, nv.lib(), nv.stream_handle()
        n = images.shape[0]
        h, w = out["size"]
        if self.loss_scale is None:
            self.loss_scale = self._initial_scale(n, h, w)
        S = self.loss_scale
        inv = 1.0 / S
        ctx.flat_g.zero_()
        # ---- loss roots --------------------------------------------------------------------------------------------
        fp = out["fake_packed"]
        g_img = self._grad(fp.data_ptr(), shape=(n, h // 2, w // 2, 16))
        coef_mse = S * self.lambda_mse * 2.0 * 0.25 / (n * 3 * h * w)
        nv.check(L.crdr_mse_backward(fp.data_ptr(), fp.shape[-1], images.data_ptr(), n, h // 2, w // 2, h, w, coef_mse,
                                     g_img.hi.data_ptr(), 16, st))
        if self.perceptual_weight:
            # the stand-in perceptual term is a second squared error: the same gradient direction, rescaled on the device
            # (its weight is multiplied by the device-resident beta in stage 3)
            g_img.hi.mul_((1.0 + self._pfac * (self.perceptual_weight * 2.0 / (self.lambda_mse * 0.5))).to(torch.float16))
        if image_grad is not None:
            nv.check(L.crdr_planes_grad_to_phases(image_grad.hi.data_ptr(), image_grad.c, n, h // 2, w // 2, image_grad_scale,
                                                  g_img.hi.data_ptr(), 16, st))
        self._rate_coef = S / (math.log(2.0) * n * h * w)   # d(mean bpp) / d(-ln L) per element; x the device-resident rate weight
        tape, ctx.tape = ctx.tape, None
        recs = list(reversed(tape))
        if phase is None:
            return self._sweep(recs, ("encoder", "decoder"), inv)
        hs_cfg = self.hs.mu[0].cfg
        cut = next((i for i, r in enumerate(recs) if r[0] == "conv" and r[1].cfg is hs_cfg), len(recs))
        self._pending = recs[cut:]
        self._sweep(recs[:cut], ("decoder",), inv)

    def _sweep(self, recs, gain_sets, inv):
        ctx = self.ctx
        for rec in recs:
            kind = rec[0]
            if kind == "conv":
                self._conv_backward(*rec[1:])
            elif kind == "gate":
                self._gate_backward(*rec[1:])
            elif kind == "affine_in":
                self._affine_in_backward(*rec[1:])
            elif kind == "gauss":
                self._gauss_backward(*rec[1:])
            elif kind == "eb":
                self._eb_backward(*rec[1:])
            elif kind == "lrelu":
                self._lrelu_backward(*rec[1:])
        self._join_sides()
        ctx.finish_gain_grads(inv, gain_sets)
        if "decoder" in gain_sets:
            ctx.finish_cond_grads(inv)

    def _sums_to(self, partial, blocks, nsums, which, c, target, accumulate=True, scale=None):
        nv.check(nv.lib().crdr_colsum_finish(partial.data_ptr(), blocks, nsums, which, c, target.data_ptr(),
                                             (1.0 / self.loss_scale) if scale is None else scale, 1 if accumulate else 0,
                                             nv.stream_handle()))

    def _conv_backward(self, conv, x, kw, out, bwd_res, no_input_grad):
        ctx, L, st = conv.cfg.train, nv.lib(), nv.stream_handle()
        op = conv.op
        out_f32 = kw.get("out_f32")
        if out is not None and kw.get("want_planes", True):
            G, gcoff = self._grads.get(out.hi.data_ptr()), kw.get("out_coff", 0)
        else:
            G, gcoff = self._grads.get(out_f32.data_ptr()), kw.get("out_f32_coff", 0)
        if G is None:
            return          # nothing downstream depends on this output
        cout = op.cout
        relu, mode = bool(kw.get("relu")), kw.get("mode", nv.EPI_NONE)
        scale, shift = kw.get("scale"), kw.get("shift")
        add_vec = kw.get("add_vec")
        affine = scale is not None or shift is not None
        assert not (add_vec is not None and relu and affine), "ReLU + conditioning bias + gain in one epilogue is not used by the path"
        tanh = mode == nv.EPI_HALF_TANH
        need_dv = relu or affine or tanh
        ho, wo = G.h, G.w
        m = G.n * ho * wo
        blocks = max(1, min(2048, m // 32))
        partial = self._partial(blocks, 3, cout)
        d = nv.EpiBwdDesc()
        d.g = nv.Planes(G.hi.data_ptr(), None, G.c, gcoff)
        d.m, d.c, d.relu = m, cout, 1 if relu else 0
        if relu or affine:
            d.out = out.planes(kw.get("out_coff", 0))
        d.scale, d.shift = nv.ptr(scale), nv.ptr(shift)
        if add_vec is not None and relu:
            d.add_vec = add_vec.data_ptr()
        if tanh:
            res32 = kw["res"]
            d.f32_out, d.f32_res = out_f32.data_ptr(), res32.data_ptr()
            d.f32_cs, d.f32_coff = out_f32.shape[-1], kw.get("out_f32_coff", 0)
            assert res32.shape[-1] == out_f32.shape[-1] and kw.get("res_coff", 0) == kw.get("out_f32_coff", 0)
        if need_dv:
            dv = Act.empty(G.n, ho, wo, cout, two=False, device=self.device)
            d.dv = dv.planes(0)
            dv_coff = 0
        else:
            dv, dv_coff = G, gcoff
        if mode == nv.EPI_RESIDUAL:
            res = kw["res"]
            Gr = self._grad(res.hi.data_ptr(), like=res)
            d.dres = nv.Planes(Gr.hi.data_ptr(), None, Gr.c, kw.get("res_coff", 0))
        elif tanh:
            Tres, coff = bwd_res
            Gr = self._grad(Tres.hi.data_ptr(), like=Tres)
            d.dres = nv.Planes(Gr.hi.data_ptr(), None, Gr.c, coff)
        d.partial, d.blocks = partial.data_ptr(), blocks
        nv.check(L.crdr_epilogue_backward(C.byref(d), st))
        segs = kw.get("segs") or [(0, op.cin)]
        if need_dv:
            self._keep.append(dv)       # read by the side stream after this function returns
        with self._side() as ws:
            self._param_grads(conv, op, x, kw, dv, dv_coff, cout, partial, blocks, scale, shift, affine, ws)
            if add_vec is not None:     # conditioning bias: after the ReLU (sum of g) or, without one, like the bias (sum of dv)
                self._sums_to(partial, blocks, 3, 1 if relu else 0, op.cout, ctx.gain_grad_slot(add_vec), scale=1.0)
        # dgrad
        if no_input_grad:
            return
        if conv.dgrad is None:
            # (a convolution whose matrix is a transform of the parameter -- phase-packed last layer, output-channel halves /
            # padded head of the discriminator -- gets its adjoint from the parameter through both index maps)
            conv.dgrad = bw.DgradSet(conv.master, op.transposed, op.stride, op.padding, op.kh, segs, cin_real=op.cin_real,
                                     pre=conv.packed_transform)
            conv.dgrad.repack()
        Gx = self._grad(x.hi.data_ptr(), like=x)
        if need_dv:
            conv.dgrad.run(dv, Gx)
        else:
            conv.dgrad.run(G, Gx, dv_coff=gcoff, dv_c=cout)

    def _param_grads(self, conv, op, x, kw, dv, dv_coff, cout, partial, blocks, scale, shift, affine, ws):
        """Bias / gain column sums and the weight gradient of one convolution record (side stream)."""
        ctx = conv.cfg.train
        if not ctx.param_grads:
            return          # this network's parameters are frozen in this sweep (discriminator during the generator step)
        if conv.bias_master is not None:
            if conv.bias_buf is not None:      # derived bias vector (phase-packed / sliced / padded): scatter through its index map
                tmp = torch.zeros(cout, dtype=torch.float32, device=self.device)
                self._sums_to(partial, blocks, 3, 0, cout, tmp, accumulate=False)
                ctx.grad_view(conv.bias_master).view(-1).index_add_(0, conv.bmap_dst, tmp[conv.bmap_src])
            else:
                self._sums_to(partial, blocks, 3, 0, op.cout, ctx.grad_view(conv.bias_master))
        if affine:
            if shift is not None:
                self._sums_to(partial, blocks, 3, 1, op.cout, ctx.gain_grad_slot(shift), scale=1.0)
            if scale is not None:
                self._sums_to(partial, blocks, 3, 2, op.cout, ctx.gain_grad_slot(scale), scale=1.0)
        # wgrad
        segs = kw.get("segs") or [(0, op.cin)]
        k2 = op.kh * op.kw
        taps = [(i - op.padding, j - op.padding) for i in range(op.kh) for j in range(op.kw)]
        if conv.wmap is not None:
            # transformed weight (im2col'd first layer, phase-packed last layer): gradient in the ConvOp's layout, then
            # scattered to the parameter through the transform's index map
            tmp = torch.empty((op.cout, op.cin, op.kh, op.kw), dtype=torch.float32, device=self.device)
            bw.wgrad(dv, dv_coff, op.cout, x, 0, op.cin, taps, op.stride, tmp, op.cin * k2, k2, 1, scale=1.0 / self.loss_scale, ws=ws)
            ctx.grad_view(conv.master).view(-1).index_add_(0, conv.wmap_dst, tmp.view(-1)[conv.wmap_src])
        else:
            gw = ctx.grad_view(conv.master).view(-1)
            start = 0
            for off, ln in segs:
                real = min(ln, op.cin_real - start)
                if real > 0:
                    if op.transposed:   # parameter [ci, co, kh, kw]:  S = x (a = ci), B = dv (b = co)
                        bw.wgrad(x, off, real, dv, dv_coff, op.cout, taps, op.stride, gw[start * op.cout * k2:], op.cout * k2, k2, 1,
                                 scale=1.0 / self.loss_scale, accumulate=True, ws=ws)
                    else:               # parameter [co, ci, kh, kw]:  S = dv (a = co), B = x (b = ci)
                        bw.wgrad(dv, dv_coff, op.cout, x, off, real, taps, op.stride, gw[start * k2:], op.cin_real * k2, k2, 1,
                                 scale=1.0 / self.loss_scale, accumulate=True, ws=ws)
                start += ln

    def _lrelu_backward(self, out, slope):
        """LeakyReLU between two convolutions: the gradient of its output is rescaled in place (the convolution records that
        wrote `out` come next in the reverse sweep and read it as the gradient of their result)."""
        G = self._grads.get(out.hi.data_ptr())
        if G is None:
            return
        d = nv.EpiBwdDesc()
        d.g, d.out, d.dv = G.planes(0), out.planes(0), G.planes(0)
        d.m, d.c, d.relu, d.leaky_slope, d.blocks = out.pixels, out.c, 1, slope, max(1, min(2048, out.pixels // 32))
        nv.check(nv.lib().crdr_epilogue_backward(C.byref(d), nv.stream_handle()))

    def _gate_backward(self, x, t, a, out, scale, shift):
        G = self._grads.get(out.hi.data_ptr())
        if G is None:
            return
        ctx, L, st = self.ctx, nv.lib(), nv.stream_handle()
        Gx, Gt, Ga = self._grad(x.hi.data_ptr(), like=x), self._grad(t.hi.data_ptr(), like=t), self._grad(a.hi.data_ptr(), like=a)
        blocks = max(1, min(512, x.pixels // 64))
        partial = self._partial(blocks, 2, x.c)
        d = nv.GateDesc()
        d.x, d.t, d.a = x.planes(0), t.planes(0), a.planes(0)
        d.m, d.c, d.scale, d.shift = x.pixels, x.c, nv.ptr(scale), nv.ptr(shift)
        d.g, d.dx, d.dt, d.da = G.planes(0), Gx.planes(0), Gt.planes(0), Ga.planes(0)
        d.partial, d.blocks = partial.data_ptr(), blocks
        nv.check(L.crdr_gate_backward(C.byref(d), st))
        if shift is not None:
            self._sums_to(partial, blocks, 2, 0, x.c, ctx.gain_grad_slot(shift), scale=1.0)
        if scale is not None:
            self._sums_to(partial, blocks, 2, 1, x.c, ctx.gain_grad_slot(scale), scale=1.0)

    def _affine_in_backward(self, x, scale, shift, yhat32):
        """g_s input: x = y_hat * scale + shift (the first InterpChAtt of the decoder) -> d(y_hat) into the support
        tensor's gradient (y_hat channels)."""
        G = self._grads.get(x.hi.data_ptr())
        if G is None:
            return
        T = self._T
        GT = self._grad(T.hi.data_ptr(), like=T)
        affine = scale is not None
        m = x.pixels
        blocks = max(1, min(512, m // 64))
        partial = self._partial(blocks, 3, x.c)
        d = nv.EpiBwdDesc()
        d.g, d.m, d.c = G.planes(0), m, x.c
        if affine:
            d.out, d.scale, d.shift = x.planes(0), nv.ptr(scale), nv.ptr(shift)
        d.dres = nv.Planes(GT.hi.data_ptr(), None, GT.c, self.charm.off_y)
        d.partial, d.blocks = partial.data_ptr(), blocks
        nv.check(nv.lib().crdr_epilogue_backward(C.byref(d), nv.stream_handle()))
        if affine:
            self._sums_to(partial, blocks, 3, 1, x.c, self.ctx.gain_grad_slot(shift), scale=1.0)
            self._sums_to(partial, blocks, 3, 2, x.c, self.ctx.gain_grad_slot(scale), scale=1.0)

    def _gauss_backward(self, s0, cnt, T, y32, ms, noise):
        ch = self.charm
        GT = self._grad(T.hi.data_ptr(), like=T)
        Gy = self._grad(self._y_act.hi.data_ptr(), like=self._y_act)
        n, hgt, wid = T.n, T.h, T.w
        Gms = self._grad(ms.data_ptr(), shape=(n, hgt, wid, ms.shape[-1]))
        d = nv.GaussBwdDesc()
        d.y, d.y_cs, d.y_coff = y32.data_ptr(), ch.yc, s0 * ch.sc
        d.noise, d.ms = noise.data_ptr(), ms.data_ptr()
        d.ms_cs, d.mu_coff, d.sigma_coff = ms.shape[-1], s0 * ch.sc, ch.yc + s0 * ch.sc
        d.n, d.hw, d.c, d.c_total, d.nchw_coff = n, hgt * wid, ch.sc * cnt, ch.yc, s0 * ch.sc
        d.scale_bound, d.lik_bound, d.coef = self.gp.bound, 1e-9, self._rate_coef
        d.coef_scale = self._rate_w.data_ptr()
        d.gpre = nv.Planes(GT.hi.data_ptr(), None, GT.c, ch.off_tmp + s0 * ch.sc)
        d.dy = nv.Planes(Gy.hi.data_ptr(), None, Gy.c, s0 * ch.sc)
        d.dmu = nv.Planes(Gms.hi.data_ptr(), None, Gms.c, s0 * ch.sc)
        d.dsigma = nv.Planes(Gms.hi.data_ptr(), None, Gms.c, ch.yc + s0 * ch.sc)
        nv.check(nv.lib().crdr_gauss_backward(C.byref(d), nv.stream_handle()))

    def _eb_backward(self, z32, zhat, noise_z):
        """z path: straight-through z_hat (entropy_bottleneck.py:23-30) + the rate term of the factorised prior."""
        ctx = self.ctx
        S = self.loss_scale
        n, hz, wz, c = z32.shape
        names = [k for k in self.eb_sd if k.startswith(("_matrix", "_bias", "_factor"))]
        leaves = {k: self.eb_sd[k].detach().requires_grad_(True) for k in names}
        with torch.enable_grad():
            z = z32.detach().permute(0, 3, 1, 2).contiguous().requires_grad_(True)     # NCHW
            v = (z + noise_z).permute(1, 0, 2, 3).reshape(c, 1, -1)
            lik = _eb_likelihood(leaves, v)
            lik = lik + (torch.clamp(lik, min=1e-9) - lik).detach()       # LowerBound: every (negative) gradient passes
            loss = -(self._rate_coef / S) * self._rate_w[0] * torch.log(lik).sum()
            grads = torch.autograd.grad(loss, [z] + [leaves[k] for k in names])
        for k, g in zip(names, grads[1:]):
            ctx.grads["entropy_model_z." + k].add_(g)
        dz = grads[0].permute(0, 2, 3, 1) * S                              # NHWC, loss-scaled
        Gzh = self._grads.get(zhat.hi.data_ptr())
        if Gzh is not None:
            dz = dz + Gzh.hi.float()
        Gz = self._grad(z32.data_ptr(), shape=(n, hz, wz, c))
        Gz.hi.copy_(dz.clamp(-65504.0, 65504.0).half())

    # ------------------------------------------------------------------ the step
    def aux_step(self):
        """BaseModel.aux_loss (base_model.py:65-76) + its own Adam (rate_distortion_trainer.py:96-100); 576 parameters."""
        sd = self.eb_sd
        qv = sd["quantiles"].detach().requires_grad_(True)
        with torch.enable_grad():
            h = qv
            for i in range(5):
                h = torch.matmul(F.softplus(sd[f"_matrix{i}"].detach()), h) + sd[f"_bias{i}"].detach()
                if i < 4:
                    h = h + torch.tanh(sd[f"_factor{i}"].detach()) * torch.tanh(h)
            loss = torch.abs(h - self.ctx.buffers["entropy_model_z.target"]).sum()
            (g,) = torch.autograd.grad(loss, [qv])
        self.ctx.grads["entropy_model_z.quantiles"].copy_(g)
        return loss.detach()

    def all_reduce_grads(self, bucket_bytes=64 << 20):
        """Data parallelism: average the flat gradient buffer over the ranks in fixed-size buckets (NCCL over NVLink;
        SURVEY 8e: 127.7 M parameters = 510.9 MB per step).  A no-op for a single process."""
        from .sharding import allreduce_mean_flat
        return allreduce_mean_flat(self.ctx.flat_g, bucket_bytes, self.pg)

    def set_lr(self, lr=None, aux_lr=None):
        """Scheduler hook (MultiStepLR of crdr_stage_2.yaml): the rates live on the device."""
        if lr is not None:
            self.lr = lr
            self._lr[0] = lr
        if aux_lr is not None:
            self.aux_lr = aux_lr
            self._lr[1] = aux_lr

    def optimizer_step(self):
        ctx, L, st = self.ctx, nv.lib(), nv.stream_handle()
        self.step_count += 1
        self._step_dev += 1.0 - self._skip.to(torch.float64)     # a skipped step does not advance Adam's bias correction
        b1, b2 = self.betas
        bc = torch.stack([1.0 - b1 ** self._step_dev, torch.sqrt(1.0 - b2 ** self._step_dev)], dim=1).to(torch.float32)   # [1, 2]
        self._hyper[:, 0] = self._lr
        self._hyper[:, 1:3] = bc
        self._hyper[:, 3] = self._skip
        clip_ptr = None
        if self.clip:
            nv.check(L.crdr_sum_squares(ctx.flat_g.data_ptr(), ctx.n_main_padded, self._sumsq.data_ptr(), self._sumsq[1024:].data_ptr(), st))
            self._clip_coef.copy_(torch.clamp(self.clip / (torch.sqrt(self._sumsq[1024]) + 1e-6), max=1.0).reshape(1))
            clip_ptr = self._clip_coef.data_ptr()
        nv.check(L.crdr_adam_step(ctx.flat_p.data_ptr(), ctx.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), ctx.n_main_padded,
                                  self.lr, b1, b2, self.eps, 0, clip_ptr, 1.0, self._hyper[0].data_ptr(), st))
        if ctx.total > ctx.n_main_padded:   # aux parameters (quantiles): their own learning rate, no clipping
            o, cnt = ctx.n_main_padded, ctx.total - ctx.n_main_padded
            nv.check(L.crdr_adam_step(ctx.flat_p[o:].data_ptr(), ctx.flat_g[o:].data_ptr(), self.m[o:].data_ptr(), self.v[o:].data_ptr(),
                                      cnt, self.aux_lr, b1, b2, self.eps, 0, None, 1.0, self._hyper[1].data_ptr(), st))
        ctx.repack()

    def _core_forward(self, images, q, noise):
        """Training-mode forward + loss values: capture-safe (no host synchronisation, no collective)."""
        self._out = self.forward(images, q, noise)      # beta: the device scalar set by train_step
        return self._losses_device(images, self._out, q, decide=False)

    def _core_backward(self, images, phase=None):
        self.backward(images, self._out, phase=phase)
        return self.aux_step() if phase in (None, 2) else None

    def train_step(self, images, q=None, noise=None, generator=None, beta=None):
        """One optimisation step on a batch of [-1, 1] crops (device fp32 NCHW).  Returns the loss dict (device scalars;
        overwritten by the next step of the same shape and quality level when CUDA graphs are on).

        The step is ~1 900 small launches whose host enqueue (47 ms) costs several times their device time, so from the
        second call of a (shape, quality level) on, the step replays four captured CUDA graphs -- F = forward + loss
        values, B1 = backward of g_s and ChARM, B2 = backward of h_s / h_a / g_a, O = clip + Adam + re-packing -- with the
        data-parallel exchanges between them: the qbpp mean of the rate switch after F, the all-reduce of the decoder and
        context-model gradients (90 % of the 510 MB) started after B1 so that it overlaps B2, the rest after B2."""
        from .sharding import broadcast_from_rank0
        n, _, h, w = images.shape
        if q is None:
            qt = torch.randint(self.model.rate_level, (1,)).to(self.device, torch.float32) if self.uses_rate else torch.zeros(1, device=self.device)
            q = float(broadcast_from_rank0(qt, self.pg).item())      # one level per (global) batch, like the reference
        if self.uses_beta:
            if beta is None:    # beta_cond_interpca_hyperprior_model.py:43-45: max_beta * randint(0, 101) / 100, one per batch
                bt = torch.randint(0, 101, (1,)).to(self.device, torch.float32) * (self.gs.max_beta / 100.0)
                self._beta.copy_(broadcast_from_rank0(bt, self.pg))
            else:
                self._beta.fill_(float(beta))
        if noise is None:
            mk = lambda c, a, b: torch.rand((n, c, a, b), dtype=torch.float32, device=self.device, generator=generator) - 0.5
            noise = {"z": mk(self.zc, h // 64, w // 64), "y": mk(self.charm.yc, h // 16, w // 16)}
        key = (n, h, w, float(q))
        if not self.use_graphs or key not in self._warm:
            # eager: also the warm-up that builds the adjoint matrices, tensor maps and kernel attributes before a capture
            for attempt in range(4):
                ld = self._core_forward(images, q, noise)
                self._decide_rate()
                self._decide_skip()
                ld["aux"] = self._core_backward(images)
                try:
                    nv.status_check()       # before the update: a step whose gradients were clamped is redone
                    break
                except nv.NativeError as e:
                    if "overflow" not in str(e) or attempt == 3:
                        raise
                    self._on_overflow()
            self.all_reduce_grads()
            self.optimizer_step()
            self._warm.add(key)
            return self._finish_losses(ld)
        hit = self._graphs.get(key)
        if hit is None:
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            st_in = dict(x=torch.empty_like(images), z=torch.empty_like(noise["z"]), y=torch.empty_like(noise["y"]))
            st_in["x"].copy_(images); st_in["z"].copy_(noise["z"]); st_in["y"].copy_(noise["y"])
            torch.cuda.current_stream().synchronize()
            gf, gb, gb2, go = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            counts, l0 = [], nv.LAUNCH_COUNT[0]
            with torch.cuda.graph(gf, pool=self._pool):
                ld = self._core_forward(st_in["x"], q, {"z": st_in["z"], "y": st_in["y"]})
            with torch.cuda.graph(gb, pool=self._pool):
                self._core_backward(st_in["x"], phase=1)
            with torch.cuda.graph(gb2, pool=self._pool):
                ld["aux"] = self._core_backward(st_in["x"], phase=2)
            count = self.step_count
            with torch.cuda.graph(go, pool=self._pool):
                self.optimizer_step()
            self.step_count = count        # the capture itself executes nothing
            hit = self._graphs[key] = (gf, gb, gb2, go, st_in, ld, nv.LAUNCH_COUNT[0] - l0, int(q) if self.uses_rate else 0)
        gf, gb, gb2, go, st_in, ld, launches, level = hit
        st_in["x"].copy_(images, non_blocking=True)
        st_in["z"].copy_(noise["z"], non_blocking=True)
        st_in["y"].copy_(noise["y"], non_blocking=True)
        gf.replay()
        self._level = level
        self._decide_rate()
        self._decide_skip()
        from .sharding import allreduce_finish_mean, allreduce_sum_async
        gb.replay()                 # loss roots, g_s, ChARM: decoder.* and context_model.* gradients are final ...
        handles = allreduce_sum_async(self.ctx.flat_g, self._early_ranges, 64 << 20, self.pg)   # ... and travel while ...
        gb2.replay()                # ... h_s, the factorised prior, h_a and g_a run
        handles += allreduce_sum_async(self.ctx.flat_g, self._late_ranges, 64 << 20, self.pg)
        allreduce_finish_mean(handles, self.ctx.flat_g, self.pg)
        go.replay()
        self.step_count += 1
        nv.LAUNCH_COUNT[0] += launches
        try:
            nv.status_check()
        except nv.NativeError as e:
            if "overflow" not in str(e):
                raise
            self._on_overflow()             # this update used clamped gradients (finite); the next steps run at a lower scale
        return self._finish_losses(dict(ld))

    def training_state(self):
        """Everything needed to resume (base_trainer.save with keep_training_state: model parameters + optimiser state):
        CPU tensors keyed by parameter name, Adam step count, loss scale."""
        c = self.ctx
        take = lambda flat: {k: flat[o:o + c.params[k].numel()].view(c.params[k].shape).detach().cpu().clone()
                             for k, o in zip(c.names, c.offsets)}
        return {"params": take(c.flat_p), "exp_avg": take(self.m), "exp_avg_sq": take(self.v), "step": int(round(float(self._step_dev))),
                "step_count": self.step_count, "loss_scale": self.loss_scale, "lr": self.lr, "aux_lr": self.aux_lr}

    def load_training_state(self, state):
        c = self.ctx
        for src, flat in ((state["params"], c.flat_p), (state["exp_avg"], self.m), (state["exp_avg_sq"], self.v)):
            for k, o in zip(c.names, c.offsets):
                flat[o:o + c.params[k].numel()].view(c.params[k].shape).copy_(src[k].to(self.device))
        self._step_dev.fill_(float(state["step"]))
        self.step_count, self.loss_scale = int(state["step_count"]), state["loss_scale"]
        self.set_lr(state["lr"], state["aux_lr"])
        self._graphs.clear()
        self._warm.clear()
        c.repack()

    def sync_to_model(self):
        """Copy the trained parameters back into the nn.Module (checkpointing: state_dict layout of the reference)."""
        with torch.no_grad():
            for k, p in self.model.named_parameters():
                p.copy_(self.ctx.params[k].to(p.device))


class GanCodecTrainer(CodecTrainer):
    """Stage 3 of the reference (MultirateBetaCondHrrGanRateDistortionTrainer.optimize_parameters,
    multirate_hr_rgan_beta_cond_rate_distortion_trainer.py:13-114) on the CUDA engines (LPIPS as its weight-free stand-in):

      G step:  l = distortion + rate + beta * (perceptual + adv),
               adv = lambda_gan / 2 * (BCE(D(rel) - D(fake), 0) + BCE(D(fake) - D(rel), 1))
               where `rel` is the reconstruction at the next higher quality level (no gradient; the real image at the top
               level) and D = sub-discriminator int(q); the discriminator's parameters are frozen, its input gradient is
               added to the MSE gradient of the reconstruction.
      D step:  0.5 * BCE(D(real) - D(fake).detach(), 1) + 0.5 * BCE(D(fake.detach()) - D(real).detach(), 0), Adam on the
               active sub-discriminator only (parameters without a gradient are skipped by torch.optim.Adam).

    The discriminator runs through the same tape / reverse-sweep machinery as the codec (3x3 convolutions, LeakyReLU,
    tcgen05 wgrad, dgrad); the logit maps are 16 x 16 per crop, so the BCE terms are evaluated with torch on the device."""

    def __init__(self, model, discriminator, device="cuda:0", lambda_gan=0.000390625, d_lr=1e-4, relative_score_rate_delta=1,
                 **kw):
        kw.setdefault("target_rate", (0.0,) * 5)
        kw.setdefault("rate_lambda_a", (3.4, 1.3, 0.4, 0.12, 0.05))
        super().__init__(model, device=device, **kw)
        self.discriminator = discriminator
        self.dctx = TrainContext(discriminator, device)
        subs = len(discriminator.subD_list)
        with torch.cuda.device(self.device):
            self.D = [discriminator.subD_list[k].lower(device, sd=self.dctx.live_sd(f"subD_list.{k}"), train=self.dctx) for k in range(subs)]
        self.dctx.repack()
        # flat-buffer segment of every sub-discriminator (named_parameters order: subD_list.0.*, subD_list.1.*, ...)
        self._dseg = []
        for k in range(subs):
            idx = [i for i, n in enumerate(self.dctx.names) if n.startswith(f"subD_list.{k}.")]
            lo = self.dctx.offsets[idx[0]]
            hi = self.dctx.offsets[idx[-1] + 1] if idx[-1] + 1 < len(self.dctx.offsets) else self.dctx.total
            self._dseg.append((lo, hi))
        self.dm, self.dv = torch.zeros_like(self.dctx.flat_p), torch.zeros_like(self.dctx.flat_p)
        self.d_steps = [0] * subs
        self.lambda_gan, self.d_lr, self.delta = float(lambda_gan), float(d_lr), int(relative_score_rate_delta)
        self._dstep_dev = torch.zeros(subs, dtype=torch.float64, device=self.device)
        self._dhyper = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.use_gan_graphs = True

    # ------------------------------------------------------------------ discriminator passes
    def d_forward(self, k, images, tape=True, input_grad=False):
        """Sub-discriminator k on [n, 3, h, w] fp32 device images.  Returns (logit map [n, h/16, w/16] fp32, tape, logits Act,
        input planes Act)."""
        n, _, h, w = images.shape
        x = Act.empty(n, h, w, 8, two=False, device=self.device)
        nv.check(nv.lib().crdr_image_to_planes(images.data_ptr(), n, h, w, h, w, x.planes(0), nv.stream_handle()))
        self.dctx.tape = [] if tape else None
        logits = self.D[k].run(x, no_input_grad=not input_grad)
        rec, self.dctx.tape = self.dctx.tape, None
        return logits.hi[..., 0].float(), rec, logits, x

    def d_backward(self, tape, logits, dpred, param_grads=True, peak_bound=None):
        """Reverse sweep of one discriminator pass from d(loss)/d(logit map) (unscaled fp32 [n, h', w']).  The pass uses
        its own power-of-two loss scale, chosen so that the root gradient sits at 2^6 in fp16 (the adversarial term is
        1e-4 of the MSE term: under the codec's scale its gradients would live in fp16's subnormals); returns that scale.
        peak_bound: an analytic bound of max |dpred| (no host round trip: needed inside a CUDA-graph capture)."""
        peak = float(dpred.abs().max().item()) if peak_bound is None else float(peak_bound)
        scale = 2.0 ** math.floor(math.log2(64.0 / peak)) if peak > 0 else 1.0
        G = self._grad(logits.hi.data_ptr(), like=logits)
        G.hi.zero_()
        G.hi[..., 0] = (dpred * scale).clamp(-65504.0, 65504.0).half()
        saved, self.loss_scale = self.loss_scale, scale
        self.dctx.param_grads = param_grads
        try:
            for rec in reversed(tape):
                if rec[0] == "conv":
                    self._conv_backward(*rec[1:])
                elif rec[0] == "lrelu":
                    self._lrelu_backward(*rec[1:])
            self._join_sides()
        finally:
            self.dctx.param_grads = True
            self.loss_scale = saved
        return scale

    @staticmethod
    def _bce(x, target):
        return F.binary_cross_entropy_with_logits(x, torch.full_like(x, target))

    # ------------------------------------------------------------------ the step
    def generator_backward(self, images, q, noise, beta, rel, bounded=False):
        """Forward + losses + backward of the generator step (parameters of the discriminator frozen): fills ctx.flat_g with
        the gradient of distortion + rate + beta * (perceptual stand-in + adv).  rel: the relative-score images (no gradient).
        beta: float (stored in the device scalar) or None (keep the device scalar).  bounded: size the discriminator pass's loss
        scale from an analytic bound instead of the measured peak (capture-safe)."""
        k = int(q)
        if self.loss_scale is None:
            self.loss_scale = self._initial_scale(images.shape[0], images.shape[2], images.shape[3])
        out = self.forward(images, q, noise, beta=beta)
        self._pfac.copy_(self._beta)                     # l_total = dist + rate + beta * (perceptual + adv)
        fake = out["fake_images"]
        ld = self._losses_device(images, out, q)
        real_d, _, _, _ = self.d_forward(k, rel, tape=False)
        fake_g, dtape, dlogits, dplanes = self.d_forward(k, fake, tape=True, input_grad=True)
        with torch.enable_grad():
            fg = fake_g.detach().requires_grad_(True)
            adv = self.lambda_gan * 0.5 * (self._bce(real_d - fg, 0.0) + self._bce(fg - real_d, 1.0))
            (dfg,) = torch.autograd.grad(self._beta[0] * adv, [fg])
        self.dctx.flat_g.zero_()
        # |d(beta adv)/d logit| <= beta_max * lambda_gan / (number of logits): each BCE term's derivative is a sigmoid / count
        bound = self.gs.max_beta * self.lambda_gan / fake_g.numel() if bounded else None
        d_scale = self.d_backward(dtape, dlogits, dfg, param_grads=False, peak_bound=bound)
        g_in = self._grads[dplanes.hi.data_ptr()]          # d(beta * adv) / d(fake image) as 8-channel planes, x d_scale
        self.backward(images, out, image_grad=g_in, image_grad_scale=self.loss_scale / d_scale)
        return ld, adv, fake

    def _gan_core(self, images, q, noise, noise_rel, bounded):
        """The whole stage-3 step on device-resident scalars (beta, rate weight, skip flag, Adam schedules): no host round
        trip, so it can be enqueued eagerly or captured into one CUDA graph."""
        k = int(q)
        # ---- relative-score image: the reconstruction one quality level up (no gradient), or the real image at the top level
        if q + self.delta > self.model.rate_level - 1:
            rel = images
        else:
            rel = self.forward(images, q + self.delta, noise_rel, beta=None)["fake_images"]
            self.ctx.tape = None
        # ================================================================== train G
        ld, adv, fake = self.generator_backward(images, q, noise, None, rel, bounded=bounded)
        ld["aux"] = self.aux_step()
        self.all_reduce_grads()
        self.optimizer_step()               # honours the skip flag (nan / inf / huge loss: the reference returns before any update)
        ld = self._finish_losses(ld)
        ld["adv"] = adv.detach()
        # ================================================================== train D
        # (gradient buffers are keyed by the storage address of the activation they belong to: the generator step's
        # activations are released now, so its entries must go before new tensors can land on the same addresses)
        self._grads, self._keep = {}, []
        self.dctx.flat_g.zero_()
        fake_d, ftape, flogits, _ = self.d_forward(k, fake, tape=True)
        real_p, rtape, rlogits, _ = self.d_forward(k, images, tape=True)
        with torch.enable_grad():
            rp, fp = real_p.detach().requires_grad_(True), fake_d.detach().requires_grad_(True)
            l_d_real = 0.5 * self._bce(rp - fp.detach(), 1.0)
            l_d_fake = 0.5 * self._bce(fp - rp.detach(), 0.0)
            drp, dfp = torch.autograd.grad(l_d_real + l_d_fake, [rp, fp])
        bound = 0.5 / real_p.numel() if bounded else None
        self.d_backward(rtape, rlogits, drp, peak_bound=bound)
        self.d_backward(ftape, flogits, dfp, peak_bound=bound)
        from .sharding import allreduce_mean_flat
        allreduce_mean_flat(self.dctx.flat_g, 64 << 20, self.pg)
        self.d_optimizer_step(k)
        ld.update(d_real=l_d_real.detach(), d_fake=l_d_fake.detach(), out_d_real=real_p.mean(), out_d_fake=fake_d.mean())
        return ld

    def train_step(self, images, q=None, noise=None, generator=None, beta=None):
        """One generator + discriminator update.  Like CodecTrainer.train_step, the first call of a (shape, quality level)
        is enqueued eagerly (it also builds the adjoint matrices of the codec and of sub-discriminator int(q)), the second
        captures the whole step into ONE CUDA graph and from then on the step is a replay (single process; under
        torch.distributed the collectives keep the step eager)."""
        from .sharding import broadcast_from_rank0
        n, _, h, w = images.shape
        if q is None:
            qt = torch.randint(self.model.rate_level, (1,)).to(self.device, torch.float32)
            q = float(broadcast_from_rank0(qt, self.pg).item())
        if beta is None:
            bt = torch.randint(0, 101, (1,)).to(self.device, torch.float32) * (self.gs.max_beta / 100.0)
            self._beta.copy_(broadcast_from_rank0(bt, self.pg))
        else:
            self._beta.fill_(float(beta))
        mk = lambda c, a, b: torch.rand((n, c, a, b), dtype=torch.float32, device=self.device, generator=generator) - 0.5
        if noise is None:
            noise = {"z": mk(self.zc, h // 64, w // 64), "y": mk(self.charm.yc, h // 16, w // 16)}
        noise_rel = {"z": mk(self.zc, h // 64, w // 64), "y": mk(self.charm.yc, h // 16, w // 16)}
        if self.loss_scale is None:
            self.loss_scale = self._initial_scale(n, h, w)
        key = (n, h, w, float(q))
        if not self.use_gan_graphs or self._distributed() or key not in self._warm:
            ld = self._gan_core(images, q, noise, noise_rel, bounded=False)
            nv.status_check()
            self._warm.add(key)
            return ld
        hit = self._graphs.get(key)
        if hit is None:
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            st_in = dict(x=images.clone(), z=noise["z"].clone(), y=noise["y"].clone(), rz=noise_rel["z"].clone(), ry=noise_rel["y"].clone())
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            count, dcount, l0 = self.step_count, list(self.d_steps), nv.LAUNCH_COUNT[0]
            with torch.cuda.graph(g, pool=self._pool):
                ld = self._gan_core(st_in["x"], q, {"z": st_in["z"], "y": st_in["y"]}, {"z": st_in["rz"], "y": st_in["ry"]}, bounded=True)
            self.step_count, self.d_steps = count, dcount          # the capture itself executes nothing
            hit = self._graphs[key] = (g, st_in, ld, nv.LAUNCH_COUNT[0] - l0, int(q))
        g, st_in, ld, launches, k = hit
        for name, src in (("x", images), ("z", noise["z"]), ("y", noise["y"]), ("rz", noise_rel["z"]), ("ry", noise_rel["y"])):
            st_in[name].copy_(src, non_blocking=True)
        g.replay()
        self.step_count += 1
        self.d_steps[k] += 1
        nv.LAUNCH_COUNT[0] += launches
        nv.status_check()
        return dict(ld)

    def d_optimizer_step(self, k):
        """Adam on the active sub-discriminator's segment; step count, learning rate and the skip flag are device resident."""
        lo, hi = self._dseg[k]
        self.d_steps[k] += 1
        c = self.dctx
        b1, b2 = self.betas
        self._dstep_dev[k] += 1.0 - self._skip[0].to(torch.float64)
        t = self._dstep_dev[k]
        self._dhyper[0:1].fill_(self.d_lr)      # (a Python scalar assigned to an element would be a host copy: not capturable)
        self._dhyper[1] = (1.0 - b1 ** t).to(torch.float32)
        self._dhyper[2] = torch.sqrt(1.0 - b2 ** t).to(torch.float32)
        self._dhyper[3] = self._skip[0]
        nv.check(nv.lib().crdr_adam_step(c.flat_p[lo:].data_ptr(), c.flat_g[lo:].data_ptr(), self.dm[lo:].data_ptr(), self.dv[lo:].data_ptr(),
                                         hi - lo, self.d_lr, b1, b2, self.eps, 0, None, 1.0, self._dhyper.data_ptr(), nv.stream_handle()))
        c.repack()

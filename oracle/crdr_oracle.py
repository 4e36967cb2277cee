"""ORACLE -- test infrastructure only (never imported by the product package `crdr_b200`).

CPU restatement (plain PyTorch fp32 ops) of the reference's codec hot path as pure functions over a
checkpoint ``state_dict`` in the reference layout.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it.

Followed reference code (paths relative to /root/reference):
  g_a   src/models/subnet/autoencoder/elic_interpca_autoencoder.py:22-56, elic_autoencoder.py:42-72
  g_s   src/models/subnet/autoencoder/elic_interpca_beta_cond_autoencoder.py:42-162
  blocks src/models/layer/elic_layers.py:23-52, cheng_nlam.py:5-47,
        interp_channel_attention.py:39-73, fourier_cond.py:12-37
  h_a/h_s src/models/subnet/hyperprior/minnen20_hyperprior.py:9-58
  ChARM src/models/subnet/context_model/minnen20_charm_context_model.py:88-240
  entropy wrappers src/models/subnet/entropy_model/*.py  (arithmetic: CompressAI 1.2.4, restated in
        oracle/shims/compressai -- PARITY UNPINNED against the real package)
  model src/models/comp_model/beta_cond_interpca_hyperprior_charm_model.py:34-149, base_model.py:35-57,145-167,
        hyperprior_model.py:80-85,120-136;  header src/utils/codec_utils.py:81-125

Pinned here against the reference itself: tests/test_oracle_vs_reference.py imports the unmodified reference
modules from /root/reference (when present) and requires bit-equal tensors and byte-equal streams.
"""
import math
import os
import struct
import sys

import numpy as np
import torch
import torch.nn.functional as F

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)

from compressai.ans import RansDecoder  # noqa: E402  (oracle restatement, see oracle/shims)
from compressai.entropy_models import EntropyBottleneck, GaussianConditional  # noqa: E402
from compressai.models import get_scale_table  # noqa: E402

CFG = dict(rate_level=5, num_slices=10, max_support=5, slice_ch=32, max_beta=5.12, L=10, stride=64, zc=192, yc=320,
           scale_bound=0.11)


def sub(sd, prefix):
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def conv(sd, name, x, stride=1, pad=0):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=pad)


def deconv(sd, name, x, stride=2, pad=2, opad=1):
    return F.conv_transpose2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=pad, output_padding=opad)


def interp_gain(sd, i, x, q):
    if f"interp_ca_list.{i}.weight" not in sd:   # stage-1 transforms (ElicEncoder / ElicDecoder) have no InterpChAtt
        return x
    w, b = sd[f"interp_ca_list.{i}.weight"], sd[f"interp_ca_list.{i}.bias"]
    ind = torch.tensor([q], dtype=torch.float)
    l = torch.floor(ind)
    r = torch.minimum(l + 1.0, torch.tensor(w.shape[0] - 1))
    alpha = (r - ind).reshape(-1, 1, 1, 1, 1)
    wv = (w[l.long()] * alpha + w[r.long()] * (1 - alpha)).squeeze(1)
    bv = (b[l.long()] * alpha + b[r.long()] * (1 - alpha)).squeeze(1)
    return F.softplus(wv) * x + bv


def bottleneck(sd, x, names=("conv.0", "conv.2", "conv.4"), cond=None):
    t = torch.relu(conv(sd, names[0], x))
    if cond is not None:
        t = t + conv(sd, "proj_1", cond)
    t = torch.relu(conv(sd, names[1], t, pad=1))
    if cond is not None:
        t = t + conv(sd, "proj_2", cond)
    t = conv(sd, names[2], t)
    if cond is not None:
        t = t + conv(sd, "proj_3", cond)
    return t + x


def block_group(sd, x, cond=None):
    for i in range(3):
        x = bottleneck(sub(sd, f"block{i}"), x, cond=cond)
    return x


def nlam(sd, x):
    names = ("c1", "c2", "c3")
    t = x
    for i in range(3):
        t = bottleneck(sub(sd, f"trunk_block.{i}"), t, names)
    a = x
    for i in range(3):
        a = bottleneck(sub(sd, f"attention_block.{i}"), a, names)
    return x + t * torch.sigmoid(conv(sd, "conv", a))


def g_a(sd, x, q):
    e = sub(sd, "encoder")
    x = interp_gain(e, 0, conv(e, "conv1", x, 2, 2), q)
    x = interp_gain(e, 1, block_group(sub(e, "block1"), x), q)
    x = interp_gain(e, 2, conv(e, "conv2", x, 2, 2), q)
    x = interp_gain(e, 3, block_group(sub(e, "block2"), x), q)
    x = interp_gain(e, 4, nlam(sub(e, "attn2"), x), q)
    x = interp_gain(e, 5, conv(e, "conv3", x, 2, 2), q)
    x = interp_gain(e, 6, block_group(sub(e, "block3"), x), q)
    x = interp_gain(e, 7, conv(e, "conv4", x, 2, 2), q)
    return interp_gain(e, 8, nlam(sub(e, "attn4"), x), q)


def h_a(sd, y):
    h = sub(sd, "hyperencoder")
    t = torch.relu(conv(h, "conv1", y, 1, 1))
    t = torch.relu(conv(h, "conv2", t, 2, 2))
    return conv(h, "conv3", t, 2, 2)


def h_s(sd, z_hat):
    outs = []
    for br in ("hd_mu", "hd_std"):
        b = sub(sd, "hyperdecoder." + br)
        t = torch.relu(deconv(b, "conv1", z_hat))
        t = torch.relu(deconv(b, "conv2", t))
        outs.append(deconv(b, "conv3", t, stride=1, pad=1, opad=0))
    return torch.cat(outs, dim=1)


def beta_cond(sd, beta):
    d = sub(sd, "decoder")
    if "mlp.0.weight" not in d:                  # stage-1 / stage-2 decoders: no beta conditioning
        return None
    freq = torch.pow(torch.Tensor([2]), torch.arange(CFG["L"]))
    nb = (torch.Tensor([beta]).float() / CFG["max_beta"] - 0.5) * 2
    emb = torch.cat([torch.sin(nb * freq), torch.cos(nb * freq)], dim=0).unsqueeze(0)
    c = F.linear(torch.relu(F.linear(emb, d["mlp.0.weight"], d["mlp.0.bias"])), d["mlp.2.weight"], d["mlp.2.bias"])
    return c.unsqueeze(-1).unsqueeze(-1)


def g_s(sd, y_hat, q, beta):
    d = sub(sd, "decoder")
    c = beta_cond(sd, beta)
    x = nlam(sub(d, "attn1"), interp_gain(d, 0, y_hat, q))
    x = deconv(d, "conv1", interp_gain(d, 1, x, q))
    x = block_group(sub(d, "block1"), interp_gain(d, 2, x, q), c)
    x = deconv(d, "conv2", interp_gain(d, 3, x, q))
    x = nlam(sub(d, "attn2"), interp_gain(d, 4, x, q))
    x = block_group(sub(d, "block2"), interp_gain(d, 5, x, q), c)
    x = deconv(d, "conv3", interp_gain(d, 6, x, q))
    x = block_group(sub(d, "block3"), interp_gain(d, 7, x, q), c)
    return deconv(d, "conv4", interp_gain(d, 8, x, q))


def slice_net(sd, x):
    t = torch.relu(conv(sd, "model.0", x, 1, 2))
    t = torch.relu(conv(sd, "model.2", t, 1, 2))
    return conv(sd, "model.4", t, 1, 1)


def entropy_models(sd):
    """CompressAI-restatement objects loaded with the checkpoint's entropy parameters and fresh tables."""
    eb = EntropyBottleneck(CFG["zc"])
    own = eb.state_dict()
    for k in own:
        key = "entropy_model_z." + k
        if key in sd and not k.startswith("_quantized") and k not in ("_offset", "_cdf_length"):
            own[k] = sd[key].detach().clone().float()
    eb.load_state_dict(own)
    eb.update(force=True)
    gc = GaussianConditional(None, scale_bound=CFG["scale_bound"])
    gc.update_scale_table(get_scale_table(), force=True)
    return eb, gc


def gc_eval(gc, y_slice, mu, sigma):
    y_hat = gc.quantize(y_slice, "dequantize", mu)
    _, lik = gc(y_slice, sigma, means=mu, training=False)
    return y_hat, lik


def charm(sd, y, hyper_out, gc, mode="forward", y_string=None):
    """mode 'forward' -> (y_hat, likelihood, mu, sigma); 'decompress' -> (y_hat, symbols)."""
    cm = sub(sd, "context_model")
    S, K = CFG["num_slices"], CFG["max_support"]
    hyper_mean, hyper_scale = torch.chunk(hyper_out, 2, dim=1)
    y_slices = torch.chunk(y, S, dim=1) if y is not None else [None] * S
    if mode == "decompress":
        cdf, lens, offs = gc._quantized_cdf.tolist(), gc._cdf_length.tolist(), gc._offset.tolist()
        dec = RansDecoder()
        dec.set_stream(y_string)
    hats, liks, mus, sigmas, syms = [], [], [], [], []
    for s in range(S):
        support = hats[:K]
        mean_support = torch.cat([hyper_mean] + support, dim=1)
        scale_support = torch.cat([hyper_scale] + support, dim=1)
        mu = slice_net(sub(cm, f"mean_slice_transforms.{s}"), mean_support)
        sigma = slice_net(sub(cm, f"scale_slice_transforms.{s}"), scale_support)
        if mode == "decompress":
            idx = gc.build_indexes(sigma)
            vals = dec.decode_stream(idx.reshape(-1).int().tolist(), cdf, lens, offs)
            sym = torch.Tensor(vals).reshape(sigma.size())
            y_hat_s = gc.dequantize(sym, mu)
            syms.append(sym)
        else:
            y_hat_s, lik = gc_eval(gc, y_slices[s], mu, sigma)
            liks.append(lik)
        lrp = slice_net(sub(cm, f"lrp_slice_transforms.{s}"), torch.cat([mean_support, y_hat_s], dim=1))
        hats.append(y_hat_s + 0.5 * torch.tanh(lrp))
        mus.append(mu)
        sigmas.append(sigma)
    y_hat = torch.cat(hats, dim=1)
    if mode == "decompress":
        return y_hat, torch.cat(syms, dim=1).int()
    return y_hat, torch.cat(liks, dim=1), torch.cat(mus, dim=1), torch.cat(sigmas, dim=1)


def pad_image(x, stride=64):
    _, _, h, w = x.shape
    ph, pw = int(np.ceil(h / stride) * stride - h), int(np.ceil(w / stride) * stride - w)
    return x if ph == 0 and pw == 0 else F.pad(x, (0, pw, 0, ph), mode="reflect")


def bits_of(lik):
    return float(-(torch.log(lik).sum()) / np.log(2))


@torch.no_grad()
def analysis(sd, x, q, eb, gc):
    """Everything the encoder computes before range coding (image [1,3,H,W] in [-1,1])."""
    xp = pad_image(x.float())
    y = g_a(sd, xp, q)
    z = h_a(sd, y)
    z_hat, z_lik = eb(z, training=False)
    z_sym = eb.quantize(z, "symbols", eb._get_medians().reshape(1, -1, 1, 1))
    hyper = h_s(sd, z_hat)
    y_hat, y_lik, mu, sigma = charm(sd, y, hyper, gc)
    y_idx = gc.build_indexes(sigma)
    y_sym = gc.quantize(y, "symbols", mu)
    return dict(y=y, z=z, z_hat=z_hat, z_lik=z_lik, z_sym=z_sym, hyper=hyper, y_hat=y_hat, y_lik=y_lik, mu=mu,
                sigma=sigma, y_idx=y_idx, y_sym=y_sym)


@torch.no_grad()
def compress(sd, x, q, eb=None, gc=None):
    if eb is None:
        eb, gc = entropy_models(sd)
    _, _, h, w = x.shape
    a = analysis(sd, x, q, eb, gc)
    z_str = eb.compress(a["z"])[0]
    y_str = gc.compress(a["y"], a["y_idx"], means=a["mu"])[0]
    if q is None:   # single-rate models: HeaderHandler (codec_utils.py:22-39)
        header = struct.pack("<HHB", h, w, int(torch.max(torch.abs(a["y_hat"]))))
    else:
        header = struct.pack("<HHBB", h, w, int(torch.max(torch.abs(a["y_hat"]))), int(float(q) * 16))
    a.update(string_list=[header, z_str, y_str], pred_y_bit=bits_of(a["y_lik"]), pred_z_bit=bits_of(a["z_lik"]))
    a["pred_y_bpp"], a["pred_z_bpp"] = a["pred_y_bit"] / (h * w), a["pred_z_bit"] / (h * w)
    return a


@torch.no_grad()
def decompress(sd, string_list, beta, eb=None, gc=None):
    if eb is None:
        eb, gc = entropy_models(sd)
    if len(string_list[0]) == 5:
        (h, w, _), q = struct.unpack("<HHB", string_list[0][:5]), None
    else:
        h, w, _, q16 = struct.unpack("<HHBB", string_list[0][:6])
        q = q16 / 16
    s = CFG["stride"]
    hz, wz = int(np.ceil(h / s)), int(np.ceil(w / s))
    z_hat = eb.decompress([string_list[1]], (hz, wz))
    hyper = h_s(sd, z_hat)
    y_hat, y_sym = charm(sd, None, hyper, gc, mode="decompress", y_string=string_list[2])
    img = g_s(sd, y_hat, q, beta)[:, :, :h, :w].clamp(-1, 1)
    return img, z_hat, y_hat, y_sym


def eb_likelihood_at(eb, v):
    """EntropyBottleneck likelihood (lower-bounded) at arbitrary points v [N, C, H, W] (CompressAI EntropyBottleneck.forward
    after its quantize step: permute to C x 1 x (N H W), _likelihood, likelihood_lower_bound, permute back)."""
    n, c, h, w = v.shape
    vals = v.permute(1, 0, 2, 3).contiguous().reshape(c, 1, -1)
    lik = eb.likelihood_lower_bound(eb._likelihood(vals))
    return lik.reshape(c, n, h, w).permute(1, 0, 2, 3).contiguous()


def ste_round(x):
    """src/models/subnet/entropy_model/ste_round.py:4-5 (straight-through: the rounding residual carries no gradient)."""
    return (torch.round(x) - x).detach() + x


@torch.no_grad()
def forward_train(sd, x, q, beta, noise, eb, gc, forced_y_symbols=None):
    """Forward VALUES of model.forward(..., is_train=True) (beta_cond_interpca_hyperprior_charm_model.py:34-78;
    minnen20_charm_context_model.py:88-141; ste_gaussian_conditional.py:20-27; entropy_bottleneck.py:23-30) with the
    uniform noise given explicitly: noise = {"z": [N, zc, h/64, w/64], "y": [N, yc, h/16, w/16]} in [-1/2, 1/2).
    The reference draws the same values inside CompressAI's quantize(mode="noise"): z first (in C x 1 x (N H W) order),
    then one draw per slice.
    forced_y_symbols (tests only): integer rounding decisions round(y - mu) [N, yc, h/16, w/16] to use instead of this
    function's own torch.round -- an fp32-rounding-level difference in mu can flip a tie, and a flipped symbol changes
    every later slice; forcing the decisions lets a gradient comparison isolate the backward arithmetic."""
    S, K = CFG["num_slices"], CFG["max_support"]
    y = g_a(sd, x, q)
    z = h_a(sd, y)
    med = eb._get_medians().reshape(1, -1, 1, 1)
    z_hat = ste_round(z - med) + med
    z_lik = eb_likelihood_at(eb, z + noise["z"])
    hyper = h_s(sd, z_hat)
    cm = sub(sd, "context_model")
    hyper_mean, hyper_scale = torch.chunk(hyper, 2, dim=1)
    y_slices, n_slices = torch.chunk(y, S, dim=1), torch.chunk(noise["y"], S, dim=1)
    hats, liks, qliks = [], [], []
    for s in range(S):
        support = hats[:K]
        mean_support = torch.cat([hyper_mean] + support, dim=1)
        scale_support = torch.cat([hyper_scale] + support, dim=1)
        mu = slice_net(sub(cm, f"mean_slice_transforms.{s}"), mean_support)
        sigma = slice_net(sub(cm, f"scale_slice_transforms.{s}"), scale_support)
        liks.append(gc.likelihood_lower_bound(gc._likelihood(y_slices[s] + n_slices[s], sigma, mu)))
        if forced_y_symbols is None:
            y_hat_s = ste_round(y_slices[s] - mu) + mu
        else:
            r = y_slices[s] - mu
            y_hat_s = (forced_y_symbols[:, s * CFG["slice_ch"]:(s + 1) * CFG["slice_ch"]].to(r.dtype) - r).detach() + r + mu
        qliks.append(gc(y_slices[s], sigma, means=mu, training=False)[1])
        lrp = slice_net(sub(cm, f"lrp_slice_transforms.{s}"), torch.cat([mean_support, y_hat_s], dim=1))
        hats.append(y_hat_s + 0.5 * torch.tanh(lrp))
    y_hat = torch.cat(hats, dim=1)
    fake = g_s(sd, y_hat, q, beta)
    _, z_qlik = eb(z, training=False)
    return {"fake_images": fake, "likelihoods": {"y": torch.cat(liks, dim=1), "z": z_lik}, "latent_code": {"y": y, "z": z},
            "quantized_code": {"y": y_hat, "z": z_hat}, "q_likelihoods": {"y": torch.cat(qliks, dim=1), "z": z_qlik}}


def discriminator(sd, x, slope=0.2):
    """CLIC21GVAEDiscriminator.forward with norm_type none (src/models/discriminator/clic21_gvae_discriminator.py:27-50):
    sd = state dict of ONE sub-discriminator (keys model.{0,2,...}.weight / .bias); strides 1, 2, 1, 2, ..., head 1."""
    idx = sorted(int(k.split(".")[1]) for k in sd if k.endswith(".weight"))
    for n, i in enumerate(idx):
        last = n == len(idx) - 1
        x = F.conv2d(x, sd[f"model.{i}.weight"], sd[f"model.{i}.bias"], stride=1 if (last or n % 2 == 0) else 2, padding=1)
        if not last:
            x = F.leaky_relu(x, slope)
    return x


def to_uint8(img):
    """img_utils.torch2npimg truncation semantics (img_utils.py:30-42)."""
    return ((img + 1.0) / 2.0 * 255.0).numpy().astype(np.uint8)


def psnr_u8(real, fake):
    r, f = to_uint8(real).astype(np.float32), to_uint8(fake).astype(np.float32)
    return 10.0 * math.log10(255.0 ** 2 / float(np.mean((r - f) ** 2)))

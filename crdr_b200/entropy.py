"""Entropy models of the codec path: parameter / table holders with the reference checkpoint layout.

These replace the CompressAI 1.2.4 classes the reference subclasses
(src/models/subnet/entropy_model/entropy_bottleneck.py:12-30, gaussian_conditional.py:17-24,
ste_gaussian_conditional.py:9-27).  The per-element arithmetic (quantise, likelihood, CDF index) runs in
``libcrdr_sm100.so``; the sequential range coder in ``libcrdr_rans.so``.  What stays here is the one-off
table construction (`update`) and the packing of the factorised-prior parameters for the kernel.
"""
import math

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import rans
from .registry import ENTROPYMODEL_REGISTRY


class _Bound(nn.Module):
    """Holds the `bound` buffer CompressAI's LowerBound registers (checkpoint compatibility)."""

    def __init__(self, value):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(value)]))


class _CoderTables(nn.Module):
    likelihood_floor = 1e-9
    coder_precision = 16

    def __init__(self):
        super().__init__()
        self.likelihood_lower_bound = _Bound(self.likelihood_floor)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._tables = None

    def _store_tables(self, pmf, tail_mass, pmf_length, max_length):
        """Quantise every row of `pmf` (+ its tail mass) to a 16-bit CDF."""
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        pmf, tail_mass = pmf.detach().float().cpu(), tail_mass.detach().float().cpu()
        for i in range(len(pmf_length)):
            row = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i].reshape(1)))
            q = rans.pmf_to_quantized_cdf(row.numpy(), self.coder_precision)
            cdf[i, : q.size] = torch.from_numpy(q)
        return cdf

    def coder_tables(self):
        """rans.Tables view of the current buffers (cached until the next update / load)."""
        if self._tables is None:
            if self._quantized_cdf.numel() == 0:
                raise RuntimeError("entropy model tables are empty: call update()/codec_setup() first")
            self._tables = rans.Tables(self._quantized_cdf.cpu().numpy(), self._cdf_length.cpu().numpy(),
                                       self._offset.cpu().numpy())
        return self._tables

    def invalidate(self):
        self._tables = None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # an in-place load changes the CDF buffers under a cached rans.Tables: rebuild it on next use
        self.invalidate()
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _resize_buffers_from(self, state_dict, prefix, names):
        # checkpoints carry filled tables while a fresh module has empty ones (base_model.py:80-96)
        for name in names:
            key = prefix + name
            if key in state_dict:
                buf = getattr(self, name)
                if buf.numel() == 0:
                    buf.resize_(state_dict[key].size())
        self.invalidate()


@ENTROPYMODEL_REGISTRY.register()
class EntropyBottleneck(_CoderTables):
    """Factorised prior over z (CompressAI EntropyBottleneck semantics, filters (3,3,3,3), init_scale 10)."""

    def __init__(self, channels, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **_):
        super().__init__()
        self.channels, self.filters = int(channels), tuple(int(f) for f in filters)
        self.init_scale, self.tail_mass = float(init_scale), float(tail_mass)
        widths = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / widths[i + 1]))
            m = torch.Tensor(self.channels, widths[i + 1], widths[i])
            m.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(m))
            b = torch.Tensor(self.channels, widths[i + 1], 1)
            nn.init.uniform_(b, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(b))
            if i < len(self.filters):
                f = torch.Tensor(self.channels, widths[i + 1], 1)
                nn.init.zeros_(f)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(f))
        self.quantiles = nn.Parameter(torch.Tensor(self.channels, 1, 3))
        self.quantiles.data = torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(self.channels, 1, 1)
        t = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-t, 0, t]))

    def medians(self):
        return self.quantiles[:, 0, 1].detach()

    def _logits(self, x, detach=True, on_cpu=False):
        """Cumulative logits of the per-channel 1-3-3-3-3-1 network; x: [C, 1, K]."""
        def fetch(name):
            p = getattr(self, name)
            p = p.detach() if detach else p
            return p.float().cpu() if on_cpu else p
        h = x
        for i in range(len(self.filters) + 1):
            h = torch.matmul(F.softplus(fetch(f"_matrix{i:d}")), h) + fetch(f"_bias{i:d}")
            if i < len(self.filters):
                h = h + torch.tanh(fetch(f"_factor{i:d}")) * torch.tanh(h)
        return h

    def loss(self):
        """Auxiliary quantile loss (BaseModel.aux_loss, base_model.py:65-76)."""
        return torch.abs(self._logits(self.quantiles, detach=True) - self.target).sum()

    @torch.no_grad()
    def update(self, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        q = self.quantiles.detach().float().cpu()
        med = q[:, 0, 1]
        minima = torch.clamp(torch.ceil(med - q[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(q[:, 0, 2] - med).int(), min=0)
        pmf_start = med - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
        lower = self._logits(samples - 0.5, on_cpu=True)
        upper = self._logits(samples + 0.5, on_cpu=True)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        dev = self._offset.device
        self._quantized_cdf = self._store_tables(pmf, tail, pmf_length, max_length).to(dev)
        self._offset = (-minima).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)
        self.invalidate()
        return True

    @torch.no_grad()
    def kernel_params(self, device):
        """[C, 58] fp32: softplus(matrix) / bias / tanh(factor) per layer, as crdr_eb_desc documents."""
        cols = []
        for i in range(len(self.filters) + 1):
            m = F.softplus(getattr(self, f"_matrix{i:d}").detach().float())
            cols.append(m.reshape(self.channels, -1))
            cols.append(getattr(self, f"_bias{i:d}").detach().float().reshape(self.channels, -1))
            if i < len(self.filters):
                cols.append(torch.tanh(getattr(self, f"_factor{i:d}").detach().float()).reshape(self.channels, -1))
        p = torch.cat(cols, dim=1)
        assert p.shape[1] == 58 and self.filters == (3, 3, 3, 3)
        return p.to(device).contiguous(), self.medians().float().to(device).contiguous()


@ENTROPYMODEL_REGISTRY.register()
class SteEntropyBottleneck(EntropyBottleneck):
    """Reference name (entropy_bottleneck.py:18-30); evaluation behaviour is identical.  The noise likelihood / STE code of
    training mode are values of crdr_eb_quantize with a noise tensor; their gradients are taken in train.CodecTrainer."""


@ENTROPYMODEL_REGISTRY.register()
class GaussianMeanScaleConditional(_CoderTables):
    """Conditional Gaussian over y with scale lower bound and a 64-level CDF table
    (CompressAI GaussianConditional(scale_table=None, scale_bound=...))."""

    def __init__(self, scale_bound=None, tail_mass=1e-9, **_):
        super().__init__()
        if scale_bound is None or scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.tail_mass = float(tail_mass)
        self.lower_bound_scale = _Bound(scale_bound)
        self.register_buffer("scale_table", torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @torch.no_grad()
    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.scale_table.device
        self.scale_table = torch.Tensor(tuple(float(s) for s in scale_table)).to(dev)
        self.update()
        return True

    @staticmethod
    def _phi(x):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * x)

    @torch.no_grad()
    def update(self):
        table = self.scale_table.detach().float().cpu()
        mult = -scipy.stats.norm.ppf(self.tail_mass / 2)
        center = torch.ceil(table * mult).int()
        length = 2 * center + 1
        max_length = int(torch.max(length).item())
        k = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
        s = table.unsqueeze(1)
        upper, lower = self._phi((0.5 - k) / s), self._phi((-0.5 - k) / s)
        dev = self._offset.device
        self._quantized_cdf = self._store_tables(upper - lower, 2 * lower[:, :1], length, max_length).to(dev)
        self._offset = (-center).to(dev)
        self._cdf_length = (length + 2).to(dev)
        self.invalidate()


@ENTROPYMODEL_REGISTRY.register()
class SteGaussianMeanScaleConditional(GaussianMeanScaleConditional):
    """Reference name (ste_gaussian_conditional.py:9-27)."""

    def __init__(self, scale_bound=None, entropy_quant_type="noise", **kw):
        super().__init__(scale_bound=scale_bound, **kw)
        assert entropy_quant_type == "noise"
        self.entropy_quant_type = entropy_quant_type


def get_scale_table(lo=0.11, hi=256, levels=64):
    """compressai.models.get_scale_table (called at hyperprior_model.py:123)."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))

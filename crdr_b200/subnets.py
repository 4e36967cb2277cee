"""Parameter-holding sub-networks under the reference's registry names.

Each class reproduces the *parameter tree* of its reference counterpart (same attribute names, shapes and
default initialisers, so ``state_dict()`` keys / checkpoints are interchangeable -- layout pinned by
tests/golden/state_dict_layout_crdr.json) but carries no PyTorch forward: ``lower()`` compiles the current
weights into a kernel-launch engine from ``crdr_b200.codec``.  Constructor kwargs are exactly the yaml keys of
config/_base_/model/beta_cond_interp_ca_elic_charm.yaml.
"""
import numpy as np
import torch
import torch.nn as nn

from . import codec
from .registry import (CONTEXTMODEL_REGISTRY, DECODER_REGISTRY, ENCODER_REGISTRY, HYPERDECODER_REGISTRY,
                       HYPERENCODER_REGISTRY)


def _conv(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2)


def _deconv(cin, cout, k=5, stride=2):
    return nn.ConvTranspose2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, output_padding=stride - 1)


class _Holder(nn.Module):
    """Namespace of sub-modules; never called."""

    def __init__(self, **children):
        super().__init__()
        for name, child in children.items():
            self.add_module(name, child)

    def forward(self, *a, **k):
        raise RuntimeError("parameter holder: the forward path is the lowered CUDA engine (model.engine())")


def _indexed(*modules_at):
    """Children named by integer position (matches the reference's nn.Sequential key names)."""
    return _Holder(**{str(i): m for i, m in modules_at})


def _bottleneck(ch, mid, cond_ch=None):
    kids = dict(conv=_indexed((0, _conv(ch, mid, 1)), (2, _conv(mid, mid, 3)), (4, _conv(mid, ch, 1))))
    if cond_ch:
        kids.update(proj_1=_conv(cond_ch, mid, 1), proj_2=_conv(cond_ch, mid, 1), proj_3=_conv(cond_ch, ch, 1))
    return _Holder(**kids)


def _block_group(ch, mid, num_blocks, cond_ch=None):
    return _Holder(**{f"block{i}": _bottleneck(ch, mid, cond_ch) for i in range(num_blocks)})


def _nlam(ch):
    res = lambda: _Holder(c1=_conv(ch, ch // 2, 1), c2=_conv(ch // 2, ch // 2, 3), c3=_conv(ch // 2, ch, 1))
    return _Holder(trunk_block=_indexed(*[(i, res()) for i in range(3)]),
                   attention_block=_indexed(*[(i, res()) for i in range(3)]),
                   conv=_conv(ch, ch, 1))


class _Gain(nn.Module):
    """InterpChAtt parameters (interp_channel_attention.py:17-37)."""

    def __init__(self, ch, rate_level, actv="identity", use_interp=False, use_bias=False):
        super().__init__()
        if not (actv == "softplus" and use_interp and use_bias):
            raise NotImplementedError("only the shipped InterpChAtt setting (softplus, interp, bias) is lowered")
        self.weight = nn.Parameter(torch.ones(rate_level, 1, ch, 1, 1) * float(np.log(np.e - 1)))
        self.bias = nn.Parameter(torch.zeros(rate_level, 1, ch, 1, 1))


def _gain_list(channels, rate_level, ca_kwargs):
    return nn.ModuleList([_Gain(c, rate_level, **ca_kwargs) for c in channels])


@ENCODER_REGISTRY.register()
class ElicEncoder(nn.Module):
    """Stage-1 analysis transform without InterpChAtt (elic_autoencoder.py:33-72; config/_base_/model/elic_charm.yaml)."""

    def __init__(self, in_ch=3, out_ch=192, main_ch=192, block_mid_ch=192, num_blocks=3, res_in_res=False):
        super().__init__()
        if in_ch != 3 or num_blocks != 3 or res_in_res:
            raise NotImplementedError("lowered analysis transform expects RGB input, 3 blocks per stage, no res_in_res")
        m = main_ch
        self.conv1 = _conv(in_ch, m, 5, 2)
        self.block1 = _block_group(m, block_mid_ch, num_blocks)
        self.conv2 = _conv(m, m, 5, 2)
        self.block2 = _block_group(m, block_mid_ch, num_blocks)
        self.attn2 = _nlam(m)
        self.conv3 = _conv(m, m, 5, 2)
        self.block3 = _block_group(m, block_mid_ch, num_blocks)
        self.conv4 = _conv(m, out_ch, 5, 2)
        self.attn4 = _nlam(out_ch)
        self.num_downscale, self.latent_ch = 4, out_ch

    def lower(self, device, sd=None, **kw):
        return codec.AnalysisEngine((sd if sd is not None else dict(self.state_dict())), device, **kw)


@ENCODER_REGISTRY.register()
class ElicInterpCaEncoder(nn.Module):
    def __init__(self, rate_level, in_ch=3, out_ch=192, main_ch=192, block_mid_ch=192, num_blocks=3, ca_kwargs={}):
        super().__init__()
        if in_ch != 3 or num_blocks != 3:
            raise NotImplementedError("lowered analysis transform expects RGB input and 3 blocks per stage")
        m = main_ch
        self.conv1 = _conv(in_ch, m, 5, 2)
        self.block1 = _block_group(m, block_mid_ch, num_blocks)
        self.conv2 = _conv(m, m, 5, 2)
        self.block2 = _block_group(m, block_mid_ch, num_blocks)
        self.attn2 = _nlam(m)
        self.conv3 = _conv(m, m, 5, 2)
        self.block3 = _block_group(m, block_mid_ch, num_blocks)
        self.conv4 = _conv(m, out_ch, 5, 2)
        self.attn4 = _nlam(out_ch)
        self.interp_ca_list = _gain_list([m] * 7 + [out_ch] * 2, rate_level, ca_kwargs)
        self.num_downscale, self.latent_ch, self.rate_level = 4, out_ch, rate_level

    def lower(self, device, sd=None, **kw):
        return codec.AnalysisEngine((sd if sd is not None else dict(self.state_dict())), device, **kw)


def _n002_init(module):
    # decoder `weight_init: True` (elic_interpca_beta_cond_autoencoder.py:30-40,147-148)
    if isinstance(module, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
        module.weight.data.normal_(0.0, 0.02)
        module.bias.data.fill_(0)


@DECODER_REGISTRY.register()
class ElicDecoder(nn.Module):
    """Stage-1 synthesis transform (elic_autoencoder.py:75-119)."""

    def __init__(self, in_ch=192, out_ch=3, main_ch=192, block_mid_ch=192, num_blocks=3, use_tanh=True,
                 pixel_shuffle=False, res_in_res=False):
        super().__init__()
        if pixel_shuffle or res_in_res or num_blocks != 3 or out_ch != 3:
            raise NotImplementedError("only the shipped decoder variant (ConvTranspose up-sampling) is lowered")
        m = main_ch
        self.attn1 = _nlam(in_ch)
        self.conv1 = _deconv(in_ch, m)
        self.block1 = _block_group(m, block_mid_ch, num_blocks)
        self.conv2 = _deconv(m, m)
        self.attn2 = _nlam(m)
        self.block2 = _block_group(m, block_mid_ch, num_blocks)
        self.conv3 = _deconv(m, m)
        self.block3 = _block_group(m, block_mid_ch, num_blocks)
        self.conv4 = _deconv(m, out_ch)
        self.use_tanh = use_tanh

    def lower(self, device, sd=None, **kw):
        return codec.SynthesisEngine((sd if sd is not None else dict(self.state_dict())), use_tanh=self.use_tanh, device=device, **kw)


@DECODER_REGISTRY.register()
class ElicInterpCaDecoder(ElicDecoder):
    """Stage-2 synthesis transform: ElicDecoder with an InterpChAtt in front of every layer
    (elic_interpca_autoencoder.py:59-97)."""

    def __init__(self, rate_level, in_ch=192, out_ch=3, main_ch=192, block_mid_ch=192, num_blocks=3, use_tanh=True,
                 pixel_shuffle=False, ca_kwargs={}):
        super().__init__(in_ch=in_ch, out_ch=out_ch, main_ch=main_ch, block_mid_ch=block_mid_ch, num_blocks=num_blocks,
                         use_tanh=use_tanh, pixel_shuffle=pixel_shuffle)
        self.interp_ca_list = _gain_list([in_ch] * 2 + [main_ch] * 7, rate_level, ca_kwargs)
        self.rate_level = rate_level


@DECODER_REGISTRY.register()
class ElicInterpCaBetaCondDecoder(nn.Module):
    def __init__(self, rate_level, L=10, max_beta=5.12, cond_ch=512, use_pi=True, include_x=False,
                 weight_init=False, in_ch=192, out_ch=3, main_ch=192, block_mid_ch=192, num_blocks=3, use_tanh=True,
                 pixel_shuffle=False, res_in_res=False, ca_kwargs={}):
        super().__init__()
        if pixel_shuffle or res_in_res or num_blocks != 3 or out_ch != 3:
            raise NotImplementedError("only the shipped decoder variant (ConvTranspose up-sampling) is lowered")
        m = main_ch
        self.attn1 = _nlam(in_ch)
        self.conv1 = _deconv(in_ch, m)
        self.block1 = _block_group(m, block_mid_ch, num_blocks, cond_ch)
        self.conv2 = _deconv(m, m)
        self.attn2 = _nlam(m)
        self.block2 = _block_group(m, block_mid_ch, num_blocks, cond_ch)
        self.conv3 = _deconv(m, m)
        self.block3 = _block_group(m, block_mid_ch, num_blocks, cond_ch)
        self.conv4 = _deconv(m, out_ch)
        self.interp_ca_list = _gain_list([in_ch] * 2 + [m] * 7, rate_level, ca_kwargs)
        enc_ch = 2 * L + 1 if include_x else 2 * L
        self.mlp = _indexed((0, nn.Linear(enc_ch, cond_ch)), (2, nn.Linear(cond_ch, cond_ch)))
        if weight_init:
            self.apply(_n002_init)
        self.hparams = dict(max_beta=max_beta, L=L, use_pi=use_pi, include_x=include_x, use_tanh=use_tanh)
        self.max_beta, self.rate_level = max_beta, rate_level

    def lower(self, device, sd=None, **kw):
        return codec.SynthesisEngine((sd if sd is not None else dict(self.state_dict())), device=device, **self.hparams, **kw)


@HYPERENCODER_REGISTRY.register()
class Minnen20HyperEncoder(nn.Module):
    def __init__(self, bottleneck_y=320, bottleneck_z=192):
        super().__init__()
        self.conv1 = _conv(bottleneck_y, 320, 3)
        self.conv2 = _conv(320, 256, 5, 2)
        self.conv3 = _conv(256, bottleneck_z, 5, 2)
        self.num_downscale, self.latent_ch = 2, bottleneck_z

    def lower(self, device, sd=None, **kw):
        return codec.HyperAnalysisEngine((sd if sd is not None else dict(self.state_dict())), device, **kw)


@HYPERDECODER_REGISTRY.register()
class Minnen20HyperDecoder(nn.Module):
    def __init__(self, bottleneck_z=192, hyper_out_ch=640):
        super().__init__()
        assert hyper_out_ch % 2 == 0
        branch = lambda: _Holder(conv1=_deconv(bottleneck_z, 192), conv2=_deconv(192, 256),
                                 conv3=_deconv(256, hyper_out_ch // 2, k=3, stride=1))
        self.hd_mu, self.hd_std = branch(), branch()
        self.out_ch = hyper_out_ch

    def lower(self, device, sd=None, **kw):
        return codec.HyperSynthesisEngine((sd if sd is not None else dict(self.state_dict())), device, **kw)


def _slice_net(cin, cout):
    return _Holder(model=_indexed((0, _conv(cin, 224, 5)), (2, _conv(224, 128, 5)), (4, _conv(128, cout, 3))))


@CONTEXTMODEL_REGISTRY.register()
class Minnen20CharmContextModel(nn.Module):
    def __init__(self, num_slices, bottleneck_y, hyper_out_ch, max_support_slices=5, slice_transform_kwargs={},
                 crop_gaussian_params=False):
        super().__init__()
        assert bottleneck_y % num_slices == 0, \
            f"bottleneck_y % num_slices must be 0, but got {bottleneck_y} and {num_slices}"
        assert max_support_slices == -1 or 1 <= max_support_slices <= num_slices
        if slice_transform_kwargs or crop_gaussian_params:
            raise NotImplementedError("non-default SliceTransform options are not lowered")
        sc, hc = bottleneck_y // num_slices, hyper_out_ch // 2
        self.slice_ch, self.num_slices, self.max_support_slices, self.hyper_ch = sc, num_slices, max_support_slices, hc
        self.mean_slice_transforms = nn.ModuleList()
        self.scale_slice_transforms = nn.ModuleList()
        self.lrp_slice_transforms = nn.ModuleList()
        for s in range(num_slices):
            sup = sc * (s if max_support_slices == -1 else min(s, max_support_slices))
            self.mean_slice_transforms.append(_slice_net(sup + hc, sc))
            self.scale_slice_transforms.append(_slice_net(sup + hc, sc))
            self.lrp_slice_transforms.append(_slice_net(sup + hc + sc, sc))

    def lower(self, device, sd=None, **kw):
        return codec.CharmEngine((sd if sd is not None else dict(self.state_dict())), self.num_slices, self.slice_ch, self.hyper_ch,
                                 self.max_support_slices, device, **kw)

"""Discriminators of the stage-3 (GAN) training under the reference's registry names: parameter trees with the reference's
state_dict keys (clic21_gvae_discriminator.py:12-50, module_list_discriminator.py:13-29); the arithmetic runs in the
lowered engine (``codec.DiscriminatorEngine``).  Only ``norm_type: none`` (what config/crdr_stage_3.yaml uses) is lowered."""
import torch.nn as nn

from . import codec
from .registry import DISCRIMINATOR_REGISTRY
from .subnets import _indexed


@DISCRIMINATOR_REGISTRY.register()
class CLIC21GVAEDiscriminator(nn.Module):
    """conv3x3(s1) LReLU, conv3x3(s2) LReLU, then (conv s1, conv s2) x (num_downscale - 1) with channel doubling up to
    8 * main_ch, and a 3x3 head: `model.{0,2,...}` are the convolutions, the odd indices the LeakyReLU(0.2) layers."""

    def __init__(self, in_ch=3, out_ch=1, main_ch=64, norm_type="BN", num_downscale=4):
        super().__init__()
        if norm_type != "none":
            raise NotImplementedError("only norm_type: none is lowered (config/crdr_stage_3.yaml)")
        if in_ch != 3 or out_ch != 1:
            raise NotImplementedError("the lowered discriminator scores RGB images with one logit map")
        convs, strides = [(in_ch, main_ch), (main_ch, main_ch)], [1, 2]
        c = main_ch
        for _ in range(num_downscale - 1):
            o = min(c * 2, main_ch * 8)
            convs += [(c, o), (o, o)]
            strides += [1, 2]
            c = o
        convs.append((c, out_ch))
        strides.append(1)
        self.strides = strides
        self.model = _indexed(*[(2 * i, nn.Conv2d(ci, co, kernel_size=3, stride=s, padding=1))
                                for i, ((ci, co), s) in enumerate(zip(convs, strides))])

    def lower(self, device, sd=None, **kw):
        return codec.DiscriminatorEngine(sd if sd is not None else dict(self.state_dict()), self.strides, device, **kw)


@DISCRIMINATOR_REGISTRY.register()
class ModuleListDiscriminator(nn.Module):
    """One sub-discriminator per quality level; a call uses sub-discriminator int(rate_ind) only
    (module_list_discriminator.py:25-29)."""

    def __init__(self, _subd_type, _num_subd, **kwargs):
        super().__init__()
        self.subD_list = nn.ModuleList(DISCRIMINATOR_REGISTRY.get(_subd_type)(**kwargs) for _ in range(_num_subd))


def build_discriminator(discriminator_opt):
    """models/discriminator/__init__.py:15-30."""
    kw = dict(discriminator_opt)
    return DISCRIMINATOR_REGISTRY.get(kw.pop("type"))(**kw)

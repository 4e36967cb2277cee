"""Seeded checkpoints and images shared by the tests, smoke() and bench.py (no network, no files)."""
import math
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def crdr_opt(device="cuda:0", config="crdr.yaml"):
    from crdr_b200.config import BaseConfig
    return BaseConfig.fromfile(os.path.join(ROOT, "config", config), device=device, is_train=False)


def calibrate_(sd, seed=0):
    """Random-init weights give a degenerate entropy path (|y| << 1, every sigma below the 0.11 bound).
    Spread the predicted scales log-uniformly over the table and raise the latent gain so symbols are non-trivial."""
    g = torch.Generator().manual_seed(seed + 99)
    for k, v in sd.items():
        if k.startswith("context_model.scale_slice_transforms.") and k.endswith("model.4.bias"):
            u = torch.rand(v.shape, generator=g)
            v.copy_(torch.exp(math.log(0.05) + u * (math.log(24.0) - math.log(0.05))))
        if k.startswith("context_model.mean_slice_transforms.") and k.endswith("model.4.bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 2.0)
    if "encoder.interp_ca_list.8.weight" in sd:
        sd["encoder.interp_ca_list.8.weight"].add_(30.0)   # softplus(w) ~ 30.5
    else:   # stage-1 encoder (no InterpChAtt): raise the latent gain through the last gate convolution's trunk instead
        for k in ("encoder.attn4.trunk_block.2.c3.weight", "encoder.attn4.trunk_block.2.c3.bias"):
            sd[k].mul_(60.0)
    sd["entropy_model_z.quantiles"][:, 0, 0] = -6.0
    sd["entropy_model_z.quantiles"][:, 0, 2] = 7.0
    sd["entropy_model_z.quantiles"][:, 0, 1] = torch.rand(sd["entropy_model_z.quantiles"].shape[0], generator=g) - 0.5
    return sd


def build_model(seed=0, calibrated=True, device="cuda:0", config="crdr.yaml"):
    """(model, state_dict on CPU).  The model's parameters stay on the CPU; its engine lives on `device`."""
    from crdr_b200.model import build_comp_model
    torch.manual_seed(seed)
    model = build_comp_model(crdr_opt(device, config))
    if calibrated:
        with torch.no_grad():
            calibrate_(dict(model.state_dict()), seed)
    model.codec_setup()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return model, sd


def image(n, h, w, seed=7, smooth=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, h, w, generator=g) * 2 - 1
    if smooth:  # box-filtered noise: latents are less degenerate than for white noise
        x = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (2, 2, 2, 2), mode="reflect"), 5, stride=1) * 2.5
        x = x.clamp(-1, 1)
    return x.contiguous()

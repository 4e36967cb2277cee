"""ORACLE shim: import-only stand-ins (MS-SSIM is not on the codec path)."""
import torch


def ssim(*a, **k):
    raise NotImplementedError("pytorch_msssim is not available offline")


def ms_ssim(*a, **k):
    raise NotImplementedError("pytorch_msssim is not available offline")


class MS_SSIM(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise NotImplementedError("pytorch_msssim is not available offline")


class SSIM(MS_SSIM):
    pass

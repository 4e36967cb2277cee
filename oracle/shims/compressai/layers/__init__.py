"""ORACLE: import-only stand-in for compressai.layers.GDN (used only by
out-of-scope reference nets that `import src` auto-imports)."""
import torch
import torch.nn as nn


class GDN(nn.Module):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta = nn.Parameter(torch.ones(in_channels))
        self.gamma = nn.Parameter(gamma_init * torch.eye(in_channels))

    def forward(self, x):
        C = x.size(1)
        norm = torch.nn.functional.conv2d(x ** 2, self.gamma.abs().reshape(C, C, 1, 1), self.beta.abs())
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm

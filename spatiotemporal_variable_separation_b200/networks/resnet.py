"""Latent time-steppers (counterpart of /root/reference/var_sep/networks/resnet.py).

One application advances the dynamic code T by one frame: x <- x + f(x) per block, with NO
runtime step-size multiplier — ``gain_resnet`` only scales the initial weights.
"""
import torch.nn as nn

from .. import ops
from .conv import ConvBlock, make_conv_block  # noqa: F401  (re-exported like upstream)
from .mlp import MLP


class MLPResBlock(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.mlp = MLP(input_size, hidden_size, input_size, 3)

    def step(self, h, groups=1):
        residual = self.mlp.run(h, groups)
        return ops.add_act(h, residual), residual

    def forward(self, x):
        h, r = self.step(ops.to_internal(x))
        return ops.to_external(h).view(len(x), -1), ops.to_external(r).view(len(x), -1)


class MLPResnet(nn.Module):
    def __init__(self, input_size, n_blocks, hidden_size):
        super().__init__()
        self.in_size = input_size
        self.n_blocks = n_blocks
        self.blocks = nn.ModuleList([MLPResBlock(input_size, hidden_size) for _ in range(n_blocks)])

    def step(self, h, groups=1):
        residuals = []
        for blk in self.blocks:
            h, r = blk.step(h, groups)
            residuals.append(r)
        return h, residuals

    def forward(self, x, return_res=True):
        h, res = self.step(ops.to_internal(x))
        out = ops.to_external(h).view(len(x), -1)
        if return_res:
            return out, [ops.to_external(r).view(len(x), -1) for r in res]
        return out


class ConvResBlock(nn.Module):
    def __init__(self, in_c, out_c, nf=64):
        super().__init__()
        self.conv = nn.Sequential(
            ConvBlock(nn.Conv2d(in_c, nf, 3, padding=1), 'leaky_relu'),
            ConvBlock(nn.Conv2d(nf, nf, 3, padding=1), 'leaky_relu'),
            ConvBlock(nn.Conv2d(nf, out_c, 3, padding=1), 'none'))
        self.up = ConvBlock(nn.Conv2d(in_c, out_c, 3, padding=1), 'none') if in_c != out_c else nn.Identity()

    def step(self, h, groups=1):
        residual = h
        for blk in self.conv:
            residual = blk(residual, groups)
        base = h if isinstance(self.up, nn.Identity) else self.up(h, groups)
        return ops.add_act(base, residual), residual

    def forward(self, x):
        h, r = self.step(ops.to_internal(x))
        return ops.to_external(h), ops.to_external(r)


class ConvResnet(nn.Module):
    def __init__(self, in_c, n_blocks=1, nf=64):
        super().__init__()
        self.n_blocks = n_blocks
        self.resblock_modules = nn.ModuleList([ConvResBlock(in_c, in_c, nf=nf) for _ in range(n_blocks)])

    def step(self, h, groups=1):
        residuals = []
        for blk in self.resblock_modules:
            h, r = blk.step(h, groups)
            residuals.append(r)
        return h, residuals

    def forward(self, x, return_res=True):
        h, res = self.step(ops.to_internal(x))
        if return_res:
            return ops.to_external(h), [ops.to_external(r) for r in res]
        return ops.to_external(h)

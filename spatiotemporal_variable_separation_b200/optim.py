"""Fused Adam over a flat parameter arena (replaces torch.optim.Adam of main.py:145 / train.py:162).

All parameters of the model are re-homed as views of ONE fp32 buffer, their gradients as views of a
second one.  Consequences:
  * the optimizer is one HBM-bound kernel launch (16 B read + 12 B written per parameter) instead of
    ~6 foreach launches per tensor;
  * weight-gradient kernels accumulate straight into the arena (``p._vs_grad``), so autograd issues
    no per-parameter add kernels and ``zero_grad`` is a single memset;
  * data-parallel training all-reduces a handful of large contiguous buckets (parallel.py).
Parameters whose gradient is never produced (ResNet18.bn_out) see g = 0, hence
m = v = 0 and an exactly-zero update — the same end state as torch.optim.Adam skipping them.
"""
import torch

from . import _lib as L
from . import ops
from ._lib import ptr


class FusedAdam:
    def __init__(self, params, lr=4e-4, betas=(0.9, 0.99), eps=1e-8):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no parameters'
        dev = self.params[0].device
        L.require_cuda(self.params[0])
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4            # keep every tensor 16-byte aligned
        self.numel = total
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.grad_scale = 1.0
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                view = self.flat_p[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                g = self.flat_g[off:off + n].view(p.shape)
                p.grad = g
                p._vs_grad = g
        ops.invalidate_packed()

    # torch.optim-like surface used by train()
    @property
    def param_groups(self):
        return [{'params': self.params, 'lr': self.lr, 'betas': self.betas, 'eps': self.eps}]

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()

    def step(self):
        self.step_dev += 1
        L.call('vs_adam_step', ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.numel,
             self.lr, self.betas[0], self.betas[1], self.eps, float(self.grad_scale), 0, ptr(self.step_dev), L.stream())
        ops.invalidate_packed()
        ops.repack_all()

    def grad_of(self, p):
        return p._vs_grad

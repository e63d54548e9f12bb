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
        # one parameter group, kept as a persistent dict so that learning-rate schedules can write to it
        self._groups = [dict(params=self.params, lr=float(lr), initial_lr=float(lr), betas=(float(betas[0]), float(betas[1])),
                             eps=float(eps), weight_decay=0, amsgrad=False)]
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 7) // 8 * 8            # every tensor starts on a 32-byte boundary (8 fp32 / 16 bytes of bf16)
        self.numel = total
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        # the learning rate the kernel reads: device-resident so that a captured CUDA graph follows a schedule
        self.lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self._lr_uploaded = float(lr)
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
        ops.invalidate_params(self.params)
        self._pack_cache = {}

    # torch.optim-like surface used by train()
    @property
    def param_groups(self):
        return self._groups

    @property
    def lr(self):
        return self._groups[0]['lr']

    @lr.setter
    def lr(self, value):
        self._groups[0]['lr'] = float(value)

    @property
    def betas(self):
        return self._groups[0]['betas']

    @property
    def eps(self):
        return self._groups[0]['eps']

    def zero_grad(self, set_to_none=False):
        ops.join_wgrad(self.flat_p.device)
        self.flat_g.zero_()

    def sync_lr(self):
        """Upload the host learning rate if a scheduler changed it (one 4-byte fill; never inside a graph capture, where
        the value would be baked in — ``train.GraphedStep`` calls this before every replay instead)."""
        if self._lr_uploaded != self.lr:
            self.lr_dev.fill_(self.lr)
            self._lr_uploaded = self.lr

    def step(self):
        ops.join_wgrad(self.flat_p.device)             # weight gradients enqueued on the side stream (ops._OnWgradStream)
        if not (self.flat_p.is_cuda and torch.cuda.is_current_stream_capturing()):
            self.sync_lr()
        self.step_dev += 1
        L.call('vs_adam_step', ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.numel,
             self.lr, self.betas[0], self.betas[1], self.eps, float(self.grad_scale), 0, ptr(self.step_dev),
             ptr(self.lr_dev), L.stream())
        # the kernel wrote the weights through raw pointers: every packed copy of OUR parameters is stale now; refresh
        # the ones that exist in one launch (a table owned by this optimizer — no process-global registry)
        ops.invalidate_params(self.params)
        ops.repack_params(self.params, self._pack_cache)

    def grad_of(self, p):
        return p._vs_grad

    # ---- checkpointing in torch.optim.Adam's own format (the reference saves no optimizer state) ----
    def state_dict(self):
        """Same layout as ``torch.optim.Adam(...).state_dict()`` over the same parameter list, so training can resume
        under either optimizer: per parameter index ``step`` (fp32 scalar), ``exp_avg``, ``exp_avg_sq``."""
        step = float(self.step_dev.item())
        state = {}
        if step > 0:
            for i, (p, off) in enumerate(zip(self.params, self.offsets)):
                n = p.numel()
                state[i] = {'step': torch.tensor(step),
                            'exp_avg': self.exp_avg[off:off + n].view(p.shape).clone(),
                            'exp_avg_sq': self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        group = {k: v for k, v in self._groups[0].items() if k != 'params'}
        group['params'] = list(range(len(self.params)))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        groups = sd['param_groups']
        assert len(groups) == 1 and len(groups[0]['params']) == len(self.params), 'parameter list does not match'
        g = groups[0]
        assert not g.get('amsgrad', False) and not g.get('weight_decay', 0), 'amsgrad / weight decay are not implemented'
        self._groups[0].update(lr=float(g['lr']), betas=(float(g['betas'][0]), float(g['betas'][1])), eps=float(g['eps']))
        if 'initial_lr' in g:
            self._groups[0]['initial_lr'] = float(g['initial_lr'])
        steps = set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            st = sd['state'].get(i)
            if st is None:                       # a parameter torch.optim.Adam never saw a gradient for
                continue
            n = p.numel()
            self.exp_avg[off:off + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps.add(int(float(st['step'])))
        assert len(steps) <= 1, f'parameters at different step counts: {sorted(steps)}'
        self.step_dev.fill_(steps.pop() if steps else 0)


class MultiStepLR:
    """``torch.optim.lr_scheduler.MultiStepLR`` (main.py:146-147) for ``FusedAdam``: lr = initial_lr * gamma^(number of
    milestones reached), one ``step()`` per epoch (train.py:166-167).  Works on any object with ``param_groups``.
    (``FusedAdam`` reads the learning rate from device memory, so a captured CUDA graph follows the schedule.)"""

    def __init__(self, optimizer, milestones, gamma=0.1, last_epoch=0):
        self.optimizer, self.milestones, self.gamma = optimizer, sorted(int(m) for m in milestones), float(gamma)
        self.last_epoch = int(last_epoch)
        for g in optimizer.param_groups:
            g.setdefault('initial_lr', g['lr'])
        self._apply()

    def _apply(self):
        k = sum(1 for m in self.milestones if m <= self.last_epoch)
        for g in self.optimizer.param_groups:
            g['lr'] = g['initial_lr'] * self.gamma ** k

    def step(self):
        self.last_epoch += 1
        self._apply()

    def get_last_lr(self):
        return [g['lr'] for g in self.optimizer.param_groups]

    def state_dict(self):
        return {'milestones': list(self.milestones), 'gamma': self.gamma, 'last_epoch': self.last_epoch}

    def load_state_dict(self, sd):
        self.milestones, self.gamma, self.last_epoch = sorted(sd['milestones']), float(sd['gamma']), int(sd['last_epoch'])
        self._apply()

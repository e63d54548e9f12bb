"""CPU restatement of one training step.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/var_sep/train.py line by line, with the host random
draw ``t_random`` (train.py:72-75) injected so both sides take the same slice.
"""
import math

import torch
import torch.nn.functional as F


def zero_order_loss(s_old, s_new, skipco):
    """train.py:38-42."""
    if skipco:
        s_old = torch.cat([s_old[0].flatten()] + [x.flatten() for x in s_old[1]])
        s_new = torch.cat([s_new[0].flatten()] + [x.flatten() for x in s_new[1]])
    return (s_old - s_new).pow(2).mean()


def draw_t_random(rng, nt_cond, n_frames, offset):
    """train.py:72-75 — upper bound is exclusive, one more value when offset != 0."""
    hi = n_frames if offset == 0 else n_frames + 1
    return int(rng.randint(nt_cond, hi))


def ae_loss(net, cond, target, nt_cond, offset, skipco, t_random):
    """train.py:45-88."""
    full = torch.cat([cond, target], dim=1)
    s_old = net.Es(full[:, :nt_cond], return_skip=skipco)
    s_new = net.Es(full[:, -nt_cond:], return_skip=skipco)
    t_rand = net.Et(full[:, t_random - nt_cond:t_random])
    if skipco:
        recon = net.decoder(s_old[0], t_rand, s_old[1])
    else:
        recon = net.decoder(s_old, t_rand)
    return F.mse_loss(full[:, t_random - offset], recon), s_new, s_old


def step_losses(net, cond, target, cfg, t_random):
    """train.py:116-149.  Returns dict(total, ae, s, pred, t, forecasts, t_codes)."""
    nt_cond, nt_pred, offset, skipco = cfg['nt_cond'], cfg['nt_pred'], cfg['offset'], cfg['skipco']
    assert offset == nt_cond or offset == 0                       # train.py:103
    lamb_t = 0 if cfg.get('no_s') else cfg['lamb_t']              # train.py:99-101
    ae, s_recent, s_old = ae_loss(net, cond, target, nt_cond, offset, skipco, t_random)
    s_inv = zero_order_loss(s_old, s_recent, skipco)
    full = torch.cat([cond, target], dim=1)
    forecasts, t_codes, _, _ = net.get_forecast(cond, nt_pred + offset, init_s_code=s_old)
    f_off = nt_cond if offset == 0 else 0
    pred = F.mse_loss(forecasts, full[:, f_off:])
    if cfg['architecture'] == 'encoderSST':                        # average_tloss, main.py:162
        t_reg = 0.5 * t_codes[:, 0].pow(2).view(full.shape[0], -1).mean()
    else:
        t_reg = 0.5 * torch.sum(t_codes[:, 0].pow(2), dim=1).mean()
    total = cfg['lamb_ae'] * ae + cfg['lamb_s'] * s_inv + cfg['lamb_pred'] * pred + lamb_t * t_reg
    return dict(total=total, ae=ae, s=s_inv, pred=pred, t=t_reg, forecasts=forecasts, t_codes=t_codes)


class Adam:
    """torch.optim.Adam(lr, betas) as configured by main.py:145 — eps 1e-8, no
    weight decay, no amsgrad; parameters whose grad is None are skipped."""

    def __init__(self, named_params, lr, betas, eps=1e-8):
        self.params = [p for _, p in named_params]
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.state = {}

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        for i, p in enumerate(self.params):
            if p.grad is None:
                continue
            st = self.state.setdefault(i, dict(t=0, m=torch.zeros_like(p), v=torch.zeros_like(p)))
            st['t'] += 1
            g = p.grad
            st['m'].mul_(self.b1).add_(g, alpha=1 - self.b1)
            st['v'].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1 = 1 - self.b1 ** st['t']
            bc2 = 1 - self.b2 ** st['t']
            denom = (st['v'].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(st['m'], denom, value=-self.lr / bc1)


def train_step(net, opt, cond, target, cfg, t_random):
    """zero_grad -> losses -> backward -> Adam (train.py:115-162, no AMP)."""
    net.train = True
    opt.zero_grad()
    out = step_losses(net, cond, target, cfg, t_random)
    out['total'].backward()
    opt.step()
    return out

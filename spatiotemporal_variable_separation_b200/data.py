"""Synthetic input sequences of each configuration's shape and value statistics.

No dataset can be downloaded here, and the metric is defined on synthetic data: Moving-MNIST-shaped
sequences are two procedural 28x28 glyphs bouncing elastically in a 64x64 frame (positions U{0..36},
velocities U{-4..4}, summed and clipped to [0,1] as /root/reference/var_sep/data/moving_mnist.py:131-253
does with real digits); the other datasets are matched in shape and range only.  Everything is
vectorised torch code that runs on the device it is asked for.
"""
import torch


def _glyphs(n, gen, device):
    """n random 28x28 'digit-like' strokes in [0,1]: a few thick line segments, blurred."""
    yy, xx = torch.meshgrid(torch.arange(28, device=device, dtype=torch.float32),
                            torch.arange(28, device=device, dtype=torch.float32), indexing='ij')
    img = torch.zeros(n, 28, 28, device=device)
    for _ in range(3):
        p0 = 4 + 20 * torch.rand(n, 2, generator=gen, device=device)
        p1 = 4 + 20 * torch.rand(n, 2, generator=gen, device=device)
        d = p1 - p0
        t = ((yy[None] - p0[:, 0, None, None]) * d[:, 0, None, None] + (xx[None] - p0[:, 1, None, None]) * d[:, 1, None, None])
        t = (t / (d.pow(2).sum(1)[:, None, None] + 1e-6)).clamp(0, 1)
        dist2 = (yy[None] - p0[:, 0, None, None] - t * d[:, 0, None, None]) ** 2 + \
                (xx[None] - p0[:, 1, None, None] - t * d[:, 1, None, None]) ** 2
        img = torch.maximum(img, torch.exp(-dist2 / 3.0))
    return img


def moving_glyphs(batch, n_frames, channels=1, size=64, n_object=2, seed=0, device='cpu'):
    gen = torch.Generator(device=device).manual_seed(seed)
    lim = size - 28
    out = torch.zeros(batch, n_frames, size, size, device=device)
    ar = torch.arange(size, device=device)
    for _ in range(n_object):
        g = _glyphs(batch, gen, device)
        pos = torch.randint(0, lim + 1, (batch, 2), generator=gen, device=device).float()
        vel = torch.randint(-4, 5, (batch, 2), generator=gen, device=device).float()
        t = torch.arange(n_frames, device=device, dtype=torch.float32)
        raw = pos[:, None, :] + vel[:, None, :] * t[None, :, None]            # [B,T,2]
        period = 2 * lim
        m = torch.remainder(raw, period)
        p = torch.where(m > lim, period - m, m).round().long()                  # elastic bounce
        yy = ar[None, None, :] - p[:, :, 0, None]                               # [B,T,size]
        xx = ar[None, None, :] - p[:, :, 1, None]
        my, mx = (yy >= 0) & (yy < 28), (xx >= 0) & (xx < 28)
        rows = torch.gather(g[:, None].expand(-1, n_frames, -1, -1), 2,
                            yy.clamp(0, 27)[..., None].expand(-1, -1, -1, 28))   # [B,T,size,28]
        patch = torch.gather(rows, 3, xx.clamp(0, 27)[:, :, None, :].expand(-1, -1, size, -1))
        out += patch * (my[..., None] & mx[:, :, None, :])
    out = out.clamp_(0, 1)
    return out[:, :, None].expand(-1, -1, channels, -1, -1).contiguous()


def synthetic_batch(cfg, batch=None, device='cpu', seed=0):
    """One batch [B, nt_cond + nt_pred, C, H, W] float32 = cat(cond, target) for configuration ``cfg``."""
    B = batch or cfg['batch_size']
    T = cfg['nt_cond'] + cfg['nt_pred']
    C, H, W = cfg['shape']
    gen = torch.Generator(device=device).manual_seed(seed)
    data = cfg['data']
    if data == 'mnist':
        return moving_glyphs(B, T, C, H, seed=seed, device=device)
    if data == 'wave':
        # smooth radial waves, min-max scaled to [0,1] per sequence (data/wave_eq.py:56-57)
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H, device=device), torch.linspace(0, 1, W, device=device), indexing='ij')
        c = torch.rand(B, 3, 2, generator=gen, device=device)
        k = 10 + 20 * torch.rand(B, 3, generator=gen, device=device)
        t = torch.arange(T, device=device, dtype=torch.float32)
        r = ((yy[None, None] - c[:, :, 0, None, None]) ** 2 + (xx[None, None] - c[:, :, 1, None, None]) ** 2).sqrt()
        u = torch.sin(k[:, :, None, None, None] * r[:, :, None] - 0.6 * t[None, None, :, None, None]).sum(1)
        lo, hi = u.amin((1, 2, 3), keepdim=True), u.amax((1, 2, 3), keepdim=True)
        return ((u - lo) / (hi - lo + 1e-6))[:, :, None].contiguous()
    if data == 'sst':
        return torch.randn(B, T, C, H, W, generator=gen, device=device)
    x = torch.rand(B, T, C, H, W, generator=gen, device=device)
    return x * x if data == 'taxibj' else x

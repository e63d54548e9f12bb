"""Synthetic input sequences of each configuration's shape and value statistics, and the device-side
Moving-MNIST training generator (``MovingSequences``: the reference's bounce semantics, one kernel per batch).

No dataset can be downloaded here, and the metric is defined on synthetic data: Moving-MNIST-shaped
sequences are two procedural 28x28 glyphs bouncing elastically in a 64x64 frame (positions U{0..36},
velocities U{-4..4}, summed and clipped to [0,1] as /root/reference/var_sep/data/moving_mnist.py:131-253
does with real digits); the other datasets are matched in shape and range only.  Everything is
vectorised torch code that runs on the device it is asked for.
"""
import numpy as np
import torch

from . import _lib as L


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


class MovingSequences:
    """Training batches of bouncing glyphs generated ON THE DEVICE: counterpart of ``MovingMNIST`` in
    /root/reference/var_sep/data/moving_mnist.py (train branch of ``__getitem__`` :112-130, ``_compute_trajectory``
    :131-170, ``_process_collision`` :172-253; deterministic variant) used the way main.py:111-114 uses its DataLoader.

    The reference builds every sample with Python loops over digits and frames on the host (a few hundred samples per
    second and core); here the host only draws five integers per object — glyph index, start position, speed — from
    numpy's global RNG **in the reference's order**, so that after ``np.random.seed(s)`` the batches are bit-identical
    to ``batch_size`` consecutive ``dataset[i]`` calls of the reference, and one kernel (``vs_moving_sequences``) renders
    all frames of the batch straight into device memory.  Iterating yields ``(cond, target)`` like the reference loader.

    ``glyphs``: uint8 [G, h, w] (the MNIST digits in the reference; any glyph bank here)."""

    def __init__(self, glyphs, frame_size=64, nt_cond=5, seq_len=15, max_speed=4, num_digits=2, batch_size=128,
                 batches_per_epoch=None, device='cuda'):
        g = torch.as_tensor(np.ascontiguousarray(glyphs))
        assert g.dtype == torch.uint8 and g.dim() == 3, 'glyphs: uint8 [G, h, w]'
        self.device = torch.device(device)
        self.glyphs = g.to(self.device)
        self.n_glyphs, self.gh, self.gw = (int(v) for v in g.shape)
        self.frame_size, self.nt_cond, self.seq_len = frame_size, nt_cond, seq_len
        self.max_speed, self.num_digits, self.batch_size = max_speed, num_digits, batch_size
        # the reference's __len__ is an arbitrary 200000 samples per epoch (moving_mnist.py:103-110)
        self.batches_per_epoch = batches_per_epoch if batches_per_epoch is not None else 200000 // batch_size
        lo = np.array([0, 0, 0, -max_speed, -max_speed])
        hi = np.array([self.n_glyphs, frame_size - self.gh + 1, frame_size - self.gw + 1, max_speed + 1, max_speed + 1])
        self._lo, self._hi = lo, hi

    def draw(self, batch_size=None):
        """int32 [B, num_digits, 5] = (glyph, sx, sy, dx, dy) per object: one vectorised ``np.random.randint`` call that
        consumes the global RNG stream exactly like the reference's scalar calls (moving_mnist.py:119-120, 147-150)."""
        B = batch_size or self.batch_size
        n = B * self.num_digits
        v = np.random.randint(np.tile(self._lo, n), np.tile(self._hi, n))
        return v.reshape(B, self.num_digits, 5).astype(np.int32)

    def render(self, objs):
        """objs (host int32 [B, n, 5]) -> frames [B, seq_len, 1, F, F] fp32 on the device."""
        L.require_cuda(self.glyphs)
        objs_d = torch.as_tensor(np.ascontiguousarray(objs, dtype=np.int32)).to(self.device, non_blocking=True)
        B = objs_d.shape[0]
        frames = torch.empty((B, self.seq_len, 1, self.frame_size, self.frame_size), device=self.device, dtype=torch.float32)
        L.call('vs_moving_sequences', self.glyphs, self.n_glyphs, self.gh, self.gw, objs_d, objs_d.shape[1], B, self.seq_len,
               self.frame_size, frames, L.stream())
        return frames

    def batch(self, batch_size=None):
        frames = self.render(self.draw(batch_size))
        return frames[:, :self.nt_cond], frames[:, self.nt_cond:]

    def __len__(self):
        return self.batches_per_epoch

    def __iter__(self):
        for _ in range(self.batches_per_epoch):
            yield self.batch()


def procedural_glyphs(n=256, size=28, seed=0):
    """A bank of digit-like uint8 strokes (MNIST cannot be downloaded here)."""
    gen = torch.Generator().manual_seed(seed)
    return (255 * _glyphs(n, gen, 'cpu')).round().to(torch.uint8).numpy()

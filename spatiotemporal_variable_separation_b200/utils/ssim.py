"""SSIM on the device (counterpart of /root/reference/var_sep/utils/ssim.py:81-149).

``ssim_loss`` keeps the reference signature and semantics (despite its name it returns the SSIM *index*: the map with
``reduction='none'``, its mean or sum otherwise).  The five Gaussian-window sums and the SSIM formula run in one
kernel per call (``vs_ssim_mse_planes``: one CTA per (image, channel) plane); a user-supplied ``kernel`` must hold the
same window for every channel, as the reference's own ``_fspecial_gaussian`` does."""
import torch

from .. import _lib as L
from .._lib import ptr


def _fspecial_gaussian(size, channel, sigma):
    """Normalised size x size Gaussian window (the window ssim.py:84-92 builds as a softmax of negative squared
    distances), one copy per channel: [channel, 1, size, size]."""
    offsets = torch.arange(size, dtype=torch.float32) - (size - 1) / 2.0
    g = torch.exp(-offsets.pow(2) / (2.0 * sigma ** 2))
    window = torch.outer(g, g)
    window = window / window.sum()
    return window.expand(channel, 1, size, size).contiguous()


def plane_metrics(input, target, max_val, filter_size=11, k1=0.01, k2=0.03, sigma=1.5, kernel=None, want_map=False):
    """input / target [N, C, H, W] fp32 -> (ssim_mean [N, C], mse_mean [N, C], ssim_map [N, C, H-fs+1, W-fs+1] or None)."""
    if input.size() != target.size():
        raise ValueError('Expected input size ({}) to match target size ({}).'.format(input.size(0), target.size(0)))
    L.require_cuda(input, target)
    N, C, H, W = input.shape
    if kernel is None:
        kernel = _fspecial_gaussian(filter_size, 1, sigma)
    fs = kernel.shape[-1]
    win = kernel.reshape(-1, fs * fs)[0].to(device=input.device, dtype=torch.float32).contiguous()
    x, y = input.float().contiguous(), target.float().contiguous()
    ssim_mean = torch.empty((N, C), device=x.device, dtype=torch.float32)
    mse_mean = torch.empty((N, C), device=x.device, dtype=torch.float32)
    ssim_map = torch.empty((N, C, H - fs + 1, W - fs + 1), device=x.device, dtype=torch.float32) if want_map else None
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    L.call('vs_ssim_mse_planes', ptr(x), ptr(y), N * C, H, W, ptr(win), fs, c1, c2, ptr(ssim_map), ptr(ssim_mean),
           ptr(mse_mean), L.stream())
    return ssim_mean, mse_mean, ssim_map


def ssim_loss(input, target, max_val, filter_size=11, k1=0.01, k2=0.03, sigma=1.5, kernel=None, size_average=None,
              reduce=None, reduction='mean'):
    if size_average is not None or reduce is not None:
        from torch.nn import _reduction as _Reduction
        reduction = _Reduction.legacy_get_string(size_average, reduce)
    dim = input.dim()
    if dim == 2:
        input, target = input[None, None], target[None, None]
    elif dim == 3:
        input, target = input[None], target[None]
    elif dim != 4:
        raise ValueError('Expected 2, 3, or 4 dimensions (got {})'.format(dim))
    ssim_mean, _, ssim_map = plane_metrics(input, target, max_val, filter_size, k1, k2, sigma, kernel,
                                           want_map=reduction == 'none')
    if reduction == 'none':
        return ssim_map
    # every plane has the same number of window positions: the mean over the map is the mean of the plane means
    if reduction == 'mean':
        return ssim_mean.mean()
    fs = kernel.shape[-1] if kernel is not None else filter_size
    return ssim_mean.sum() * float((input.shape[-2] - fs + 1) * (input.shape[-1] - fs + 1))

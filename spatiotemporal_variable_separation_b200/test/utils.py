"""Evaluation helpers (counterpart of /root/reference/var_sep/test/utils.py).

``load_model`` rebuilds the four networks from the experiment's ``params.json`` (main.py:105-106) through the factory and
loads the checkpoint files of ``xp_dir`` — ``state_dict``s written by this package or whole-module pickles written by
the reference (utils/helper.py).  ``_ssim_wrapper`` keeps the reference's shape contract; ``eval_metrics`` is the
per-sequence MSE / PSNR / SSIM block of test/mnist/test.py:136-141 with the SSIM map and both frame reductions on the
device."""
import argparse
import os

import torch

from ..networks.factory import build_model
from ..options import config_from_args
from ..utils import helper
from ..utils.ssim import plane_metrics


def load_model(args, epoch_number=None):
    """test/utils.py:8-16.  ``args`` needs ``xp_dir`` and ``device``; the architecture flags are read from
    ``xp_dir/params.json`` when ``args`` does not carry them."""
    xp_dir = args['xp_dir'] if isinstance(args, dict) else args.xp_dir
    cfg_path = os.path.join(xp_dir, 'params.json')
    merged = dict(helper.load_json(cfg_path)) if os.path.exists(cfg_path) else {}
    merged.update({k: v for k, v in (args if isinstance(args, dict) else vars(args)).items() if v is not None})
    cfg = config_from_args(argparse.Namespace(**merged))
    sep_net = build_model(cfg, merged.get('device', 'cuda'))
    helper.load(merged['xp_dir'], sep_net, epoch_number)
    sep_net.eval()
    return sep_net


def _ssim_wrapper(pred, gt):
    """[B, T, C, H, W] x 2 -> [B, T, C]: SSIM index averaged over the window positions (test/utils.py:19-24)."""
    bsz, nt_pred = pred.shape[0], pred.shape[1]
    img_shape = pred.shape[2:]
    ssim, _, _ = plane_metrics(pred.reshape(bsz * nt_pred, *img_shape), gt.reshape(bsz * nt_pred, *img_shape), max_val=1.)
    return ssim.view(bsz, nt_pred, img_shape[0])


def eval_metrics(x_pred, x_target):
    """test/mnist/test.py:136-141: per-sequence ``mse``, ``psnr`` and ``ssim`` ([B] each) of forecasts against targets
    ([B, T, C, H, W] in [0, 1])."""
    bsz, nt = x_pred.shape[0], x_pred.shape[1]
    img_shape = x_pred.shape[2:]
    ssim, mse, _ = plane_metrics(x_pred.reshape(bsz * nt, *img_shape), x_target.reshape(bsz * nt, *img_shape), max_val=1.)
    ssim, mse = ssim.view(bsz, nt, img_shape[0]), mse.view(bsz, nt, img_shape[0])
    return {'mse': mse.mean(2).mean(1), 'psnr': (10 * torch.log10(1 / mse)).mean(2).mean(1), 'ssim': ssim.mean(2).mean(1)}

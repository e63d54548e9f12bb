"""Evaluation metrics: SSIM / MSE / PSNR of forecasts.

CPU part: the executable specification (tests/emu.py) and the package wrappers over it against values recorded from
the reference (tests/golden/gen_metrics_golden.py), and the arithmetic of the CUDA kernel itself — csrc/metrics_core.h
compiled for the host with g++ — against the specification.  GPU part: the kernel against the specification and the
golden values."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import detfill
from spatiotemporal_variable_separation_b200 import _lib as L
from spatiotemporal_variable_separation_b200 import ops
from spatiotemporal_variable_separation_b200.test import utils as eval_utils
from spatiotemporal_variable_separation_b200.utils import helper
from spatiotemporal_variable_separation_b200.utils.ssim import _fspecial_gaussian, plane_metrics, ssim_loss
from tests import emu, harness
from tests.summ import subsample

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'eval', '_metrics', 'metrics.npz')
CASES = [('m1', 3, 4, (1, 64, 64), 'blobs'), ('c3', 2, 3, (3, 64, 64), 'uniform'), ('t2', 2, 2, (2, 32, 32), 'uniform')]


def pair(tag, B, T, shape, kind):
    a = detfill.frames('metrics:a:' + tag, B, T, shape, kind=kind)
    b = detfill.frames('metrics:b:' + tag, B, T, shape, kind=kind)
    return (0.7 * a + 0.3 * b).float(), a.float()


def check_wrappers_against_golden(device):
    g = np.load(GOLD)
    for tag, B, T, shape, kind in CASES:
        pred, gt = pair(tag, B, T, shape, kind)
        pred, gt = pred.to(device), gt.to(device)
        np.testing.assert_allclose(eval_utils._ssim_wrapper(pred, gt).cpu().numpy(), g[tag + '_ssim'], rtol=0, atol=2e-5)
        m = eval_utils.eval_metrics(pred, gt)
        np.testing.assert_allclose(m['mse'].cpu().numpy(), g[tag + '_mse'], rtol=1e-5)
        np.testing.assert_allclose(m['psnr'].cpu().numpy(), g[tag + '_psnr'], rtol=1e-5)
        np.testing.assert_allclose(m['ssim'].cpu().numpy(), g[tag + '_ssim_seq'], rtol=0, atol=2e-5)
        fp, fg = pred.reshape(B * T, *shape), gt.reshape(B * T, *shape)
        assert abs(float(ssim_loss(fp, fg, max_val=1.)) - float(g[tag + '_loss_mean'])) < 2e-5
        assert abs(float(ssim_loss(fp, fg, max_val=1., reduction='sum')) - float(g[tag + '_loss_sum'])) < 2e-5 * abs(float(g[tag + '_loss_sum'])) + 1e-3
        smap = ssim_loss(fp, fg, max_val=1., reduction='none')
        assert list(smap.shape) == [B * T, shape[0], shape[1] - 10, shape[2] - 10]
        # per-pixel values are conditioned by sigma^2 = E[x^2] - mu^2 against c2 = 9e-4: two fp32 evaluations of the
        # reference formula in different summation orders differ by up to 7e-5 on the sparse 'm1' case (4.5e-5 from
        # fp64), while the means over a plane agree to 1e-7
        np.testing.assert_allclose(subsample(smap), g[tag + '_map_sub'], rtol=0, atol=5e-4)
        assert abs(float(ssim_loss(fp, fg, max_val=1., filter_size=7, sigma=1.0)) - float(g[tag + '_k7'])) < 2e-5
        assert abs(float(ssim_loss(fg, fg, max_val=1.)) - 1.0) < 1e-5                      # identical images


def test_specification_and_wrappers_match_reference_golden():
    with emu.install():
        check_wrappers_against_golden('cpu')


def test_wrapper_errors_match_reference():
    with emu.install():
        with pytest.raises(ValueError):
            ssim_loss(torch.zeros(2, 1, 16, 16), torch.zeros(3, 1, 16, 16), max_val=1.)
        with pytest.raises(ValueError):
            ssim_loss(torch.zeros(2, 1, 1, 16, 16), torch.zeros(2, 1, 1, 16, 16), max_val=1.)


_HARNESS = r'''
#include "metrics_core.h"
extern "C" void ssim_plane(const float* X, const float* Y, const float* K, int H, int W, int fs, float c1, float c2, float* out) {
    const int OH = H - fs + 1, OW = W - fs + 1;
    for (int o = 0; o < OH * OW; ++o) out[o] = vs_ssim_at(X, Y, K, W, fs, o / OW, o % OW, c1, c2);
}
'''


def test_kernel_arithmetic_compiled_for_the_host(tmp_path):
    """The per-pixel function the CUDA kernel calls (csrc/metrics_core.h), built with g++ and run over whole planes,
    against the specification: pins the window indexing and the formula without a GPU."""
    src = tmp_path / 'h.cpp'
    src.write_text(_HARNESS)
    so = tmp_path / 'h.so'
    inc = os.path.join(ROOT, 'spatiotemporal_variable_separation_b200', 'csrc')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-ffp-contract=off', f'-I{inc}', str(src), '-o', str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ssim_plane.argtypes = [fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, fp]
    for (H, W, fs, sigma) in [(64, 64, 11, 1.5), (32, 48, 7, 1.0), (16, 16, 15, 2.0)]:
        torch.manual_seed(H + fs)
        x, y = torch.rand(H, W), torch.rand(H, W)
        k = _fspecial_gaussian(fs, 1, sigma).reshape(-1).contiguous()
        out = torch.empty((H - fs + 1) * (W - fs + 1))
        lib.ssim_plane(*[ctypes.cast(t.data_ptr(), fp) for t in (x, y, k)], H, W, fs, 1e-4, 9e-4, ctypes.cast(out.data_ptr(), fp))
        smap, smean, mmean = torch.empty_like(out), torch.empty(1), torch.empty(1)
        emu.emu_call('vs_ssim_mse_planes', x, y, 1, H, W, k, fs, 1e-4, 9e-4, smap, smean, mmean, None)
        assert float((out - smap).abs().max()) < 2e-5, (H, W, fs)


def test_load_model_round_trip(tmp_path):
    """test/utils.py:8-16: networks rebuilt from params.json + the checkpoint files give the same eval forecast."""
    cfg = harness.load_golden('mnist-small')['cfg']
    ops.set_compute_dtype(torch.float32)
    with emu.install():
        from tests.test_host_emulated import build_filled
        net = build_filled(cfg).eval()
        helper.save(str(tmp_path), net)
        json.dump({k: v for k, v in cfg.items() if k not in ('shape', 'last_activation')}, open(tmp_path / 'params.json', 'w'))
        again = eval_utils.load_model({'xp_dir': str(tmp_path), 'device': 'cpu'})
        assert not again.training
        cond, _ = harness.inputs(cfg)
        with torch.no_grad():
            f1 = net.get_forecast(cond, 7)[0]
            f2 = again.get_forecast(cond, 7)[0]
        assert torch.equal(f1, f2)


@pytest.mark.gpu
def test_kernel_matches_specification_and_golden_on_gpu():
    torch.manual_seed(3)
    for (planes, H, W, fs, sigma) in [(7, 64, 64, 11, 1.5), (5, 32, 32, 11, 1.5), (3, 40, 24, 7, 1.0), (2, 15, 15, 15, 2.0)]:
        x, y = torch.rand(planes, H, W), torch.rand(planes, H, W)
        y = 0.6 * x + 0.4 * y
        k = _fspecial_gaussian(fs, 1, sigma).reshape(-1).contiguous()
        n_out = (H - fs + 1) * (W - fs + 1)
        for want_map in (True, False):
            outs = {}
            for dev in ('cuda', 'cpu'):
                smap = torch.full((planes * n_out,), float('nan'), device=dev) if want_map else None
                smean, mmean = torch.empty(planes, device=dev), torch.empty(planes, device=dev)
                args = [x.to(dev), y.to(dev), planes, H, W, k.to(dev), fs, 1e-4, 9e-4, smap, smean, mmean]
                if dev == 'cuda':
                    L.call('vs_ssim_mse_planes', *args, L.stream())
                    torch.cuda.synchronize()
                else:
                    emu.emu_call('vs_ssim_mse_planes', *args, None)
                outs[dev] = (smap, smean, mmean)
            if want_map:
                assert float((outs['cuda'][0].cpu() - outs['cpu'][0]).abs().max()) < 2e-4      # dense random planes: 6e-6 from fp64
            assert float((outs['cuda'][1].cpu() - outs['cpu'][1]).abs().max()) < 2e-5
            np.testing.assert_allclose(outs['cuda'][2].cpu().numpy(), outs['cpu'][2].numpy(), rtol=1e-5)
    check_wrappers_against_golden('cuda')

"""Parity on what is benchmarked: the FULL-SIZE BASELINE configurations (mnist DCGAN B=128 nf=64, wave MLP B=128,
taxibj VGG B=100) through the package on the GPU against the oracle run live on the host (fp64 = truth).

fp32 mode — the north-star bound, per tensor, by relative L2 error (not norms):
    losses, forecasts, t_codes:   err <= max(2e-5, 2 * err_ref)
    every parameter gradient:     err <= max(1e-4 * |g|, 2 * err_ref)
  err_ref is the reference's OWN fp32-vs-fp64 error for the same tensor, taken as the larger of its two stock
  executions (oneDNN on the host, cuDNN/cuBLAS with TF32 off on the device): at B >= 100 the gradient error of ANY fp32
  evaluation is dominated by LeakyReLU / max-pool units whose pre-activation lies within rounding distance of the kink
  (measured: 4e-4 median on mnist, 2e-3 on taxibj for the reference itself), so it is a noisy quantity.
bf16 mode — the stated looser bound, calibrated on the reference's own mixed-precision path (torch.autocast(bfloat16),
  the bf16 counterpart of --torch_amp: fp32 master weights, bf16 convolutions, fp32 BatchNorm):
    losses <= 2e-3, forecasts / t_codes <= max(5e-3, 1.5 * err_amp)
    gradients: median over tensors <= 1.25 * median_amp, every tensor <= max(5e-2, 2 * err_amp)
  (measured on the B200, profiles/r02_fullsize_parity.json: mnist median 4.8e-2 vs 5.4e-2 for the reference's autocast,
  max 0.17 vs 0.26; taxibj 0.35 vs 0.38 — a 23-layer train-mode BatchNorm chain amplifies bf16 rounding in BOTH).
"""
import numpy as np
import pytest
import torch

from spatiotemporal_variable_separation_b200 import ops
from tests import fullsize

pytestmark = pytest.mark.gpu
CASES = ['mnist', 'wave', 'taxibj']


@pytest.fixture(autouse=True)
def _restore_dtype():
    prev = ops.compute_dtype()
    yield
    ops.set_compute_dtype(prev)


@pytest.mark.parametrize('name', CASES)
def test_fp32_full_size_step_matches_fp64_oracle(name):
    truth = fullsize.oracle_run(name, 'float64')
    ref32 = fullsize.oracle_run(name, 'float32')
    ref32c = fullsize.oracle_run(name, 'float32_cuda')
    ref_rep = fullsize.compare(ref32, truth)
    refc_rep = fullsize.compare(ref32c, truth)
    rep = fullsize.compare(fullsize.cuda_run(name, torch.float32), truth, ref32, ref32c)
    assert rep['loss'] <= 2e-5, rep['loss_terms']
    for k in ('forecasts', 't_codes'):
        assert rep[k] <= max(2e-5, 2 * max(ref_rep[k], refc_rep[k])), (k, rep[k], ref_rep[k], refc_rep[k])
    for n, (err, err_ref) in rep['grads'].items():
        assert err <= max(1e-4, 2 * err_ref), (n, err, err_ref)
    assert all(v <= 1e-5 for v in rep['zero_grads'].values()), rep['zero_grads']


@pytest.mark.parametrize('name', CASES)
def test_bf16_full_size_step_within_stated_bound(name):
    truth = fullsize.oracle_run(name, 'float64')
    amp = fullsize.oracle_run(name, 'autocast_bf16')
    amp_rep = fullsize.compare(amp, truth)
    rep = fullsize.compare(fullsize.cuda_run(name, torch.bfloat16), truth, amp)
    assert rep['loss'] <= 2e-3, rep['loss_terms']
    for k in ('forecasts', 't_codes'):
        assert rep[k] <= max(5e-3, 1.5 * amp_rep[k]), (k, rep[k], amp_rep[k])
    errs = np.array([e for e, _ in rep['grads'].values()])
    amps = np.array([a for _, a in rep['grads'].values()])
    assert np.median(errs) <= 1.25 * np.median(amps), (np.median(errs), np.median(amps))
    for n, (err, err_amp) in rep['grads'].items():
        assert err <= max(5e-2, 2 * err_amp), (n, err, err_amp)

"""Pin the CPU oracle against the fixtures generated from the unmodified reference.

Tolerances: the oracle and the reference run the same ATen CPU kernels on the
same values, in the same order, so fp32 agreement is expected at the 1e-6
level (thread-count differences give 2.5e-7 on forecasts / 1.8e-6 on gradients,
SURVEY section 8c).  The bounds below are 1e-5 relative for losses/forecasts and
5e-5 relative (on [norm, probe-dot]) for gradients and two-step states.
"""
import numpy as np
import pytest
import torch

from oracle import shapes, step
from tests import harness
from tests.summ import summarize, subsample, rel_err

NAMES = harness.golden_names()
ALL_NAMES = NAMES + harness.extra_golden_names()


def test_goldens_exist():
    assert {'mnist-small', 'wave-small', 'taxibj-small', 'sst-small', 'chairs-small'} <= set(NAMES)


@pytest.mark.parametrize('name', ALL_NAMES)
def test_state_dict_layout_matches_reference(name):
    g = harness.load_golden(name)
    ours = shapes.model_shapes(g['cfg'])
    for part in harness.PARTS:
        ref = g['shapes'][part]
        assert list(ours[part].keys()) == list(ref.keys()), part
        for k in ref:
            assert list(ours[part][k]) == ref[k], (part, k)


@pytest.mark.parametrize('name', ALL_NAMES)
def test_losses_forecasts_and_grads(name):
    g = harness.load_golden(name)
    cfg = g['cfg']
    torch.set_num_threads(8)
    net, out, grads = harness.oracle_step(cfg, torch.float32, int(g['np_seed']))
    ours = np.array([float(out[k].detach()) for k in ('ae', 's', 'pred', 't', 'total')])
    np.testing.assert_allclose(ours, g['loss32'], rtol=1e-5, atol=1e-7)
    assert list(out['forecasts'].shape) == list(g['forecast_shape'])
    assert rel_err(subsample(out['forecasts']), g['forecast_sub']) < 1e-5
    assert rel_err(out['t_codes'].detach().numpy(), g['t_codes']) < 1e-5
    gmax = max(np.nan_to_num(g['grad32'][:, 0]).max(), 1e-30)
    for n, ref in zip(g['grad_names'], g['grad32']):
        gr = grads[str(n)]
        if np.isnan(ref[0]):
            assert gr is None, n
            continue
        s = summarize(str(n), gr)
        # absolute floor for the mathematically-zero gradients (biases feeding a train-mode BN)
        assert np.all(np.abs(s - ref) <= 5e-5 * np.abs(ref[0]) + 1e-6 * gmax), (n, s, ref)


@pytest.mark.parametrize('name', ALL_NAMES)
def test_two_adam_steps_match_reference_train_loop(name):
    g = harness.load_golden(name)
    cfg = g['cfg']
    torch.set_num_threads(8)
    net = harness.oracle_net(cfg)
    opt = step.Adam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
    cond, target = harness.inputs(cfg)
    for t_random in harness.t_random_sequence(cfg, int(g['np_seed']), 2):
        step.train_step(net, opt, cond, target, cfg, t_random)
    state = {f'{part}.{k}': v for part in harness.PARTS for k, v in net.P[part].items()
             if isinstance(v, torch.Tensor)}
    for n, ref in zip(g['after2_names'], g['after2']):
        s = summarize(str(n), state[str(n)])
        assert np.all(np.abs(s - ref) <= 5e-5 * max(abs(ref[0]), 1e-12)), (n, s, ref)


def test_fp64_oracle_agrees_with_fp64_reference():
    g = harness.load_golden('mnist-small')
    _, out, _ = harness.oracle_step(g['cfg'], torch.float64, int(g['np_seed']))
    ours = np.array([float(out[k].detach()) for k in ('ae', 's', 'pred', 't', 'total')])
    np.testing.assert_allclose(ours, g['loss64'], rtol=1e-12)


@pytest.mark.parametrize('name', harness.eval_golden_names())
def test_eval_rollout_and_content_swap(name):
    """eval-mode forecast over three training horizons, restart from a given T code and
    content swap through init_s_code, oracle against the reference's recorded values (1e-5)."""
    g = harness.load_eval_golden(name)
    torch.set_num_threads(8)
    net = harness.oracle_net(g['cfg']).requires_grad_(False)
    net.train = False
    harness.check_eval_rollout(g, net, g['cfg']['skipco'], rtol=1e-5)

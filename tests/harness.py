"""Shared test harness: build the oracle for a config, run it, load goldens."""
import glob
import json
import os

import numpy as np
import torch

from oracle import detfill, functional, shapes, step
from spatiotemporal_variable_separation_b200 import configs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PARTS = ('Es', 'Et', 'decoder', 't_resnet')


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def extra_golden_names():
    """Flag combinations beyond the BASELINE configurations (tests/golden/extra/, CPU checks only so far)."""
    return sorted('extra/' + os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'extra', '*.npz')))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    g['cfg'] = json.loads(str(g['cfg']))
    g['shapes'] = json.loads(str(g['shapes']))
    return g


def eval_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'eval', '*.npz')))


def load_eval_golden(name):
    """Eval-mode long-horizon rollout / content swap recorded from the reference (tests/golden/gen_eval_golden.py)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'eval', name + '.npz'), allow_pickle=False))
    g['cfg'] = json.loads(str(g['cfg']))
    return g


def check_eval_rollout(g, net, skipco, device='cpu', rtol=2e-5):
    """``net``: anything with the reference's Es / Et / get_forecast surface, already in eval mode.  Compares the
    long-horizon forecast, its latent rollout and last residual, the restart from ``init_t_code`` and the content
    swap through ``init_s_code`` with the reference's values.  Returns the worst relative error."""
    from tests.summ import subsample, rel_err
    cfg = g['cfg']
    cond, target = inputs(cfg)
    full = torch.cat([cond, target], 1).to(device)
    cond = cond.to(device)
    n_long, n_swap = int(g['n_long']), int(g['n_swap'])
    with torch.no_grad():
        f_long, t_long, s_code, res = net.get_forecast(cond, n_long)
        s_other = net.Es(full[:, -cfg['nt_cond']:], return_skip=skipco)
        f_swap, _, _, _ = net.get_forecast(cond, n_swap, init_s_code=s_other)
        f_init, _, _, _ = net.get_forecast(cond, n_swap, init_t_code=net.Et(cond))
    assert list(f_long.shape) == list(g['long_shape']) and list(f_swap.shape) == list(g['swap_shape'])
    s_first = s_code[0] if isinstance(s_code, (tuple, list)) else s_code
    errs = {'long': rel_err(subsample(f_long.contiguous()), g['long_sub']),
            't_codes': rel_err(t_long.detach().cpu().numpy(), g['t_long']),
            's_code': rel_err(subsample(s_first.contiguous()), g['s_code']),
            'residual': rel_err(res[-1][-1].detach().cpu().numpy(), g['last_residual']),
            'swap': rel_err(subsample(f_swap.contiguous()), g['swap_sub']),
            'sum': abs(float(f_long.double().sum()) - g['long_sum'][0]) / g['long_sum'][1],
            'restart': rel_err(f_init.detach().cpu().numpy(), f_long[:, :n_swap].detach().cpu().numpy())}
    for k, v in errs.items():
        assert v < rtol, (k, v, errs)
    return max(errs.values())


def oracle_net(cfg, dtype=torch.float32):
    """Oracle network with the deterministic name-keyed weights."""
    sh = shapes.model_shapes(cfg)
    # values are defined in fp32 and widened (exactly what gen_golden.py does with net.to(dtype))
    P = {part: {k: v.to(dtype) for k, v in detfill.fill_state(sh[part], part + '.').items()} for part in PARTS}
    for part in PARTS:                      # integer buffers stay int64
        for k in P[part]:
            if k.endswith('num_batches_tracked'):
                P[part][k] = torch.zeros((), dtype=torch.int64)
    if cfg.get('no_s'):
        P['Es'] = {'__code_size__': cfg['code_size_s']}
    net = functional.Net(cfg, P['Es'], P['Et'], P['decoder'], P['t_resnet'])
    return net.requires_grad_(True)


def inputs(cfg, dtype=torch.float32):
    x = detfill.frames('x:' + cfg['name'], cfg['batch_size'], cfg['nt_cond'] + cfg['nt_pred'], cfg['shape'],
                       kind=configs.input_kind(cfg)).to(dtype)
    return x[:, :cfg['nt_cond']].contiguous(), x[:, cfg['nt_cond']:].contiguous()


def t_random_sequence(cfg, seed, n):
    """The values np.random.randint yields inside the reference's ae_loss (train.py:72-75)
    for n consecutive steps after np.random.seed(seed)."""
    rng = np.random.RandomState(seed)
    return [step.draw_t_random(rng, cfg['nt_cond'], cfg['nt_cond'] + cfg['nt_pred'], cfg['offset'])
            for _ in range(n)]


def oracle_step(cfg, dtype=torch.float32, seed=4242):
    """One loss+backward of the oracle; returns (net, out dict, grads by name)."""
    net = oracle_net(cfg, dtype)
    cond, target = inputs(cfg, dtype)
    t_random = t_random_sequence(cfg, seed, 1)[0]
    out = step.step_losses(net, cond, target, cfg, t_random)
    out['total'].backward()
    grads = {n: p.grad for n, p in net.parameters()}
    return net, out, grads

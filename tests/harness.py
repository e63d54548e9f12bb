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


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    g['cfg'] = json.loads(str(g['cfg']))
    g['shapes'] = json.loads(str(g['shapes']))
    return g


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

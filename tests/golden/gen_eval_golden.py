"""Golden fixtures of the EVAL-mode long-horizon rollout, from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/gen_eval_golden.py       # writes tests/golden/eval/<config>.npz

Per parity configuration the reference networks are built and filled exactly as in ``gen_golden.py`` (including
non-trivial BatchNorm running statistics), switched to ``eval()`` as ``test/utils.py:8-16`` does, and under
``torch.no_grad()``

* ``get_forecast(cond, n_long)`` with a horizon three times the training one (``test/mnist/test.py:99-133``),
* the content swap ``get_forecast(cond, n_swap, init_s_code=Es(other window))`` (``README.md:112-116``),
* the same forecast restarted from ``init_t_code=Et(cond)``

are recorded as sub-sampled forecasts, exact sums, the full latent rollout and the last residual.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from gen_golden import build_reference, inputs                           # noqa: E402  (also puts /root/reference on the path)
from spatiotemporal_variable_separation_b200 import configs              # noqa: E402
from tests.summ import subsample                                         # noqa: E402

OUT = os.path.join(HERE, 'eval')


def horizons(cfg):
    n_train = cfg['nt_pred'] + cfg['offset']
    return 3 * n_train, n_train


@torch.no_grad()
def generate(name, extra='', tag=None):
    cfg = configs.preset(name, small=True, extra=extra)
    if tag:
        cfg['name'] = tag
    net = build_reference(cfg).eval()
    cond, target = inputs(cfg)
    full = torch.cat([cond, target], 1)
    n_long, n_swap = horizons(cfg)
    f_long, t_long, s_code, res = net.get_forecast(cond, n_long)
    other = full[:, -cfg['nt_cond']:]
    s_other = net.Es(other, return_skip=cfg['skipco'])
    f_swap, t_swap, _, _ = net.get_forecast(cond, n_swap, init_s_code=s_other)
    f_init, _, _, _ = net.get_forecast(cond, n_swap, init_t_code=net.Et(cond))
    assert torch.equal(f_init, f_long[:, :n_swap])
    s_first = s_code[0] if isinstance(s_code, (tuple, list)) else s_code
    out = dict(cfg=json.dumps(cfg), n_long=n_long, n_swap=n_swap,
               long_shape=np.array(f_long.shape), long_sub=subsample(f_long).astype(np.float32),
               long_sum=np.array([float(f_long.double().sum()), float(f_long.double().abs().sum())]),
               t_long=t_long.numpy(), s_code=subsample(s_first).astype(np.float32),
               last_residual=res[-1][-1].numpy(),
               swap_shape=np.array(f_swap.shape), swap_sub=subsample(f_swap).astype(np.float32),
               swap_sum=np.array([float(f_swap.double().sum()), float(f_swap.double().abs().sum())]))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, cfg['name'] + '.npz')
    np.savez_compressed(path, **out)
    print(f'{cfg["name"]}: {tuple(f_long.shape)} long, {tuple(f_swap.shape)} swapped -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)')


if __name__ == '__main__':
    torch.set_num_threads(8)
    for n in ['mnist', 'wave', 'taxibj', 'sst', 'chairs']:
        generate(n)
    generate('mnist', extra='--mixing mul --code_size_s 6 --n_blocks 2 --offset 0', tag='mnist-small-mul')
    generate('mnist', extra='--skipco', tag='mnist-small-skipco')
    generate('mnist', extra='--no_s', tag='mnist-small-no_s')

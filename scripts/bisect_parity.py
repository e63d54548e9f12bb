"""Which flag of a small configuration moves the fp32 gradient error of the CUDA path?  Runs the oracle (fp64 / fp32) on
the host and the CUDA path in fp32 for variants of the mnist-small flags; prints the worst per-tensor relative L2 errors."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatiotemporal_variable_separation_b200 import configs, ops
from tests import fullsize, harness
from tests.test_host_emulated import build_filled, run_step
from oracle import step

BASE = configs.SMALL_FLAGS['mnist']
VARIANTS = {'base': '', 'offset0': '--offset 0', 'mul': '--mixing mul --code_size_s 6', 'nblocks2': '--n_blocks 2',
            'mul+offset0': '--mixing mul --code_size_s 6 --offset 0', 'all': '--mixing mul --code_size_s 6 --offset 0 --n_blocks 2',
            'cs6': '--code_size_s 6'}
ops.set_compute_dtype(torch.float32)
import itertools
CASES = [('all', VARIANTS['all'], 7), ('all', VARIANTS['all'], 6), ('mul+nb2', '--mixing mul --code_size_s 6 --n_blocks 2', 7), ('mul+offset0', VARIANTS['mul+offset0'], 7), ('nblocks2', VARIANTS['nblocks2'], 7), ('offset0+nb2', '--offset 0 --n_blocks 2', 7)]
for name, extra, tr in CASES:
    cfg = configs.preset('mnist', small=True, extra=extra)
    cfg['name'] = 'mnist-small-mul' if name == 'all' else 'bisect-' + name
    t_random = tr if tr is not None else cfg['nt_cond'] + 1
    name = f'{name}@t{t_random}'
    runs = {}
    for dt in ('float64', 'float32'):
        net = harness.oracle_net(cfg, getattr(torch, dt))
        cond, target = harness.inputs(cfg, getattr(torch, dt))
        out = step.step_losses(net, cond, target, cfg, t_random)
        out['total'].backward()
        runs[dt] = {n: p.grad.detach() for n, p in net.parameters() if p.grad is not None}
        runs[dt + '_fwd'] = (out['t_codes'].detach(), out['forecasts'].detach())
    net = build_filled(cfg, 'cuda').train()
    out = run_step(net, cfg, t_random, 'cuda')
    out['total'].backward()
    ours = {f'{part}.{k}': p.grad.detach().cpu() for part in harness.PARTS for k, p in getattr(net, part).named_parameters() if p.grad is not None}
    fw = runs['float64_fwd']; fr = runs['float32_fwd']
    print(f'   forward rel-L2 vs fp64: t_codes ours {fullsize.rel_l2(out["t_codes"].detach().cpu(), fw[0]):.2e} ref32 {fullsize.rel_l2(fr[0], fw[0]):.2e}; forecasts ours {fullsize.rel_l2(out["forecasts"].detach().cpu(), fw[1]):.2e} ref32 {fullsize.rel_l2(fr[1], fw[1]):.2e}')
    gmax = max(float(g.norm()) for g in runs['float64'].values())
    rows = []
    for n, g64 in runs['float64'].items():
        if float(g64.norm()) < 1e-6 * gmax:
            continue
        rows.append((fullsize.rel_l2(ours[n], g64), fullsize.rel_l2(runs['float32'][n], g64), n))
    rows.sort(reverse=True)
    print(f'{name:12s} ours max {rows[0][0]:.2e} median {np.median([r[0] for r in rows]):.2e} | ref32 max {max(r[1] for r in rows):.2e}'
          f' | worst: ' + ', '.join(f'{n} {a:.1e}/{b:.1e}' for a, b, n in rows[:4]), flush=True)

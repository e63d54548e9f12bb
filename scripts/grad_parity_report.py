"""Per-golden distribution of the gradient error of the fp32 path against the reference's fp64 gradients, next to the
reference's own fp32-vs-fp64 error (tests/golden/*.npz).  `--device cuda` runs libvarsep_sm100a.so, `--device cpu` the
host logic over tests/emu.py.  Writes one JSON object (default gpurun_out/grad_parity.json)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatiotemporal_variable_separation_b200 import ops  # noqa: E402
from tests import emu, harness  # noqa: E402
from tests.summ import summarize  # noqa: E402
from tests.test_host_emulated import NAMES, build_filled, run_step  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--device', default='cuda')
ap.add_argument('--out', default='gpurun_out/grad_parity.json')
args = ap.parse_args()
ops.set_compute_dtype(torch.float32)
report = {}
for name in NAMES:
    g = harness.load_golden(name)
    cfg = g['cfg']
    ctx = emu.install() if args.device == 'cpu' else None
    if ctx is not None:
        ctx.__enter__()
    net = build_filled(cfg, args.device).train()
    t_random = harness.t_random_sequence(cfg, int(g['np_seed']), 1)[0]
    out = run_step(net, cfg, t_random, args.device)
    out['total'].backward()
    grads = {f'{part}.{k}': (p.grad.cpu() if p.grad is not None else None)
             for part in harness.PARTS for k, p in getattr(net, part).named_parameters()}
    if ctx is not None:
        ctx.__exit__(None, None, None)
    gmax = np.nanmax(g['grad64'][:, 0])
    ours, ref, names = [], [], []
    for n, r32, r64 in zip(g['grad_names'], g['grad32'], g['grad64']):
        if np.isnan(r32[0]) or r64[0] < 1e-6 * gmax:       # unused / mathematically-zero gradients
            continue
        s = summarize(str(n), grads[str(n)])
        ours.append(float(np.abs(s - r64).max() / abs(r64[0])))
        ref.append(float(np.abs(r32 - r64).max() / abs(r64[0])))
        names.append(str(n))
    ours, ref = np.array(ours), np.array(ref)
    loss = np.array([float(out[k].detach()) for k in ('ae', 's', 'pred', 't', 'total')])
    report[name] = {
        'tensors': int(len(ours)),
        'ours_vs_fp64': {'median': float(np.median(ours)), 'max': float(ours.max()), 'over_1e-4': int((ours > 1e-4).sum())},
        'reference_fp32_vs_fp64': {'median': float(np.median(ref)), 'max': float(ref.max()), 'over_1e-4': int((ref > 1e-4).sum())},
        'loss_rel_err_vs_reference_fp32': float(np.abs(loss - g['loss32']).max() / np.abs(g['loss32']).max()),
        'worst_tensors': [[names[i], float('%.3e' % ours[i]), float('%.3e' % ref[i])] for i in np.argsort(-ours)[:12]],
    }
    print(name, json.dumps(report[name]))
os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
json.dump({'device': args.device, 'metric': 'max over [norm, random projection] of |ours - ref64| / |g|_ref64, per parameter tensor',
           'goldens': report}, open(args.out, 'w'), indent=1)

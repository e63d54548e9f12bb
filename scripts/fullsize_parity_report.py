"""Measured parity of the FULL-SIZE BASELINE configurations on the GPU: the oracle run live on the host (fp64 truth, fp32
reference rounding) against the CUDA path in fp32 and bf16 — losses, forecasts, latent rollout and every parameter
gradient by relative L2 error.  Writes gpurun_out/fullsize_parity.json (copied to profiles/ per round).

    python scripts/fullsize_parity_report.py [--configs mnist wave taxibj] [--out gpurun_out/fullsize_parity.json]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fullsize  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--configs', nargs='+', default=['mnist', 'wave', 'taxibj'])
ap.add_argument('--out', default='gpurun_out/fullsize_parity.json')
ap.add_argument('--batch', type=int, default=None)
args = ap.parse_args()
report = {}
for name in args.configs:
    t0 = time.time()
    truth = fullsize.oracle_run(name, 'float64', args.batch)
    ref32 = fullsize.oracle_run(name, 'float32', args.batch)
    t1 = time.time()
    refbf = fullsize.oracle_run(name, 'autocast_bf16', args.batch)
    ref32c = fullsize.oracle_run(name, 'float32_cuda', args.batch)
    entry = {'oracle_seconds': t1 - t0, 'reference_fp32_vs_fp64': fullsize.summary(fullsize.compare(ref32, truth)),
             'reference_fp32_cudnn_vs_fp64': fullsize.summary(fullsize.compare(ref32c, truth, ref32)),
             'reference_autocast_bf16_vs_fp64': fullsize.summary(fullsize.compare(refbf, truth))}
    for label, dt, ref, refb in (('cuda_fp32_vs_fp64', torch.float32, ref32, ref32c),
                                 ('cuda_bf16_vs_fp64', torch.bfloat16, refbf, None)):
        rep = fullsize.compare(fullsize.cuda_run(name, dt, args.batch), truth, ref, refb)
        entry[label] = fullsize.summary(rep)
        entry[label]['grads_by_tensor'] = {n: [float('%.3e' % v[0]), float('%.3e' % v[1])] for n, v in rep['grads'].items()}
    report[name] = entry
    print(name, json.dumps({k: ({kk: vv for kk, vv in v.items() if kk != 'grads_by_tensor'} if isinstance(v, dict) else v)
                            for k, v in entry.items()}), flush=True)
os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
json.dump({'metric': 'relative L2 error |x - x64| / |x64| per quantity / per parameter tensor; x64 = fp64 oracle on the host',
           'configs': report}, open(args.out, 'w'), indent=1)

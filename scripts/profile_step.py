"""One eager training step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spatiotemporal_variable_separation_b200 import configs  # noqa: E402
from spatiotemporal_variable_separation_b200.data import synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='mnist')
ap.add_argument('--dtype', default='bf16')
ap.add_argument('--steps', type=int, default=1)
args = ap.parse_args()
cfg = configs.preset(args.config)
dev = torch.device('cuda', 0)
tr = bench.build_trainer(cfg, dev, torch.bfloat16 if args.dtype == 'bf16' else torch.float32, 1, False)
full = synthetic_batch(cfg, device=dev)
tr.input_buffer(full.shape, dev).copy_(full)
t_random = cfg['nt_cond'] + 2
for _ in range(2):
    tr.run(full.shape, t_random)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    tr.run(full.shape, t_random)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled', args.steps, 'step(s)')

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
tr = bench.Trainer(cfg, dev, torch.bfloat16 if args.dtype == 'bf16' else torch.float32, 1, False)
tr.full.copy_(synthetic_batch(cfg, device=dev))
for _ in range(2):
    tr.step(7)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    tr.step(7)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled', args.steps, 'step(s)')

"""Timing and stress of the shifted-window kernel on the decoder's 128 -> 64 layer (transposed, with BatchNorm sums) and
its input gradient (direct): median time at the benchmark's size, then many back-to-back launches at varying batch sizes
(odd tile counts, few items per CTA pair) with the result of every launch compared against the first."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from spatiotemporal_variable_separation_b200 import _lib as L
from tests.test_kernels_gpu import geom
dt = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]
def tensors(N, G):
    g, P, Q = geom(dt, N, 32, 32, 64, 128, 4, 2, 1, groups=G)
    torch.manual_seed(N)
    return (g, torch.randn(N, 32, 32, 64, device='cuda').to(dt), torch.randn(N, 16, 16, 128, device='cuda').to(dt),
            (torch.randn(64, 16, 128, device='cuda') / 45).to(dt), (torch.randn(128, 16, 64, device='cuda') / 32).to(dt),
            torch.zeros(G * 64 * 2, dtype=torch.float64, device='cuda'))
g, big, small, w_tr, w_di, stats = tensors(1664, 13)
t_tr = timeit(lambda: L.call('vs_conv_forward', g, L.TRANSPOSED, small, w_tr, None, big, stats, L.stream()))
t_di = timeit(lambda: L.call('vs_conv_forward', g, L.DIRECT, big, w_di, None, small, None, L.stream()))
print(f'N=1664: transposed+stats {t_tr:7.1f} us   direct {t_di:7.1f} us', flush=True)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for N, G in ((1664, 13), (301, 7), (75, 5), (37, 1), (150, 3), (999, 9)):
    g, big, small, w_tr, w_di, stats = tensors(N, G)
    out_tr, out_di = torch.empty_like(big), torch.empty_like(small)
    ref, dev = None, 0.0
    for i in range(reps):
        stats.zero_()
        L.call('vs_conv_forward', g, L.TRANSPOSED, small, w_tr, None, out_tr, stats, L.stream())
        L.call('vs_conv_forward', g, L.DIRECT, big, w_di, None, out_di, None, L.stream())
        if i % 20 == 0:
            torch.cuda.synchronize()
            cur = (out_tr.clone(), out_di.clone(), stats.clone())
            if ref is None: ref = cur
            assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]), (N, i)
            dev = max(dev, float(((cur[2] - ref[2]).abs() / ref[2].abs().clamp_min(1.0)).max()))
    torch.cuda.synchronize()
    print(f'N={N}: {reps} launches of each mode, outputs identical, BatchNorm sums within {dev:.1e} (fp32 partial sums, atomic order)', flush=True)

"""Time the latent rollout kernels alone (CUDA events), with and without the buffers saved for backward."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatiotemporal_variable_separation_b200 import _lib as L
from spatiotemporal_variable_separation_b200._lib import ptr
T, B, d, h, nb = 15, 128, 20, 512, 1
dev = 'cuda'
torch.manual_seed(0)
ws = [torch.randn(h, d, device=dev) * 0.1, torch.zeros(h, device=dev), torch.randn(h, h, device=dev) * 0.04, torch.zeros(h, device=dev),
      torch.randn(d, h, device=dev) * 0.04, torch.zeros(d, device=dev)]
wt = [ws[0].t().contiguous(), ws[1], ws[2].t().contiguous(), ws[3], ws[4].t().contiguous(), ws[5]]
codes = torch.randn(T, B, d, device=dev)
hidden = torch.empty(nb, 2, T - 1, B, h, device=dev); xin = torch.empty(nb, T - 1, B, d, device=dev); res = torch.empty_like(xin)
dres = torch.empty_like(xin); dhid = torch.empty_like(hidden)
pa, pat = L.pointer_array(ws), L.pointer_array(wt)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print('fwd with saves    %.1f us' % timeit(lambda: L.call('vs_latent_rollout_forward', ptr(codes), pa, T, B, d, h, nb, ptr(hidden), ptr(xin), ptr(res), L.stream())))
print('fwd without saves %.1f us' % timeit(lambda: L.call('vs_latent_rollout_forward', ptr(codes), pa, T, B, d, h, nb, None, None, None, L.stream())))
dc = torch.randn(T, B, d, device=dev)
print('bwd               %.1f us' % timeit(lambda: L.call('vs_latent_rollout_backward', ptr(dc), pat, T, B, d, h, nb, ptr(hidden), ptr(dres), ptr(dhid), L.stream())))
for Bv in (8, 64):
    c2 = torch.randn(T, Bv, d, device=dev)
    print('fwd without saves, B=%d  %.1f us' % (Bv, timeit(lambda: L.call('vs_latent_rollout_forward', ptr(c2), pa, T, Bv, d, h, nb, None, None, None, L.stream()))))

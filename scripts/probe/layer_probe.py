import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from spatiotemporal_variable_separation_b200 import _lib as L
from tests import emu
from tests.test_kernels_gpu import run_both, geom
from tests.fullsize import rel_l2
torch.manual_seed(0)
for N, G in ((20, 5), (40, 10), (20, 4), (24, 6), (16, 4)):
    rows, C = N * 1024, 8
    y = (torch.randn(rows, C) * 2 + 0.5)
    dout = torch.randn(rows, C)
    mean, invstd = torch.randn(G * C) * 0.1 + 0.5, torch.rand(G * C) + 0.3
    gamma, beta = torch.randn(C) * 0.1 + 1, torch.randn(C) * 0.1
    sums = torch.zeros(G * C * 2, dtype=torch.float64)
    gpu, cpu = run_both('vs_bn_act_backward_reduce', [dout, y, 0, rows, C, G, mean, invstd, gamma, beta, 2, sums, None])
    e_red = rel_l2(gpu[11].float(), cpu[11].float())
    dy = torch.zeros(rows, C)
    gpu2, cpu2 = run_both('vs_bn_act_backward_apply', [dout, y, dy, 0, rows, C, G, mean, invstd, gamma, beta, 2, cpu[11], 1, torch.zeros(C), torch.zeros(C), None])
    e_app = rel_l2(gpu2[2], cpu2[2])
    # dgrad of ConvTranspose2d(16, 8, 4, 2, 1): direct conv of dy [N,32,32,8] -> [N,16,16,16]
    g, P, Q = geom(torch.float32, N, 32, 32, 8, 16, 4, 2, 1)
    wp = torch.randn(16, 16, 8) / 11
    out = torch.zeros(N, P, Q, 16)
    gpu3, cpu3 = run_both('vs_conv_forward', [g, L.DIRECT, cpu2[2].reshape(N, 32, 32, 8), wp, None, out, None, None])
    e_dg = rel_l2(gpu3[5], cpu3[5])
    # fprop of the same layer with statistics
    gt, _, _ = geom(torch.float32, N, 32, 32, 8, 16, 4, 2, 1, groups=G)
    x = torch.randn(N, 16, 16, 16)
    wpt = torch.randn(8, 16, 16) / 16
    stats = torch.zeros(G * 8 * 2, dtype=torch.float64)
    gpu4, cpu4 = run_both('vs_conv_forward', [gt, L.TRANSPOSED, x, wpt, torch.randn(8), torch.zeros(N, 32, 32, 8), stats, None])
    print(f'N={N} G={G}: reduce {e_red:.2e} apply {e_app:.2e} dgrad {e_dg:.2e} fprop {rel_l2(gpu4[5], cpu4[5]):.2e} stats {rel_l2(gpu4[6].float(), cpu4[6].float()):.2e} path {L.load().vs_conv_forward_path(g, L.DIRECT)}', flush=True)

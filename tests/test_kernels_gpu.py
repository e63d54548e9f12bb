"""Every entry point of libvarsep_sm100a.so against the executable specification (tests/emu.py),
argument for argument, on the GPU.  fp32 storage: 2e-5 relative (fp32 accumulation, different
summation order); bf16 storage: outputs are compared after rounding the specification's result to
bf16 (1 ulp = 2^-8 relative).  Element-wise comparisons tolerate a 1e-4 fraction of outliers in the
kernels whose result depends on the sign of a rounded pre-activation (a LeakyReLU/ReLU unit whose
pre-activation rounds to the other side of zero legitimately takes the other slope).
"""
import ctypes

import numpy as np
import pytest
import torch

from spatiotemporal_variable_separation_b200 import _lib as L
from tests import emu

pytestmark = pytest.mark.gpu
DT = {'f32': torch.float32, 'bf16': torch.bfloat16}


def run_both(name, args):
    """args: list of CPU tensors / scalars / Geom.  Runs CUDA on device copies and the emulator on CPU
    copies; returns (gpu tensors moved back, cpu tensors) in argument order."""
    cpu = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
    gpu = [a.cuda() if isinstance(a, torch.Tensor) else a for a in args]
    L.call(name, *gpu[:-1], L.stream())
    torch.cuda.synchronize()
    emu.emu_call(name, *cpu)
    return [g.cpu() if isinstance(g, torch.Tensor) else g for g in gpu], cpu


def close(a, b, dtype, what, outliers=0.0, scale=None):
    a, b = a.double().flatten(), b.double().flatten()
    if dtype == torch.bfloat16:
        rtol = 2 ** -7
    else:
        rtol = 2e-5
    ref = float(b.abs().max()) if scale is None else scale
    bad = (a - b).abs() > rtol * b.abs() + rtol * ref * 1e-1 + 1e-30
    frac = float(bad.double().mean())
    assert frac <= outliers, f'{what}: {frac:.2e} of elements off (max err {float((a - b).abs().max()):.3e}, ref max {ref:.3e})'


def geom(dtype, N, H, W, C, K, R, stride, pad, groups=1, act=0, flags=0):
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    return L.Geom(0 if dtype == torch.float32 else 1, N, H, W, C, P, Q, K, R, R, stride, pad, groups, act, flags), P, Q


# (N, H, W, C, K, R, stride, pad, groups): every conv geometry class of the five configurations
CONV_CASES = [
    (6, 16, 16, 16, 32, 4, 2, 1, 2),      # DCGAN k4 s2 p1
    (4, 64, 64, 5, 8, 4, 2, 1, 1),        # first encoder layer, C % 4 != 0
    (8, 4, 4, 24, 20, 4, 1, 0, 2),        # 4x4 valid conv == Flatten+Linear / first_upconv
    (9, 1, 1, 37, 50, 1, 1, 0, 3),        # Linear, odd sizes
    (4, 12, 12, 16, 24, 3, 1, 1, 2),      # VGG / SST 3x3
    (2, 64, 64, 15, 16, 5, 2, 3, 1),      # ResNet18 stem, 64 -> 33
    (4, 17, 17, 8, 16, 3, 2, 1, 2),       # ResNet18 stride-2 block, odd size
    (4, 17, 17, 8, 16, 1, 2, 0, 1),       # ResNet18 1x1 s2 downsample
    (4, 3, 3, 32, 12, 3, 1, 0, 1),        # ResNet18 conv_out
    (4, 32, 32, 4, 64, 4, 2, 1, 4),       # last DCGAN decoder layer seen from the big side (C small)
    (130, 8, 8, 8, 8, 4, 2, 1, 2),        # rows not a multiple of the tile, groups of 65 samples
    # "thin" streaming kernels: a handful of channels on one side
    (8, 64, 64, 1, 64, 4, 2, 1, 2),       # last DCGAN decoder layer (nc=1) / its dgrad / its wgrad
    (4, 64, 64, 3, 64, 4, 2, 1, 1),       # same for nc=3 (chairs)
    (8, 64, 64, 5, 64, 4, 2, 1, 2),       # first DCGAN encoder layer (nt_cond*nc = 5)
    (6, 32, 32, 2, 64, 3, 1, 1, 2),       # last VGG decoder layer (ConvT 64->2, k3 s1 p1)
    # tcgen05 over a CTA-built im2col tile: partial pixel boxes, K below / above one 64-channel tile, two kk chunks
    (12, 40, 40, 3, 48, 4, 2, 1, 2),
    (8, 32, 32, 6, 128, 4, 2, 1, 2),
    (20, 16, 16, 7, 72, 3, 1, 1, 1),
    # transposed direction of these: GEMM + col2im kernel (nc = 2, two 64-channel chunks, non-square image)
    (6, 32, 32, 2, 128, 4, 2, 1, 1),
    (5, 16, 64, 1, 64, 4, 2, 1, 1),
    (33, 1, 1, 148, 24, 4, 1, 0, 3),      # placeholder geometry replaced below
]
# first_upconv: ConvTranspose k4 s1 p0 of a 1x1 input (big side 4x4, small side 1x1)
CONV_CASES[-1] = (33, 4, 4, 24, 148, 4, 1, 0, 3)


@pytest.mark.parametrize('dt', ['f32', 'bf16'])
@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('mode', [L.DIRECT, L.TRANSPOSED])
def test_conv_forward(case, mode, dt):
    N, H, W, C, K, R, stride, pad, G = case
    dtype = DT[dt]
    torch.manual_seed(1)
    g, P, Q = geom(dtype, N, H, W, C, K, R, stride, pad, groups=G)
    if mode == L.DIRECT:
        x, OC, IC, oshape = torch.randn(N, H, W, C), K, C, (N, P, Q, K)
    else:
        x, OC, IC, oshape = torch.randn(N, P, Q, K), C, K, (N, H, W, C)
    wp = torch.randn(OC, R * R, IC) / np.sqrt(IC * R * R)
    bias = torch.randn(OC)
    out = torch.zeros(oshape)
    stats = torch.zeros(G * OC * 2, dtype=torch.float64)
    x, wp, out = x.to(dtype), wp.to(dtype), out.to(dtype)
    gpu, cpu = run_both('vs_conv_forward', [g, mode, x, wp, bias, out, stats, None])
    close(gpu[5], cpu[5], dtype, 'conv out')
    # statistics: of the fp32 accumulator (fused epilogue of the CUDA-core kernel) or, for the kernels
    # that take them in a second pass, of the values they stored (identical for fp32 storage)
    try:
        close(gpu[6].float(), cpu[6].float(), torch.float32, 'bn stats', scale=float(cpu[6].abs().max()))
    except AssertionError:
        yq = gpu[5].float().reshape(G, -1, OC).double()
        want = torch.stack([yq.sum(1), (yq * yq).sum(1)], -1).reshape(-1)
        close(gpu[6].float(), want.float(), torch.float32, 'bn stats of stored output', scale=float(want.abs().max()))
    # fused activation epilogue, no stats
    for act in (2, 4):
        g2, _, _ = geom(dtype, N, H, W, C, K, R, stride, pad, groups=1, act=act)
        gpu, cpu = run_both('vs_conv_forward', [g2, mode, x, wp, bias, out.clone(), None, None])
        close(gpu[5], cpu[5], dtype, f'conv out act={act}', outliers=1e-4)
    # no bias
    gpu, cpu = run_both('vs_conv_forward', [g2, mode, x, wp, None, out.clone(), None, None])
    close(gpu[5], cpu[5], dtype, 'conv out no bias', outliers=1e-4)


# geometries eligible for the tcgen05 path (channel counts multiples of 64): DCGAN k4 s2 p1 in both
# directions, 3x3 stride 1, tiles spanning several images, partial tiles, 2 output-channel tiles
TC_CASES = [
    (16, 16, 16, 64, 128, 4, 2, 1, 2),
    (8, 8, 8, 128, 256, 4, 2, 1, 2),
    (6, 32, 32, 64, 64, 4, 2, 1, 3),
    (4, 16, 16, 64, 64, 3, 1, 1, 2),
    (2, 32, 32, 128, 64, 3, 1, 1, 1),
    (130, 8, 8, 64, 64, 4, 2, 1, 2),
    (3, 64, 64, 64, 64, 3, 1, 1, 1),
    (5, 4, 4, 64, 128, 3, 1, 1, 1),
    # channel counts that are only multiples of 8: partial 64-channel chunks and tail tiles in the output channels
    (4, 16, 16, 72, 96, 3, 1, 1, 2),
    (8, 8, 8, 40, 200, 4, 2, 1, 1),
    (64, 1, 1, 1200, 1200, 1, 1, 0, 2),      # the 1200-wide linear layers of the WaveEq MLP configuration
    (6, 1, 1, 20480, 1200, 1, 1, 0, 2),      # its first encoder layer
    # CTA-pair (cta_group::2) kernels: 256-column tiles, odd number of pixel tiles, partial last 256-channel tile
    (4, 8, 8, 256, 512, 4, 2, 1, 2),
    (6, 8, 8, 128, 256, 3, 1, 1, 3),
    (10, 4, 4, 512, 256, 4, 2, 1, 1),
    # partial pixel tiles (12 of 16 columns, 4 of 8 rows) with BatchNorm statistics fused in the staged epilogue
    (6, 12, 12, 64, 64, 3, 1, 1, 1),
    # shifted-window CTA-pair kernel (k4 s2 p1, 64 / 128 output channels, class grid of 16-row x 8-column tiles): the
    # decoder's 128 -> 64 layer and its input gradient, an odd tile count with one image per BatchNorm group, two tile rows
    (5, 32, 32, 64, 128, 4, 2, 1, 1),
    (3, 32, 16, 64, 128, 4, 2, 1, 3),
    (4, 64, 32, 64, 128, 4, 2, 1, 2),
]


@pytest.mark.parametrize('case', TC_CASES)
@pytest.mark.parametrize('mode', [L.DIRECT, L.TRANSPOSED])
def test_conv_forward_tensor_core(case, mode):
    """bf16 tcgen05 kernel against (a) the specification and (b) the CUDA-core kernel on the same inputs."""
    N, H, W, C, K, R, stride, pad, G = case
    dtype = torch.bfloat16
    torch.manual_seed(7)
    g, P, Q = geom(dtype, N, H, W, C, K, R, stride, pad, groups=G)
    if mode == L.DIRECT:
        x, OC, IC, oshape = torch.randn(N, H, W, C), K, C, (N, P, Q, K)
    else:
        x, OC, IC, oshape = torch.randn(N, P, Q, K), C, K, (N, H, W, C)
    wp = (torch.randn(OC, R * R, IC) / np.sqrt(IC * R * R)).to(dtype)
    bias = torch.randn(OC)
    x = x.to(dtype)
    out = torch.full(oshape, float('nan')).to(dtype)
    stats = torch.zeros(G * OC * 2, dtype=torch.float64)
    gpu, cpu = run_both('vs_conv_forward', [g, mode, x, wp, bias, out, stats, None])
    close(gpu[5], cpu[5], dtype, 'tc conv out')
    # statistics: fused in the epilogue from the fp32 accumulators when every tile lies in one BatchNorm group,
    # otherwise taken in a second pass from the bf16 values that were stored
    try:
        close(gpu[6].float(), cpu[6].float(), torch.float32, 'tc bn stats (fused)', scale=float(cpu[6].abs().max()))
    except AssertionError:
        yq = gpu[5].float().reshape(G, -1, OC).double()
        want = torch.stack([yq.sum(1), (yq * yq).sum(1)], -1).reshape(-1)
        close(gpu[6].float(), want.float(), torch.float32, 'tc bn stats (second pass)', scale=float(want.abs().max()))
    g2, _, _ = geom(dtype, N, H, W, C, K, R, stride, pad, act=2, flags=L.FLAG_FORCE_SIMT)
    g3, _, _ = geom(dtype, N, H, W, C, K, R, stride, pad, act=2)
    xs, ws, bs = x.cuda(), wp.cuda(), bias.cuda()
    o_simt, o_tc = torch.zeros(oshape, dtype=dtype, device='cuda'), torch.zeros(oshape, dtype=dtype, device='cuda')
    L.call('vs_conv_forward', g2, mode, xs, ws, bs, o_simt, None, L.stream())
    L.call('vs_conv_forward', g3, mode, xs, ws, bs, o_tc, None, L.stream())
    torch.cuda.synchronize()
    close(o_tc.cpu(), o_simt.cpu(), dtype, 'tc vs simt', outliers=1e-4)


@pytest.mark.parametrize('mode', [L.DIRECT, L.TRANSPOSED])
def test_conv_forward_many_tiles(mode):
    """Several work items per CTA pair of the shifted-window kernel (ring and accumulator stages wrap, statistics
    flushed between BatchNorm groups): the decoder's 128 -> 64 layer / its input gradient on 300 images."""
    test_conv_forward_tensor_core((300, 32, 32, 64, 128, 4, 2, 1, 2), mode)


@pytest.mark.timeout(1800)
@pytest.mark.parametrize('env', [{'VARSEP_DISABLE_SHIFT': '1'},
                                 {'VARSEP_DISABLE_SHIFT': '1', 'VARSEP_RESIDENT_OC': '3', 'VARSEP_RESIDENT_MIN_ITEMS': '1'}],
                         ids=['per_class', 'resident'])
def test_conv_forward_alternative_kernels(env):
    """Every tensor-core case again in a fresh process with the kernels the default selection no longer reaches: the
    per-class kernels on the layers the shifted-window kernel takes, and the opt-in resident-weight variant of the
    CTA-pair kernel on every eligible layer whatever its size."""
    import os
    import subprocess
    import sys
    r = subprocess.run([sys.executable, '-m', 'pytest', __file__, '-x', '-q', '-m', 'gpu', '-k', 'test_conv_forward_tensor_core',
                        '-p', 'no:cacheprovider'], env=dict(os.environ, **env), capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize('dt', ['f32', 'bf16'])
@pytest.mark.parametrize('case', CONV_CASES + TC_CASES)
def test_conv_wgrad_and_pack(case, dt):
    N, H, W, C, K, R, stride, pad, G = case
    dtype = DT[dt]
    torch.manual_seed(2)
    g, P, Q = geom(dtype, N, H, W, C, K, R, stride, pad)
    small, big = torch.randn(N, P, Q, K).to(dtype), torch.randn(N, H, W, C).to(dtype)
    dw = torch.randn(K, C, R, R)      # accumulate semantics: starts non-zero
    gpu, cpu = run_both('vs_conv_wgrad', [g, small, big, dw, None])
    close(gpu[3], cpu[3], torch.float32, 'wgrad', scale=float(cpu[3].abs().max()))
    if dtype == torch.bfloat16:     # tensor-core kernel (when eligible) against the CUDA-core kernel
        g2, _, _ = geom(dtype, N, H, W, C, K, R, stride, pad, flags=L.FLAG_FORCE_SIMT)
        d1, d2 = torch.zeros(K, C, R, R, device='cuda'), torch.zeros(K, C, R, R, device='cuda')
        L.call('vs_conv_wgrad', g, small.cuda(), big.cuda(), d1, L.stream())
        L.call('vs_conv_wgrad', g2, small.cuda(), big.cuda(), d2, L.stream())
        torch.cuda.synchronize()
        close(d1.cpu(), d2.cpu(), torch.float32, 'tc wgrad vs simt', scale=float(d2.abs().max()))
    w = torch.randn(K, C, R, R)
    for swap in (0, 1):
        out = torch.zeros(K * C * R * R).to(dtype)
        gpu, cpu = run_both('vs_pack_weight', [w, out, g.dtype, K, C, R * R, swap, None])
        assert torch.equal(gpu[1], cpu[1])


@pytest.mark.parametrize('dt', ['f32', 'bf16'])
@pytest.mark.parametrize('rows,C,G', [(96, 64, 3), (130, 20, 2), (4096, 128, 4), (50, 1, 5), (64, 7, 1)])
def test_bn_kernels(rows, C, G, dt):
    dtype = DT[dt]
    torch.manual_seed(3)
    y = (torch.randn(rows, C) * 2 + 0.5).to(dtype)
    stats = torch.stack([y.float().reshape(G, -1, C).double().sum(1),
                         (y.float().reshape(G, -1, C).double() ** 2).sum(1)], -1).reshape(-1)
    mean, invstd = torch.zeros(G * C), torch.zeros(G * C)
    rmean, rvar = torch.randn(C), torch.rand(C) + 0.5
    nbt = torch.zeros((), dtype=torch.int64) + 3
    gpu, cpu = run_both('vs_bn_finalize', [stats, G, C, rows // G, 1e-5, 0.1, mean, invstd, rmean, rvar, nbt, None])
    for i, n in ((6, 'mean'), (7, 'invstd'), (8, 'running_mean'), (9, 'running_var')):
        close(gpu[i], cpu[i], torch.float32, n)
    assert int(gpu[10]) == int(cpu[10]) == 3 + G
    mean, invstd = cpu[6], cpu[7]
    gamma, beta = torch.randn(C) * 0.1 + 1, torch.randn(C) * 0.1
    # the fused form (finalize + normalise + activation in one launch where the column kernel applies)
    m2_, i2_ = torch.zeros(G * C), torch.zeros(G * C)
    gpu, cpu2 = run_both('vs_bn_finalize_act_forward', [stats, G, C, rows // G, 1e-5, 0.1, m2_, i2_, rmean, rvar, nbt, y,
                                                        torch.zeros(rows, C).to(dtype), L.dtype_code(y), rows, gamma, beta, 2, None])
    for i, n in ((6, 'mean'), (7, 'invstd'), (8, 'running_mean'), (9, 'running_var')):
        close(gpu[i], cpu2[i], torch.float32, 'fused ' + n)
    assert int(gpu[10]) == int(cpu2[10]) == 3 + G
    close(gpu[12], cpu2[12], dtype, 'fused bn_act_forward', outliers=1e-4)
    for act in (0, 1, 2, 3, 4, 5):
        out = torch.zeros(rows, C).to(dtype)
        gpu, cpu = run_both('vs_bn_act_forward', [y, out, L.dtype_code(y), rows, C, G, mean, invstd, gamma, beta, act, None])
        close(gpu[1], cpu[1], dtype, f'bn_act_forward act={act}', outliers=1e-4)
        dout = torch.randn(rows, C).to(dtype)
        sums = torch.zeros(G * C * 2, dtype=torch.float64)
        gpu, cpu = run_both('vs_bn_act_backward_reduce',
                            [dout, y, L.dtype_code(y), rows, C, G, mean, invstd, gamma, beta, act, sums, None])
        close(gpu[11].float(), cpu[11].float(), torch.float32, f'bn reduce act={act}', scale=float(cpu[11].abs().max()))
        sums = cpu[11]
        for train in (1, 0):
            dy = torch.zeros(rows, C).to(dtype)
            dgamma, dbeta = torch.randn(C), torch.randn(C)
            gpu, cpu = run_both('vs_bn_act_backward_apply',
                                [dout, y, dy, L.dtype_code(y), rows, C, G, mean, invstd, gamma, beta, act, sums, train,
                                 dgamma, dbeta, None])
            close(gpu[2], cpu[2], dtype, f'bn apply act={act} train={train}', outliers=1e-4, scale=float(cpu[2].abs().max()))
            close(gpu[14], cpu[14], torch.float32, 'dgamma')
            close(gpu[15], cpu[15], torch.float32, 'dbeta')
    m2, i2 = torch.zeros(C), torch.zeros(C)
    gpu, cpu = run_both('vs_bn_eval_stats', [rmean, rvar, C, 1e-5, m2, i2, None])
    close(gpu[4], cpu[4], torch.float32, 'eval mean')
    close(gpu[5], cpu[5], torch.float32, 'eval invstd')


@pytest.mark.parametrize('dt', ['f32', 'bf16'])
def test_elementwise_kernels(dt):
    dtype = DT[dt]
    code = 0 if dtype == torch.float32 else 1
    torch.manual_seed(4)
    n = 10007
    a, b = torch.randn(n).to(dtype), torch.randn(n).to(dtype)
    for act in range(6):
        gpu, cpu = run_both('vs_add_act', [a, b, torch.zeros(n).to(dtype), code, n, act, None])
        close(gpu[2], cpu[2], dtype, f'add_act {act}', outliers=1e-4)
        gpu, cpu = run_both('vs_act_backward', [a, cpu[2], torch.zeros(n).to(dtype), code, n, act, None])
        close(gpu[2], cpu[2], dtype, f'act_backward {act}', outliers=1e-4)
    # concat with broadcast + its backward
    rows, src_rows = 24, 8
    src, dst = torch.randn(src_rows, 5).to(dtype), torch.randn(rows, 12).to(dtype)
    gpu, cpu = run_both('vs_copy_channels', [src, 5, src_rows, dst, 12, 4, rows, code, None])
    assert torch.equal(gpu[3], cpu[3])
    gpu, cpu = run_both('vs_slice_channels_reduce', [dst, 12, 4, rows, torch.zeros(src_rows, 5).to(dtype), 5, src_rows, code, None])
    close(gpu[4], cpu[4], dtype, 'slice_reduce')
    # mul mixing
    s, t = torch.randn(src_rows, 6).to(dtype), torch.randn(rows, 6).to(dtype)
    gpu, cpu = run_both('vs_mul_bcast', [s, src_rows, t, torch.zeros(rows, 6).to(dtype), rows, 6, code, None])
    close(gpu[3], cpu[3], dtype, 'mul_bcast')
    d = torch.randn(rows, 6).to(dtype)
    gpu, cpu = run_both('vs_mul_bcast_backward', [d, s, src_rows, t, torch.zeros(src_rows, 6).to(dtype),
                                                  torch.zeros(rows, 6).to(dtype), rows, 6, code, None])
    close(gpu[4], cpu[4], dtype, 'mul_bcast ds')
    close(gpu[5], cpu[5], dtype, 'mul_bcast dt')
    # pooling (with exact ties, as after a ReLU) and upsampling
    for (N, H, W, C, k, st, pad) in [(2, 8, 8, 6, 2, 2, 0), (2, 33, 33, 5, 3, 2, 1)]:
        x = torch.relu(torch.randn(N, H, W, C)).to(dtype)
        P = (H + 2 * pad - k) // st + 1
        gpu, cpu = run_both('vs_maxpool_forward', [x, torch.zeros(N, P, P, C).to(dtype), code, N, H, W, C, k, st, pad, None])
        assert torch.equal(gpu[1], cpu[1])
        dy = torch.randn(N, P, P, C).to(dtype)
        gpu, cpu = run_both('vs_maxpool_backward', [x, dy, torch.zeros(N, H, W, C).to(dtype), code, N, H, W, C, k, st, pad, None])
        # ties between exact zeros may route to a different zero; the ReLU upstream kills those anyway
        mask = (x.float() > 0)
        close(gpu[2].float() * mask, cpu[2].float() * mask, dtype, 'maxpool_backward')
    x = torch.randn(2, 5, 7, 3).to(dtype)
    gpu, cpu = run_both('vs_upsample2_forward', [x, torch.zeros(2, 10, 14, 3).to(dtype), code, 2, 5, 7, 3, None])
    assert torch.equal(gpu[1], cpu[1])
    dy = torch.randn(2, 10, 14, 3).to(dtype)
    gpu, cpu = run_both('vs_upsample2_backward', [dy, torch.zeros(2, 5, 7, 3).to(dtype), code, 2, 5, 7, 3, None])
    close(gpu[1], cpu[1], dtype, 'upsample2_backward')
    # layout boundary
    fr = torch.randn(3, 7, 2, 6, 5)
    gpu, cpu = run_both('vs_frames_to_nhwc', [fr, 3, 7, 2, 6, 5, 2, 4, torch.zeros(3, 6, 5, 8).to(dtype), code, None])
    assert torch.equal(gpu[8], cpu[8])
    xi = torch.randn(3, 6, 5, 4).to(dtype)
    gpu, cpu = run_both('vs_nhwc_to_nchw', [xi, code, torch.zeros(3, 4, 6, 5), 3, 4, 6, 5, None])
    assert torch.equal(gpu[2], cpu[2])
    gpu, cpu = run_both('vs_nchw_to_nhwc', [cpu[2], torch.zeros(3, 6, 5, 4).to(dtype), code, 3, 4, 6, 5, None])
    assert torch.equal(gpu[1], cpu[1])
    # ... the block-staged frame folding (full 256-pixel blocks + a tail block), windows of 5 and of 15 folded channels
    for (B_, T_, Cf, H_, W_, t0, nt) in [(5, 9, 1, 24, 24, 3, 5), (3, 8, 3, 20, 17, 2, 5), (2, 6, 5, 16, 16, 1, 4)]:
        fr = torch.randn(B_, T_, Cf, H_, W_)
        gpu, cpu = run_both('vs_frames_to_nhwc', [fr, B_, T_, Cf, H_, W_, t0, nt, torch.zeros(B_, H_, W_, nt * Cf).to(dtype), code, None])
        assert torch.equal(gpu[8], cpu[8]), (B_, T_, Cf, H_, W_, t0, nt)
    # ... single-channel / 1x1 tensors, where the layout change is a cast (vector body + scalar tail)
    for (N_, C_, H_, W_) in [(7, 1, 33, 31), (9, 37, 1, 1), (4, 1, 64, 64)]:
        xi = torch.randn(N_, H_, W_, C_).to(dtype)
        gpu, cpu = run_both('vs_nhwc_to_nchw', [xi, code, torch.zeros(N_, C_, H_, W_), N_, C_, H_, W_, None])
        assert torch.equal(gpu[2], cpu[2])
        gpu, cpu = run_both('vs_nchw_to_nhwc', [cpu[2], torch.zeros(N_, H_, W_, C_).to(dtype), code, N_, C_, H_, W_, None])
        assert torch.equal(gpu[1], cpu[1])
    # column sums (general, narrow, and the flat single-column form with a tail)
    # ... and the vectorised wide form (>= 4096 rows, 16-byte channel groups), with a row count that is not a multiple of anything
    for rows, C_ in [(1000, 37), (5000, 3), (70001, 1), (100, 1), (40003, 64), (8191, 128), (5000, 8)]:
        m = torch.randn(rows, C_).to(dtype)
        gpu, cpu = run_both('vs_colsum', [m, code, rows, C_, torch.randn(C_), None])
        close(gpu[4], cpu[4], torch.float32, 'colsum', scale=float(cpu[4].abs().max()) + float(m.float().abs().sum(0).max()) * 1e-3)


def test_loss_and_adam_kernels():
    torch.manual_seed(5)
    B, T, Ln = 4, 5, 96
    a = torch.randn(T, B, Ln).transpose(0, 1)            # [B,T,L] view, strides (L, B*L, 1)
    full = torch.randn(B, 9, Ln)
    b = full[:, 4:]
    abase, fbase = a.transpose(0, 1).contiguous().reshape(-1), full.reshape(-1)
    acc = torch.zeros(2, dtype=torch.float64)
    # the views are described by explicit strides; pass the base buffers (b starts 4 frames in)
    boff = fbase[4 * Ln:]
    gpu, cpu = run_both('vs_sqdiff_sum', [abase, Ln, B * Ln, boff, 9 * Ln, Ln, B, T, Ln, acc[0:1], None])
    want = float(((a - b) ** 2).double().sum())
    assert abs(float(gpu[9][0]) - want) < 1e-6 * want and abs(float(cpu[9][0]) - want) < 1e-6 * want
    gpu, cpu = run_both('vs_sqdiff_sum', [abase, Ln, B * Ln, None, 0, 0, B, T, Ln, acc[1:2], None])
    assert abs(float(gpu[9][0]) - float((a ** 2).double().sum())) < 1e-6 * want
    # rows that are not a multiple of four floats take the element-wise form
    a7, b7 = torch.randn(3, 2, 7), torch.randn(3, 2, 7)
    acc7 = torch.zeros(1, dtype=torch.float64)
    gpu, cpu = run_both('vs_sqdiff_sum', [a7.reshape(-1), 14, 7, b7.reshape(-1), 14, 7, 3, 2, 7, acc7, None])
    want7 = float(((a7 - b7) ** 2).double().sum())
    assert abs(float(gpu[9][0]) - want7) < 1e-6 * want7 and abs(float(cpu[9][0]) - want7) < 1e-6 * want7
    gt = torch.tensor([0.5, 2.0])
    for accumulate in (0, 1):
        da = torch.randn(T * B * Ln)
        gpu, cpu = run_both('vs_sqdiff_backward', [abase, Ln, B * Ln, boff, 9 * Ln, Ln, B, T, Ln, 0.25, gt[0:1], gt[1:2],
                                                   45.0, da, accumulate, None])
        close(gpu[13], cpu[13], torch.float32, 'sqdiff_backward')
    accd = torch.tensor([3.0, 5.0, 7.0], dtype=torch.float64)
    coef = (ctypes.c_double * 3)(0.5, 0.25, 2.0)
    lamb = (ctypes.c_double * 3)(10.0, 45.0, 0.001)
    terms = torch.zeros(4)
    gpu, cpu = run_both('vs_loss_combine', [accd, coef, lamb, 3, terms, None])
    close(gpu[4], cpu[4], torch.float32, 'loss_combine')
    # Adam: 3 steps, host and device step counters, odd length (tail path)
    n = 4 * 1000 + 3
    p, g = torch.randn(n + 1)[:n].clone(), torch.randn(n)
    P = torch.zeros(n + 5)
    for use_dev in (False, True):
        pg, gg, mg, vg = [torch.zeros(n + 1).cuda() for _ in range(4)]
        pg[:n], gg[:n] = p.cuda(), g.cuda()
        pc, mc, vc = p.clone(), torch.zeros(n), torch.zeros(n)
        step_dev = torch.zeros(1, dtype=torch.int32).cuda()
        lr_dev = torch.full((1,), 4e-4, dtype=torch.float32).cuda()
        for step in (1, 2, 3):
            step_dev += 1
            # device-resident counter AND learning rate (the host lr argument is then ignored)
            L.call('vs_adam_step', pg, gg, mg, vg, n, 123.0 if use_dev else 4e-4, 0.5, 0.99, 1e-8, 0.5,
                   0 if use_dev else step, step_dev if use_dev else None, lr_dev if use_dev else None, L.stream())
            emu.emu_call('vs_adam_step', pc, g, mc, vc, n, 4e-4, 0.5, 0.99, 1e-8, 0.5, step, None, None, None)
        torch.cuda.synchronize()
        close(pg[:n].cpu(), pc, torch.float32, 'adam param')
        close(vg[:n].cpu(), vc, torch.float32, 'adam v')
        assert float(pg[n]) == 0.0


def test_error_reporting_and_launch_count():
    before = L.launch_count()
    x = torch.zeros(8, device='cuda')
    L.call('vs_add_act', x, x, x, 0, 8, 0, L.stream())
    assert L.launch_count() == before + 1
    bad = L.Geom(0, 2, 8, 8, 4, 5, 5, 4, 3, 3, 1, 1, 1, 0, 0)     # P/Q inconsistent
    with pytest.raises(RuntimeError, match='inconsistent'):
        L.call('vs_conv_forward', bad, 0, x, x, None, x, None, L.stream())
    with pytest.raises(RuntimeError, match='dtype'):
        L.call('vs_add_act', x, x, x, 7, 8, 0, L.stream())


@pytest.mark.parametrize('T,B,d,h,nb', [(15, 128, 20, 512, 1), (6, 5, 10, 64, 2), (25, 33, 32, 512, 3), (1, 4, 20, 64, 1)])
def test_latent_rollout_kernels(T, B, d, h, nb):
    """Fused rollout forward / adjoint against the specification (fp32; 2e-5 relative per tensor)."""
    torch.manual_seed(11)
    ws = []
    for _ in range(nb):
        ws += [torch.randn(h, d) / d ** 0.5, 0.1 * torch.randn(h), torch.randn(h, h) / h ** 0.5 * 0.7, 0.1 * torch.randn(h),
               torch.randn(d, h) / h ** 0.5 * 0.3, 0.1 * torch.randn(d)]
    n = max(T - 1, 0)
    codes = torch.zeros(T, B, d)
    codes[0] = torch.randn(B, d)
    hidden, xin, res = torch.zeros(nb, 2, n, B, h), torch.zeros(nb, n, B, d), torch.zeros(nb, n, B, d)
    # CPU specification
    c_cpu, h_cpu, x_cpu, r_cpu = codes.clone(), hidden.clone(), xin.clone(), res.clone()
    arr = emu._emu_pointer_array(ws)
    emu.emu_call('vs_latent_rollout_forward', c_cpu, arr, T, B, d, h, nb, h_cpu, x_cpu, r_cpu, None)
    # CUDA
    wg = [w.cuda() for w in ws]
    c_gpu, h_gpu, x_gpu, r_gpu = codes.cuda(), hidden.cuda(), xin.cuda(), res.cuda()
    L.call('vs_latent_rollout_forward', c_gpu, L.pointer_array(wg), T, B, d, h, nb, h_gpu, x_gpu, r_gpu, L.stream())
    torch.cuda.synchronize()
    for a, b, name in ((c_gpu, c_cpu, 'codes'), (h_gpu, h_cpu, 'hidden'), (x_gpu, x_cpu, 'xin'), (r_gpu, r_cpu, 'res')):
        if b.numel():
            close(a.cpu(), b, torch.float32, 'rollout fwd ' + name, outliers=1e-4, scale=float(b.abs().max()))
    if T == 1:
        return
    dcodes = torch.randn(T, B, d)
    d_cpu, dres_cpu, dh_cpu = dcodes.clone(), torch.zeros(nb, n, B, d), torch.zeros(nb, 2, n, B, h)
    wt = [w.t().contiguous() if w.dim() == 2 else w for w in ws]        # the adjoint takes transposed weights
    emu.emu_call('vs_latent_rollout_backward', d_cpu, emu._emu_pointer_array(wt), T, B, d, h, nb, h_cpu, dres_cpu, dh_cpu, None)
    d_gpu, dres_gpu, dh_gpu = dcodes.cuda(), torch.zeros(nb, n, B, d).cuda(), torch.zeros(nb, 2, n, B, h).cuda()
    wtg = [w.cuda() for w in wt]
    L.call('vs_latent_rollout_backward', d_gpu, L.pointer_array(wtg), T, B, d, h, nb, h_cpu.cuda(), dres_gpu, dh_gpu, L.stream())
    torch.cuda.synchronize()
    close(d_gpu[0].cpu(), d_cpu[0], torch.float32, 'rollout bwd dcodes[0]', scale=float(d_cpu[0].abs().max()))
    close(dres_gpu.cpu(), dres_cpu, torch.float32, 'rollout bwd dres', scale=float(dres_cpu.abs().max()))
    close(dh_gpu.cpu(), dh_cpu, torch.float32, 'rollout bwd dhidden', outliers=1e-4, scale=float(dh_cpu.abs().max()))


_PAIR_WGRAD_SNIPPET = r'''
import sys, torch
sys.path.insert(0, %r)
from spatiotemporal_variable_separation_b200 import _lib as L
torch.manual_seed(3)
for (N, H, W, C, K, R, st, pad) in [(4, 8, 8, 256, 512, 4, 2, 1), (6, 8, 8, 128, 256, 3, 1, 1)]:
    P = (H + 2 * pad - R) // st + 1; Q = (W + 2 * pad - R) // st + 1
    small = torch.randn(N, P, Q, K, device='cuda').bfloat16(); big = torch.randn(N, H, W, C, device='cuda').bfloat16()
    d1 = torch.zeros(K, C, R, R, device='cuda'); d2 = torch.zeros_like(d1)
    L.call('vs_conv_wgrad', L.Geom(1, N, H, W, C, P, Q, K, R, R, st, pad, 1, 0, 0), small, big, d1, L.stream())
    L.call('vs_conv_wgrad', L.Geom(1, N, H, W, C, P, Q, K, R, R, st, pad, 1, 0, L.FLAG_FORCE_SIMT), small, big, d2, L.stream())
    torch.cuda.synchronize()
    err = float((d1 - d2).abs().max()) / float(d2.abs().max())
    assert err < 2e-3, err
print('pair wgrad ok')
'''


def test_wgrad_cta_pair_opt_in():
    """The cta_group::2 weight-gradient kernel is opt-in (VARSEP_ENABLE_WGRAD_PAIR=1, read once per process)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VARSEP_ENABLE_WGRAD_PAIR='1')
    r = subprocess.run([sys.executable, '-c', _PAIR_WGRAD_SNIPPET % root], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'pair wgrad ok' in r.stdout, r.stdout + r.stderr


def test_pack_weights_multi_matches_single_entry_kernel():
    """One-launch repack of many (parameter, layout) pairs == the per-parameter kernel, bit for bit: filter shapes with
    4x4 / 3x3 / 5x5 / 1x1 taps, both layouts, both dtypes, channel counts that are not multiples of the 64-wide tile."""
    import struct
    torch.manual_seed(11)
    shapes = [(64, 5, 16), (128, 64, 16), (70, 130, 16), (1, 64, 16), (64, 1, 16), (24, 40, 9), (16, 15, 25), (50, 37, 1),
              (1200, 96, 1), (512, 148, 16)]
    rows, first, keep = [], 0, []
    for (K, C, RS) in shapes:
        w = torch.randn(K, C, RS, device='cuda')
        for swap in (0, 1):
            for dt in (torch.float32, torch.bfloat16):
                out = torch.full((K * C * RS,), float('nan'), device='cuda', dtype=dt)
                want = torch.empty_like(out)
                L.call('vs_pack_weight', w, want, L.dtype_code(want), K, C, RS, swap, L.stream())
                rows.append(struct.pack('<QQiiiiii', w.data_ptr(), out.data_ptr(), K, C, RS, swap,
                                        L.VS_F32 if dt == torch.float32 else L.VS_BF16, first))
                first += (K * C * RS + 1023) // 1024
                keep.append((w, out, want, (K, C, RS, swap, dt)))
    table = torch.frombuffer(bytearray(b''.join(rows)), dtype=torch.uint8).cuda()
    L.call('vs_pack_weights_multi', table, len(rows), first, L.stream())
    torch.cuda.synchronize()
    for w, out, want, what in keep:
        assert torch.equal(out, want), what


_FUSED_CLASSES_SNIPPET = r'''
import sys, torch
sys.path.insert(0, %r)
from spatiotemporal_variable_separation_b200 import _lib as L
torch.manual_seed(5)
# transposed k4 s2 p1 with 64 output channels: full tiles, several images per tile, a partial last tile, two channel tiles
for (N, P, K, C, G) in [(6, 16, 128, 64, 3), (16, 8, 64, 64, 2), (5, 4, 64, 64, 1), (4, 8, 72, 96, 2)]:
    H = 2 * P
    x = torch.randn(N, P, P, K, device='cuda').bfloat16()
    wp = (torch.randn(C, 16, K, device='cuda') / (16 * K) ** 0.5).bfloat16()
    bias = torch.randn(C, device='cuda')
    outs, sts = [], []
    for flags in (0, L.FLAG_FORCE_SIMT):
        out = torch.full((N, H, H, C), float('nan'), device='cuda').bfloat16()
        st = torch.zeros(G * C * 2, device='cuda', dtype=torch.float64)
        before = L.launch_count()
        L.call('vs_conv_forward', L.Geom(1, N, H, H, C, P, P, K, 4, 4, 2, 1, G, 0, flags), L.TRANSPOSED, x, wp, bias, out, st, L.stream())
        torch.cuda.synchronize()
        outs.append(out.float()); sts.append(st)
    err = float((outs[0] - outs[1]).abs().max()) / float(outs[1].abs().max())
    serr = float((sts[0] - sts[1]).abs().max()) / float(sts[1].abs().max())
    assert err < 2 ** -7 and serr < 2e-3, (N, P, K, C, err, serr)
print('fused classes ok')
'''


def test_fused_class_transposed_conv_opt_in():
    """The fused-parity-class kernel (tc_convT4_kernel) is opt-in (VARSEP_ENABLE_FUSED_CLASSES=1, read once per process)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VARSEP_ENABLE_FUSED_CLASSES='1')
    r = subprocess.run([sys.executable, '-c', _FUSED_CLASSES_SNIPPET % root], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'fused classes ok' in r.stdout, r.stdout + r.stderr


# (N, P, Q, C, bn_groups): the thin transposed convolution ConvTranspose2d(64, C, 4, 2, 1) behind a BatchNorm + LeakyReLU
TAIL_CASES = [(6, 32, 32, 1, 3), (4, 16, 16, 2, 2), (5, 32, 32, 1, 1), (8, 16, 32, 1, 4)]


@pytest.mark.parametrize('case', TAIL_CASES)
def test_fused_decoder_tail_kernels(case):
    """vs_tail_forward / vs_tail_wgrad / vs_tail_bn_backward (both phases, train and eval) against their specification
    = the composition of the plain entry points they replace."""
    N, P, Q, C, G = case
    K, dtype = 64, torch.bfloat16
    H, W = 2 * P, 2 * Q
    torch.manual_seed(11)
    g = L.Geom(1, N, H, W, C, P, Q, K, 4, 4, 2, 1, 1, 4, 0)                    # output activation: sigmoid
    assert L.load().vs_tail_eligible(g) == 1
    y = (torch.randn(N, P, Q, K) * 1.5 + 0.3).to(dtype)
    mean, invstd = torch.randn(G * K) * 0.2 + 0.3, torch.rand(G * K) * 0.5 + 0.4
    gamma, beta = torch.randn(K) * 0.3 + 1.0, torch.randn(K) * 0.2
    w = torch.randn(K, C, 4, 4) / np.sqrt(K * 4)
    wp_t, wp_d = torch.zeros(K * C * 16).to(dtype), torch.zeros(K * C * 16).to(dtype)
    emu.emu_call('vs_pack_weight', w, wp_t, 1, K, C, 16, 1, None)
    emu.emu_call('vs_pack_weight', w, wp_d, 1, K, C, 16, 0, None)
    bias = torch.randn(C)
    out = torch.zeros(N, H, W, C).to(dtype)
    gpu, cpu = run_both('vs_tail_forward', [g, y, mean, invstd, gamma, beta, G, 2, wp_t, bias, out, None])
    close(gpu[10], cpu[10], dtype, 'tail forward', outliers=1e-4)
    dout = (torch.randn(N, H, W, C) * 0.1).to(dtype)
    dw = torch.randn(K, C, 4, 4)
    gpu, cpu = run_both('vs_tail_wgrad', [g, y, mean, invstd, gamma, beta, G, 2, dout, dw, None])
    close(gpu[9], cpu[9], torch.float32, 'tail wgrad', scale=float(cpu[9].abs().max()) * 50)   # bf16 operands, fp32 sums
    for train in (1, 0):
        sums = torch.zeros(G * K * 2, dtype=torch.float64)
        gpu, cpu = run_both('vs_tail_bn_backward', [g, y, mean, invstd, gamma, beta, G, 2, dout, wp_d, 0, train, sums, None,
                                                    None, None, None])
        close(gpu[12].float(), cpu[12].float(), torch.float32, 'tail bn sums', scale=float(cpu[12].abs().max()) * 50)
        dy = torch.full((N, P, Q, K), float('nan')).to(dtype)
        dgamma, dbeta = torch.randn(K), torch.randn(K)
        # phase 1 on the SAME sums on both sides (the specification's), so that only this phase is compared
        args = [g, y, mean, invstd, gamma, beta, G, 2, dout, wp_d, 1, train, cpu[12].clone(), dy, dgamma, dbeta, None]
        gpu1, cpu1 = run_both('vs_tail_bn_backward', args)
        close(gpu1[13], cpu1[13], dtype, f'tail dy train={train}', outliers=2e-3, scale=float(cpu1[13].float().abs().max()))
        close(gpu1[14], cpu1[14], torch.float32, 'tail dgamma', scale=float(cpu1[14].abs().max()))
        close(gpu1[15], cpu1[15], torch.float32, 'tail dbeta', scale=float(cpu1[15].abs().max()))

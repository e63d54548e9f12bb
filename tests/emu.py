"""Executable specification of the C ABI (include/varsep.h) in CPU PyTorch.  TEST INFRASTRUCTURE ONLY.

Two uses:
  * ``install()`` swaps ``_lib.call`` for this emulator so the *host-side* logic (module wiring,
    grouped BatchNorm bookkeeping, autograd plumbing, loss assembly, optimizer arena) can be checked
    against the oracle on a machine without a GPU (``-m "not gpu"`` tests);
  * on the GPU box every CUDA entry point is compared with the same functions, argument for argument
    (tests/test_kernels_gpu.py).
It is never imported by the product package; without it a CPU tensor makes every operator raise.
"""
import contextlib

import torch
import torch.nn.functional as F

from spatiotemporal_variable_separation_b200 import _lib as L

ACT_NAMES = {0: None, 1: 'relu', 2: 'leaky', 3: 'elu', 4: 'sigmoid', 5: 'tanh'}


def _act(z, a):
    return {0: lambda v: v, 1: F.relu, 2: lambda v: F.leaky_relu(v, 0.2), 3: F.elu, 4: torch.sigmoid,
            5: torch.tanh}[a](z)


def _act_grad_in(z, a):
    z = z.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        _act(z, a).sum().backward()
    return z.grad


def _act_grad_out(o, a):
    if a == 0:
        return torch.ones_like(o)
    if a == 1:
        return (o > 0).to(o.dtype)
    if a == 2:
        return torch.where(o > 0, torch.ones_like(o), torch.full_like(o, 0.2))
    if a == 3:
        return torch.where(o > 0, torch.ones_like(o), o + 1)
    if a == 4:
        return o * (1 - o)
    return 1 - o * o


def _f(t):
    return t.detach().float()


def _store(dst, val):
    dst.copy_(val.reshape(dst.shape).to(dst.dtype))


def _w_from_packed(wp, OC, RS, IC):
    return _f(wp).reshape(OC, RS, IC)


def vs_pack_weight(w, out, dtype, K, C, RS, swap, stream):
    w3 = _f(w).reshape(K, C, RS)
    _store(out, w3.permute(1, 2, 0) if swap else w3.permute(0, 2, 1))


def vs_pack_weights_multi(table, n, total_blocks, stream):
    import struct
    raw = bytes(table.numpy().tobytes())
    for i in range(n):
        src, dst, K, C, RS, swap, dtype, first = struct.unpack_from('<QQiiiiii', raw, 40 * i)
        vs_pack_weight(_TENSOR_REGISTRY[src], _TENSOR_REGISTRY[dst], dtype, K, C, RS, swap, None)


def _geo(g):
    return {k: getattr(g, k) for k, _ in L.Geom._fields_}


def vs_conv_forward(g, mode, x, wp, bias, out, stats, stream):
    g = _geo(g)
    R, S, st, pad = g['R'], g['S'], g['stride'], g['pad']
    if mode == L.DIRECT:
        xin = _f(x).reshape(g['N'], g['H'], g['W'], g['C']).permute(0, 3, 1, 2)
        w = _w_from_packed(wp, g['K'], R * S, g['C']).reshape(g['K'], R, S, g['C']).permute(0, 3, 1, 2)
        y = F.conv2d(xin, w, None, st, pad)
        OC = g['K']
    else:
        xin = _f(x).reshape(g['N'], g['P'], g['Q'], g['K']).permute(0, 3, 1, 2)
        # packed [C][RS][K] -> conv_transpose2d weight [K(in)][C(out)][R][S]
        w = _w_from_packed(wp, g['C'], R * S, g['K']).reshape(g['C'], R, S, g['K']).permute(3, 0, 1, 2)
        # output size is fixed by the geometry (H, W): the remainder rows of a strided conv come back
        # through output_padding (< stride)
        oph = g['H'] - ((g['P'] - 1) * st - 2 * pad + R)
        opw = g['W'] - ((g['Q'] - 1) * st - 2 * pad + S)
        y = F.conv_transpose2d(xin, w, None, st, pad, output_padding=(oph, opw))
        OC = g['C']
    y = y.permute(0, 2, 3, 1)
    if bias is not None:
        y = y + _f(bias)
    if stats is not None:
        G = g['groups']
        yg = y.reshape(G, -1, OC).double()
        stats.copy_((stats.reshape(G, OC, 2) + torch.stack([yg.sum(1), (yg * yg).sum(1)], -1)).reshape(stats.shape))
    _store(out, _act(y, g['act']))


def vs_conv_wgrad(g, small, big, dw, stream):
    g = _geo(g)
    sm = _f(small).reshape(g['N'], g['P'], g['Q'], g['K']).permute(0, 3, 1, 2)
    bg = _f(big).reshape(g['N'], g['H'], g['W'], g['C']).permute(0, 3, 1, 2).clone().requires_grad_(False)
    w = torch.zeros(g['K'], g['C'], g['R'], g['S'], requires_grad=True)
    with torch.enable_grad():
        y = F.conv2d(bg, w, None, g['stride'], g['pad'])
        (y * sm).sum().backward()
    dw += w.grad.reshape(dw.shape)


def vs_colsum(a, dtype, rows, C, db, stream):
    db += _f(a).reshape(rows, C).sum(0)


def vs_bn_finalize(stats, G, C, count, eps, momentum, mean, invstd, rmean, rvar, nbt, stream):
    s = stats.reshape(G, C, 2)
    mu = s[..., 0] / count
    var = (s[..., 1] / count - mu * mu).clamp_min(0)
    mean.copy_(mu.float().reshape(mean.shape))
    invstd.copy_((1.0 / torch.sqrt(var + eps)).float().reshape(invstd.shape))
    for gi in range(G):
        unb = var[gi] * count / (count - 1) if count > 1 else var[gi]
        if rmean is not None:
            rmean.copy_((1 - momentum) * rmean + momentum * mu[gi].float())
        if rvar is not None:
            rvar.copy_((1 - momentum) * rvar + momentum * unb.float())
    if nbt is not None:
        nbt += G


def vs_bn_finalize_act_forward(stats, G, C, count, eps, momentum, mean, invstd, rmean, rvar, nbt, y, out, dtype, rows, gamma,
                               beta, act, stream):
    vs_bn_finalize(stats, G, C, count, eps, momentum, mean, invstd, rmean, rvar, nbt, None)
    vs_bn_act_forward(y, out, dtype, rows, C, G, mean, invstd, gamma, beta, act, None)


def vs_bn_eval_stats(rmean, rvar, C, eps, mean, invstd, stream):
    mean.copy_(rmean)
    invstd.copy_(1.0 / torch.sqrt(rvar + eps))


def _bn_z(y, rows, C, G, mean, invstd, gamma, beta):
    yv = _f(y).reshape(G, rows // G, C)
    xh = (yv - mean.reshape(G, 1, C)) * invstd.reshape(G, 1, C)
    return xh, gamma * xh + beta


def vs_bn_act_forward(y, out, dtype, rows, C, G, mean, invstd, gamma, beta, act, stream):
    _, z = _bn_z(y, rows, C, G, mean, invstd, gamma, beta)
    _store(out, _act(z, act))


def _bn_z64(y, rows, C, G, mean, invstd, gamma, beta):
    """fp64 per-element arithmetic on the fp32-rounded statistics, as ATen's CPU batch-norm backward
    (accumulate type double): the subtraction dz - mean(dz) - xhat*mean(dz*xhat) cancels heavily for
    the layers next to the latent codes, so fp32 element arithmetic would cost ~3 digits there."""
    yv = y.detach().double().reshape(G, rows // G, C)
    xh = (yv - mean.double().reshape(G, 1, C)) * invstd.double().reshape(G, 1, C)
    return xh, gamma.double() * xh + beta.double()


def vs_bn_act_backward_reduce(dout, y, dtype, rows, C, G, mean, invstd, gamma, beta, act, sums, stream):
    xh, z = _bn_z64(y, rows, C, G, mean, invstd, gamma, beta)
    dz = dout.detach().double().reshape(G, rows // G, C) * _act_grad_in(z, act)
    sums += torch.stack([dz.sum(1), (dz * xh).sum(1)], -1).reshape(sums.shape)


def vs_bn_act_backward_apply(dout, y, dy, dtype, rows, C, G, mean, invstd, gamma, beta, act, sums, train, dgamma,
                             dbeta, stream):
    xh, z = _bn_z64(y, rows, C, G, mean, invstd, gamma, beta)
    dz = dout.detach().double().reshape(G, rows // G, C) * _act_grad_in(z, act)
    s = sums.reshape(G, 1, C, 2)
    cnt = rows // G
    scale = gamma.double() * invstd.double().reshape(G, 1, C)
    if train:
        r = scale * (dz - s[..., 0] / cnt - xh * s[..., 1] / cnt)
    else:
        r = scale * dz
    _store(dy, r)
    if dgamma is not None:
        dgamma += sums.reshape(G, C, 2)[..., 1].sum(0).float().reshape(dgamma.shape)
    if dbeta is not None:
        dbeta += sums.reshape(G, C, 2)[..., 0].sum(0).float().reshape(dbeta.shape)


# ---- fused decoder tail: the composition of the plain entry points it replaces (include/varsep.h) ------------------
def _tail_act(g, y, mean, invstd, gamma, beta, G, bn_act):
    gg = _geo(g)
    rows, K = gg['N'] * gg['P'] * gg['Q'], gg['K']
    act = torch.empty(rows, K, dtype=y.dtype)                 # rounded to the storage type, as the unfused path stores it
    vs_bn_act_forward(y, act, None, rows, K, G, mean, invstd, gamma, beta, bn_act, None)
    return act, rows, K


def _geom_with(g, **kw):
    d = _geo(g)
    d.update(kw)
    return L.Geom(*[d[k] for k, _ in L.Geom._fields_])


def vs_tail_forward(g, y, mean, invstd, gamma, beta, G, bn_act, wp, bias, out, stream):
    act, _, _ = _tail_act(g, y, mean, invstd, gamma, beta, G, bn_act)
    vs_conv_forward(g, L.TRANSPOSED, act, wp, bias, out, None, None)


def vs_tail_wgrad(g, y, mean, invstd, gamma, beta, G, bn_act, dout, dw, stream):
    act, _, _ = _tail_act(g, y, mean, invstd, gamma, beta, G, bn_act)
    vs_conv_wgrad(g, act, dout, dw, None)


def vs_tail_bn_backward(g, y, mean, invstd, gamma, beta, G, bn_act, dout, wp_direct, phase, train, sums, dy, dgamma,
                        dbeta, stream):
    gg = _geo(g)
    rows, K = gg['N'] * gg['P'] * gg['Q'], gg['K']
    dact = torch.empty(rows, K, dtype=torch.float32)          # the fp32 accumulator tile, never rounded to bf16
    vs_conv_forward(_geom_with(g, act=0, groups=1), L.DIRECT, dout, wp_direct, None, dact, None, None)
    if phase == 0:
        vs_bn_act_backward_reduce(dact, y, None, rows, K, G, mean, invstd, gamma, beta, bn_act, sums, None)
    else:
        vs_bn_act_backward_apply(dact, y, dy, None, rows, K, G, mean, invstd, gamma, beta, bn_act, sums, train, dgamma,
                                 dbeta, None)


def vs_act_backward(dout, out, dx, dtype, n, act, stream):
    _store(dx, _f(dout) * _act_grad_out(_f(out), act))


def vs_add_act(a, b, out, dtype, n, act, stream):
    _store(out, _act(_f(a) + _f(b), act))


def vs_copy_channels(src, src_C, src_rows, dst, dst_C, dst_off, rows, dtype, stream):
    s = src.detach().reshape(src_rows, src_C)
    d = dst.view(rows, dst_C)
    d[:, dst_off:dst_off + src_C] = s.repeat(rows // src_rows, 1)


def vs_slice_channels_reduce(ddst, dst_C, dst_off, rows, dsrc, src_C, src_rows, dtype, stream):
    d = _f(ddst).reshape(rows // src_rows, src_rows, dst_C)[:, :, dst_off:dst_off + src_C].sum(0)
    _store(dsrc, d)


def vs_mul_bcast(s, s_rows, t, out, rows, C, dtype, stream):
    _store(out, _f(s).reshape(1, s_rows, C) * _f(t).reshape(rows // s_rows, s_rows, C))


def vs_mul_bcast_backward(dout, s, s_rows, t, ds, dt, rows, C, dtype, stream):
    d = _f(dout).reshape(rows // s_rows, s_rows, C)
    if ds is not None:
        _store(ds, (d * _f(t).reshape(rows // s_rows, s_rows, C)).sum(0))
    if dt is not None:
        _store(dt, d * _f(s).reshape(1, s_rows, C))


def vs_maxpool_forward(x, y, dtype, N, H, W, C, k, stride, pad, stream):
    _store(y, F.max_pool2d(_f(x).reshape(N, H, W, C).permute(0, 3, 1, 2), k, stride, pad).permute(0, 2, 3, 1))


def vs_maxpool_backward(x, dy, dx, dtype, N, H, W, C, k, stride, pad, stream):
    xv = _f(x).reshape(N, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    with torch.enable_grad():
        y = F.max_pool2d(xv, k, stride, pad)
        y.backward(_f(dy).reshape(N, y.shape[2], y.shape[3], C).permute(0, 3, 1, 2))
    _store(dx, xv.grad.permute(0, 2, 3, 1))


def vs_upsample2_forward(x, y, dtype, N, H, W, C, stream):
    v = x.detach().reshape(N, H, W, C)
    _store(y, v.repeat_interleave(2, 1).repeat_interleave(2, 2))


def vs_upsample2_backward(dy, dx, dtype, N, H, W, C, stream):
    _store(dx, _f(dy).reshape(N, H, 2, W, 2, C).sum((2, 4)))


def vs_frames_to_nhwc(frames, B, T, Cf, H, W, t0, nt, out, dtype, stream):
    v = frames.detach().reshape(B, T, Cf, H, W)[:, t0:t0 + nt].reshape(B, nt * Cf, H, W).permute(0, 2, 3, 1)
    _store(out, v)


def vs_nhwc_to_nchw(x, dtype, out, N, C, H, W, stream):
    _store(out, _f(x).reshape(N, H, W, C).permute(0, 3, 1, 2))


def vs_nchw_to_nhwc(x, out, dtype, N, C, H, W, stream):
    _store(out, _f(x).reshape(N, C, H, W).permute(0, 2, 3, 1))


def _strided(t, sb, st, B, T, Ln):
    return torch.as_strided(t.detach(), (B, T, Ln), (sb, st, 1), t.storage_offset())


def vs_sqdiff_sum(a, a_sb, a_st, b, b_sb, b_st, B, T, Ln, acc, stream):
    d = _strided(a, a_sb, a_st, B, T, Ln)
    if b is not None:
        d = d - _strided(b, b_sb, b_st, B, T, Ln)
    acc += (d * d).double().sum()


def vs_sqdiff_backward(a, a_sb, a_st, b, b_sb, b_st, B, T, Ln, scale, g_term, g_total, lamb, da, accumulate, stream):
    d = _strided(a, a_sb, a_st, B, T, Ln)
    if b is not None:
        d = d - _strided(b, b_sb, b_st, B, T, Ln)
    sc = scale * ((float(g_term[0]) if g_term is not None else 0.0) +
                  (lamb * float(g_total[0]) if g_total is not None else 0.0))
    dst = torch.as_strided(da, (B, T, Ln), (a_sb, a_st, 1), da.storage_offset())
    if accumulate:
        dst += sc * d
    else:
        dst.copy_(sc * d)


def vs_loss_combine(acc, coef, lamb, n, terms, stream):
    t = torch.tensor([float(acc[i]) * coef[i] for i in range(n)], dtype=torch.float32)
    terms[:n] = t
    terms[n] = sum(float(lamb[i]) * t[i] for i in range(n))


def vs_adam_step(p, g, m, v, n, lr, b1, b2, eps, grad_scale, step_host, step_dev, lr_dev, stream):
    step = int(step_dev[0]) if step_dev is not None else step_host
    if lr_dev is not None:
        lr = float(lr_dev[0])
    gs = g * grad_scale
    m.mul_(b1).add_(gs, alpha=1 - b1)
    v.mul_(b2).addcmul_(gs, gs, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr / bc1)


def vs_moving_sequences(glyphs, n_glyphs, gh, gw, objs, n_obj, B, T, F, frames, stream):
    from oracle import moving_mnist
    out = moving_mnist.render(glyphs.numpy().reshape(n_glyphs, gh, gw), objs.numpy().reshape(B, n_obj, 5), T, F)
    frames.copy_(torch.from_numpy(out).reshape(frames.shape))


_TABLE = {k: v for k, v in globals().items() if k.startswith('vs_')}


def emu_call(name, *args):
    with torch.no_grad():
        _TABLE[name](*args)


@contextlib.contextmanager
def install():
    """Route the package's C-ABI calls to the emulator (CPU tensors allowed) for the duration."""
    from spatiotemporal_variable_separation_b200 import ops
    saved = (L.call, L.stream, L.require_cuda, L.launch_count, L.pointer_array, ops.repack_params)
    L.call, L.stream, L.require_cuda, L.launch_count = emu_call, (lambda: None), (lambda *a: None), (lambda: -1)
    L.pointer_array = _emu_pointer_array
    ops.repack_params = lambda params, cache: None      # the table holds raw device pointers; the emulator repacks lazily per call instead
    try:
        yield
    finally:
        L.call, L.stream, L.require_cuda, L.launch_count, L.pointer_array, ops.repack_params = saved


# ---- latent rollout (appended after _TABLE is built: register explicitly) -------------------------
def _ptrs_to_tensors(arr, registry):
    return [registry[int(a)] for a in arr]


_TENSOR_REGISTRY = {}


def _register(ts):
    for t in ts:
        _TENSOR_REGISTRY[t.data_ptr()] = t


def _blocks(w_host, nb):
    ws = _ptrs_to_tensors(w_host, _TENSOR_REGISTRY)
    return [ws[6 * j:6 * j + 6] for j in range(nb)]


def vs_latent_rollout_forward(codes, w_host, T, B, d, h, nb, hidden, xin, res, stream):
    x = codes[0].clone()
    for t in range(1, T):
        for j, (w1, b1, w2, b2, w3, b3) in enumerate(_blocks(w_host, nb)):
            if xin is not None:
                xin[j, t - 1] = x
            u = F.relu(F.linear(x, w1, b1))
            v = F.relu(F.linear(u, w2, b2))
            r = F.linear(v, w3, b3)
            if hidden is not None:
                hidden[j, 0, t - 1], hidden[j, 1, t - 1] = u, v
            if res is not None:
                res[j, t - 1] = r
            x = x + r
        codes[t] = x


def vs_latent_rollout_backward(dcodes, w_host, T, B, d, h, nb, hidden, dres, dhidden, stream):
    blocks = _blocks(w_host, nb)          # TRANSPOSED weights: W1^T [d][h], W2^T [h][h], W3^T [h][d] (flat buffers)
    g = dcodes[T - 1].clone()
    for t in range(T - 1, 0, -1):
        for j in range(nb - 1, -1, -1):
            w1t, _, w2t, _, w3t, _ = blocks[j]
            w1t, w2t, w3t = w1t.reshape(d, h), w2t.reshape(h, h), w3t.reshape(h, d)
            dres[j, t - 1] = g
            da2 = (g @ w3t.t()) * (hidden[j, 1, t - 1] > 0)
            dhidden[j, 1, t - 1] = da2
            da1 = (da2 @ w2t.t()) * (hidden[j, 0, t - 1] > 0)
            dhidden[j, 0, t - 1] = da1
            g = g + da1 @ w1t.t()
        g = g + dcodes[t - 1]
    dcodes[0] = g


_TABLE.update({'vs_latent_rollout_forward': vs_latent_rollout_forward,
               'vs_latent_rollout_backward': vs_latent_rollout_backward})
_orig_pointer_array = L.pointer_array


def _emu_pointer_array(tensors):
    _register(tensors)
    return [t.data_ptr() for t in tensors]


# ---- evaluation metrics (appended after _TABLE is built: registered explicitly) -----------------------
def vs_ssim_mse_planes(pred, target, planes, H, W, kernel, fs, c1, c2, ssim_map, ssim_mean, mse_mean, stream):
    """utils/ssim.py:95-116 on `planes` single-channel images + the per-plane mean squared error."""
    import torch.nn.functional as F
    x = pred.reshape(-1)[:planes * H * W].reshape(planes, 1, H, W).float()
    y = target.reshape(-1)[:planes * H * W].reshape(planes, 1, H, W).float()
    k = kernel.reshape(-1)[:fs * fs].reshape(1, 1, fs, fs).float()
    mu1, mu2 = F.conv2d(x, k), F.conv2d(y, k)
    s11 = F.conv2d(x * x, k) - mu1 ** 2
    s22 = F.conv2d(y * y, k) - mu2 ** 2
    s12 = F.conv2d(x * y, k) - mu1 * mu2
    v1, v2 = 2 * s12 + c2, s11 + s22 + c2
    ssim = ((2 * mu1 * mu2 + c1) * v1) / ((mu1 ** 2 + mu2 ** 2 + c1) * v2)
    if ssim_map is not None:
        ssim_map.reshape(-1)[:ssim.numel()].copy_(ssim.reshape(-1))
    ssim_mean.reshape(-1)[:planes].copy_(ssim.mean(dim=(1, 2, 3)))
    mse_mean.reshape(-1)[:planes].copy_(((x - y) ** 2).mean(dim=(1, 2, 3)))


_TABLE['vs_ssim_mse_planes'] = vs_ssim_mse_planes

"""Host-side operators: thin autograd wrappers around the C ABI.

Every arithmetic operation of the hot path is a kernel of libvarsep_sm100a.so;
torch.autograd only wires the backward calls together and PyTorch tensors are
used as device-memory handles.  Internal activations are NHWC tensors
``[N, H, W, C]`` of the compute dtype (fp32 = parity mode, bf16 = tensor-core mode).
"""
from collections import namedtuple

import torch

from . import _lib as L
from ._lib import ptr

_compute_dtype = torch.float32


def set_compute_dtype(dtype):
    """torch.float32 (bit-for-bit fp32 accumulation, parity mode) or torch.bfloat16 (tcgen05 path)."""
    global _compute_dtype
    assert dtype in (torch.float32, torch.bfloat16)
    _compute_dtype = dtype


def compute_dtype():
    return _compute_dtype


# ------------------------------------------------------------------------------------------------
# packed-weight cache.  Kernels read weights as [OC][taps][IC] in the compute dtype; the fp32
# torch parameter stays the master copy.  Every cache entry lives on the parameter object itself
# (``w.__dict__``), so it dies with the module; nothing process-global refers to a parameter.  An
# entry is current while its tag (tensor version counter, address, per-parameter epoch) matches;
# ``invalidate_params`` bumps the epoch of the parameters an optimizer has just written through raw
# pointers (the fused Adam step does not bump version counters).
# ------------------------------------------------------------------------------------------------
_PACK_ROWS_PER_LAUNCH = 1024            # rows of the table one vs_pack_weights_multi launch keeps in shared memory


def _epoch(t):
    return t.__dict__.get('_vs_epoch', 0) if t is not None else 0


def _tag(w):
    return (w._version, w.data_ptr(), _epoch(w))


def invalidate_params(params):
    """Mark every packed / padded / folded copy derived from ``params`` stale."""
    for p in params:
        p.__dict__['_vs_epoch'] = p.__dict__.get('_vs_epoch', 0) + 1


def packed_weight(w, K, C, RS, swap, dtype):
    """Packed copy of parameter ``w``; cached on the parameter object itself (so the cache dies with
    the module and can never alias another tensor that later reuses the same address)."""
    cache = w.__dict__.setdefault('_vs_pack', {})
    key = (K, C, RS, bool(swap), dtype)
    tag = _tag(w)
    hit = cache.get(key)
    if hit is not None and hit[0] == tag:
        return hit[1]
    out = hit[1] if hit is not None else torch.empty(K * C * RS, device=w.device, dtype=dtype)
    L.call('vs_pack_weight', ptr(w), ptr(out), L.dtype_code(out), K, C, RS, int(swap), L.stream())
    cache[key] = (tag, out)
    return out


def padded_rows(w, Kp):
    """fp32 copy of parameter ``w`` with dim 0 zero-padded to ``Kp`` rows (a persistent shadow, refreshed when the
    parameter changes).  TMA needs a 16-byte channel pitch, so a transposed convolution whose input channel count
    is not a multiple of 8 (the decoder's first up-convolution reads the 148-channel [S|T] code) runs on the
    tensor cores with zero input channels appended; the packed copies are made from this shadow."""
    hit = w.__dict__.get('_vs_padrows')
    tag = _tag(w)
    if hit is not None and hit[0].shape[0] == Kp:
        if hit[1] != tag:
            hit[0][:w.shape[0]].copy_(w.detach())
            invalidate_params([hit[0]])
            w.__dict__['_vs_padrows'] = (hit[0], tag)
        return hit[0]
    shadow = torch.zeros((Kp,) + tuple(w.shape[1:]), device=w.device, dtype=torch.float32)
    shadow[:w.shape[0]].copy_(w.detach())
    w.__dict__['_vs_padrows'] = (shadow, tag)
    return shadow


def repack_params(params, cache):
    """Refresh every packed copy that exists for ``params`` (and for their zero-padded shadows) with one kernel launch
    per 1024 table rows, and mark them current, so that the next step's forward / backward find cache hits instead of
    ~30 small pack launches.  Called by ``FusedAdam.step`` with the optimizer's OWN parameters; ``cache`` is a dict the
    caller keeps (the device-resident table lives as long as the optimizer, not as long as the process)."""
    import struct
    live = []
    for p in params:
        if not p.is_cuda:
            continue
        pad = p.__dict__.get('_vs_padrows')
        if pad is not None:
            shadow = pad[0]
            shadow[:p.shape[0]].copy_(p.detach())
            p.__dict__['_vs_padrows'] = (shadow, _tag(p))
            for key, (_, out) in shadow.__dict__.get('_vs_pack', {}).items():
                live.append((shadow, key, out))
        for key, (_, out) in p.__dict__.get('_vs_pack', {}).items():
            live.append((p, key, out))
    if not live:
        return
    ptrs = [(w.data_ptr(), out.data_ptr()) for w, _, out in live]
    if cache.get('ptrs') != ptrs:
        launches = []
        for lo in range(0, len(live), _PACK_ROWS_PER_LAUNCH):
            rows, first = [], 0
            for w, (K, C, RS, swap, dtype), out in live[lo:lo + _PACK_ROWS_PER_LAUNCH]:
                rows.append(struct.pack('<QQiiiiii', w.data_ptr(), out.data_ptr(), K, C, RS, int(swap),
                                        L.VS_F32 if dtype == torch.float32 else L.VS_BF16, first))
                first += (K * C * RS + L.VS_PACK_BLOCK_ELEMS - 1) // L.VS_PACK_BLOCK_ELEMS
            blob = torch.frombuffer(bytearray(b''.join(rows)), dtype=torch.uint8).to(live[0][0].device)
            launches.append((blob, len(rows), first))
        cache.clear()
        cache.update(ptrs=ptrs, launches=launches)
    for blob, n, blocks in cache['launches']:
        L.call('vs_pack_weights_multi', ptr(blob), n, blocks, L.stream())
    for w, key, out in live:
        w.__dict__['_vs_pack'][key] = (_tag(w), out)


# ------------------------------------------------------------------------------------------------
# zero-initialised fp64 scratch (BatchNorm statistics / backward sums / loss accumulators).  One training step needs
# ~25 small zeroed buffers; ``begin_step`` clears one arena with a single memset and the operators take consecutive
# slices of it.  A slice is never handed out twice between two ``begin_step`` calls, so code that runs outside the
# training step (or after the arena is exhausted) simply falls back to ``torch.zeros``.
# ------------------------------------------------------------------------------------------------
_zero_arena = {'buf': None, 'off': 0}
_ZERO_ARENA_DOUBLES = 1 << 19          # 4 MB


def begin_step(device):
    """Called once at the start of a training step (train.step_losses)."""
    a = _zero_arena
    if a['buf'] is None or a['buf'].device != torch.device(device):
        a['buf'] = torch.empty(_ZERO_ARENA_DOUBLES, device=device, dtype=torch.float64)
    a['buf'].zero_()
    a['off'] = 0


def zeros_f64(n, device):
    a = _zero_arena
    buf = a['buf']
    if buf is not None and buf.device == torch.device(device) and a['off'] + n <= buf.numel():
        out = buf[a['off']:a['off'] + n]
        a['off'] += (n + 1) & ~1          # keep slices 16-byte aligned
        return out
    return torch.zeros(n, device=device, dtype=torch.float64)


# ------------------------------------------------------------------------------------------------
# "gradient ready" notifications for data-parallel bucketing (parallel.GradReducer): forward operators announce the
# parameters they use, backward operators announce them again once their gradient kernels are enqueued.
# ------------------------------------------------------------------------------------------------
_grad_hooks = [None, None]


def set_grad_ready_hook(on_use, on_ready):
    _grad_hooks[0], _grad_hooks[1] = on_use, on_ready


# ------------------------------------------------------------------------------------------------
# weight gradients off the critical path.  The backward chain of a layer is  BatchNorm backward -> input gradient
# (needed by the layer below) -> weight gradient (needed only by the optimizer).  When the gradient goes straight into
# the optimizer's arena, the weight-gradient kernels are enqueued on a per-device side stream: they then run next to the
# BatchNorm backward (HBM bound) and input-gradient kernels of the layers below instead of in front of them.  The stream
# that called ``backward()`` waits for the side stream when the pass ends (engine callback); consumers that run INSIDE the
# pass (the early gradient-exchange buckets) call ``join_wgrad`` themselves.
# ------------------------------------------------------------------------------------------------
_wgrad_streams = {}
_wgrad_dirty = set()
_wgrad_overlap = [True]


def set_wgrad_overlap(flag):
    _wgrad_overlap[0] = bool(flag)


class _OnWgradStream:
    """``with _OnWgradStream(buffer, *operands):`` — launches inside go to the side stream when ``buffer`` is an arena view
    (``_grad_buffer`` returned no autograd tensor) on a CUDA device; the operands are kept alive for that stream."""

    def __init__(self, gbuf, *operands):
        dev = operands[0].device
        self.on = _wgrad_overlap[0] and gbuf[1] is None and dev.type == 'cuda'
        self.dev, self.operands = dev, operands

    def __enter__(self):
        if not self.on:
            return self
        ws = _wgrad_streams.get(self.dev)
        if ws is None:
            ws = _wgrad_streams[self.dev] = torch.cuda.Stream(device=self.dev)
        ws.wait_stream(torch.cuda.current_stream(self.dev))
        self.ctx = torch.cuda.stream(ws)
        self.ctx.__enter__()
        for t in self.operands:
            t.record_stream(ws)
        if self.dev not in _wgrad_dirty:
            _wgrad_dirty.add(self.dev)
            # when this backward pass ends, the stream it was called on waits for the side stream: `loss.backward()` keeps
            # its meaning (gradients complete in stream order) for every consumer, not only for FusedAdam / GradReducer
            try:
                torch.autograd.Variable._execution_engine.queue_callback(join_wgrad)
            except RuntimeError:
                pass                                     # not inside a backward pass (direct call of a backward helper)
        return self

    def __exit__(self, *exc):
        if self.on:
            self.ctx.__exit__(*exc)
        return False


def join_wgrad(device=None, stream=None):
    """Make ``stream`` (default: the current one) wait for the weight gradients enqueued on the side stream."""
    for dev in list(_wgrad_dirty):
        if device is not None and torch.device(device) != dev:
            continue
        (stream or torch.cuda.current_stream(dev)).wait_stream(_wgrad_streams[dev])
        if stream is None:
            _wgrad_dirty.discard(dev)


ConvCfg = namedtuple('ConvCfg', 'kind K C R S stride pad act groups training has_bn eps momentum flags')


def _geom(cfg, dtype, N, H, W, P, Q, act, groups):
    return L.Geom(L.VS_F32 if dtype == torch.float32 else L.VS_BF16, N, H, W, cfg.C, P, Q, cfg.K, cfg.R, cfg.S,
                  cfg.stride, cfg.pad, groups, act, cfg.flags)


def _conv_forward(ctx, x, weight, bias, cfg, act_code, groups, stats_n):
    """The convolution launch shared by ``ConvBlockFn`` and ``DecoderTailFn``.  Records on ``ctx`` what the backward
    needs (``cfg_fwd`` the layer's own geometry, ``cfg`` the launch geometry — K zero-padded to a 16-byte channel pitch
    when needed —, ``dims``, ``mode``).  ``stats_n`` > 0: a zeroed fp64 [stats_n] buffer receives the per-(group,
    channel) sums of the output (BatchNorm batch statistics).  Returns (x as launched, y, stats or None)."""
    x = x.contiguous()
    N = x.shape[0]
    if cfg.kind == 'conv':
        H, W = x.shape[1], x.shape[2]
        assert x.shape[3] == cfg.C, (x.shape, cfg)
        P = (H + 2 * cfg.pad - cfg.R) // cfg.stride + 1
        Q = (W + 2 * cfg.pad - cfg.S) // cfg.stride + 1
        out_shape, OC, mode = (N, P, Q, cfg.K), cfg.K, L.DIRECT
    else:
        P, Q = x.shape[1], x.shape[2]
        assert x.shape[3] == cfg.K, (x.shape, cfg)
        H = (P - 1) * cfg.stride - 2 * cfg.pad + cfg.R
        W = (Q - 1) * cfg.stride - 2 * cfg.pad + cfg.S
        out_shape, OC, mode = (N, H, W, cfg.C), cfg.C, L.TRANSPOSED
    dt = x.dtype
    ctx.cfg_fwd = cfg
    wsrc = weight
    if mode == L.TRANSPOSED and dt == torch.bfloat16 and cfg.K % 8 != 0 and cfg.K >= 32 and cfg.C % 8 == 0:
        # input channels padded with zeros to a 16-byte pitch so that the layer runs on the tensor cores
        Kp = (cfg.K + 7) // 8 * 8
        xp = torch.zeros((N, P, Q, Kp), device=x.device, dtype=dt)
        L.call('vs_copy_channels', ptr(x), cfg.K, N * P * Q, ptr(xp), Kp, 0, N * P * Q, L.dtype_code(xp), L.stream())
        x, wsrc, cfg = xp, padded_rows(weight, Kp), cfg._replace(K=Kp)
    wp = packed_weight(wsrc, cfg.K, cfg.C, cfg.R * cfg.S, mode == L.TRANSPOSED, dt)
    y = torch.empty(out_shape, device=x.device, dtype=dt)
    ctx.cfg, ctx.dims, ctx.mode = cfg, (N, H, W, P, Q, OC), mode      # cfg: geometry of the launches (K padded)
    stats = zeros_f64(stats_n, x.device) if stats_n else None
    g = _geom(cfg, dt, N, H, W, P, Q, act_code, groups)
    L.call('vs_conv_forward', g, mode, ptr(x), ptr(wp), ptr(bias), ptr(y), ptr(stats), L.stream())
    return x, y, stats


def _conv_backward(ctx, x, weight, dy, p_weight, p_bias, need_dx, need_dw, need_db, bias_is_dead):
    """Input, weight and bias gradients of the convolution recorded by ``_conv_forward`` from the gradient ``dy`` of its
    output.  Returns (dx, dw for autograd, db for autograd)."""
    cfg, (N, H, W, P, Q, OC), mode = ctx.cfg, ctx.dims, ctx.mode
    dt = dy.dtype
    rows = dy.numel() // OC
    g = _geom(cfg, dt, N, H, W, P, Q, 0, 1)
    cfg0 = ctx.cfg_fwd                      # the layer's own geometry (differs from cfg when K was padded)
    padded = cfg0.K != cfg.K
    dx = None
    if need_dx:
        g0 = _geom(cfg0, dt, N, H, W, P, Q, 0, 1)
        dx = torch.empty(x.shape[:-1] + (cfg0.K,), device=x.device, dtype=dt) if padded else torch.empty_like(x)
        back_mode = L.TRANSPOSED if mode == L.DIRECT else L.DIRECT
        wp = packed_weight(weight, cfg0.K, cfg0.C, cfg0.R * cfg0.S, back_mode == L.TRANSPOSED, dt)
        L.call('vs_conv_forward', g0, back_mode, ptr(dy), ptr(wp), None, ptr(dx), None, L.stream())
    dw = db = None
    if need_dw:
        dw = _grad_buffer(p_weight)
        small, big = (dy, x) if cfg.kind == 'conv' else (x, dy)
        if padded:
            # gradient of the zero-padded weight; its first K rows are the layer's gradient
            dwp = torch.zeros((cfg.K,) + tuple(weight.shape[1:]), device=x.device, dtype=torch.float32)
            L.call('vs_conv_wgrad', g, ptr(small), ptr(big), ptr(dwp), L.stream())
            dw[0].add_(dwp[:cfg0.K].view_as(dw[0]))
        else:
            with _OnWgradStream(dw, small, big):
                L.call('vs_conv_wgrad', g, ptr(small), ptr(big), ptr(dw[0]), L.stream())
    if p_bias is not None and need_db:
        db = _grad_buffer(p_bias)
        if not bias_is_dead:
            # (eval-mode BatchNorm is an affine map: the bias gradient is the plain column sum of dy)
            with _OnWgradStream(db, dy):
                L.call('vs_colsum', ptr(dy), L.dtype_code(dy), rows, OC, ptr(db[0]), L.stream())
        # else: BatchNorm's backward returns a dy whose per-(group, channel) sum is exactly zero, so the
        # bias gradient is mathematically 0 (the reference computes rounding noise there);
        # the (zero-initialised) buffer is left untouched instead of streaming dy once more.
    return dx, dw[1] if dw else None, db[1] if db else None


class ConvBlockFn(torch.autograd.Function):
    """conv/convT/linear -> [BatchNorm (grouped batch stats or running stats)] -> activation.

    Replaces one ``make_conv_block`` Sequential (conv.py:41-60) or one Linear (+ReLU of the next
    MLP block, mlp.py:24-41).  Geometry terms: see include/varsep.h (big[N,H,W,C] <-> small[N,P,Q,K]).
    """

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, rmean, rvar, nbt, cfg):
        L.require_cuda(x, weight)
        act = L.ACT[cfg.act]
        if cfg.has_bn:
            G = cfg.groups if cfg.training else 1
            OC = cfg.K if cfg.kind == 'conv' else cfg.C
            x, y, stats = _conv_forward(ctx, x, weight, bias, cfg, 0, G, G * OC * 2 if cfg.training else 0)
            rows = y.numel() // OC
            mean = torch.empty(G * OC, device=x.device, dtype=torch.float32)
            invstd = torch.empty_like(mean)
            out = torch.empty_like(y)
            if cfg.training:
                # statistics -> mean / invstd / running-stat EMA and normalise + activation in one launch
                L.call('vs_bn_finalize_act_forward', ptr(stats), G, OC, rows // G, cfg.eps, cfg.momentum, ptr(mean),
                       ptr(invstd), ptr(rmean), ptr(rvar), ptr(nbt), ptr(y), ptr(out), L.dtype_code(y), rows, ptr(gamma),
                       ptr(beta), act, L.stream())
            else:
                L.call('vs_bn_eval_stats', ptr(rmean), ptr(rvar), OC, cfg.eps, ptr(mean), ptr(invstd), L.stream())
                L.call('vs_bn_act_forward', ptr(y), ptr(out), L.dtype_code(y), rows, OC, G, ptr(mean), ptr(invstd),
                     ptr(gamma), ptr(beta), act, L.stream())
            ctx.save_for_backward(x, weight, y, mean, invstd, gamma, beta)
            ctx.G = G
        else:
            x, out, _ = _conv_forward(ctx, x, weight, bias, cfg, act, 1, 0)
            ctx.save_for_backward(x, weight, out)
        # python references to the leaf parameters (their ``_vs_grad`` arena views are looked up in backward)
        ctx.params = (weight, bias, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, dout):
        cfg, (N, H, W, P, Q, OC) = ctx.cfg, ctx.dims
        dout = dout.contiguous()
        rows = dout.numel() // OC
        act = L.ACT[cfg.act]
        dgamma = dbeta = None
        p_weight, p_bias, p_gamma, p_beta = ctx.params
        if cfg.has_bn:
            x, weight, y, mean, invstd, gamma, beta = ctx.saved_tensors
            G = ctx.G
            sums = zeros_f64(G * OC * 2, dout.device)
            # the per-channel sums  sum(dz), sum(dz * xhat)  feed dy in training mode and ARE the affine gradients
            # (dbeta, dgamma) in either mode: fine-tuning with frozen statistics (eval() + grad) still trains gamma / beta
            affine = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
            if cfg.training or affine:
                L.call('vs_bn_act_backward_reduce', ptr(dout), ptr(y), L.dtype_code(y), rows, OC, G, ptr(mean),
                     ptr(invstd), ptr(gamma), ptr(beta), act, ptr(sums), L.stream())
                if affine:
                    dgamma, dbeta = _grad_buffer(p_gamma), _grad_buffer(p_beta)
            dy = torch.empty_like(y)
            L.call('vs_bn_act_backward_apply', ptr(dout), ptr(y), ptr(dy), L.dtype_code(y), rows, OC, G, ptr(mean),
                 ptr(invstd), ptr(gamma), ptr(beta), act, ptr(sums), int(cfg.training),
                 ptr(dgamma[0]) if dgamma else None, ptr(dbeta[0]) if dbeta else None, L.stream())
        else:
            x, weight, out = ctx.saved_tensors
            if act != 0:
                dy = torch.empty_like(out)
                L.call('vs_act_backward', ptr(dout), ptr(out), ptr(dy), L.dtype_code(out), out.numel(), act, L.stream())
            else:
                dy = dout
        dx, dw, db = _conv_backward(ctx, x, weight, dy, p_weight, p_bias, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                    ctx.needs_input_grad[2], cfg.has_bn and cfg.training)
        if _grad_hooks[1] is not None:
            _grad_hooks[1](ctx.params)
        return (dx, dw, db, dgamma[1] if dgamma else None, dbeta[1] if dbeta else None, None, None, None, None)


# ------------------------------------------------------------------------------------------------
# fused decoder tail: [conv + BatchNorm + activation] -> thin ConvTranspose2d(nf, nc <= 2, 4, 2, 1) + output activation
# (DCGAN64Decoder.conv[2], conv[3]; conv.py:255-263).  The normalised tensor between the two layers — the largest
# activation of the step, 268 MB for the Moving-MNIST configuration — and its gradient are never written to HBM:
# BatchNorm + activation ride the operand path of the thin layer's kernels, and the thin layer's input gradient is
# recomputed in the epilogues of the two BatchNorm-backward passes (include/varsep.h, "fused decoder tail").
# ------------------------------------------------------------------------------------------------
def _tail_eligible(geom):
    """1 when libvarsep's fused tail kernels accept the thin transposed convolution ``geom`` (host-side query)."""
    return L.load().vs_tail_eligible(geom) == 1


def _tail_fused_backward():
    import os
    return os.environ.get('VARSEP_TAIL_FUSED_BACKWARD', '0') == '1'


def tail_geom(cfg3, dtype, N, P, Q):
    H = (P - 1) * cfg3.stride - 2 * cfg3.pad + cfg3.R
    W = (Q - 1) * cfg3.stride - 2 * cfg3.pad + cfg3.S
    return _geom(cfg3, dtype, N, H, W, P, Q, L.ACT[cfg3.act], 1), H, W


class DecoderTailFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, w2, b2, gamma, beta, rmean, rvar, nbt, w3, b3, cfg2, cfg3):
        L.require_cuda(x, w2, w3)
        G = cfg2.groups if cfg2.training else 1
        OC = cfg2.K if cfg2.kind == 'conv' else cfg2.C
        x, y, stats = _conv_forward(ctx, x, w2, b2, cfg2, 0, G, G * OC * 2 if cfg2.training else 0)
        rows = y.numel() // OC
        mean = torch.empty(G * OC, device=x.device, dtype=torch.float32)
        invstd = torch.empty_like(mean)
        if cfg2.training:
            L.call('vs_bn_finalize', ptr(stats), G, OC, rows // G, cfg2.eps, cfg2.momentum, ptr(mean), ptr(invstd),
                   ptr(rmean), ptr(rvar), ptr(nbt), L.stream())
        else:
            L.call('vs_bn_eval_stats', ptr(rmean), ptr(rvar), OC, cfg2.eps, ptr(mean), ptr(invstd), L.stream())
        N, P3, Q3 = y.shape[0], y.shape[1], y.shape[2]
        g3, H3, W3 = tail_geom(cfg3, y.dtype, N, P3, Q3)
        wp3 = packed_weight(w3, cfg3.K, cfg3.C, cfg3.R * cfg3.S, True, y.dtype)
        out = torch.empty((N, H3, W3, cfg3.C), device=x.device, dtype=y.dtype)
        L.call('vs_tail_forward', g3, ptr(y), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), G, L.ACT[cfg2.act], ptr(wp3),
               ptr(b3), ptr(out), L.stream())
        ctx.save_for_backward(x, w2, y, mean, invstd, gamma, beta, w3, out)
        ctx.G, ctx.cfg2, ctx.cfg3, ctx.g3 = G, cfg2, cfg3, g3
        ctx.params = (w2, b2, gamma, beta, w3, b3)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w2, y, mean, invstd, gamma, beta, w3, out = ctx.saved_tensors
        p_w2, p_b2, p_gamma, p_beta, p_w3, p_b3 = ctx.params
        cfg2, cfg3, g3, G = ctx.cfg2, ctx.cfg3, ctx.g3, ctx.G
        OC = y.shape[-1]
        act2, act3 = L.ACT[cfg2.act], L.ACT[cfg3.act]
        dout = dout.contiguous()
        # ---- the thin layer: activation backward, bias gradient, weight gradient (act rebuilt from y on the operand path)
        if act3 != 0:
            dz3 = torch.empty_like(out)
            L.call('vs_act_backward', ptr(dout), ptr(out), ptr(dz3), L.dtype_code(out), out.numel(), act3, L.stream())
        else:
            dz3 = dout
        dw3 = db3 = None
        if p_b3 is not None and ctx.needs_input_grad[9]:
            db3 = _grad_buffer(p_b3)
            with _OnWgradStream(db3, dz3):
                L.call('vs_colsum', ptr(dz3), L.dtype_code(dz3), dz3.numel() // cfg3.C, cfg3.C, ptr(db3[0]), L.stream())
        if ctx.needs_input_grad[8]:
            dw3 = _grad_buffer(p_w3)
            with _OnWgradStream(dw3, y, dz3, mean, invstd):
                L.call('vs_tail_wgrad', g3, ptr(y), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), G, act2, ptr(dz3),
                       ptr(dw3[0]), L.stream())
        # ---- BatchNorm backward.  Default: the thin layer's input gradient is materialised once (bf16) and the two column
        # kernels stream it.  ``VARSEP_TAIL_FUSED_BACKWARD=1`` recomputes it inside both passes instead
        # (vs_tail_bn_backward: correct, less HBM traffic, but its one-row-per-thread epilogue costs more issue slots than
        # the traffic it saves — 281 + 269 us against 109 + 105 + 126 us on the B200).
        wp3d = packed_weight(w3, cfg3.K, cfg3.C, cfg3.R * cfg3.S, False, y.dtype)
        sums = zeros_f64(G * OC * 2, dout.device)
        dgamma = dbeta = None
        affine = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        if affine:
            dgamma, dbeta = _grad_buffer(p_gamma), _grad_buffer(p_beta)
        dy = torch.empty_like(y)
        if _tail_fused_backward():
            bn_args = (g3, ptr(y), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), G, act2, ptr(dz3), ptr(wp3d))
            if cfg2.training or affine:
                L.call('vs_tail_bn_backward', *bn_args, 0, int(cfg2.training), ptr(sums), None, None, None, L.stream())
            L.call('vs_tail_bn_backward', *bn_args, 1, int(cfg2.training), ptr(sums), ptr(dy),
                   ptr(dgamma[0]) if dgamma else None, ptr(dbeta[0]) if dbeta else None, L.stream())
        else:
            N, P3, Q3 = y.shape[0], y.shape[1], y.shape[2]
            rows = N * P3 * Q3
            g3d = L.Geom(g3.dtype, N, g3.H, g3.W, g3.C, P3, Q3, g3.K, g3.R, g3.S, g3.stride, g3.pad, 1, 0, 0)
            dact = torch.empty_like(y)
            L.call('vs_conv_forward', g3d, L.DIRECT, ptr(dz3), ptr(wp3d), None, ptr(dact), None, L.stream())
            if cfg2.training or affine:
                L.call('vs_bn_act_backward_reduce', ptr(dact), ptr(y), L.dtype_code(y), rows, OC, G, ptr(mean), ptr(invstd),
                       ptr(gamma), ptr(beta), act2, ptr(sums), L.stream())
            L.call('vs_bn_act_backward_apply', ptr(dact), ptr(y), ptr(dy), L.dtype_code(y), rows, OC, G, ptr(mean), ptr(invstd),
                   ptr(gamma), ptr(beta), act2, ptr(sums), int(cfg2.training),
                   ptr(dgamma[0]) if dgamma else None, ptr(dbeta[0]) if dbeta else None, L.stream())
        # ---- the convolution of the BatchNorm block
        dx, dw2, db2 = _conv_backward(ctx, x, w2, dy, p_w2, p_b2, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                      ctx.needs_input_grad[2], cfg2.training)
        if _grad_hooks[1] is not None:
            _grad_hooks[1](ctx.params)
        return (dx, dw2, db2, dgamma[1] if dgamma else None, dbeta[1] if dbeta else None, None, None, None,
                dw3[1] if dw3 else None, db3[1] if db3 else None, None, None)


def decoder_tail(x, block, last, groups=1):
    """``last(block(x))`` for a BatchNorm ConvBlock ``block`` followed by the thin transposed convolution ``last``
    (networks.conv.DCGAN64Decoder), fused when the kernels accept the geometry; the two plain launches otherwise."""
    conv2, bn2 = block[0], block[1]
    w2 = conv2.weight
    K2, C2, R2, S2 = tuple(w2.shape)
    cfg2 = ConvCfg(block.kind, K2, C2, R2, S2, conv2.stride[0], conv2.padding[0], block.act, groups, bool(bn2.training), True,
                   bn2.eps, bn2.momentum, 0)
    K3, C3, R3, S3 = tuple(last.weight.shape)
    cfg3 = ConvCfg('convT', K3, C3, R3, S3, last.stride[0], last.padding[0], last._act, 1, False, False, 0.0, 0.0, 0)
    N = x.shape[0]
    if block.kind == 'convT':
        P3 = (x.shape[1] - 1) * cfg2.stride - 2 * cfg2.pad + R2
        Q3 = (x.shape[2] - 1) * cfg2.stride - 2 * cfg2.pad + S2
    else:
        P3 = (x.shape[1] + 2 * cfg2.pad - R2) // cfg2.stride + 1
        Q3 = (x.shape[2] + 2 * cfg2.pad - S2) // cfg2.stride + 1
    folded = not bn2.training and _fold_eval_bn and not torch.is_grad_enabled()
    if folded or block.act not in (None, 'none', 'identity', 'relu', 'leaky_relu') or \
            not _tail_eligible(tail_geom(cfg3, x.dtype, N, P3, Q3)[0]):
        return last(block(x, groups), groups)
    if _grad_hooks[0] is not None and torch.is_grad_enabled():
        _grad_hooks[0]((w2, conv2.bias, bn2.weight, bn2.bias, last.weight, last.bias))
    return DecoderTailFn.apply(x, w2, conv2.bias, bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var,
                               bn2.num_batches_tracked, last.weight, last.bias, cfg2, cfg3)


def _grad_buffer(p):
    """(buffer the kernel accumulates into, what backward returns to autograd).

    Parameters owned by a ``FlatState`` (see optim.py) expose ``_vs_grad``: a persistent fp32 view of
    the flat gradient arena.  Kernels then accumulate straight into it and autograd gets ``None``
    (no per-parameter add kernels); otherwise a fresh zeroed tensor is returned the normal way."""
    g = getattr(p, '_vs_grad', None)
    if g is not None:
        return g, None
    z = torch.zeros_like(p, dtype=torch.float32)
    return z, z


# ------------------------------------------------------------------------------------------------
# inference-time BatchNorm folding (opt-in)
# ------------------------------------------------------------------------------------------------
_fold_eval_bn = False


def set_eval_bn_folding(flag):
    """In ``eval()`` mode under ``torch.no_grad()`` BatchNorm uses fixed running statistics, so
    ``act(BN(conv(x)))`` == ``act(conv'(x))`` with  w' = w * s,  b' = (b - running_mean) * s + beta,
    s = gamma / sqrt(running_var + eps)  per output channel: one launch (conv with fused bias + activation)
    instead of three, and the pre-BatchNorm tensor is never written.  Off by default: the folded weights round
    differently (1e-6 class in fp32), and it has not been timed yet."""
    global _fold_eval_bn
    _fold_eval_bn = bool(flag)


def _folded_conv(conv, bn, kind):
    """Persistent folded (weight, bias) of a conv + eval-mode BatchNorm pair, refreshed when any input changes."""
    w, b = conv.weight, conv.bias
    tag = (_tag(w), None if b is None else _tag(b), _tag(bn.weight), _tag(bn.bias),
           bn.running_mean._version, bn.running_var._version)
    hit = bn.__dict__.get('_vs_fold')
    if hit is not None and hit[0] == tag:
        return hit[1], hit[2]
    with torch.no_grad():
        s = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        oc_dim = 1 if kind == 'convT' else 0                    # ConvTranspose weights are [C_in, C_out, kh, kw]
        shape = [1] * w.dim()
        shape[oc_dim] = -1
        wf = w.detach().float() * s.view(shape)
        bf = (-bn.running_mean.float() if b is None else b.detach().float() - bn.running_mean.float()) * s + bn.bias.float()
        if hit is not None and hit[1].shape == wf.shape:        # keep the tensors (and their packed copies' cache keys)
            hit[1].copy_(wf)
            hit[2].copy_(bf)
            wf, bf = hit[1], hit[2]
    bn.__dict__['_vs_fold'] = (tag, wf, bf)
    return wf, bf


def conv_block(x, conv, bn=None, act=None, kind='conv', groups=1, wshape=None, flags=0):
    """Run ``conv`` (nn.Conv2d / nn.ConvTranspose2d / nn.Linear used as parameter containers)
    followed by the optional BatchNorm module ``bn`` and activation name ``act`` on NHWC ``x``."""
    w = conv.weight
    if wshape is None:
        if w.dim() == 2:
            wshape = (w.shape[0], w.shape[1], 1, 1)
        else:
            wshape = tuple(w.shape)
    K, C, R, S = wshape
    stride = conv.stride[0] if hasattr(conv, 'stride') else 1
    pad = conv.padding[0] if hasattr(conv, 'padding') else 0
    training = bool(bn.training) if bn is not None else False
    if bn is not None and not training and _fold_eval_bn and not torch.is_grad_enabled():
        wf, bf = _folded_conv(conv, bn, kind)
        cfg = ConvCfg(kind, K, C, R, S, stride, pad, act, 1, False, False, 0.0, 0.0, flags)
        return ConvBlockFn.apply(x, wf, bf, None, None, None, None, None, cfg)
    cfg = ConvCfg(kind, K, C, R, S, stride, pad, act, groups, training, bn is not None,
                  bn.eps if bn is not None else 0.0, bn.momentum if bn is not None else 0.0, flags)
    if _grad_hooks[0] is not None and torch.is_grad_enabled():
        _grad_hooks[0]((w, conv.bias, bn.weight, bn.bias) if bn is not None else (w, conv.bias))
    if bn is not None:
        return ConvBlockFn.apply(x, w, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                 bn.num_batches_tracked, cfg)
    return ConvBlockFn.apply(x, w, conv.bias, None, None, None, None, None, cfg)


# ------------------------------------------------------------------------------------------------
# layout boundary
# ------------------------------------------------------------------------------------------------
class ToInternalFn(torch.autograd.Function):
    """fp32 NCHW [N,C,H,W] (reference layout) -> NHWC [N,H,W,C] of the compute dtype."""

    @staticmethod
    def forward(ctx, x, dtype):
        L.require_cuda(x)
        x = x.contiguous().float()
        N, Cc, H, W = x.shape
        out = torch.empty((N, H, W, Cc), device=x.device, dtype=dtype)
        L.call('vs_nchw_to_nhwc', ptr(x), ptr(out), L.dtype_code(out), N, Cc, H, W, L.stream())
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        N, H, W, Cc = d.shape
        out = torch.empty((N, Cc, H, W), device=d.device, dtype=torch.float32)
        L.call('vs_nhwc_to_nchw', ptr(d), L.dtype_code(d), ptr(out), N, Cc, H, W, L.stream())
        return out, None


class ToExternalFn(torch.autograd.Function):
    """NHWC compute dtype -> fp32 NCHW."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, Cc = x.shape
        ctx.dtype = x.dtype
        out = torch.empty((N, Cc, H, W), device=x.device, dtype=torch.float32)
        L.call('vs_nhwc_to_nchw', ptr(x), L.dtype_code(x), ptr(out), N, Cc, H, W, L.stream())
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous().float()
        N, Cc, H, W = d.shape
        out = torch.empty((N, H, W, Cc), device=d.device, dtype=ctx.dtype)
        L.call('vs_nchw_to_nhwc', ptr(d), ptr(out), L.dtype_code(out), N, Cc, H, W, L.stream())
        return out


def to_internal(x):
    """[N,C,H,W] or [N,C] fp32 -> NHWC compute dtype."""
    if x.dim() == 2:
        x = x.view(x.shape[0], x.shape[1], 1, 1)
    return ToInternalFn.apply(x, _compute_dtype)


def to_external(x):
    return ToExternalFn.apply(x)


def frames_window(frames, t0, nt, out=None, row0=0):
    """frames [B,T,C,H,W] fp32 -> time-folded NHWC [B,H,W,nt*C] (conv.py:90); no gradient (data).
    With ``out``/``row0`` the window is written into rows [row0, row0+B) of a larger batch."""
    L.require_cuda(frames)
    frames = frames.contiguous()
    B, T, Cf, H, W = frames.shape
    if out is None:
        out = torch.empty((B, H, W, nt * Cf), device=frames.device, dtype=_compute_dtype)
        row0 = 0
    dst = out[row0:row0 + B]
    L.call('vs_frames_to_nhwc', ptr(frames), B, T, Cf, H, W, t0, nt, ptr(dst), L.dtype_code(out), L.stream())
    return out


# ------------------------------------------------------------------------------------------------
# element-wise
# ------------------------------------------------------------------------------------------------
class AddActFn(torch.autograd.Function):
    """act(a + b): residual connections (resnet.py:29,69; conv.py:465-466)."""

    @staticmethod
    def forward(ctx, a, b, act):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        L.call('vs_add_act', ptr(a), ptr(b), ptr(out), L.dtype_code(a), a.numel(), L.ACT[act], L.stream())
        ctx.act = L.ACT[act]
        if ctx.act:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        if ctx.act:
            out, = ctx.saved_tensors
            dx = torch.empty_like(out)
            L.call('vs_act_backward', ptr(d), ptr(out), ptr(dx), L.dtype_code(out), out.numel(), ctx.act, L.stream())
            d = dx
        return d, d, None


def add_act(a, b, act=None):
    return AddActFn.apply(a, b, act)


class ConcatFn(torch.autograd.Function):
    """Channel concatenation of NHWC tensors; an input with fewer rows is broadcast over the groups
    (one reference call's S code / skip tensors feeding G batched decoder calls)."""

    @staticmethod
    def forward(ctx, *xs):
        xs = [x.contiguous() for x in xs]
        rows = max(x.numel() // x.shape[-1] for x in xs)
        big = max(xs, key=lambda x: x.numel() // x.shape[-1])
        Ct = sum(x.shape[-1] for x in xs)
        out = torch.empty(big.shape[:-1] + (Ct,), device=big.device, dtype=big.dtype)
        off, meta = 0, []
        for x in xs:
            c, r = x.shape[-1], x.numel() // x.shape[-1]
            L.call('vs_copy_channels', ptr(x), c, r, ptr(out), Ct, off, rows, L.dtype_code(out), L.stream())
            meta.append((tuple(x.shape), c, r, off))
            off += c
        ctx.meta, ctx.rows, ctx.Ct = meta, rows, Ct
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        grads = []
        for i, (shape, c, r, off) in enumerate(ctx.meta):
            if not ctx.needs_input_grad[i]:
                grads.append(None)
                continue
            g = torch.empty(shape, device=d.device, dtype=d.dtype)
            L.call('vs_slice_channels_reduce', ptr(d), ctx.Ct, off, ctx.rows, ptr(g), c, r, L.dtype_code(d), L.stream())
            grads.append(g)
        return tuple(grads)


def concat_channels(*xs):
    return ConcatFn.apply(*xs)


class MulBcastFn(torch.autograd.Function):
    """s * t with s broadcast over groups ('mul' mixing: conv.py:223, mlp_encdec.py:47)."""

    @staticmethod
    def forward(ctx, s, t):
        s, t = s.contiguous(), t.contiguous()
        Cc = t.shape[-1]
        out = torch.empty_like(t)
        L.call('vs_mul_bcast', ptr(s), s.numel() // Cc, ptr(t), ptr(out), t.numel() // Cc, Cc, L.dtype_code(t), L.stream())
        ctx.save_for_backward(s, t)
        return out

    @staticmethod
    def backward(ctx, d):
        s, t = ctx.saved_tensors
        d = d.contiguous()
        Cc = t.shape[-1]
        ds = torch.empty_like(s) if ctx.needs_input_grad[0] else None
        dt = torch.empty_like(t) if ctx.needs_input_grad[1] else None
        L.call('vs_mul_bcast_backward', ptr(d), ptr(s), s.numel() // Cc, ptr(t), ptr(ds), ptr(dt), t.numel() // Cc, Cc,
             L.dtype_code(t), L.stream())
        return ds, dt


def mul_bcast(s, t):
    return MulBcastFn.apply(s, t)


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        x = x.contiguous()
        N, H, W, Cc = x.shape
        P, Q = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty((N, P, Q, Cc), device=x.device, dtype=x.dtype)
        L.call('vs_maxpool_forward', ptr(x), ptr(y), L.dtype_code(x), N, H, W, Cc, k, stride, pad, L.stream())
        ctx.save_for_backward(x)
        ctx.geo = (k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, d):
        x, = ctx.saved_tensors
        d = d.contiguous()
        N, H, W, Cc = x.shape
        dx = torch.empty_like(x)
        L.call('vs_maxpool_backward', ptr(x), ptr(d), ptr(dx), L.dtype_code(x), N, H, W, Cc, *ctx.geo, L.stream())
        return dx, None, None, None


def maxpool(x, k=2, stride=2, pad=0):
    return MaxPoolFn.apply(x, k, stride, pad)


class Upsample2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, Cc = x.shape
        y = torch.empty((N, 2 * H, 2 * W, Cc), device=x.device, dtype=x.dtype)
        L.call('vs_upsample2_forward', ptr(x), ptr(y), L.dtype_code(x), N, H, W, Cc, L.stream())
        return y

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        N, H2, W2, Cc = d.shape
        dx = torch.empty((N, H2 // 2, W2 // 2, Cc), device=d.device, dtype=d.dtype)
        L.call('vs_upsample2_backward', ptr(d), ptr(dx), L.dtype_code(d), N, H2 // 2, W2 // 2, Cc, L.stream())
        return dx


def upsample2(x):
    return Upsample2Fn.apply(x)


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def _btl(t):
    """Describe an fp32 tensor as [B][T][L] with a contiguous inner run of L elements."""
    if t.dim() >= 3:
        inner = t[0, 0]
        if inner.is_contiguous() or inner.numel() == 1:
            return t.shape[0], t.shape[1], inner.numel(), t.stride(0), t.stride(1)
    if not t.is_contiguous():
        raise ValueError('loss operand must be contiguous or a [B,T,...] view with contiguous frames')
    return 1, 1, t.numel(), 0, 0


class SplitFirstGroupFn(torch.autograd.Function):
    """frames [G, B, ...] -> (frames[0], frames[1:]).  Autograd's own select / slice backward would zero-fill two
    full-size buffers, copy each gradient into its own and add them (two 33 MB fills + a 100 MB add per step for the
    mnist decoder output); here both gradients are copied into ONE buffer."""

    @staticmethod
    def forward(ctx, frames):
        ctx.meta = (frames.shape, frames.dtype, frames.device)
        return frames[0], frames[1:]

    @staticmethod
    def backward(ctx, g_first, g_rest):
        shape, dtype, device = ctx.meta
        out = torch.empty(shape, device=device, dtype=dtype)
        if g_first is None:
            out[0].zero_()
        else:
            out[0].copy_(g_first)
        if g_rest is None:
            out[1:].zero_()
        else:
            out[1:].copy_(g_rest)
        return out


def split_first_group(frames):
    return SplitFirstGroupFn.apply(frames)


class LossTermsFn(torch.autograd.Function):
    """All squared-error terms of the step in one pass each, combined on the device.

    ``pairs[i]`` is a list of (a, b) operand pairs summed into term i (b may be None: sum of a^2);
    term_i = coef_i * sum;  output = [term_0 .. term_{n-1}, sum_i lamb_i * term_i].
    Replaces F.mse_loss / pow-mean / sum reductions of train.py:42,86,139,145-149.
    Gradients flow to every ``a`` and to ``b`` when it requires grad.
    """

    @staticmethod
    def forward(ctx, spec, *tensors):
        # spec: list of (coef, lamb, [(ia, ib or -1), ...]) indexing into tensors
        n = len(spec)
        dev = tensors[0].device
        acc = torch.zeros(n, device=dev, dtype=torch.float64)
        descs = []
        for i, (coef, lamb, pairs) in enumerate(spec):
            for ia, ib in pairs:
                a = tensors[ia]
                b = tensors[ib] if ib >= 0 else None
                assert a.dtype == torch.float32 and (b is None or b.dtype == torch.float32)
                # a gradient is written with the operand's own strides: those must tile a dense buffer
                if ctx.needs_input_grad[1 + ia] and not _dense_strides(a):
                    a = a.contiguous()
                if b is not None and ctx.needs_input_grad[1 + ib] and not _dense_strides(b):
                    b = b.contiguous()
                Ba, Ta, La, asb, ast = _btl(a)
                if b is not None:
                    Bb, Tb, Lb, bsb, bst = _btl(b)
                    if (Ba, Ta, La) != (Bb, Tb, Lb):
                        # fall back to a common flat description
                        a, b = a.contiguous(), b.contiguous()
                        Ba, Ta, La, asb, ast = 1, 1, a.numel(), 0, 0
                        bsb, bst = 0, 0
                else:
                    bsb = bst = 0
                slot = C_double_ptr(acc, i)
                L.call('vs_sqdiff_sum', ptr(a), asb, ast, ptr(b), bsb, bst, Ba, Ta, La, slot, L.stream())
                descs.append((i, ia, ib, a, b, (Ba, Ta, La, asb, ast, bsb, bst)))
        terms = torch.empty(n + 1, device=dev, dtype=torch.float32)
        coef = (L.C.c_double * n)(*[s[0] for s in spec])
        lamb = (L.C.c_double * n)(*[s[1] for s in spec])
        L.call('vs_loss_combine', ptr(acc), coef, lamb, n, ptr(terms), L.stream())
        ctx.spec, ctx.descs, ctx.n = spec, descs, n
        ctx.shapes = [(t.shape, t.stride(), t.requires_grad) for t in tensors]
        return terms

    @staticmethod
    def backward(ctx, dterms):
        dterms = dterms.contiguous()
        n = ctx.n
        grads = [None] * len(ctx.shapes)
        g_total = dterms[n:n + 1]
        for (i, ia, ib, a, b, (B, T, Ln, asb, ast, bsb, bst)) in ctx.descs:
            coef, lamb, _ = ctx.spec[i]
            g_term = dterms[i:i + 1]
            for which, idx in ((0, ia), (1, ib)):
                if idx < 0 or not ctx.needs_input_grad[1 + idx]:
                    continue
                tgt = a if which == 0 else b
                first = grads[idx] is None
                if first:
                    # gradient laid out exactly like the operand (dense strides were enforced in forward)
                    grads[idx] = torch.empty_strided(tgt.shape, tgt.stride(), device=tgt.device, dtype=torch.float32)
                gbuf = grads[idx]
                sgn = 2.0 * coef if which == 0 else -2.0 * coef
                if which == 0:
                    L.call('vs_sqdiff_backward', ptr(a), asb, ast, ptr(b), bsb, bst, B, T, Ln, sgn, g_term, g_total,
                         lamb, ptr(gbuf), 0 if first else 1, L.stream())
                else:
                    # d/db (a-b)^2 = -2 (a-b): same kernel, destination indexed like b
                    L.call('vs_sqdiff_backward', ptr(b), bsb, bst, ptr(a), asb, ast, B, T, Ln, -sgn, g_term, g_total,
                         lamb, ptr(gbuf), 0 if first else 1, L.stream())
        return (None,) + tuple(grads)


def _dense_strides(t):
    """True if t's strides are a permutation of a dense layout (every element of the buffer is covered)."""
    order = sorted(range(t.dim()), key=lambda i: -t.stride(i))
    expect = 1
    for i in reversed(order):
        if t.shape[i] != 1 and t.stride(i) != expect:
            return False
        expect *= t.shape[i]
    return True


def C_double_ptr(t, i):
    return t[i:i + 1]


def loss_terms(spec, tensors):
    return LossTermsFn.apply(spec, *tensors)


# ------------------------------------------------------------------------------------------------
# latent rollout (MLP residual stepper), one launch per direction
# ------------------------------------------------------------------------------------------------
class LatentRolloutFn(torch.autograd.Function):
    """codes[t+1] = stepper(codes[t]) for t < T-1 (model.py:78-83 / resnet.py:42-50) as one kernel; the
    backward is the adjoint recurrence as one kernel plus three (T-1)*B-row weight-gradient GEMMs per block.
    Inputs: t0 [B,d] fp32 and, per block, W1,b1,W2,b2,W3,b3.  Returns (codes [T,B,d], res [nb,T-1,B,d]).
    ``res`` (the reference's ``t_residuals``, model.py:80-82) is returned for inspection only and is marked
    non-differentiable: the training objective (train.py:120-149) never differentiates through it."""

    @staticmethod
    def forward(ctx, T, t0, *weights):
        L.require_cuda(t0)
        nb = len(weights) // 6
        B, d = t0.shape
        h = weights[0].shape[0]
        dev = t0.device
        codes = torch.empty((T, B, d), device=dev, dtype=torch.float32)
        codes[0].copy_(t0)
        n = max(T - 1, 0)
        hidden = torch.empty((nb, 2, n, B, h), device=dev, dtype=torch.float32)
        xin = torch.empty((nb, n, B, d), device=dev, dtype=torch.float32)
        res = torch.empty((nb, n, B, d), device=dev, dtype=torch.float32)
        ws = [w.detach().contiguous() for w in weights]
        assert all(w.dtype == torch.float32 for w in ws)
        L.call('vs_latent_rollout_forward', ptr(codes), L.pointer_array(ws), T, B, d, h, nb, ptr(hidden), ptr(xin), ptr(res),
               L.stream())
        ctx.save_for_backward(hidden, xin, *ws)
        ctx.dims, ctx.params = (T, B, d, h, nb), weights
        ctx.mark_non_differentiable(res)
        return codes, res

    @staticmethod
    def backward(ctx, dcodes, _dres):
        T, B, d, h, nb = ctx.dims
        hidden, xin = ctx.saved_tensors[:2]
        ws = ctx.saved_tensors[2:]
        dcodes = dcodes.contiguous().clone()
        n = T - 1
        dev = dcodes.device
        dres = torch.empty((nb, n, B, d), device=dev, dtype=torch.float32)
        dhidden = torch.empty((nb, 2, n, B, h), device=dev, dtype=torch.float32)
        if n > 0:
            # the adjoint products need W^T: packed fp32 transposes, cached per parameter like every packed weight
            wt = []
            for j in range(nb):
                w1, b1, w2, b2, w3, b3 = ctx.params[6 * j:6 * j + 6]
                wt += [packed_weight(w1, h, d, 1, True, torch.float32), ws[6 * j + 1],
                       packed_weight(w2, h, h, 1, True, torch.float32), ws[6 * j + 3],
                       packed_weight(w3, d, h, 1, True, torch.float32), ws[6 * j + 5]]
            L.call('vs_latent_rollout_backward', ptr(dcodes), L.pointer_array(wt), T, B, d, h, nb, ptr(hidden), ptr(dres),
                   ptr(dhidden), L.stream())
        grads = []
        rows = n * B

        def lin_grads(p_w, p_b, small, big, K, Cc):
            # dW[K][Cc] += small[rows,K]^T big[rows,Cc];  db[K] += column sums of small
            gw, gb = _grad_buffer(p_w), _grad_buffer(p_b)
            if rows > 0:
                g = L.Geom(L.VS_F32, rows, 1, 1, Cc, 1, 1, K, 1, 1, 1, 0, 1, 0, 0)
                with _OnWgradStream(gw if gb[1] is None else gb, small, big):
                    L.call('vs_conv_wgrad', g, ptr(small), ptr(big), ptr(gw[0]), L.stream())
                    L.call('vs_colsum', ptr(small), L.VS_F32, rows, K, ptr(gb[0]), L.stream())
            return gw[1], gb[1]

        for j in range(nb):
            w1, b1, w2, b2, w3, b3 = ctx.params[6 * j:6 * j + 6]
            g1 = lin_grads(w1, b1, dhidden[j, 0], xin[j], h, d)
            g2 = lin_grads(w2, b2, dhidden[j, 1], hidden[j, 0], h, h)
            g3 = lin_grads(w3, b3, dres[j], hidden[j, 1], d, h)
            grads += [g1[0], g1[1], g2[0], g2[1], g3[0], g3[1]]
        if _grad_hooks[1] is not None:
            _grad_hooks[1](ctx.params)
        return (None, dcodes[0]) + tuple(grads)


def latent_rollout(t0, stepper, T):
    """t0 [B,d] fp32; stepper: networks.resnet.MLPResnet.  -> codes [T,B,d] fp32, res [nb,T-1,B,d]."""
    weights = []
    for blk in stepper.blocks:
        lins = [m[-1] for m in blk.mlp.module]
        assert len(lins) == 3
        for lin in lins:
            weights += [lin.weight, lin.bias]
    if _grad_hooks[0] is not None and torch.is_grad_enabled():
        _grad_hooks[0](weights)
    return LatentRolloutFn.apply(T, t0, *weights)

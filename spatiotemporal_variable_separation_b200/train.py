"""Training step (counterpart of /root/reference/var_sep/train.py).

``zero_order_loss``, ``ae_loss`` and ``train`` keep the reference signatures.  ``step_losses`` is
the GPU-organised body of one step (train.py:116-149): the same arithmetic, but the two Es calls,
the two Et calls and the 1 + nt_pred + offset decoder calls are each ONE grouped launch sequence
(BatchNorm statistics per call), and all loss terms are reduced by fused kernels.
"""
import contextlib
import os

import numpy as np
import torch
from tqdm import tqdm

from . import ops
from .networks.model import _external_codes
from .utils.helper import save


def zero_order_loss(s_code_old, s_code_new, skipco):
    """train.py:38-42 — mean((S_old - S_new)^2) over the code (and, with skipco, all skip tensors)."""
    if skipco:
        olds = [s_code_old[0]] + list(s_code_old[1])
        news = [s_code_new[0]] + list(s_code_new[1])
    else:
        olds, news = [s_code_old], [s_code_new]
    n = sum(t.numel() for t in olds)
    tensors = [t.float() for t in olds + news]
    k = len(olds)
    return ops.loss_terms([(1.0 / n, 1.0, [(i, k + i) for i in range(k)])], tensors)[0]


def draw_t_random(nt_cond, n_frames, offset):
    """train.py:72-75 — host RNG (numpy global state), exclusive upper bound."""
    if offset == 0:
        return int(np.random.randint(nt_cond, n_frames))
    return int(np.random.randint(nt_cond, n_frames + 1))


def ae_loss(cond, target, sep_net, nt_cond, offset, skipco, t_random=None):
    """train.py:45-88 through the public module calls (sequential form)."""
    full_data = torch.cat([cond, target], dim=1)
    s_code_old = sep_net.Es(full_data[:, :nt_cond], return_skip=skipco)
    s_code_new = sep_net.Es(full_data[:, -nt_cond:], return_skip=skipco)
    if t_random is None:
        t_random = draw_t_random(nt_cond, full_data.size(1), offset)
    t_code_random = sep_net.Et(full_data[:, t_random - nt_cond:t_random])
    if skipco:
        reconstruction = sep_net.decoder(s_code_old[0], t_code_random, skip=s_code_old[1])
    else:
        reconstruction = sep_net.decoder(s_code_old, t_code_random)
    supervision = full_data[:, t_random - offset]
    loss = ops.loss_terms([(1.0 / reconstruction.numel(), 1.0, [(0, 1)])], [reconstruction, supervision])[0]
    return loss, s_code_new, s_code_old


def step_losses(sep_net, full_data, nt_cond, nt_pred, offset, skipco, lamb_ae, lamb_s, lamb_t, lamb_pred,
                average_tloss=False, t_random=None, reducer=None, overlap_encoders=True):
    """One forward pass of the training objective.  ``full_data`` = cat(cond, target) [B,L,C,H,W] fp32.

    Returns dict(total, ae, s, pred, t, forecasts, t_codes); ``total.backward()`` then produces every
    parameter gradient.  Call order inside each network matches train.py:120-133 so the BatchNorm
    running statistics evolve identically (Es: old,new; Et: random,cond; decoder: AE, forecasts).
    """
    assert offset == nt_cond or offset == 0                                           # train.py:103
    B, n_frames = full_data.shape[0], full_data.shape[1]
    if t_random is None:
        t_random = draw_t_random(nt_cond, n_frames, offset)
    if full_data.is_cuda:
        ops.begin_step(full_data.device)          # one memset for all the small zeroed buffers of this step
    if reducer is not None:
        reducer.begin_step()                      # gradient buckets leave during backward, in completion order
    # ---- encoders: two calls each, batched as two BatchNorm groups.  Es and Et are independent until the decoder, and
    # their launches (and the 64-CTA latent rollout that follows Et) each fill only part of the GPU, so the content
    # encoder runs on a side stream next to the dynamic encoder + rollout; autograd replays the same split in backward.
    side = _side_stream(full_data) if overlap_encoders else None
    if side is not None:
        main = torch.cuda.current_stream(full_data.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            s_both = sep_net.Es.encode(sep_net.encoder_input(full_data, [0, n_frames - nt_cond]), 2, skipco)
    else:
        s_both = sep_net.Es.encode(sep_net.encoder_input(full_data, [0, n_frames - nt_cond]), 2, skipco)
    t_both = sep_net.Et.encode(sep_net.encoder_input(full_data, [t_random - nt_cond, 0]), 2)
    if side is not None:
        main.wait_stream(side)
        for t in ([s_both[0]] + list(s_both[1]) if skipco else [s_both]):
            t.record_stream(main)                                     # produced on `side`, consumed on `main`
    if skipco:
        s_code, s_skips = s_both
        s_old, s_new = s_code[:B], s_code[B:]
        skip_old, skip_new = [s[:B] for s in s_skips], [s[B:] for s in s_skips]
    else:
        s_old, s_new, skip_old, skip_new = s_both[:B], s_both[B:], None, None
    t_rand, t_cond = t_both[:B], t_both[B:]
    # ---- rollout + AE reconstruction and all forecasts in one grouped decode
    forecasts, t_codes, _, _, recon = sep_net.forecast_internal(s_old, skip_old, t_cond, nt_pred + offset, B,
                                                                extra_t=t_rand)
    # ---- losses
    supervision = full_data[:, t_random - offset]
    f_off = nt_cond if offset == 0 else 0
    target = full_data[:, f_off:]
    olds = [_loss_operand(s_old)] + ([_loss_operand(s) for s in skip_old] if skipco else [])
    news = [_loss_operand(s_new)] + ([_loss_operand(s) for s in skip_new] if skipco else [])
    t0 = _external_codes(t_cond).reshape(B, -1)
    tensors = [recon, supervision, forecasts, target, t0] + olds + news
    k = len(olds)
    n_s = sum(t.numel() for t in olds)
    # train.py:145-148: mean over everything, or sum over dim 1 (channels) then mean over the rest — for flat codes the
    # latter is 0.5/B, for feature-map codes [B,C,H,W] it is 0.5/(B*H*W)
    t_coef = 0.5 / t0.numel() if average_tloss else 0.5 * t_cond.shape[-1] / t0.numel()
    spec = [(1.0 / recon.numel(), lamb_ae, [(0, 1)]),
            (1.0 / n_s, lamb_s, [(5 + i, 5 + k + i) for i in range(k)]),
            (1.0 / forecasts.numel(), lamb_pred, [(2, 3)]),
            (t_coef, lamb_t, [(4, -1)])]
    terms = ops.loss_terms(spec, tensors)
    return dict(total=terms[4], ae=terms[0], s=terms[1], pred=terms[2], t=terms[3], terms=terms,
                forecasts=forecasts, t_codes=t_codes, t_random=t_random)


_SIDE_STREAMS = {}


def _side_stream(t):
    """One persistent side stream per device (None on CPU tensors, i.e. under the test emulator)."""
    if not t.is_cuda:
        return None
    dev = t.device.index
    if dev not in _SIDE_STREAMS:
        _SIDE_STREAMS[dev] = torch.cuda.Stream(device=t.device)
    return _SIDE_STREAMS[dev]


def _loss_operand(h):
    """internal tensor -> fp32 operand of the loss kernels (layout is irrelevant for a sum over all elements)."""
    if h.dtype == torch.float32:
        return h.contiguous()
    return ops.to_external(h.reshape(-1, *h.shape[-3:]))


class GraphedStep:
    """One whole training step — ``zero_grad`` + the objective of train.py:116-149 + ``backward`` + (data-parallel)
    gradient all-reduce + fused Adam — replayed as a captured CUDA graph.

    The only host-side decision inside a step is the draw ``t_random`` (train.py:72-75), which selects frame windows,
    so there is one graph per (batch shape, t_random) — at most ``nt_pred + 1`` per shape, all sharing one memory pool
    and one static input buffer per shape.  Everything else a step mutates lives in device memory the graph addresses
    directly: the parameter / gradient / moment arenas of ``FusedAdam``, its step counter and **learning rate**
    (``MultiStepLR`` therefore needs no re-capture), BatchNorm running statistics and ``num_batches_tracked``.

    ``graph=False`` runs the same body eagerly (debugging, or shapes seen once)."""

    def __init__(self, sep_net, optimizer, nt_cond, nt_pred, offset, skipco, lamb_ae, lamb_s, lamb_t, lamb_pred,
                 average_tloss=False, reducer=None, graph=True, overlap_encoders=True):
        self.net, self.opt, self.reducer = sep_net, optimizer, reducer
        self.args = (nt_cond, nt_pred, offset, skipco, lamb_ae, lamb_s, lamb_t, lamb_pred, average_tloss)
        self.nt_cond, self.nt_pred, self.offset = nt_cond, nt_pred, offset
        self.graph, self.overlap_encoders = bool(graph), overlap_encoders
        self.inputs, self.terms, self.graphs, self.pool = {}, {}, {}, None
        self._capture_stream = None
        self.dtype = ops.compute_dtype()                         # the graphs bake the compute dtype in

    # ---- static buffers ------------------------------------------------------------------------------------
    def input_buffer(self, shape, device):
        """The static [B, nt_cond+nt_pred, C, H, W] fp32 buffer the graphs of this batch shape read."""
        key = tuple(shape)
        if key not in self.inputs:
            self.inputs[key] = torch.zeros(key, device=device, dtype=torch.float32)
            self.terms[key] = torch.zeros(5, device=device, dtype=torch.float32)
        return self.inputs[key]

    def _body(self, key, t_random):
        self.opt.zero_grad()
        out = step_losses(self.net, self.inputs[key], *self.args, t_random=t_random, reducer=self.reducer,
                          overlap_encoders=self.overlap_encoders)
        out['total'].backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        self.terms[key].copy_(out['terms'].detach())

    # ---- one step ---------------------------------------------------------------------------------------------
    def run(self, shape, t_random=None):
        """Step on whatever the static input buffer of ``shape`` holds.  Returns the static 5-vector
        [ae, s, pred, t, total] (device memory, overwritten by the next step of this shape)."""
        key = tuple(shape)
        if t_random is None:
            t_random = draw_t_random(self.nt_cond, key[1], self.offset)
        prev = ops.compute_dtype()
        ops.set_compute_dtype(self.dtype)
        guard = torch.cuda.device(self.opt.flat_p.device) if self.opt.flat_p.is_cuda else contextlib.nullcontext()
        guard.__enter__()
        try:
            if not self.graph:
                self._body(key, t_random)
                return self.terms[key]
            g = self.graphs.get((key, t_random))
            if g is None:
                # warm up on a side stream (lazy allocations, packed-weight caches), then capture.  The step's own chain
                # is captured on a HIGH-priority stream: the weight-gradient side stream (ops._OnWgradStream) and the
                # content encoder's stream then fill the SMs the critical chain leaves free instead of competing with it
                if self._capture_stream is None:
                    prio = int(os.environ.get('VARSEP_STEP_STREAM_PRIORITY', '-1'))
                    self._capture_stream = torch.cuda.Stream(priority=prio)
                s = self._capture_stream
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._body(key, t_random)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self.pool, stream=s):
                    self._body(key, t_random)
                if self.pool is None:
                    self.pool = g.pool()
                self.graphs[(key, t_random)] = g
                return self.terms[key]             # the capture itself does not execute; the warm-up step did
            self.opt.sync_lr()
            g.replay()
            return self.terms[key]
        finally:
            guard.__exit__(None, None, None)
            ops.set_compute_dtype(prev)

    def __call__(self, cond, target, t_random=None):
        """``cond`` [B,nt_cond,C,H,W], ``target`` [B,nt_pred,C,H,W] on the host (ideally pinned) or the device."""
        shape = (cond.shape[0], cond.shape[1] + target.shape[1]) + tuple(cond.shape[2:])
        buf = self.input_buffer(shape, self.opt.flat_p.device)
        buf[:, :cond.shape[1]].copy_(cond, non_blocking=True)
        buf[:, cond.shape[1]:].copy_(target, non_blocking=True)
        return self.run(shape, t_random)

    def close(self):
        """Drop the captured graphs (they hold NCCL work when data-parallel: do this before destroying the group)."""
        self.graphs.clear()
        self.pool = None


class DevicePrefetcher:
    """Iterates a loader of host batches, keeping ONE batch in flight: while step k computes, batch k+1 travels
    host -> device on a copy stream into one of two staging buffers (pinned source memory makes the copy asynchronous;
    pageable memory still works, synchronously).  Yields device tensors that are valid until the next ``__next__``
    (main.py:111-114 uses a DataLoader with pin_memory; the reference then copies synchronously inside the loop)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        it = iter(self.loader)
        slots = [None, None]
        ready = [torch.cuda.Event(), torch.cuda.Event()]

        def stage(k):
            try:
                batch = next(it)
            except StopIteration:
                return False
            single = isinstance(batch, torch.Tensor)
            batch = (batch,) if single else tuple(batch)
            with torch.cuda.stream(self.copy_stream):
                if slots[k % 2] is None or any(a.shape != b.shape for a, b in zip(slots[k % 2][0], batch)):
                    slots[k % 2] = ([torch.empty(b.shape, device=self.device, dtype=b.dtype) for b in batch], single)
                for dst, src in zip(slots[k % 2][0], batch):
                    dst.copy_(src, non_blocking=True)
                ready[k % 2].record(self.copy_stream)
            return True

        k = 0
        more = stage(0)
        while more:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ready[k % 2])
            tensors, single = slots[k % 2]
            # the other slot was consumed by work already enqueued on the compute stream: the copy stream may reuse it
            # once that work is done
            self.copy_stream.wait_stream(cur)
            more = stage(k + 1)
            yield tensors[0] if single else tuple(tensors)
            k += 1


def train(xp_dir, train_loader, device, sep_net, optimizer, scheduler, use_apex_amp, use_torch_amp, epochs, lamb_ae,
          lamb_s, lamb_t, lamb_pred, offset, nt_cond, nt_pred, no_s, skipco, chkpt_interval, average_tloss,
          reducer=None, graph=True):
    """train.py:91-175.  Either AMP flag selects the bf16 tensor-core compute path (the analogue of the
    reference's fp16 autocast; bf16 needs no loss scaling).  With a ``FusedAdam`` optimizer every step is a CUDA-graph
    replay (``GraphedStep``) fed by a one-batch-ahead host->device prefetch; any other optimizer runs the same step
    eagerly.  ``reducer`` (parallel.GradReducer) makes the loop data-parallel."""
    from .optim import FusedAdam
    prev_dtype = ops.compute_dtype()
    if use_apex_amp or use_torch_amp:
        ops.set_compute_dtype(torch.bfloat16)
    if no_s:
        lamb_t = 0
        print("No regularization on T as there is no S")
    assert offset == nt_cond or offset == 0
    device = torch.device(device)
    fused = isinstance(optimizer, FusedAdam) and device.type == 'cuda'
    try:
        if fused:
            stepper = GraphedStep(sep_net, optimizer, nt_cond, nt_pred, offset, skipco, lamb_ae, lamb_s, lamb_t,
                                  lamb_pred, average_tloss, reducer=reducer, graph=graph)
        pb = tqdm(total=epochs * len(train_loader), ncols=0)
        for epoch in range(epochs):
            sep_net.train()
            batches = DevicePrefetcher(train_loader, device) if device.type == 'cuda' else train_loader
            for cond, target in batches:
                if fused:
                    stepper(cond, target)
                else:
                    cond, target = cond.to(device, non_blocking=True), target.to(device, non_blocking=True)
                    optimizer.zero_grad()
                    full_data = torch.cat([cond, target], dim=1)
                    out = step_losses(sep_net, full_data, nt_cond, nt_pred, offset, skipco, lamb_ae, lamb_s, lamb_t,
                                      lamb_pred, average_tloss, reducer=reducer)
                    out['total'].backward()
                    if reducer is not None:
                        reducer.finish()
                    optimizer.step()
                pb.update()
            if scheduler is not None:
                scheduler.step()
            if chkpt_interval is not None and (epoch + 1) % chkpt_interval == 0:
                save(xp_dir, sep_net, epoch_number=epoch + 1)
    except KeyboardInterrupt:
        pass
    finally:
        if fused:
            stepper.close()
        ops.set_compute_dtype(prev_dtype)
    save(xp_dir, sep_net)

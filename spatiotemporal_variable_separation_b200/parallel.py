"""Data-parallel training over the GPUs of one node (new functionality: the reference is single-device,
options.py:45).

Semantics: every rank holds a full replica and processes its own shard of the global batch (weak
scaling: `batch_size` sequences per GPU).  BatchNorm statistics stay per replica, so each replica's
forward pass is bit-identical to the single-GPU path on the same shard.  The only exchange step is the
gradient reduction: the flat gradient arena of ``FusedAdam`` is summed across ranks with NCCL
(NVLink 5 / NVSwitch) in three buckets, each launched on a side stream as soon as autograd has
finished the part of the model it covers — decoder, then the latent stepper, then the encoders — so
the transfer overlaps the rest of backward; the 1/world factor is folded into the Adam kernel.
The host draw ``t_random`` must be identical on every rank (seed numpy identically, or broadcast it).
"""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, sep_net, opt, group=None, overlap=True):
        self.opt, self.group, self.overlap = opt, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        opt.grad_scale = 1.0 / self.world
        # arena ranges of the four networks (parameters were registered Es, Et, decoder, t_resnet)
        index = {id(p): (off, p.numel()) for p, off in zip(opt.params, opt.offsets)}
        self.ranges = {}
        for name in ('Es', 'Et', 'decoder', 't_resnet'):
            spans = [index[id(p)] for p in getattr(sep_net, name).parameters() if id(p) in index]
            if spans:
                lo = min(o for o, _ in spans)
                hi = max(o + (n + 3) // 4 * 4 for o, n in spans)
                self.ranges[name] = (lo, hi)
        self.comm_stream = torch.cuda.Stream() if (overlap and opt.flat_g.is_cuda) else None
        self.pending = []
        self.done = set()

    # ---- bucket launch ------------------------------------------------------------------------
    def _reduce(self, names):
        if self.world == 1:
            return
        names = [n for n in names if n in self.ranges and n not in self.done]
        if not names:
            return
        self.done.update(names)
        lo = min(self.ranges[n][0] for n in names)
        hi = max(self.ranges[n][1] for n in names)
        bucket = self.opt.flat_g[lo:hi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.pending.append(dist.all_reduce(bucket, group=self.group, async_op=True))
        else:
            dist.all_reduce(bucket, group=self.group)

    # ---- hooks placed by the training step -------------------------------------------------------
    def after(self, tensor, names):
        """Reduce the gradient buckets ``names`` once the gradient w.r.t. ``tensor`` exists, i.e. once
        autograd has run every operator downstream of it."""
        if self.world > 1 and self.overlap and tensor.requires_grad:
            tensor.register_hook(lambda g: (self._reduce(names), g)[1])

    def finish(self):
        """After ``backward()``: reduce whatever is left and make the compute stream wait for all buckets."""
        self._reduce(('decoder', 't_resnet'))
        self._reduce(('Es', 'Et'))
        for work in self.pending:
            work.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.pending.clear()
        self.done.clear()


def broadcast_model(sep_net, src=0, group=None):
    """Identical replicas at start (parameters and BatchNorm buffers)."""
    for t in list(sep_net.parameters()) + list(sep_net.buffers()):
        dist.broadcast(t.data, src, group=group)

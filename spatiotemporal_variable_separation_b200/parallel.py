"""Data-parallel training over the GPUs of one node (new functionality: the reference is single-device,
options.py:45).

Semantics: every rank holds a full replica and processes its own shard of the global batch (weak
scaling: `batch_size` sequences per GPU).  BatchNorm statistics stay per replica, so each replica's
forward pass is bit-identical to the single-GPU path on the same shard.  The only exchange step is the
gradient reduction: the flat gradient arena of ``FusedAdam`` is summed across ranks with NCCL
(NVLink 5 / NVSwitch); the 1/world factor is folded into the Adam kernel.

Buckets follow the order in which backward COMPLETES gradients, not the order of registration: the arena
range of every network (Es, Et, decoder, t_resnet) is cut, from its END (the layers whose weight gradients
are finished first), into contiguous buckets of at most ``bucket_bytes``.  The backward operators announce
every finished parameter gradient (``ops.set_grad_ready_hook``); a bucket leaves on a side stream the moment
its last parameter is announced, so the transfer overlaps the rest of backward and only the last small
bucket (the first layers of the encoders, a few MB) is exposed after it.  A parameter used several times in
one step (the convolutional stepper of the SST configuration) counts as finished after its last use.

The host draw ``t_random`` must be identical on every rank (seed numpy identically, or broadcast it).
"""
import torch
import torch.distributed as dist

from . import _lib as L
from . import ops


class PeerExchange:
    """The gradient arenas of all ranks of ONE node mapped into every process (CUDA IPC over NVLink / NVSwitch) and the
    hand-written all-reduce over them (csrc/peer.cu: ``vs_peer_barrier`` + ``vs_peer_allreduce``): rank r sums slice r of
    all arenas in rank order and writes the result into all arenas.  Graph-capturable (three kernel launches with static
    arguments; the barrier epoch lives in device memory), deterministic, replicas bit-identical.

    Raises if the arenas cannot be shared (ranks on different nodes, no peer access): the caller then uses NCCL."""

    def __init__(self, arena, group=None):
        assert arena.is_cuda and arena.dtype == torch.float32 and arena.is_contiguous()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError('peer exchange: at most 8 ranks (one NVSwitch domain)')
        self.n = arena.numel()
        if self.n % 4:
            raise RuntimeError('peer exchange: arena length must be a multiple of 4 elements')
        dev = arena.device
        self.flags = torch.zeros(16, device=dev, dtype=torch.int32)
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int32)
        torch.cuda.synchronize(dev)
        # export: (device index, cudaIpc handle of the allocation, ..., offset of this tensor's storage in it)
        mine = {'arena': (arena.untyped_storage()._share_cuda_(), arena.storage_offset(), arena.numel()),
                'flags': (self.flags.untyped_storage()._share_cuda_(), self.flags.storage_offset(), self.flags.numel())}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._keep = []                                  # peers' mapped storages must outlive the pointers
        self.arenas, self.flag_sets = [], []
        for r, rec in enumerate(everyone):
            if r == self.rank:
                self.arenas.append(arena)
                self.flag_sets.append(self.flags)
                continue
            for key, dtype, out in (('arena', torch.float32, self.arenas), ('flags', torch.int32, self.flag_sets)):
                handle, offset, numel = rec[key]
                storage = torch.UntypedStorage._new_shared_cuda(*handle)
                t = torch.empty(0, dtype=dtype, device=storage.device).set_(storage, offset, (numel,))
                self._keep.append(storage)
                out.append(t)
        for t in self.arenas:
            if t.device != dev:
                # peer access from this process's device to the peer's device is enabled by torch on the first copy
                torch.empty(4, device=dev).copy_(t[:4])
                if not torch.cuda.can_device_access_peer(dev.index, t.device.index):
                    raise RuntimeError(f'peer exchange: {dev} cannot access {t.device}')
        torch.cuda.synchronize(dev)
        self._arena_ptrs = L.pointer_array(self.arenas)
        self._flag_ptrs = L.pointer_array(self.flag_sets)
        self.local = arena
        dist.barrier(group=group)

    def barrier(self):
        with torch.cuda.device(self.local.device):
            lib = L.load()
            rc = lib.vs_peer_barrier(self._flag_ptrs, self.rank, self.world, self.epoch.data_ptr(),
                                     torch.cuda.current_stream(self.local.device).cuda_stream)
        if rc:
            raise RuntimeError(f'vs_peer_barrier failed: {lib.vs_last_error().decode()}')

    def all_reduce(self, max_blocks=0):
        """Sum of all ranks' arenas into every arena, on the current stream of the local device."""
        self.barrier()                                   # every rank's gradients are complete
        with torch.cuda.device(self.local.device):
            lib = L.load()
            rc = lib.vs_peer_allreduce(self._arena_ptrs, self.rank, self.world, self.n, int(max_blocks),
                                       torch.cuda.current_stream(self.local.device).cuda_stream)
        if rc:
            raise RuntimeError(f'vs_peer_allreduce failed: {lib.vs_last_error().decode()}')
        self.barrier()                                   # every rank's slice has landed in this arena


class GradReducer:
    def __init__(self, sep_net, opt, group=None, overlap=True, bucket_bytes=13 << 20, split=('Es', 'Et'),
                 early=('decoder', 't_resnet'), transport='peer'):
        self.opt, self.group, self.overlap = opt, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        opt.grad_scale = 1.0 / self.world
        # transport 'peer': the hand-written all-reduce over NVLink peer memory, once, after backward (measured on 8 B200:
        # shorter than the exposed part of any overlapped NCCL schedule, and it takes no SM from the backward kernels);
        # 'nccl': bucketed ncclAllReduce overlapped with backward (also the fallback when arenas cannot be shared)
        self.peer = None
        self.transport = transport
        if transport == 'peer' and self.world > 1 and opt.flat_g.is_cuda:
            try:
                self.peer = PeerExchange(opt.flat_g, group)
            except Exception as e:                       # different nodes, no P2P, IPC disabled ...
                import warnings
                warnings.warn(f'peer-memory gradient exchange unavailable ({e}); using NCCL')
            # all ranks must agree (a rank that failed would wait in NCCL while the others wait on peer flags)
            ok = torch.tensor([1 if self.peer is not None else 0], device=opt.flat_g.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok) == 0:
                self.peer = None
        index = {id(p): (off, p.numel()) for p, off in zip(opt.params, opt.offsets)}
        # Only the networks in ``split`` are cut into several buckets.  The decoder travels as ONE bucket that leaves when
        # its backward is over: its kernels are persistent and sized to all 148 SMs, so an all-reduce running next to them
        # (NCCL occupies SMs) pushes a few of their CTAs into a second wave and costs more than the overlap gains
        # (measured at 2 GPUs: 4.28 ms per step with 8 MB buckets everywhere against 4.17 ms with one decoder bucket).  The
        # encoders' backward is a chain of short, partially-filled launches: their deep layers (80 % of the parameters,
        # finished first) leave early and only the shallow layers' few MB are exposed after backward.
        cap_all = max(int(bucket_bytes) // 4, 1)
        # ---- buckets: per network, contiguous arena ranges cut from the end of the network's range
        self.buckets = []                      # dict(lo, hi, params=set(id), name)
        self.bucket_of = {}
        covered = set()
        for name in ('decoder', 't_resnet', 'Et', 'Es'):
            net = getattr(sep_net, name, None)
            if net is None:
                continue
            ps = [p for p in net.parameters() if id(p) in index and id(p) not in covered]
            ps.sort(key=lambda p: index[id(p)][0])
            cap = cap_all if name in split else 1 << 62
            cur = None
            for p in reversed(ps):
                off, n = index[id(p)]
                end = off + (n + 3) // 4 * 4
                if cur is None or cur['hi'] - off > cap or end != cur['lo']:
                    cur = dict(lo=off, hi=end, params=set(), name=f'{name}[{len(self.buckets)}]', early=name in early)
                    self.buckets.append(cur)
                cur['lo'] = min(cur['lo'], off)
                cur['params'].add(id(p))
                self.bucket_of[id(p)] = cur
                covered.add(id(p))
        # parameters outside the four networks (none today) travel in one trailing bucket
        rest = [p for p in opt.params if id(p) not in covered]
        if rest:
            lo = min(index[id(p)][0] for p in rest)
            hi = max(index[id(p)][0] + (index[id(p)][1] + 3) // 4 * 4 for p in rest)
            b = dict(lo=lo, hi=hi, params={id(p) for p in rest}, name='rest', early=False)
            self.buckets.append(b)
            for p in rest:
                self.bucket_of[id(p)] = b
        self.comm_stream = torch.cuda.Stream(device=opt.flat_g.device) if (overlap and opt.flat_g.is_cuda) else None
        self.pending, self.done = [], set()
        self._uses, self._left, self._streams = {}, {}, {}
        self._armed = False

    # ---- armed for one step by the training step (train.step_losses) ------------------------------------------
    def begin_step(self):
        """Called before the forward pass of a step: count parameter uses in forward, release buckets in backward."""
        self.pending.clear()
        self.done.clear()
        self._uses.clear()
        self._streams.clear()
        self._left = {id(b): len(b['params']) for b in self.buckets}
        self._armed = self.world > 1 and self.overlap and self.peer is None
        ops.set_grad_ready_hook(self._on_use if self._armed else None, self._on_ready if self._armed else None)

    def _on_use(self, params):
        for p in params:
            if p is not None:
                self._uses[id(p)] = self._uses.get(id(p), 0) + 1

    def _on_ready(self, params):
        for p in params:
            if p is None:
                continue
            k = id(p)
            b = self.bucket_of.get(k)
            if b is None or id(b) in self.done:
                continue
            left = self._uses.get(k, 1) - 1
            self._uses[k] = left
            if left > 0:
                continue
            if p.is_cuda:                                   # the stream this gradient was produced on
                s = torch.cuda.current_stream(p.device)
                self._streams.setdefault(id(b), {})[s.cuda_stream] = s
            self._left[id(b)] -= 1
            if self._left[id(b)] == 0 and b['early']:
                self._launch(b)

    # ---- bucket launch ----------------------------------------------------------------------------------------
    def _launch(self, b):
        if self.world == 1 or id(b) in self.done:
            return
        self.done.add(id(b))
        bucket = self.opt.flat_g[b['lo']:b['hi']]
        if self.comm_stream is not None:
            streams = self._streams.get(id(b)) or {}
            cur = torch.cuda.current_stream(bucket.device)
            streams.setdefault(cur.cuda_stream, cur)
            for s in streams.values():
                self.comm_stream.wait_stream(s)
            with torch.cuda.stream(self.comm_stream):
                self.pending.append(dist.all_reduce(bucket, group=self.group, async_op=True))
        else:
            dist.all_reduce(bucket, group=self.group)

    # kept for callers of the round-1 interface: a tensor-gradient hook that flushes whole networks
    def after(self, tensor, names):
        return None

    def finish(self):
        """After ``backward()``: reduce whatever is left (parameters that received no gradient this step keep their
        bucket waiting) and make the compute stream wait for all buckets."""
        ops.set_grad_ready_hook(None, None)
        self._armed = False
        if self.peer is not None:
            self.peer.all_reduce()
            self.done.update(id(b) for b in self.buckets)
            return
        # merge the leftovers into as few contiguous calls as possible
        left = sorted((b for b in self.buckets if id(b) not in self.done), key=lambda b: b['lo'])
        merged = []
        for b in left:
            if merged and merged[-1]['hi'] == b['lo']:
                merged[-1]['hi'] = b['hi']
                merged[-1]['ids'].append(id(b))
            else:
                merged.append(dict(lo=b['lo'], hi=b['hi'], ids=[id(b)]))
        for m in merged:
            self.done.update(m['ids'])
            if self.world == 1:
                continue
            bucket = self.opt.flat_g[m['lo']:m['hi']]
            if self.comm_stream is not None:
                self.comm_stream.wait_stream(torch.cuda.current_stream(bucket.device))
                with torch.cuda.stream(self.comm_stream):
                    self.pending.append(dist.all_reduce(bucket, group=self.group, async_op=True))
            else:
                dist.all_reduce(bucket, group=self.group)
        for work in self.pending:
            work.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.opt.flat_g.device).wait_stream(self.comm_stream)
        self.pending.clear()


def broadcast_model(sep_net, src=0, group=None):
    """Identical replicas at start (parameters and BatchNorm buffers)."""
    for t in list(sep_net.parameters()) + list(sep_net.buffers()):
        dist.broadcast(t.data, src, group=group)

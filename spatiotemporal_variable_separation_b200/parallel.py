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


def _cudart():
    import ctypes
    return ctypes.CDLL('libcudart.so.12')


def _ipc_open(handle_bytes, dev):
    """cudaIpcOpenMemHandle with ``dev`` CURRENT: the peer allocation is mapped into this device's address space and peer
    access dev -> exporting device is enabled (cudaIpcMemLazyEnablePeerAccess).  (torch's own ``_new_shared_cuda`` opens
    the handle under the EXPORTING device's guard — fine for copies, an illegal address for a kernel running on ``dev``.)
    -> base address of the peer allocation as an int."""
    import ctypes

    class Handle(ctypes.Structure):
        _fields_ = [('reserved', ctypes.c_char * 64)]

    rt = _cudart()
    rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), Handle, ctypes.c_uint]
    h = Handle()
    ctypes.memmove(ctypes.byref(h), bytes(handle_bytes), 64)
    out = ctypes.c_void_p()
    with torch.cuda.device(dev):
        rc = rt.cudaIpcOpenMemHandle(ctypes.byref(out), h, ctypes.c_uint(1))      # 1 = cudaIpcMemLazyEnablePeerAccess
    if rc != 0 or not out.value:
        rt.cudaGetLastError()
        raise RuntimeError(f'cudaIpcOpenMemHandle failed on {dev}: cuda error {rc}')
    return int(out.value)


class PeerExchange:
    """The gradient arenas of all ranks of ONE node mapped into every process (CUDA IPC over NVLink / NVSwitch) and the
    hand-written all-reduce over them (csrc/peer.cu: ``vs_peer_barrier`` + ``vs_peer_allreduce``): rank r sums slice r of
    all arenas in rank order and writes the result into all arenas.  Graph-capturable (three kernel launches with static
    arguments; the barrier epoch lives in device memory), deterministic, replicas bit-identical.

    Raises if the arenas cannot be shared (ranks on different nodes, no peer access): the caller then uses NCCL."""

    def __init__(self, arena, group=None, wire_dtype=torch.float32):
        assert arena.is_cuda and arena.dtype == torch.float32 and arena.is_contiguous()
        assert wire_dtype in (torch.float32, torch.bfloat16)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError('peer exchange: at most 8 ranks (one NVSwitch domain)')
        self.n = arena.numel()
        if self.n % 8:
            raise RuntimeError('peer exchange: arena length must be a multiple of 8 elements')
        dev = arena.device
        # bf16 wire format: gradients travel as bf16 (half the NVLink bytes), sums are taken in fp32
        self.wire = torch.zeros(self.n, device=dev, dtype=torch.bfloat16) if wire_dtype == torch.bfloat16 else None
        self.flags = torch.zeros(16, device=dev, dtype=torch.int32)
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int32)
        torch.cuda.synchronize(dev)

        def export(t):
            # torch's export of the caching-allocator block that holds ``t``: (device, 64-byte cudaIpcMemHandle of the
            # block's base, ..., byte offset of the storage inside the block, ...)
            sh = t.untyped_storage()._share_cuda_()
            # (newer torch prefixes the 64 handle bytes with a version byte and a type byte: 'c' = plain cudaMalloc block)
            raw = bytes(sh[1])
            if len(raw) > 64 and raw[-65:-64] not in (b'c',):
                raise RuntimeError('peer exchange: the arena lives in an expandable segment (no cudaIpc handle)')
            return {'device': int(sh[0]), 'handle': raw[-64:], 'offset': int(sh[3]) + t.storage_offset() * t.element_size()}

        mine = {'arena': export(self.wire if self.wire is not None else arena), 'flags': export(self.flags), 'n': self.n,
                'wire': str(wire_dtype)}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if any(rec['n'] != self.n or rec['wire'] != str(wire_dtype) for rec in everyone):
            raise RuntimeError('peer exchange: arenas of different sizes / wire formats')
        opened = {}                                      # one mapping per (peer, allocation block)
        arena_ptrs, flag_ptrs = [], []
        for r, rec in enumerate(everyone):
            for key, local, out in (('arena', self.wire if self.wire is not None else arena, arena_ptrs),
                                    ('flags', self.flags, flag_ptrs)):
                if r == self.rank:
                    out.append(local.data_ptr())
                    continue
                e = rec[key]
                if not torch.cuda.can_device_access_peer(dev.index, e['device']):
                    raise RuntimeError(f'peer exchange: {dev} cannot access cuda:{e["device"]}')
                k = (r, e['handle'])
                if k not in opened:
                    opened[k] = _ipc_open(e['handle'], dev)
                out.append(opened[k] + e['offset'])
        self._opened = list(opened.values())
        self._arena_base = arena_ptrs
        self._arena_ptrs = (L.C.c_void_p * self.world)(*arena_ptrs)
        self._range_ptrs = {}
        self._flag_ptrs = (L.C.c_void_p * self.world)(*flag_ptrs)
        self.local = arena
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)

    def close(self):
        """Unmap the peers' allocations (after every rank has stopped using them)."""
        rt = _cudart()
        import ctypes
        with torch.cuda.device(self.local.device):
            for base in self._opened:
                rt.cudaIpcCloseMemHandle(ctypes.c_void_p(base))
        self._opened = []

    def barrier(self):
        with torch.cuda.device(self.local.device):
            lib = L.load()
            rc = lib.vs_peer_barrier(self._flag_ptrs, self.rank, self.world, self.epoch.data_ptr(),
                                     torch.cuda.current_stream(self.local.device).cuda_stream)
        if rc:
            raise RuntimeError(f'vs_peer_barrier failed: {lib.vs_last_error().decode()}')

    def all_reduce(self, lo=0, hi=None, max_blocks=0):
        """Sum of all ranks' arenas (elements [lo, hi), multiples of 4) into every arena, on the current stream of the
        local device.  ``max_blocks`` caps the grid (a partial exchange running next to backward kernels)."""
        hi = self.n if hi is None else hi
        assert 0 <= lo <= hi <= self.n and lo % 8 == 0 and (hi - lo) % 8 == 0
        if hi == lo:
            return
        esize = 2 if self.wire is not None else 4
        if (lo, hi) not in self._range_ptrs:
            self._range_ptrs[(lo, hi)] = (L.C.c_void_p * self.world)(*[b + esize * lo for b in self._arena_base])
        ptrs = self._range_ptrs[(lo, hi)]
        lib = L.load()
        stream = torch.cuda.current_stream(self.local.device).cuda_stream
        if self.wire is not None:
            L.call('vs_grad_compress', self.local[lo:hi], self.wire[lo:hi], hi - lo, L.stream())
        self.barrier()                                   # every rank's gradients in this range are complete (and compressed)
        with torch.cuda.device(self.local.device):
            fn = lib.vs_peer_allreduce_bf16 if self.wire is not None else lib.vs_peer_allreduce
            rc = fn(ptrs, self.rank, self.world, hi - lo, int(max_blocks), stream)
        if rc:
            raise RuntimeError(f'vs_peer_allreduce failed: {lib.vs_last_error().decode()}')
        self.barrier()                                   # every rank's slice has landed in this rank's buffer
        if self.wire is not None:
            L.call('vs_grad_expand', self.wire[lo:hi], self.local[lo:hi], hi - lo, L.stream())


class GradReducer:
    def __init__(self, sep_net, opt, group=None, overlap=True, bucket_bytes=13 << 20, split=('Es', 'Et'),
                 early=('decoder', 't_resnet'), transport='peer', early_blocks=24, wire_dtype=None):
        self.opt, self.group, self.overlap = opt, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        opt.grad_scale = 1.0 / self.world
        # transport 'peer': the hand-written all-reduce over NVLink peer memory, once, after backward (measured on 8 B200:
        # shorter than the exposed part of any overlapped NCCL schedule, and it takes no SM from the backward kernels);
        # 'nccl': bucketed ncclAllReduce overlapped with backward (also the fallback when arenas cannot be shared)
        self.peer = None
        self.transport, self.early_blocks = transport, early_blocks
        if transport == 'peer' and self.world > 1 and opt.flat_g.is_cuda:
            try:
                # gradients computed in bf16 arithmetic travel as bf16 (the usual compressed exchange of bf16 training: half
                # the NVLink bytes, fp32 sums); the fp32 parity mode exchanges fp32
                if wire_dtype is None:
                    wire_dtype = torch.bfloat16 if ops.compute_dtype() == torch.bfloat16 else torch.float32
                self.peer = PeerExchange(opt.flat_g, group, wire_dtype)
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
        # (NCCL transport, round-1 build, 2 GPUs: 4.28 ms per step with 8 MB buckets everywhere against 4.17 ms with one decoder bucket).  The
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
                end = off + (n + 7) // 8 * 8
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
            hi = max(index[id(p)][0] + (index[id(p)][1] + 7) // 8 * 8 for p in rest)
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
        self._armed = self.world > 1 and self.overlap
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
            ops.join_wgrad(bucket.device, self.comm_stream)      # the weight gradients themselves run on a side stream
            with torch.cuda.stream(self.comm_stream):
                if self.peer is not None:
                    # a few CTAs only: the exchange runs next to the rest of backward and must not push the persistent
                    # kernels' CTAs into a second wave
                    self.peer.all_reduce(b['lo'], b['hi'], max_blocks=self.early_blocks)
                else:
                    self.pending.append(dist.all_reduce(bucket, group=self.group, async_op=True))
        elif self.peer is not None:
            self.peer.all_reduce(b['lo'], b['hi'])
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
        ops.join_wgrad(self.opt.flat_g.device)
        if self.peer is not None and self.comm_stream is not None:
            # the early partial exchanges share the flags / epoch with the final one: strictly one after the other
            torch.cuda.current_stream(self.opt.flat_g.device).wait_stream(self.comm_stream)
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
            if self.peer is not None:
                self.peer.all_reduce(m['lo'], m['hi'])
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

"""2-rank probe of parallel.PeerExchange: torchrun --nproc-per-node 2 scripts/probe/peer_probe.py"""
import os, sys, traceback
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
from spatiotemporal_variable_separation_b200.parallel import PeerExchange
try:
    n = 11_000_000 // 4 * 4
    arena = torch.full((n,), float(rank + 1), device='cuda')
    arena[:8] = torch.arange(8, device='cuda') * (rank + 1)
    px = PeerExchange(arena)
    print(rank, 'arenas', [hex(v) for v in px._arena_ptrs], 'flags', [hex(v) for v in px._flag_ptrs], flush=True)
    px.barrier(); torch.cuda.synchronize(); print(rank, 'barrier ok', int(px.epoch), px.flags[:4].tolist(), flush=True)
    from spatiotemporal_variable_separation_b200 import _lib as L
    lib = L.load()
    rc = lib.vs_peer_allreduce(px._arena_ptrs, px.rank, px.world, px.n, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); print(rank, 'allreduce kernel ok', rc, flush=True)
    dist.barrier()
    arena.fill_(float(rank + 1)); arena[:8] = torch.arange(8, device='cuda') * (rank + 1)
    torch.cuda.synchronize(); dist.barrier()
    px.all_reduce()
    torch.cuda.synchronize()
    w = dist.get_world_size()
    want = sum(range(1, w + 1))
    print(rank, 'ok', arena[:8].tolist(), float(arena[-1]), want, bool((arena[8:] == want).all()), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): px.all_reduce()
    e0.record()
    for _ in range(20): px.all_reduce()
    e1.record(); torch.cuda.synchronize()
    print(rank, 'all_reduce of', n * 4 / 1e6, 'MB:', e0.elapsed_time(e1) / 20 * 1e3, 'us', flush=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): px.all_reduce()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g): px.all_reduce()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    print(rank, 'graph replay ok', flush=True)
except Exception:
    traceback.print_exc()
dist.barrier(); dist.destroy_process_group()

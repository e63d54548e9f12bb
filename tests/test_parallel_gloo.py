"""Data-parallel host logic on CPU: world_size 2, gloo backend, C ABI emulated (tests/emu.py).

Checks the contract of SURVEY section 8e: after the bucketed all-reduce and the 1/world scaling folded into
Adam, every rank holds the MEAN of the per-shard gradients (BatchNorm statistics per replica), replicas
stay bit-identical after the optimizer step, and the overlapped hook path reduces every bucket exactly once.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from spatiotemporal_variable_separation_b200 import configs, ops, train as vs_train
    from spatiotemporal_variable_separation_b200.optim import FusedAdam
    from spatiotemporal_variable_separation_b200.parallel import GradReducer, broadcast_model
    from tests import emu, harness
    from tests.test_host_emulated import build_filled
    cfg = configs.preset('mnist', small=True)
    cfg['name'] = 'mnist-small'
    ops.set_compute_dtype(torch.float32)
    with emu.install():
        net = build_filled(cfg).train()
        broadcast_model(net)
        opt = FusedAdam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
        red = GradReducer(net, opt, overlap=True, bucket_bytes=16 << 10, early=('decoder', 't_resnet', 'Es', 'Et'))   # small buckets, all leave early
        cond, target = harness.inputs(cfg)
        full = torch.cat([cond, target], 1)
        shard = full[rank * 2:(rank + 1) * 2]                       # 2 sequences per rank
        def run(reducer):
            opt.zero_grad()
            out = vs_train.step_losses(net, shard, cfg['nt_cond'], cfg['nt_pred'], cfg['offset'], False, cfg['lamb_ae'],
                                       cfg['lamb_s'], cfg['lamb_t'], cfg['lamb_pred'], False, 6, reducer)
            out['total'].backward()

        run(None)                          # this rank's own gradient (train-mode BN: batch statistics only)
        local = opt.flat_g.clone()
        run(red)                           # same step with the bucket hooks active
        # every bucket whose parameters all received a gradient left DURING backward, in completion order:
        # the decoder's and the stepper's first
        left_early = [b['name'] for b in red.buckets if id(b) in red.done]
        assert any(n.startswith('decoder') for n in left_early) and any(n.startswith('t_resnet') for n in left_early), left_early
        assert len(left_early) >= len(red.buckets) - 2, (left_early, [b['name'] for b in red.buckets])
        red.finish()
        assert len(red.done) == len(red.buckets)
        summed = opt.flat_g.clone()
        opt.step()
        torch.save({'local': local, 'summed': summed, 'params': opt.flat_p.clone(), 'scale': opt.grad_scale},
                   os.path.join(out_dir, f'rank{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_bucketed_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f'rank{r}.pt') for r in (0, 1))
    assert r0['scale'] == r1['scale'] == 0.5
    want = r0['local'] + r1['local']
    assert torch.allclose(r0['summed'], want, rtol=1e-6, atol=1e-7)
    assert torch.equal(r0['summed'], r1['summed'])
    assert torch.equal(r0['params'], r1['params'])          # replicas stay identical
    assert not torch.equal(r0['local'], r1['local'])         # the shards really differed

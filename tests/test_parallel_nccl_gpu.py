"""Data-parallel path on hardware: 2 ranks, NCCL, the real libvarsep_sm100a.so.

Needs two GPUs (`gpurun --gpus 2`); skipped on a single-GPU box.  Checks:
  * the reduced gradient arena equals the SUM of the two ranks' single-GPU gradients on their own shards (Adam then
    applies 1/world: the mean), with BatchNorm statistics per replica;
  * every bucket but the last ones left during backward (completion order), and replicas are bit-identical after
    `step()`;
  * the same through `train.GraphedStep` with the all-reduces captured inside the CUDA graph, over several steps.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, transport):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from spatiotemporal_variable_separation_b200 import configs, ops, train as vs_train
    from spatiotemporal_variable_separation_b200.optim import FusedAdam
    from spatiotemporal_variable_separation_b200.parallel import GradReducer, broadcast_model
    from tests import harness
    from tests.test_host_emulated import build_filled
    cfg = configs.preset('mnist', small=True)
    cfg['name'] = 'mnist-small'
    ops.set_compute_dtype(torch.float32)
    net = build_filled(cfg, dev).train()
    broadcast_model(net)
    opt = FusedAdam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
    red = GradReducer(net, opt, overlap=True, bucket_bytes=16 << 10, early=('decoder', 't_resnet', 'Es', 'Et'),
                      transport='nccl' if transport == 'nccl' else 'peer',
                      wire_dtype=torch.bfloat16 if transport == 'peer-bf16' else torch.float32)
    assert (red.peer is not None) == (transport != 'nccl')
    cond, target = harness.inputs(cfg)
    full = torch.cat([cond, target], 1)
    shard = full[rank * 2:(rank + 1) * 2].to(dev)

    def run(reducer):
        opt.zero_grad()
        out = vs_train.step_losses(net, shard, cfg['nt_cond'], cfg['nt_pred'], cfg['offset'], False, cfg['lamb_ae'],
                                   cfg['lamb_s'], cfg['lamb_t'], cfg['lamb_pred'], False, 6, reducer)
        out['total'].backward()

    run(None)
    local = opt.flat_g.clone()
    run(red)
    early = [b['name'] for b in red.buckets if id(b) in red.done]
    red.finish()
    torch.cuda.synchronize()
    summed = opt.flat_g.clone()
    opt.step()
    after_one = opt.flat_p.clone()
    # ---- the graphed stepper with captured all-reduces, three more steps
    stepper = vs_train.GraphedStep(net, opt, cfg['nt_cond'], cfg['nt_pred'], cfg['offset'], False, cfg['lamb_ae'],
                                   cfg['lamb_s'], cfg['lamb_t'], cfg['lamb_pred'], reducer=red, graph=True)
    c, t = shard[:, :cfg['nt_cond']], shard[:, cfg['nt_cond']:]
    for tr in (6, 7, 6, 6):
        stepper(c, t, tr)
    torch.cuda.synchronize()
    stepper.close()
    torch.save({'local': local.cpu(), 'summed': summed.cpu(), 'params': after_one.cpu(), 'graphed': opt.flat_p.cpu().clone(),
                'scale': opt.grad_scale, 'early': early, 'n_buckets': len(red.buckets)}, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize('transport', ['peer', 'peer-bf16', 'nccl'])
def test_gradient_exchange_world2(tmp_path, transport):
    """'peer': the hand-written all-reduce over NVLink peer memory (csrc/peer.cu); 'nccl': bucketed ncclAllReduce."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), transport), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f'rank{r}.pt') for r in (0, 1))
    assert r0['scale'] == r1['scale'] == 0.5
    want = r0['local'] + r1['local']
    # the training path accumulates BatchNorm statistics / weight gradients with float atomics: the second run of
    # the same step reproduces the first to ~1e-6 of the largest gradient, not bit for bit
    scale = float(want.abs().max())
    if transport == 'peer-bf16':      # every rank's contribution and the sum are rounded to bf16 on the wire (2^-9 each)
        assert float((r0['summed'] - want).abs().max()) <= 2 ** -7 * scale
        assert float((r0['summed'] - want).norm() / want.norm()) <= 2 ** -8
    else:
        assert float((r0['summed'] - want).abs().max()) <= 2e-5 * scale
    assert torch.equal(r0['summed'], r1['summed'])            # the all-reduce leaves identical sums on both ranks
    assert torch.equal(r0['params'], r1['params'])            # replicas stay bit-identical after Adam
    assert torch.equal(r0['graphed'], r1['graphed'])          # ... and after graph-replayed steps
    assert not torch.equal(r0['local'], r1['local'])          # the shards really differed
    if transport == 'nccl':
        assert len(r0['early']) >= r0['n_buckets'] - 2 and any(n.startswith('decoder') for n in r0['early'])

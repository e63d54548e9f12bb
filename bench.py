"""Benchmark of the hot path: training sequences/second of the S/T-separation model.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config mnist] [--dtype bf16|fp32]
    python bench.py --impl reference ...      # the reference's CPU arithmetic (oracle port) on the host cores

One "step" = zero_grad + the full objective of var_sep/train.py:116-149 (2x Es, 2x Et, latent rollout,
1 + nt_pred + offset decoder calls, four loss terms) + backward + Adam, on one synthetic batch of the
BASELINE configuration (default: configs[1], Moving-MNIST-shaped DCGAN, batch 128 per GPU).
Prints ONE JSON line (see README / DESIGN.md section "Measurement" for every key).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# forward MACs*2 of every conv/linear per sequence, fwd+bwd (FlopCounterMode on the reference)
FLOP_PER_SEQ = {'mnist': 12.419e9, 'wave': 1.778e9, 'taxibj': 22.314e9, 'sst': 130.965e9, 'chairs': 15.635e9}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].startswith('Active') for r in self.rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm),
                'window': 'warm-up + timed steps of the device-resident run (GPU under the same load throughout)'}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, budget_s, steps=None, warmup=1, batch=None):
    """Sequences/second of the reference arithmetic (oracle port: same ATen CPU kernels, same order as
    var_sep/train.py) with all host threads.  Returns (seq_per_s, dict describing the sample)."""
    from oracle import detfill, functional, shapes, step as ostep
    from spatiotemporal_variable_separation_b200.data import synthetic_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sh = shapes.model_shapes(cfg)
    P = {part: detfill.fill_state(sh[part], part + '.') for part in ('Es', 'Et', 'decoder', 't_resnet')}
    net = functional.Net(cfg, P['Es'], P['Et'], P['decoder'], P['t_resnet']).requires_grad_(True)
    opt = ostep.Adam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
    B = batch or cfg['batch_size']
    full = synthetic_batch(cfg, batch=B, device='cpu')
    cond, target = full[:, :cfg['nt_cond']], full[:, cfg['nt_cond']:]
    t_random = cfg['nt_cond'] + 1

    def one():
        t0 = time.perf_counter()
        ostep.train_step(net, opt, cond, target, cfg, t_random)
        return time.perf_counter() - t0

    for _ in range(warmup):
        one()
    times = []
    while (steps is not None and len(times) < steps) or (steps is None and sum(times) < budget_s and len(times) < 50):
        times.append(one())
    rate = B * len(times) / sum(times)
    return rate, {'cores': cores, 'threads': torch.get_num_threads(), 'batch': B, 'steps': len(times),
                  'ms_per_step': 1e3 * sum(times) / len(times)}


def run_reference(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # bounded sample: shrink the batch so that (steps + warmup) CPU steps finish within a few minutes
    B = cfg['batch_size']
    probe_rate, _ = cpu_reference_rate(cfg, 0, steps=1, warmup=1, batch=16)
    budget = 150.0
    while B > 16 and (args.steps + args.warmup) * B / probe_rate > budget:
        B //= 2
    rate, info = cpu_reference_rate(cfg, budget, steps=args.steps, warmup=args.warmup, batch=B)
    sample = f'{info["steps"]} steps of batch {B} (full config batch {cfg["batch_size"]}) after {args.warmup} warm-up'
    line = {'metric': 'train_sequences_per_sec', 'value': rate, 'unit': 'sequences/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': info['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': workload_name(cfg), 'batch_per_step': B},
            'cpu_baseline': {'value': rate, 'unit': 'sequences/s', 'cores': info['cores'], 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': rate, 'unit': 'sequences/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def workload_name(cfg):
    return (f'{cfg["data"]} {cfg["architecture"]}/{cfg["decoder_architecture"] or cfg["architecture"]} '
            f'{"x".join(map(str, cfg["shape"]))} nt_cond {cfg["nt_cond"]} nt_pred {cfg["nt_pred"]} '
            f'batch {cfg["batch_size"]}/GPU')


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Trainer:
    """Model + fused Adam + (optionally) one captured CUDA graph per value of the host draw t_random."""

    def __init__(self, cfg, device, dtype, world, use_graph, overlap=True):
        from spatiotemporal_variable_separation_b200 import ops
        from spatiotemporal_variable_separation_b200.networks.factory import build_model
        from spatiotemporal_variable_separation_b200.optim import FusedAdam
        self.cfg, self.device, self.world, self.use_graph = cfg, device, world, use_graph
        ops.set_compute_dtype(dtype)
        torch.manual_seed(0)
        self.net = build_model(cfg, device).train()
        if world > 1:
            import torch.distributed as dist
            for p in self.net.parameters():          # identical replicas
                dist.broadcast(p.data, 0)
            for b in self.net.buffers():
                dist.broadcast(b, 0)
        self.opt = FusedAdam(self.net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
        from spatiotemporal_variable_separation_b200.parallel import GradReducer
        self.reducer = GradReducer(self.net, self.opt, overlap=overlap) if world > 1 else None
        B, T = cfg['batch_size'], cfg['nt_cond'] + cfg['nt_pred']
        self.full = torch.zeros(B, T, *cfg['shape'], device=device)        # static input buffer
        self.terms = torch.zeros(5, device=device)
        self.graphs, self.pool = {}, None

    def _step_body(self, t_random):
        from spatiotemporal_variable_separation_b200 import train as vs_train
        c = self.cfg
        self.opt.zero_grad()
        out = vs_train.step_losses(self.net, self.full, c['nt_cond'], c['nt_pred'], c['offset'], c['skipco'],
                                   c['lamb_ae'], c['lamb_s'], 0 if c['no_s'] else c['lamb_t'], c['lamb_pred'],
                                   c['architecture'] == 'encoderSST', t_random, self.reducer)
        out['total'].backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        self.terms.copy_(out['terms'].detach())

    def step(self, t_random):
        if not self.use_graph:
            self._step_body(t_random)
            return
        g = self.graphs.get(t_random)
        if g is None:
            # warm up on a side stream, then capture (one graph per value of the host draw)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._step_body(t_random)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.pool):
                self._step_body(t_random)
            if self.pool is None:
                self.pool = g.pool()
            self.graphs[t_random] = g
        g.replay()


def run_ours(args, cfg):
    from spatiotemporal_variable_separation_b200 import _lib
    from spatiotemporal_variable_separation_b200.data import synthetic_batch
    from spatiotemporal_variable_separation_b200.train import draw_t_random
    import numpy as np
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    tr = Trainer(cfg, device, dtype, world, not args.no_graph, overlap=not args.no_overlap)
    B, n_frames = cfg['batch_size'], cfg['nt_cond'] + cfg['nt_pred']

    # a pool of synthetic batches: pinned host copies (e2e) and device-resident copies (value)
    n_pool = 4
    host = [synthetic_batch(cfg, device='cpu', seed=100 * rank + i).pin_memory() for i in range(n_pool)]
    dev = [h.to(device) for h in host]
    np.random.seed(1234)          # same t_random sequence on every rank
    draws = [draw_t_random(cfg['nt_cond'], n_frames, cfg['offset']) for _ in range(2 * (args.steps + args.warmup) + 64)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # one eager step to count kernel launches per step and to populate caches
    use_graph = tr.use_graph
    tr.use_graph = False
    tr.full.copy_(dev[0])
    tr.step(draws[0])
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    tr.step(draws[0])
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0
    tr.use_graph = use_graph
    if use_graph:                  # capture every graph outside the timed regions
        for t in sorted(set(draws)):
            tr.step(t)
        torch.cuda.synchronize()

    # e2e: this step's batch travels host -> device (pinned memory, copy stream) while the previous step computes;
    # the step itself then starts with a device-to-device move into the graph's static input buffer
    copy_stream = torch.cuda.Stream()
    staging = [torch.empty_like(tr.full) for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            staging[i % 2].copy_(host[i % n_pool], non_blocking=True)
            staged[i % 2].record(copy_stream)

    def timed(n_warm, n_steps, e2e, offset):
        losses = []

        def one(i, k):
            if e2e:
                torch.cuda.current_stream().wait_event(staged[k % 2])
                tr.full.copy_(staging[k % 2], non_blocking=True)
                copy_stream.wait_stream(torch.cuda.current_stream())      # staging[k % 2] is free again after this copy
                prefetch(k + 1)
            else:
                tr.full.copy_(dev[i % n_pool], non_blocking=True)
            tr.step(draws[offset + k])
            if e2e:
                losses.append(tr.terms.cpu())        # device->host read of the step's loss terms (syncs)

        if e2e:
            prefetch(0)
        for i in range(n_warm):
            one(i, i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(n_steps):
            one(i, n_warm + i)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, losses

    with ClockSampler(local) as clocks:
        time.sleep(0.3)                       # let nvidia-smi start streaming before the load begins
        ms, _ = timed(args.warmup, args.steps, False, 0)
        if len(clocks.rows) < 3:              # very short runs: keep the GPU under the same load until 3 samples exist
            t_end = time.time() + 2.0
            while len(clocks.rows) < 3 and time.time() < t_end:
                timed(0, max(args.steps, 10), False, 0)
    ms_e2e, losses = timed(max(3, args.warmup // 2), args.steps, True, args.steps + args.warmup)
    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    assert all(torch.isfinite(l).all() for l in losses), 'non-finite loss'

    # ---- per-kernel roofline of the dominant kernel family (conv forward/dgrad launches), measured live:
    # CUDA events around every vs_conv_forward call of two extra eager steps on the launching stream
    roof = conv_roofline(tr, dev, draws, dtype)

    if rank != 0:
        _shutdown(tr, world)
        return
    peaks, src = measured_peaks()
    flop_seq = FLOP_PER_SEQ[cfg['data']]
    line = {
        'metric': 'train_sequences_per_sec', 'value': value, 'unit': 'sequences/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': workload_name(cfg), 'global_batch': world * B, 'parallelism': f'dp{world}',
                   'cuda_graph': bool(use_graph),
                   'l2': 'no flush: the per-step working set (activations + weights + Adam state, > 2 GB) is far '
                         'larger than the 126 MB L2 and each step consumes a different input batch'},
        'e2e': {'value': e2e_value, 'unit': 'sequences/s', 'h2d_bytes_per_step': host[0].numel() * 4,
                'd2h_bytes_per_step': 5 * 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches_per_step * args.steps,
        'launches_per_step': launches_per_step,
        'model_tflops': value * flop_seq / 1e12 / world,
        'tensor_frac_of_step': value * flop_seq / 1e12 / world / peaks['bf16_tflops_sustained'],
        'clocks': clocks.summary(),
        'final_loss_terms': [float(x) for x in losses[-1]],
    }
    roof['peak'] = peaks['bf16_tflops_sustained']
    roof['frac'] = roof['achieved'] / roof['peak']
    roof['peak_source'] = f'{src} bf16_tflops_sustained (kernel timed inside a long step)'
    line['roofline'] = roof
    if world == 1 and not args.no_cpu_baseline:
        rate, info = cpu_reference_rate(cfg, args.cpu_budget)
        line['cpu_baseline'] = {'value': rate, 'unit': 'sequences/s', 'cores': info['cores'], 'kind': 'port',
                                'sample': f'{info["steps"]} full-batch ({info["batch"]}) steps after 1 warm-up, '
                                          f'{info["ms_per_step"]:.0f} ms/step, {info["threads"]} threads'}
    print(json.dumps(line), flush=True)
    _shutdown(tr, world)


def _shutdown(tr, world):
    """Collective teardown on EVERY rank: drop the captured graphs (they hold NCCL work) before the group."""
    tr.graphs.clear()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def conv_roofline(tr, dev, draws, dtype):
    """Roofline of the dominant kernel, the tcgen05 tap GEMM (tc_conv_kernel and its CTA-pair variant
    tc_conv_pair_kernel: fprop and dgrad of every eligible conv layer).  CUDA events on the launching stream around each vs_conv_forward call of two extra eager steps (each pair enqueued behind a short spin kernel, so that the interval is the kernel's duration and not its launch latency); only the
    calls that the library routes to the tensor-core kernel (vs_conv_forward_path == 1) are counted.  Algorithmic
    FLOPs per launch = 2*N*P*Q*K*C*R*S (for a stride-2 transposed convolution the parity decomposition multiplies no
    structurally-zero tap, so this is also the executed count)."""
    from spatiotemporal_variable_separation_b200 import _lib
    records = []
    orig = _lib.call
    lib = _lib.load()

    def timed_call(name, *a):
        if name != 'vs_conv_forward' or lib.vs_conv_forward_path(a[0], a[1]) != 1:
            return orig(name, *a)
        g = a[0]
        flops = 2.0 * g.N * g.P * g.Q * g.K * g.C * g.R * g.S
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # an eager step is CPU bound between launches: without work ahead of it in the stream the first event fires at
        # once and the interval would include the launch latency of the kernel.  A ~20 us spin kernel keeps the stream
        # busy while the event pair and the convolution are enqueued behind it.
        torch.cuda._sleep(40000)
        e0.record()
        orig(name, *a)
        e1.record()
        records.append((e0, e1, flops))

    use_graph, tr.use_graph = tr.use_graph, False
    _lib.call = timed_call
    try:
        for i in range(2):
            tr.full.copy_(dev[i % len(dev)])
            tr.step(draws[i])
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
        tr.use_graph = use_graph
    tot_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in records)
    tot_flop = sum(f for _, _, f in records)
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get('tc_conv_kernel', {}).get('dram_bytes_per_launch')
    return {'bound': 'tensor', 'kernel': 'tc_conv_kernel + tc_conv_pair_kernel (tcgen05 tap GEMM: every fprop / dgrad launch of the eligible conv layers)',
            'achieved': tot_flop / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0, 'unit': 'TFLOP/s',
            'launches': len(records) // 2, 'avg_launch_ms': tot_ms / max(len(records), 1),
            'flop_per_launch': tot_flop / max(len(records), 1), 'traffic': traffic,
            'traffic_note': 'avg dram__bytes_read+write per tc_conv_kernel launch from the ncu pass in profiles/ (bytes)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='mnist', choices=['mnist', 'wave', 'taxibj', 'sst', 'chairs'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-overlap', action='store_true', help='all-reduce after backward instead of overlapped buckets')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    args = ap.parse_args()
    from spatiotemporal_variable_separation_b200 import configs
    cfg = configs.preset(args.config, extra=f'--batch_size {args.batch}' if args.batch else '')
    if args.impl == 'reference':
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == '__main__':
    main()

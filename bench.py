"""Benchmark of the hot path: training sequences/second of the S/T-separation model.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config mnist] [--dtype bf16|fp32]
    python bench.py --impl reference ...      # the reference's CPU arithmetic (oracle port) on the host cores
    python bench.py --impl reference-gpu ...  # the same port moved to cuda:0 (ATen / cuDNN / cuBLAS kernels of this image,
                                              # eager fp32 and under torch.autocast(bfloat16)): the Blackwell-library bar

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
def _oracle_trainer(cfg, B, device):
    """The reference's step (oracle port: torch.nn.functional calls in the order of var_sep/train.py, torch-style Adam) with
    the deterministic weights, on ``device``.  Returns a function running one full train step."""
    from oracle import detfill, functional, shapes, step as ostep
    from spatiotemporal_variable_separation_b200.data import synthetic_batch
    sh = shapes.model_shapes(cfg)
    P = {part: {k: v.to(device) for k, v in detfill.fill_state(sh[part], part + '.').items()}
         for part in ('Es', 'Et', 'decoder', 't_resnet')}
    net = functional.Net(cfg, P['Es'], P['Et'], P['decoder'], P['t_resnet']).requires_grad_(True)
    opt = ostep.Adam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
    full = synthetic_batch(cfg, batch=B, device='cpu').to(device)
    cond, target = full[:, :cfg['nt_cond']], full[:, cfg['nt_cond']:]
    t_random = cfg['nt_cond'] + 1
    return lambda: ostep.train_step(net, opt, cond, target, cfg, t_random)


def cpu_reference_rate(cfg, budget_s, steps=None, warmup=1, batch=None):
    """Sequences/second of the reference arithmetic (oracle port: same ATen CPU kernels, same order as
    var_sep/train.py) with all host threads.  Returns (seq_per_s, dict describing the sample)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = batch or cfg['batch_size']
    step = _oracle_trainer(cfg, B, 'cpu')

    def one():
        t0 = time.perf_counter()
        step()
        return time.perf_counter() - t0

    for _ in range(warmup):
        one()
    times = []
    while (steps is not None and len(times) < steps) or (steps is None and sum(times) < budget_s and len(times) < 50):
        times.append(one())
    rate = B * len(times) / sum(times)
    return rate, {'cores': cores, 'threads': torch.get_num_threads(), 'batch': B, 'steps': len(times),
                  'ms_per_step': 1e3 * sum(times) / len(times)}


def gpu_library_rate(cfg, steps=5, warmup=3):
    """The second bar of BASELINE.md section 3.6: the reference's step on ONE B200 through this image's ATen / cuDNN /
    cuBLAS kernels (the oracle port with its tensors on cuda:0, i.e. what `python -m var_sep.main --device 0` executes),
    eager fp32 (TF32 off, the torch default) and under torch.autocast(bfloat16) — the counterpart of the reference's
    --torch_amp path (main.py:159, train.py:151-155).  CUDA events, full config batch, same synthetic batch."""
    out = {}
    for mode in ('fp32', 'autocast_bf16'):
        step = _oracle_trainer(cfg, cfg['batch_size'], 'cuda:0')

        def one():
            if mode == 'fp32':
                step()
            else:
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    step()

        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {'value': cfg['batch_size'] / (ms * 1e-3), 'ms_per_step': ms}
        del step
        torch.cuda.empty_cache()
    out.update(unit='sequences/s', kind='oracle port on cuda:0 (ATen/cuDNN/cuBLAS of torch %s), eager, %d steps after %d warm-up'
               % (torch.__version__, steps, warmup))
    return out


def bench_config(cfg):
    """The ``config`` object of the JSON line — identical in every arm (the driver compares them)."""
    return {'workload': workload_name(cfg), 'batch_per_gpu': cfg['batch_size'],
            'l2': 'no flush: the per-step working set (activations + weights + Adam state, > 2 GB) is far '
                  'larger than the 126 MB L2 and each step consumes a different input batch'}


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
            'config': bench_config(cfg),
            'cpu_baseline': {'value': rate, 'unit': 'sequences/s', 'cores': info['cores'], 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': rate, 'unit': 'sequences/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def run_reference_gpu(args, cfg):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    assert torch.cuda.is_available(), '--impl reference-gpu needs a CUDA device'
    r = gpu_library_rate(cfg, steps=max(args.steps, 1), warmup=max(args.warmup, 1))
    line = {'metric': 'train_sequences_per_sec', 'value': r['autocast_bf16']['value'], 'unit': 'sequences/s', 'n_gpus': 1,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['autocast_bf16']['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 autocast', 'data': 'synthetic',
            'impl': 'reference-gpu', 'config': bench_config(cfg), 'gpu_library_baseline': r, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def workload_name(cfg):
    return (f'{cfg["data"]} {cfg["architecture"]}/{cfg["decoder_architecture"] or cfg["architecture"]} '
            f'{"x".join(map(str, cfg["shape"]))} nt_cond {cfg["nt_cond"]} nt_pred {cfg["nt_pred"]} '
            f'batch {cfg["batch_size"]}/GPU')


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_trainer(cfg, device, dtype, world, use_graph, overlap=True, bucket_mb=None, early=None, transport='peer', early_blocks=None, wire=None):
    """Model + fused Adam + the package's graphed stepper (train.GraphedStep) — what train() itself builds."""
    from spatiotemporal_variable_separation_b200 import ops, train as vs_train
    from spatiotemporal_variable_separation_b200.networks.factory import build_model
    from spatiotemporal_variable_separation_b200.optim import FusedAdam
    from spatiotemporal_variable_separation_b200.parallel import GradReducer, broadcast_model
    ops.set_compute_dtype(dtype)
    torch.manual_seed(0)
    net = build_model(cfg, device).train()
    if world > 1:
        broadcast_model(net)          # identical replicas
    opt = FusedAdam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
    kw = {'bucket_bytes': int(bucket_mb * (1 << 20))} if bucket_mb else {}
    if early is not None:
        kw['early'] = tuple(n for n in early.split(',') if n)
    if early_blocks is not None:
        kw['early_blocks'] = early_blocks
    if wire is not None:
        kw['wire_dtype'] = torch.bfloat16 if wire == 'bf16' else torch.float32
    reducer = GradReducer(net, opt, overlap=overlap, transport=transport, **kw) if world > 1 else None
    c = cfg
    return vs_train.GraphedStep(net, opt, c['nt_cond'], c['nt_pred'], c['offset'], c['skipco'], c['lamb_ae'], c['lamb_s'],
                                0 if c['no_s'] else c['lamb_t'], c['lamb_pred'], c['architecture'] == 'encoderSST',
                                reducer=reducer, graph=use_graph)


def run_ours(args, cfg):
    from spatiotemporal_variable_separation_b200 import _lib
    from spatiotemporal_variable_separation_b200.data import synthetic_batch
    from spatiotemporal_variable_separation_b200.train import DevicePrefetcher, draw_t_random
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
    use_graph = not args.no_graph
    tr = build_trainer(cfg, device, dtype, world, use_graph, overlap=not args.no_overlap, bucket_mb=args.bucket_mb, early=args.dp_early, transport=args.transport, early_blocks=args.early_blocks, wire=args.wire)
    B, n_frames, nc = cfg['batch_size'], cfg['nt_cond'] + cfg['nt_pred'], cfg['nt_cond']
    shape = (B, n_frames) + tuple(cfg['shape'])
    static_in = tr.input_buffer(shape, device)

    # a pool of synthetic batches: pinned host copies (e2e) and device-resident copies (value)
    n_pool = 4
    host = [synthetic_batch(cfg, device='cpu', seed=100 * rank + i).pin_memory() for i in range(n_pool)]
    dev = [h.to(device) for h in host]
    # what a DataLoader(pin_memory=True) yields: contiguous pinned (cond, target) pairs
    host_pairs = [(h[:, :nc].contiguous().pin_memory(), h[:, nc:].contiguous().pin_memory()) for h in host]
    np.random.seed(1234)          # same t_random sequence on every rank
    draws = [draw_t_random(cfg['nt_cond'], n_frames, cfg['offset']) for _ in range(2 * (args.steps + args.warmup) + 64)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # one eager step to count kernel launches per step and to populate caches
    tr.graph = False
    static_in.copy_(dev[0])
    tr.run(shape, draws[0])
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    tr.run(shape, draws[0])
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0
    tr.graph = use_graph
    if use_graph:                  # capture every graph outside the timed regions
        for t in sorted(set(draws)):
            tr.run(shape, t)
        torch.cuda.synchronize()

    def timed_resident(n_warm, n_steps):
        """``value``: inputs already resident in HBM (a device-to-device move into the static input buffer)."""
        for i in range(n_warm):
            static_in.copy_(dev[i % n_pool], non_blocking=True)
            tr.run(shape, draws[i])
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(n_steps):
            static_in.copy_(dev[i % n_pool], non_blocking=True)
            tr.run(shape, draws[n_warm + i])
        ev1.record()
        barrier()
        return _max_over_ranks(ev0.elapsed_time(ev1), world, device)

    def timed_e2e(n_warm, n_steps, offset):
        """``e2e``: the call a user makes — train.DevicePrefetcher over a loader of pinned HOST batches (batch k+1
        travels host->device on a copy stream while step k computes), train.GraphedStep.__call__(cond, target), and a
        device->host read of the step's five loss terms, every step."""
        loader = [host_pairs[i % n_pool] for i in range(n_warm + n_steps)]
        losses, ev0, ev1 = [], torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k, (cond, target) in enumerate(DevicePrefetcher(loader, device)):
            if k == n_warm:
                barrier()
                ev0.record()
            terms = tr(cond, target, draws[offset + k])
            losses.append(terms.cpu())            # device->host read of the step's loss terms (synchronises)
        ev1.record()
        barrier()
        return _max_over_ranks(ev0.elapsed_time(ev1), world, device), losses

    with ClockSampler(local) as clocks:
        time.sleep(0.3)                       # let nvidia-smi start streaming before the load begins
        ms = timed_resident(args.warmup, args.steps)
        if len(clocks.rows) < 3:              # very short runs: keep the GPU under the same load until 3 samples exist
            t_end = time.time() + 2.0
            while len(clocks.rows) < 3 and time.time() < t_end:
                timed_resident(0, max(args.steps, 10))
    ms_e2e, losses = timed_e2e(max(3, args.warmup // 2), args.steps, args.steps + args.warmup)
    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    assert all(torch.isfinite(l).all() for l in losses), 'non-finite loss'

    # ---- per-kernel rooflines, measured live: CUDA events around the launches of two extra eager steps
    roof, roof_hbm = kernel_rooflines(tr, shape, dev, draws, dtype)

    if rank != 0:
        _shutdown(tr, world)
        return
    peaks, src = measured_peaks()
    flop_seq = FLOP_PER_SEQ[cfg['data']]
    config = bench_config(cfg)
    line = {
        'metric': 'train_sequences_per_sec', 'value': value, 'unit': 'sequences/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': config,
        'run': {'global_batch': world * B, 'parallelism': f'dp{world}', 'cuda_graph': bool(use_graph),
                'gradient_exchange': ('none' if world == 1 else f'peer-memory all-reduce kernel (vs_peer_allreduce, wire {"bf16" if tr.reducer.peer is not None and tr.reducer.peer.wire is not None else "fp32"})'
                                      if getattr(tr.reducer, 'peer', None) is not None else 'nccl all-reduce, bucketed'),
                'stepper': 'spatiotemporal_variable_separation_b200.train.GraphedStep'},
        'e2e': {'value': e2e_value, 'unit': 'sequences/s', 'h2d_bytes_per_step': host[0].numel() * 4,
                'd2h_bytes_per_step': 5 * 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches_per_step * args.steps,
        'launches_per_step': launches_per_step,
        'model_tflops': value * flop_seq / 1e12 / world,
        'tensor_frac_of_step': value * flop_seq / 1e12 / world / peaks['bf16_tflops'],
        'clocks': clocks.summary(),
        'final_loss_terms': [float(x) for x in losses[-1]],
    }
    # the step runs unthrottled at the maximum SM clock (see `clocks`), and the kernels are timed one by one: the burst
    # figure is the honest denominator (the sustained one was measured at ~1340 MHz under a 1 kW GEMM loop)
    roof['peak'] = peaks['bf16_tflops']
    roof['frac'] = roof['achieved'] / roof['peak']
    roof['frac_of_sustained_peak'] = roof['achieved'] / peaks['bf16_tflops_sustained']
    roof['peak_source'] = f'{src} bf16_tflops (burst: kernels timed one by one at the maximum SM clock)'
    line['roofline'] = roof
    if roof_hbm is not None:
        roof_hbm['peak'] = peaks['hbm_gbs']
        roof_hbm['frac'] = roof_hbm['achieved'] / roof_hbm['peak']
        roof_hbm['peak_source'] = f'{src} hbm_gbs'
        line['roofline_hbm'] = roof_hbm
    if world == 1 and not args.no_gpu_baseline:
        try:
            line['gpu_library_baseline'] = gpu_library_rate(cfg)
        except Exception as e:                                   # e.g. out of memory next to our arenas: report, do not die
            line['gpu_library_baseline'] = {'unavailable': f'{type(e).__name__}: {e}'[:200]}
    if world == 1 and not args.no_cpu_baseline:
        rate, info = cpu_reference_rate(cfg, args.cpu_budget)
        line['cpu_baseline'] = {'value': rate, 'unit': 'sequences/s', 'cores': info['cores'], 'kind': 'port',
                                'sample': f'{info["steps"]} full-batch ({info["batch"]}) steps after 1 warm-up, '
                                          f'{info["ms_per_step"]:.0f} ms/step, {info["threads"]} threads'}
    print(json.dumps(line), flush=True)
    _shutdown(tr, world)


def run_rollout(args, cfg):
    """SURVEY section 8d 'rollout-95' / BASELINE configs[4]: eval-mode ``get_forecast(cond, 100)`` (5 conditioning frames
    -> 100 forecast frames, test/mnist/test.py:99-133), no gradient, bf16, with and without BatchNorm folded into the
    inference weights; captured as one CUDA graph.  Prints one JSON line (metric: forecast sequences / second)."""
    from spatiotemporal_variable_separation_b200 import _lib, ops
    from spatiotemporal_variable_separation_b200.data import synthetic_batch
    from spatiotemporal_variable_separation_b200.networks.factory import build_model
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    device = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    ops.set_compute_dtype(dtype)
    torch.manual_seed(0)
    net = build_model(cfg, device).eval()
    B, horizon = cfg['batch_size'], args.horizon
    cond = synthetic_batch(cfg, device=device)[:, :cfg['nt_cond']].contiguous()
    out = {}
    for fold in (False, True):
        ops.set_eval_bn_folding(fold)
        with torch.no_grad():
            for _ in range(2):
                f = net.get_forecast(cond, horizon)[0]
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            f = net.get_forecast(cond, horizon)[0]
            torch.cuda.synchronize()
            launches = _lib.launch_count() - l0
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                f = net.get_forecast(cond, horizon)[0]
            for _ in range(args.warmup):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        assert torch.isfinite(f).all()
        out['folded_bn' if fold else 'plain'] = {'value': B / (ms * 1e-3), 'ms_per_rollout': ms, 'launches': launches,
                                                  'frames_per_s': B * horizon / (ms * 1e-3), 'checksum': float(f.double().mean())}
    ops.set_eval_bn_folding(False)
    best = max(out.values(), key=lambda v: v['value'])
    line = {'metric': 'rollout_sequences_per_sec', 'value': best['value'], 'unit': 'sequences/s', 'n_gpus': 1, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': best['ms_per_rollout'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': f'{workload_name(cfg)} eval get_forecast(cond, {horizon})', 'batch_per_gpu': B},
            'variants': out, 'gpu_launches': best['launches'] * args.steps}
    if not args.no_cpu_baseline:
        from oracle import detfill, functional, shapes
        torch.set_num_threads(os.cpu_count() or 1)
        sh = shapes.model_shapes(cfg)
        P = {part: detfill.fill_state(sh[part], part + '.') for part in ('Es', 'Et', 'decoder', 't_resnet')}
        onet = functional.Net(cfg, P['Es'], P['Et'], P['decoder'], P['t_resnet'])
        onet.train = False
        c = cond[:16].cpu()
        with torch.no_grad():
            onet.get_forecast(c, 10)
            t0 = time.perf_counter()
            onet.get_forecast(c, horizon)
            dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': 16 / dt, 'unit': 'sequences/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                                'sample': f'one eval get_forecast(cond, {horizon}) of 16 sequences, {dt:.2f} s'}
    print(json.dumps(line), flush=True)


def _max_over_ranks(ms, world, device):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def _shutdown(tr, world):
    """Collective teardown on EVERY rank, then a normal return (interpreter shutdown runs: atexit hooks included).
    The captured graphs hold NCCL work, so they go first; then the process group."""
    import gc
    tr.close()
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    sys.stdout.flush()


def kernel_rooflines(tr, shape, dev, draws, dtype):
    """(1) Roofline of the dominant kernel family, the tcgen05 tap GEMM (tc_conv_kernel, its CTA-pair variant
    tc_conv_pair_kernel and the shifted-window variant tc_conv_shift_kernel: fprop and dgrad of every eligible conv layer).  CUDA events on the launching stream around each
    vs_conv_forward call of two extra eager steps (each pair enqueued behind a short spin kernel, so that the interval
    is the kernel's duration and not its launch latency); only the calls that the library routes to the tensor-core
    kernel (vs_conv_forward_path == 1) are counted.  Algorithmic FLOPs per launch = 2*N*P*Q*K*C*R*S (for a stride-2
    transposed convolution the parity decomposition multiplies no structurally-zero tap, so this is also the executed
    count).
    (2) HBM roofline of the BatchNorm family (vs_bn_act_forward / _backward_reduce / _backward_apply): the same event
    pairs; algorithmic bytes per BatchNorm layer and step = TWO passes over the layer's activation tensor (what remains
    if normalise+activation rode the consumer's operand path and the backward reduction rode the producer's epilogue),
    i.e. 2 * rows * C * elsize, against the time of ALL BatchNorm launches of the step."""
    from spatiotemporal_variable_separation_b200 import _lib
    conv_rec, bn_rec = [], []
    bn_layers = [0]
    bn_bytes = [0.0]
    orig = _lib.call
    lib = _lib.load()
    BN = ('vs_bn_act_forward', 'vs_bn_finalize_act_forward', 'vs_bn_act_backward_reduce', 'vs_bn_act_backward_apply')

    def timed_call(name, *a):
        is_conv = name == 'vs_conv_forward' and lib.vs_conv_forward_path(a[0], a[1]) == 1
        if not is_conv and name not in BN:
            return orig(name, *a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # an eager step is CPU bound between launches: without work ahead of it in the stream the first event fires at
        # once and the interval would include the launch latency of the kernel.  A ~100 us spin kernel keeps the stream
        # busy while the event pair and the kernel are enqueued behind it (four host-side enqueues take ~15-20 us;
        # an interval that starts on an idle stream would include the kernel's launch latency).
        torch.cuda._sleep(200000)
        e0.record()
        orig(name, *a)
        e1.record()
        if is_conv:
            g = a[0]
            conv_rec.append((e0, e1, 2.0 * g.N * g.P * g.Q * g.K * g.C * g.R * g.S))
        else:
            if name == 'vs_bn_act_backward_reduce':          # one per BatchNorm layer and step (dout, y, dtype, rows, C, ...)
                rows, C, dt = a[3], a[4], a[2]
                bn_layers[0] += 1
                bn_bytes[0] += 2.0 * rows * C * (4 if dt == _lib.VS_F32 else 2)
            bn_rec.append((e0, e1))

    from spatiotemporal_variable_separation_b200 import ops
    graph, tr.graph = tr.graph, False
    _lib.call = timed_call
    ops.set_wgrad_overlap(False)         # kernels are timed one by one: nothing may run next to them on another stream
    n_steps = 2
    try:
        for i in range(n_steps):
            tr.input_buffer(shape, dev[0].device).copy_(dev[i % len(dev)])
            tr.run(shape, draws[i])
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
        tr.graph = graph
        ops.set_wgrad_overlap(True)
    tot_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in conv_rec)
    tot_flop = sum(f for _, _, f in conv_rec)
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get('tc_conv_kernel', {}).get('dram_bytes_per_launch')
    roof = {'bound': 'tensor', 'kernel': 'tc_conv_kernel + tc_conv_pair_kernel + tc_conv_shift_kernel (tcgen05 tap GEMM: every fprop / dgrad launch of the eligible conv layers)',
            'achieved': tot_flop / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0, 'unit': 'TFLOP/s',
            'launches': len(conv_rec) // n_steps, 'avg_launch_ms': tot_ms / max(len(conv_rec), 1),
            'flop_per_launch': tot_flop / max(len(conv_rec), 1), 'traffic': traffic,
            'traffic_note': 'avg dram__bytes_read+write per tc_conv_kernel launch from the committed ncu --set full pass '
                            '(profiles/roofline_traffic.json; a profiler cannot run inside the timed bench)'}
    roof_hbm = None
    if bn_rec:
        bn_ms = sum(e0.elapsed_time(e1) for e0, e1 in bn_rec)
        roof_hbm = {'bound': 'hbm', 'kernel': 'bn_act_fwd_col + bn_reduce_col + bn_bwd_apply_col (BatchNorm + activation, forward and backward)',
                    'achieved': bn_bytes[0] / (bn_ms * 1e-3) / 1e9, 'unit': 'GB/s',
                    'launches': len(bn_rec) // n_steps, 'bn_layers': bn_layers[0] // n_steps,
                    'ms_per_step': bn_ms / n_steps, 'algorithmic_bytes_per_step': bn_bytes[0] / n_steps,
                    'traffic': None,
                    'note': 'algorithmic bytes = 2 tensor passes per BatchNorm layer and step (the fused ideal) over the '
                            'summed duration of every BatchNorm launch (the last decoder BatchNorm is applied on the operand '
                            'path of the fused tail kernel: its forward pass costs no launch here)'}
    return roof, roof_hbm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'reference-gpu'])
    ap.add_argument('--config', default='mnist', choices=['mnist', 'wave', 'taxibj', 'sst', 'chairs'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-overlap', action='store_true', help='all-reduce after backward instead of overlapped buckets')
    ap.add_argument('--bucket-mb', type=float, default=None, help='gradient bucket size of the encoders (parallel.GradReducer)')
    ap.add_argument('--transport', default='peer', choices=['peer', 'nccl'], help='gradient exchange: NVLink peer-memory kernel or NCCL')
    ap.add_argument('--wire', default=None, choices=['fp32', 'bf16'], help='wire format of the peer gradient exchange (default: the compute dtype)')
    ap.add_argument('--early-blocks', type=int, default=None, help='CTAs of an early (overlapped) peer exchange')
    ap.add_argument('--dp-early', default=None, help='networks whose gradient buckets may leave during backward (comma list)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    ap.add_argument('--mode', default='train', choices=['train', 'rollout'], help="'rollout': the eval-mode long-horizon forecast")
    ap.add_argument('--horizon', type=int, default=100, help='frames to forecast in --mode rollout (5 + 95)')
    args = ap.parse_args()
    from spatiotemporal_variable_separation_b200 import configs
    cfg = configs.preset(args.config, extra=f'--batch_size {args.batch}' if args.batch else '')
    if args.impl == 'reference':
        run_reference(args, cfg)
    elif args.impl == 'reference-gpu':
        run_reference_gpu(args, cfg)
    elif args.mode == 'rollout':
        run_rollout(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == '__main__':
    main()

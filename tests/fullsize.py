"""Full-size parity of the benchmarked configurations: the oracle run LIVE on the host (fp64 = truth, fp32 = the
reference's own rounding) against the CUDA path in fp32 and in bf16, on the same deterministic weights and inputs.

Everything is compared tensor by tensor with the relative L2 error  |a - b| / |b|  (not norms only): the five loss
terms, the forecasts, the latent rollout ``t_codes`` and every parameter gradient.  Used by tests/test_fullsize_gpu.py
(bounds) and scripts/fullsize_parity_report.py (the measured distributions committed under profiles/)."""
import functools

import numpy as np
import torch

from spatiotemporal_variable_separation_b200 import configs, ops
from tests import harness

LOSS_KEYS = ('ae', 's', 'pred', 't', 'total')
T_RANDOM = {'mnist': 7, 'wave': 9, 'taxibj': 6, 'chairs': 7, 'sst': 5}


def full_cfg(name, batch=None):
    cfg = configs.preset(name, extra=f'--batch_size {batch}' if batch else '')
    cfg['name'] = name + '-full'
    return cfg


@functools.lru_cache(maxsize=None)
def oracle_run(name, dtype_name, batch=None):
    """One loss + backward of the oracle at full size.  -> dict(loss [5], forecasts, t_codes, grads{name: tensor})"""
    from oracle import step
    cfg = full_cfg(name, batch)
    if dtype_name in ('autocast_bf16', 'float32_cuda'):
        # the reference's own mixed-precision path (--torch_amp: main.py:159, train.py:151-155) with bf16 as the low
        # precision type: fp32 master weights, convolutions / linears in bf16, BatchNorm and losses in fp32.  Calibrates
        # the bf16 bound: how far does the REFERENCE move from its fp64 result when it computes in bf16?
        dev = 'cuda' if torch.cuda.is_available() else 'cpu'
        net = harness.oracle_net(cfg, torch.float32)
        for P in net.P.values():
            for k, v in P.items():
                if isinstance(v, torch.Tensor):
                    P[k] = v.detach().to(dev).requires_grad_(v.requires_grad)
        cond, target = [t.to(dev) for t in harness.inputs(cfg, torch.float32)]
        # 'float32_cuda': the reference's plain fp32 step executed by cuDNN / cuBLAS instead of oneDNN (TF32 off) — a
        # second stock execution of the same arithmetic, whose distance from the CPU run shows how much of the
        # fp32-vs-fp64 gradient error is summation-order noise (activation kinks flipping) rather than a property of
        # the implementation
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.autocast(dev, dtype=torch.bfloat16, enabled=dtype_name == 'autocast_bf16'):
                out = step.step_losses(net, cond, target, cfg, T_RANDOM[name])
            out['total'].backward()
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    else:
        dtype = getattr(torch, dtype_name)
        net = harness.oracle_net(cfg, dtype)
        cond, target = harness.inputs(cfg, dtype)
        out = step.step_losses(net, cond, target, cfg, T_RANDOM[name])
        out['total'].backward()
    return dict(loss=np.array([float(out[k].detach()) for k in LOSS_KEYS]),
                forecasts=out['forecasts'].detach().float().cpu(), t_codes=out['t_codes'].detach().float().cpu(),
                grads={n: (p.grad.detach().cpu() if p.grad is not None else None) for n, p in net.parameters()})


def cuda_run(name, dtype, batch=None):
    """The same step through the package on cuda:0 (libvarsep_sm100a.so) in compute dtype ``dtype``."""
    from tests.test_host_emulated import build_filled, run_step
    cfg = full_cfg(name, batch)
    prev = ops.compute_dtype()
    ops.set_compute_dtype(dtype)
    try:
        net = build_filled(cfg, 'cuda').train()
        out = run_step(net, cfg, T_RANDOM[name], 'cuda')
        out['total'].backward()
        torch.cuda.synchronize()
        grads = {f'{part}.{k}': (p.grad.detach().cpu() if p.grad is not None else None)
                 for part in harness.PARTS for k, p in getattr(net, part).named_parameters()}
        return dict(loss=np.array([float(out[k].detach()) for k in LOSS_KEYS]),
                    forecasts=out['forecasts'].detach().cpu(), t_codes=out['t_codes'].detach().cpu(), grads=grads)
    finally:
        ops.set_compute_dtype(prev)


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def compare(run, truth, ref32=None, ref32b=None):
    """Per-quantity relative L2 error of ``run`` against ``truth`` (the fp64 oracle); with ``ref32`` (the fp32 oracle)
    also the reference's own error, tensor by tensor (the larger of two stock executions when ``ref32b`` is given).  Gradients that are mathematically zero (conv biases feeding a
    train-mode BatchNorm, parameters the step never touches) are reported separately as |g| / max|g|."""
    rep = {'loss': float(np.abs(run['loss'] - truth['loss']).max() / np.abs(truth['loss']).max()),
           'loss_terms': [float(x) for x in np.abs(run['loss'] - truth['loss']) / np.maximum(np.abs(truth['loss']), 1e-30)],
           'forecasts': rel_l2(run['forecasts'], truth['forecasts']),
           't_codes': rel_l2(run['t_codes'], truth['t_codes']), 'grads': {}, 'zero_grads': {}}
    gmax = max(float(g.double().norm()) for g in truth['grads'].values() if g is not None)
    for n, g64 in truth['grads'].items():
        ours = run['grads'].get(n)
        if g64 is None:
            assert ours is None or float(ours.abs().max()) == 0.0, n
            continue
        assert ours is not None, n
        if float(g64.double().norm()) < 1e-6 * gmax:
            rep['zero_grads'][n] = float(ours.double().norm()) / gmax
            continue
        e = rel_l2(ours, g64)
        if ref32 is None:
            rep['grads'][n] = (e, None)
        else:
            er = rel_l2(ref32['grads'][n], g64)
            if ref32b is not None:
                er = max(er, rel_l2(ref32b['grads'][n], g64))
            rep['grads'][n] = (e, er)
    return rep


def summary(rep):
    e = np.array([v[0] for v in rep['grads'].values()])
    out = {'loss': rep['loss'], 'forecasts': rep['forecasts'], 't_codes': rep['t_codes'], 'grad_tensors': len(e),
           'grad_median': float(np.median(e)), 'grad_p90': float(np.percentile(e, 90)), 'grad_max': float(e.max()),
           'grad_worst': max(rep['grads'], key=lambda n: rep['grads'][n][0]),
           'zero_grad_max': max(rep['zero_grads'].values()) if rep['zero_grads'] else 0.0}
    r = [v[1] for v in rep['grads'].values() if v[1] is not None]
    if r:
        r = np.array(r)
        out.update(ref32_median=float(np.median(r)), ref32_max=float(r.max()),
                   worst_ratio_to_bound=float(max(v[0] / max(1e-4, 2 * v[1]) for v in rep['grads'].values())))
    return out

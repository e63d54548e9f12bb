"""Forward AND backward tensor trace of one fp32 training step on the GPU against the same host code over the emulated ABI
(tests/emu.py) on the CPU: prints, operator by operator, the relative L2 difference of every output and every gradient.
    python scripts/trace_compare.py "<extra flags>" <t_random>"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatiotemporal_variable_separation_b200 import configs, ops
from tests import emu, fullsize, harness
from tests.test_host_emulated import build_filled, run_step

extra, t_random = sys.argv[1], int(sys.argv[2])
cfg = configs.preset('mnist', small=True, extra=extra); cfg['name'] = 'mnist-small-mul'
ops.set_compute_dtype(torch.float32)


def trace(device):
    rec = []
    orig = {n: getattr(ops, n) for n in ('conv_block', 'mul_bcast', 'latent_rollout', 'concat_channels', 'decoder_tail')}

    def wrap(name):
        def f(*a, **k):
            out = orig[name](*a, **k)
            outs = out if isinstance(out, tuple) else (out,)
            for i, o in enumerate(outs):
                if isinstance(o, torch.Tensor) and o.is_floating_point():
                    tag = f'{len(rec):03d} {name}[{i}] {tuple(o.shape)}'
                    rec.append([tag, o.detach().float().cpu().clone(), None])
                    if o.requires_grad:
                        slot = rec[-1]
                        o.register_hook(lambda g, slot=slot: slot.__setitem__(2, g.detach().float().cpu().clone()))
            return out
        return f
    for n in orig:
        setattr(ops, n, wrap(n))
    try:
        net = build_filled(cfg, device).train()
        out = run_step(net, cfg, t_random, device)
        out['total'].backward()
        if device == 'cuda':
            torch.cuda.synchronize()
    finally:
        for n, f in orig.items():
            setattr(ops, n, f)
    return rec


gpu = trace('cuda')
with emu.install():
    cpu = trace('cpu')
assert len(gpu) == len(cpu), (len(gpu), len(cpu))
for (tag, a, ga), (_, b, gb) in zip(gpu, cpu):
    e = fullsize.rel_l2(a, b)
    eg = fullsize.rel_l2(ga, gb) if ga is not None and gb is not None else float('nan')
    flag = '  <<<<' if e > 2e-6 or eg > 2e-5 else ''
    print(f'{tag:50s} out {e:.2e}   grad {eg:.2e}{flag}')

"""How ill-conditioned is the gradient of each small parity configuration?  The fp64 oracle is evaluated twice: on the
golden inputs and on inputs whose every weight is perturbed by a relative 1e-7 (one fp32 ulp: the size of the difference
between any two fp32 evaluation orders).  A smooth objective moves its gradient by ~1e-7; a configuration in which some
LeakyReLU / ReLU / max-pool unit sits within 1e-7 of its kink moves it by orders of magnitude more — and then NO fp32
implementation (the reference's included) can be expected within 1e-4 of the fp64 gradient on it.
Writes profiles/r02_kink_sensitivity.json."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import step
from tests import fullsize, harness

out = {}
for name in harness.golden_names():
    g = harness.load_golden(name)
    cfg = g['cfg']
    t_random = harness.t_random_sequence(cfg, int(g['np_seed']), 1)[0]
    cond, target = harness.inputs(cfg, torch.float64)

    def grads(seed, eps=0.0):
        net = harness.oracle_net(cfg, torch.float64)
        if seed is not None:
            gen = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                for _, p in net.parameters():
                    p.mul_(1 + eps * (2 * torch.rand(p.shape, generator=gen, dtype=torch.float64) - 1))
        o = step.step_losses(net, cond, target, cfg, t_random)
        o['total'].backward()
        return {n: p.grad.detach() for n, p in net.parameters() if p.grad is not None}

    base = grads(None)
    gmax = max(float(v.norm()) for v in base.values())
    out[name] = {'t_random': t_random}
    for eps in (1e-7, 1e-6):
        worst = []
        for seed in range(1, 7):
            pert = grads(seed, eps)
            errs = [(fullsize.rel_l2(pert[n], base[n]), n) for n in base if float(base[n].norm()) > 1e-6 * gmax]
            worst.append(max(errs))
        out[name][f'max_rel_l2_gradient_change_for_{eps:g}_weight_perturbation_6_seeds'] = [float('%.2e' % w[0]) for w in worst]
        out[name][f'worst_tensor_{eps:g}'] = max(worst)[1]
    print(name, out[name], flush=True)
json.dump({'what': __doc__, 'configs': out}, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r02_kink_sensitivity.json'), 'w'), indent=1)

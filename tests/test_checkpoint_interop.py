"""checkpoints travel both ways between this package and the reference, and the optimizer state
(absent upstream) is saved in torch.optim.Adam's own format.  CPU only: the C ABI is the emulator (tests/emu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from spatiotemporal_variable_separation_b200 import ops
from spatiotemporal_variable_separation_b200.optim import FusedAdam, MultiStepLR
from spatiotemporal_variable_separation_b200.utils import helper
from tests import emu, harness
from tests.test_host_emulated import build_filled, run_step

REF = '/root/reference'


def _grads_step(net, cfg, opt, t_random):
    opt.zero_grad()
    run_step(net, cfg, t_random)['total'].backward()
    opt.step()


def test_fused_adam_resumes_bit_exactly_from_its_state_dict(tmp_path):
    cfg = harness.load_golden('mnist-small')['cfg']
    ops.set_compute_dtype(torch.float32)
    with emu.install():
        a = build_filled(cfg).train()
        opt_a = FusedAdam(a.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
        sched_a = MultiStepLR(opt_a, [1], gamma=0.5)
        _grads_step(a, cfg, opt_a, 6)
        sched_a.step()
        helper.save(str(tmp_path), a)
        helper.save_training_state(str(tmp_path), opt_a, sched_a, epoch=1)
        _grads_step(a, cfg, opt_a, 7)
        # resume in a fresh model / optimizer
        b = build_filled(cfg).train()
        helper.load(str(tmp_path), b)
        opt_b = FusedAdam(b.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
        sched_b = MultiStepLR(opt_b, [1], gamma=0.5)
        assert helper.load_training_state(str(tmp_path), opt_b, sched_b) == 1
        assert opt_b.lr == opt_a.lr == cfg['lr'] * 0.5 and int(opt_b.step_dev) == 1
        _grads_step(b, cfg, opt_b, 7)
        for (n, pa), (_, pb) in zip(a.state_dict().items(), b.state_dict().items()):
            assert torch.equal(pa, pb), n
        assert torch.equal(opt_a.exp_avg, opt_b.exp_avg) and torch.equal(opt_a.exp_avg_sq, opt_b.exp_avg_sq)


def test_state_dict_is_interchangeable_with_torch_adam():
    """FusedAdam -> torch.optim.Adam -> FusedAdam: same keys / shapes, and one more step from the same gradients
    lands on the same parameters (1e-6: fused kernel vs foreach arithmetic)."""
    cfg = harness.load_golden('wave-small')['cfg']
    ops.set_compute_dtype(torch.float32)
    with emu.install():
        net = build_filled(cfg).train()
        opt = FusedAdam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
        _grads_step(net, cfg, opt, 9)
        sd = opt.state_dict()
        twins = [torch.nn.Parameter(p.detach().clone()) for p in opt.params]
        ref = torch.optim.Adam(twins, lr=cfg['lr'], betas=(cfg['beta1'], cfg['beta2']))
        ref.load_state_dict(sd)                                  # torch validates group sizes and state layout
        # same gradients for the next step on both sides
        opt.zero_grad()
        run_step(net, cfg, 8)['total'].backward()
        for t, p in zip(twins, opt.params):
            t.grad = opt.grad_of(p).detach().clone()
        opt.step()
        ref.step()
        for t, p in zip(twins, opt.params):
            assert float((t.detach() - p.detach()).abs().max()) <= 1e-6 * max(float(p.detach().abs().max()), 1e-3)
        # and back: torch's state loads into a fresh FusedAdam
        net2 = build_filled(cfg).train()
        opt2 = FusedAdam(net2.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
        opt2.load_state_dict(ref.state_dict())
        assert int(opt2.step_dev) == 2
        assert float((opt2.exp_avg - opt.exp_avg).abs().max()) <= 1e-6 * float(opt.exp_avg.abs().max())


def test_multistep_lr_follows_torch():
    w = torch.nn.Parameter(torch.zeros(3))
    ref_opt = torch.optim.Adam([w], lr=4e-4)
    ref = torch.optim.lr_scheduler.MultiStepLR(ref_opt, milestones=[2, 5, 6], gamma=0.3)

    class Holder:
        param_groups = [{'lr': 4e-4}]

    ours = MultiStepLR(Holder, [2, 5, 6], gamma=0.3)
    for _ in range(9):
        ref_opt.step()
        ref.step()
        ours.step()
        np.testing.assert_allclose(ours.get_last_lr(), ref.get_last_lr(), rtol=1e-12)


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference tree only exists in the build container')
@pytest.mark.parametrize('name', ['mnist-small', 'sst-small', 'chairs-small'])
def test_reference_pickles_load_and_our_files_load_into_the_reference(name, tmp_path):
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from gen_golden import build_reference
    from var_sep.utils import helper as ref_helper
    cfg = harness.load_golden(name)['cfg']
    ref_net = build_reference(cfg)
    with torch.no_grad():                                       # make it differ from a freshly filled model
        for p in ref_net.parameters():
            p.mul_(1.25)
    d1, d2 = tmp_path / 'ref', tmp_path / 'ours'
    d1.mkdir(); d2.mkdir()
    ref_helper.save(str(d1), ref_net)                           # whole pickled modules, helper.py:22-33
    ours = build_filled(cfg)
    helper.load(str(d1), ours)
    for part in harness.PARTS:
        for (k, v), (k2, v2) in zip(getattr(ref_net, part).state_dict().items(), getattr(ours, part).state_dict().items()):
            assert k == k2 and torch.equal(v, v2), (part, k)
    helper.save_state_dicts(str(d2), ours, epoch_number=3)     # plain state_dicts under distinct names
    fresh = build_reference(cfg)
    for attr, fname in (('Et', 'ov_Et'), ('Es', 'ov_Es'), ('decoder', 'decoder'), ('t_resnet', 't_resnet')):
        getattr(fresh, attr).load_state_dict(torch.load(str(d2 / f'{fname}_3.state_dict.pt')))
        for (k, v), (_, v2) in zip(getattr(fresh, attr).state_dict().items(), getattr(ref_net, attr).state_dict().items()):
            assert torch.equal(v, v2), (attr, k)



def test_save_writes_the_reference_format_and_the_reference_loader_procedure_works(tmp_path):
    """helper.save pickles whole modules under the reference's file names; test/utils.py:8-16 of the reference does
    ``torch.load(path).to(device)`` on them and must obtain working modules (of this package's classes) with the
    trained values, free of optimizer-arena views and per-parameter caches."""
    cfg = harness.load_golden('mnist-small')['cfg']
    ops.set_compute_dtype(torch.float32)
    with emu.install():
        net = build_filled(cfg).train()
        opt = FusedAdam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
        _grads_step(net, cfg, opt, 6)                       # parameters are now views of the flat arena, caches exist
        helper.save(str(tmp_path), net, epoch_number=2)
        cond, _ = harness.inputs(cfg)
        want = net.eval().get_forecast(cond, 3)[0]
        # the reference's loader, verbatim in spirit (torch >= 2.6 needs weights_only=False for module pickles)
        from spatiotemporal_variable_separation_b200.networks.model import SeparableNetwork
        parts = {a: torch.load(str(tmp_path / f'{f}_2.pt'), weights_only=False).to('cpu')
                 for a, f in (('Et', 'ov_Et'), ('Es', 'ov_Es'), ('decoder', 'decoder'), ('t_resnet', 't_resnet'))}
        for a, m in parts.items():
            assert type(m) is type(getattr(net, a))
            for p in m.parameters():
                assert p.untyped_storage().nbytes() == p.numel() * 4                 # compact, not the arena
                assert not any(k.startswith('_vs_') for k in p.__dict__)
        twin = SeparableNetwork(parts['Es'], parts['Et'], parts['t_resnet'], parts['decoder'], cfg['nt_cond'], cfg['skipco'])
        got = twin.eval().get_forecast(cond, 3)[0]
        assert torch.equal(got, want)
    assert os.path.getsize(tmp_path / 'ov_Es_2.pt') < 2 * sum(p.numel() * 4 for p in net.Es.parameters()) + (1 << 16)


def test_load_refuses_foreign_pickles_unless_asked(tmp_path):
    import pickle

    cfg = harness.load_golden('mnist-small')['cfg']
    net = build_filled(cfg)
    helper.save(str(tmp_path), net)
    helper.load(str(tmp_path), build_filled(cfg))               # our own module pickles pass the restricted unpickler

    class Evil:
        def __reduce__(self):
            return (os.path.join, ('pwned', 'x'))               # any callable outside the allow-list

    torch.save(Evil(), str(tmp_path / 'ov_Et.pt'))
    with pytest.raises(pickle.UnpicklingError, match='allow_pickled_modules'):
        helper.load(str(tmp_path), build_filled(cfg))
    with pytest.raises(TypeError):                              # opt-in: unpickled, then rejected for not being a module
        helper.load(str(tmp_path), build_filled(cfg), allow_pickled_modules=True)


def test_params_json_round_trip_through_load_model(tmp_path):
    """main.py:105-106 + test/utils.py:8-16: an experiment directory written by this package (params.json + the four
    checkpoint files) is rebuilt by load_model without any architecture flag on the command line."""
    from spatiotemporal_variable_separation_b200 import configs
    from spatiotemporal_variable_separation_b200.options import parser
    from spatiotemporal_variable_separation_b200.test.utils import load_model
    import shlex
    argv = ['--xp_dir', str(tmp_path), '--data_dir', '.'] + shlex.split(configs.README_FLAGS['mnist']) + \
        shlex.split(configs.SMALL_FLAGS['mnist'])
    args = parser.parse_args(argv)
    cfg = configs.preset('mnist', small=True)
    net = build_filled(cfg)
    helper.save_params(str(tmp_path), args)
    helper.save(str(tmp_path), net)
    twin = load_model({'xp_dir': str(tmp_path), 'device': 'cpu'})
    assert not twin.training
    for (k, a), (_, b) in zip(net.state_dict().items(), twin.state_dict().items()):
        assert torch.equal(a, b), k

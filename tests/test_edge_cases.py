"""Edge cases of the public surface: a one-frame forecast (no stepper call), batch size 1, eval-mode decoding,
content swap through ``init_s_code`` (test/mnist/test.py:131), skip connections in eval mode, and the
reference's assertions / error types.  Run on CPU over the ABI specification and (marked gpu) on the device."""
import numpy as np
import pytest
import torch

from spatiotemporal_variable_separation_b200 import configs, ops
from spatiotemporal_variable_separation_b200.networks import factory
from tests import emu, harness
from tests.test_host_emulated import build_filled


def _check(device):
    ops.set_compute_dtype(torch.float32)
    cfg = harness.load_golden('mnist-small')['cfg']
    onet = harness.oracle_net(cfg)
    onet.train = False
    net = build_filled(cfg, device).eval()
    cond, _ = harness.inputs(cfg)
    with torch.no_grad():
        # one-frame forecast: Es, Et and a single decoder call, no stepper
        f, t, s, res = net.get_forecast(cond.to(device), 1)
        fo, to, so, reso = onet.get_forecast(cond, 1)
        assert list(f.shape) == list(fo.shape) and res == [] and reso == []
        assert float((f.cpu() - fo).abs().max()) < 2e-5
        # batch of one sequence, long horizon, content swap with another sequence's S code
        f1, t1, s1, _ = net.get_forecast(cond[:1].to(device), 12)
        fo1, to1, so1, _ = onet.get_forecast(cond[:1], 12)
        assert float((f1.cpu() - fo1).abs().max()) < 5e-5 and float((t1.cpu() - to1).abs().max()) < 5e-5
        fs, _, _, _ = net.get_forecast(cond[1:2].to(device), 5, init_s_code=s1)
        fos, _, _, _ = onet.get_forecast(cond[1:2], 5, init_s_code=so1)
        assert float((fs.cpu() - fos).abs().max()) < 5e-5
        # given initial T code
        ft, _, _, _ = net.get_forecast(cond[:2].to(device), 3, init_t_code=t[:2, 0])
        fot, _, _, _ = onet.get_forecast(cond[:2], 3, init_t_code=to[:2, 0])
        assert float((ft.cpu() - fot).abs().max()) < 5e-5
    # eval mode must not touch the BatchNorm buffers
    assert int(net.decoder.first_upconv[1].num_batches_tracked) == 0
    # skip connections, eval mode, through the public tuple form of init_s_code
    cfg2 = harness.load_golden('mnist-small-skipco')['cfg']
    onet2 = harness.oracle_net(cfg2)
    onet2.train = False
    net2 = build_filled(cfg2, device).eval()
    c2, _ = harness.inputs(cfg2)
    with torch.no_grad():
        s_pub = net2.Es(c2.to(device), return_skip=True)
        f2, _, _, _ = net2.get_forecast(c2.to(device), 4, init_s_code=s_pub)
        fo2, _, _, _ = onet2.get_forecast(c2, 4)
    assert float((f2.cpu() - fo2).abs().max()) < 5e-5


def test_edge_cases_emulated():
    with emu.install():
        _check('cpu')


@pytest.mark.gpu
def test_edge_cases_gpu():
    _check('cuda')


def test_factory_assertions_match_reference():
    with pytest.raises(AssertionError):            # dcgan needs 64x64 (factory.py:29)
        factory.get_encoder('dcgan', [1, 32, 32], 8, 8, 3, 5, 'normal', 0.02)
    with pytest.raises(AssertionError):            # mul mixing needs equal code sizes (factory.py:52)
        factory.get_decoder('dcgan', [1, 64, 64], 4, 8, 'sigmoid', 8, 3, 'mul', False, 'normal', 0.02)
    with pytest.raises(AssertionError):            # skipco only for conv decoders (factory.py:49)
        factory.get_decoder('mlp', [1, 64, 64], 4, 4, 'sigmoid', 8, 3, 'concat', True, 'normal', 0.02)
    with pytest.raises(AssertionError):            # decoderSST is concat-only (factory.py:68)
        factory.get_decoder('decoderSST', [1, 64, 64], 4, 4, None, 8, 3, 'mul', False, 'normal', 0.02)
    with pytest.raises(NotImplementedError):       # unknown init (utils.py:100)
        factory.get_resnet(4, 1, 8, 'bogus', 1.0)
    from spatiotemporal_variable_separation_b200.networks.utils import activation_factory
    with pytest.raises(ValueError):                # unknown activation (utils.py:72)
        activation_factory('gelu')


def test_options_contract():
    """Flag names / defaults / the --gain_res prefix abbreviation of README.md:86."""
    from spatiotemporal_variable_separation_b200.options import parser
    a = parser.parse_args(['--xp_dir', 'x', '--data_dir', 'y'])
    assert (a.nt_cond, a.nt_pred, a.code_size_s, a.code_size_t, a.mixing, a.architecture) == (5, 10, 128, 20, 'concat', 'dcgan')
    assert (a.lamb_ae, a.lamb_s, a.lamb_t, a.lamb_pred, a.lr, a.beta1, a.beta2) == (10, 45, 0.001, 45, 4e-4, 0.9, 0.99)
    assert (a.gain_resnet, a.init_resnet, a.gain_encoder, a.init_encoder, a.offset, a.n_blocks) == (1.41, 'orthogonal', 0.02, 'normal', 5, 1)
    sst = configs.preset('sst')
    assert sst['gain_resnet'] == 0.71 and sst['skipco'] and sst['n_blocks'] == 2 and sst['offset'] == 0
    with pytest.raises(SystemExit):
        parser.parse_args(['--xp_dir', 'x', '--data_dir', 'y', '--torch_amp', '--apex_amp'])


def test_init_net_statistics():
    """Orthogonal init with gain g gives W W^T = g^2 I; BatchNorm gamma ~ N(1, gain)."""
    torch.manual_seed(0)
    r = factory.get_resnet(20, 1, 512, 'orthogonal', 0.71)
    w = r.blocks[0].mlp.module[2][1].weight            # [20, 512]
    assert torch.allclose(w @ w.t(), 0.71 ** 2 * torch.eye(20), atol=1e-4)
    e = factory.get_encoder('dcgan', [1, 64, 64], 16, 8, 3, 5, 'normal', 0.02)
    assert abs(float(e.conv[1][1].weight.mean()) - 1.0) < 0.02 and float(e.conv[1][0].bias.abs().max()) == 0.0


def test_split_first_group_backward_fills_one_buffer():
    """ops.split_first_group: both gradients land in ONE buffer; a missing gradient becomes zeros (pure torch plumbing)."""
    from spatiotemporal_variable_separation_b200 import ops
    x = torch.randn(4, 3, 2, 5, requires_grad=True)
    first, rest = ops.split_first_group(x)
    assert torch.equal(first, x[0]) and torch.equal(rest, x[1:])
    w1, w2 = torch.randn_like(first), torch.randn_like(rest)
    ((first * w1).sum() + (rest.transpose(0, 1) * w2.transpose(0, 1)).sum()).backward()
    assert torch.equal(x.grad[0], w1) and torch.equal(x.grad[1:], w2)
    y = torch.randn(3, 2, requires_grad=True)
    first, rest = ops.split_first_group(y)
    rest.sum().backward()
    assert torch.equal(y.grad, torch.cat([torch.zeros(1, 2), torch.ones(2, 2)]))


def test_eval_mode_batchnorm_still_trains_its_affine_parameters():
    """Fine-tuning with frozen statistics (``eval()`` with grad enabled): gamma / beta receive the gradients
    torch's batch_norm produces in eval mode, and the conv bias the plain column sum."""
    import torch.nn as nn
    from spatiotemporal_variable_separation_b200 import ops
    from tests import emu
    torch.manual_seed(0)
    conv, bn = nn.Conv2d(3, 5, 3, 1, 1), nn.BatchNorm2d(5)
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5); bn.weight.normal_(1, 0.2); bn.bias.normal_(0, 0.2)
    bn.eval()
    x = torch.randn(2, 3, 6, 6)
    ref = torch.nn.functional.leaky_relu(bn(conv(x)), 0.2)
    go = torch.randn_like(ref)
    ref.backward(go)
    want = [p.grad.clone() for p in (conv.weight, conv.bias, bn.weight, bn.bias)]
    for p in (conv.weight, conv.bias, bn.weight, bn.bias):
        p.grad = None
    ops.set_compute_dtype(torch.float32)
    with emu.install():
        out = ops.conv_block(x.permute(0, 2, 3, 1).contiguous(), conv, bn, 'leaky_relu')
        out.backward(go.permute(0, 2, 3, 1).contiguous())
    assert torch.allclose(out.permute(0, 3, 1, 2), ref, atol=1e-5)
    for p, w in zip((conv.weight, conv.bias, bn.weight, bn.bias), want):
        assert p.grad is not None and torch.allclose(p.grad, w, rtol=1e-4, atol=1e-5), (p.shape, p.grad, w)

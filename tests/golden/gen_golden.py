"""Generate the golden fixtures by running the UNMODIFIED reference on the CPU.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/gen_golden.py            # writes tests/golden/<config>.npz

For each parity configuration it builds the reference networks through the
reference's own ``factory`` (main.py:120-140), overwrites every state-dict
entry with the name-keyed deterministic values of ``oracle/detfill.py``, feeds
the deterministic synthetic sequences, and records

* ``shapes``          – state-dict keys / shapes of the four networks (json),
* ``loss32/loss64``   – ae, s, pred, t, total from the reference's own
                        ``ae_loss`` / ``zero_order_loss`` / ``get_forecast`` called
                        in the order of train.py:120-149, fp32 and fp64,
* ``forecast_sub``, ``t_codes`` – sub-sampled forecasts and the full latent rollout,
* ``grad32/grad64``   – per-parameter [L2 norm, probe dot] of the gradients,
* ``after2``          – per state-dict entry [norm, probe dot] after two steps of
                        the reference's real ``train()`` loop with torch.optim.Adam.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from oracle import detfill                                              # noqa: E402
from spatiotemporal_variable_separation_b200 import configs             # noqa: E402
from tests.summ import summarize, subsample                             # noqa: E402

from var_sep.networks.factory import get_encoder, get_decoder, get_resnet   # noqa: E402
from var_sep.networks.model import SeparableNetwork                          # noqa: E402
from var_sep.networks.utils import ConstantS                                 # noqa: E402
from var_sep import train as ref_train                                       # noqa: E402
import torch.nn.functional as F                                              # noqa: E402

NP_SEED = 4242


def build_reference(cfg):
    """main.py:120-140 with the flags of ``cfg``."""
    c = cfg
    if not c['no_s']:
        Es = get_encoder(c['architecture'], c['shape'], c['code_size_s'], c['enc_hidden_size'], c['enc_n_layers'],
                         c['nt_cond'], c['init_encoder'], c['gain_encoder'])
    else:
        Es = ConstantS(return_value=1, code_size=c['code_size_s'])
    Et = get_encoder(c['architecture'], c['shape'], c['code_size_t'], c['enc_hidden_size'], c['enc_n_layers'],
                     c['nt_cond'], c['init_encoder'], c['gain_encoder'])
    dec = get_decoder(c['decoder_architecture'] or c['architecture'], c['shape'], c['code_size_t'],
                      c['code_size_s'], c['last_activation'], c['dec_hidden_size'], c['dec_n_layers'],
                      c['mixing'], c['skipco'], c['init_encoder'], c['gain_encoder'])
    res = get_resnet(c['code_size_t'], c['n_blocks'], c['res_hidden_size'], c['init_resnet'], c['gain_resnet'],
                     c['architecture'] == 'encoderSST')
    net = SeparableNetwork(Es, Et, res, dec, c['nt_cond'], c['skipco'])
    for part in ('Es', 'Et', 'decoder', 't_resnet'):
        detfill.fill_module_(getattr(net, part), part + '.')
    return net


def inputs(cfg, dtype=torch.float32):
    kind = configs.input_kind(cfg)
    x = detfill.frames('x:' + cfg['name'], cfg['batch_size'], cfg['nt_cond'] + cfg['nt_pred'], cfg['shape'],
                       kind=kind).to(dtype)
    return x[:, :cfg['nt_cond']].contiguous(), x[:, cfg['nt_cond']:].contiguous()


def one_step(cfg, dtype):
    """Loss terms and gradients through the reference's own functions (train.py:120-149)."""
    net = build_reference(cfg).to(dtype).train()
    cond, target = inputs(cfg, dtype)
    np.random.seed(NP_SEED)
    ae, s_recent, s_old = ref_train.ae_loss(cond, target, net, cfg['nt_cond'], cfg['offset'], cfg['skipco'])
    s_inv = ref_train.zero_order_loss(s_old, s_recent, cfg['skipco'])
    full = torch.cat([cond, target], dim=1)
    forecasts, t_codes, _, _ = net.get_forecast(cond, cfg['nt_pred'] + cfg['offset'], init_s_code=s_old)
    f_off = cfg['nt_cond'] if cfg['offset'] == 0 else 0
    pred = F.mse_loss(forecasts, full[:, f_off:])
    if cfg['architecture'] == 'encoderSST':
        t_reg = 0.5 * (t_codes[:, 0].pow(2).view(full.shape[0], -1)).mean()
    else:
        t_reg = 0.5 * torch.sum(t_codes[:, 0].pow(2), dim=1).mean()
    lamb_t = 0 if cfg['no_s'] else cfg['lamb_t']
    total = cfg['lamb_ae'] * ae + cfg['lamb_s'] * s_inv + cfg['lamb_pred'] * pred + lamb_t * t_reg
    total.backward()
    losses = np.array([float(v) for v in (ae, s_inv, pred, t_reg, total)])
    grads = {}
    for part in ('Es', 'Et', 'decoder', 't_resnet'):
        for k, p in getattr(net, part).named_parameters():
            name = f'{part}.{k}'
            grads[name] = summarize(name, p.grad) if p.grad is not None else np.array([np.nan, np.nan])
    return net, losses, forecasts, t_codes, grads


def two_real_steps(cfg):
    """The reference's real train() loop, two optimizer steps on one batch."""
    net = build_reference(cfg)
    cond, target = inputs(cfg)
    opt = torch.optim.Adam(net.parameters(), lr=cfg['lr'], betas=(cfg['beta1'], cfg['beta2']))
    np.random.seed(NP_SEED)
    with tempfile.TemporaryDirectory() as xp:
        ref_train.train(xp, [(cond, target)], torch.device('cpu'), net, opt, None, False, False, 2,
                        cfg['lamb_ae'], cfg['lamb_s'], cfg['lamb_t'], cfg['lamb_pred'], cfg['offset'],
                        cfg['nt_cond'], cfg['nt_pred'], cfg['no_s'], cfg['skipco'], None,
                        cfg['architecture'] == 'encoderSST')
    out = {}
    for part in ('Es', 'Et', 'decoder', 't_resnet'):
        for k, v in getattr(net, part).state_dict().items():
            out[f'{part}.{k}'] = summarize(f'{part}.{k}', v)
    return out


def generate(name, small=True, extra='', tag=None, subdir=None):
    cfg = configs.preset(name, small=small, extra=extra)
    if tag:
        cfg['name'] = tag
    torch.manual_seed(0)
    net, loss32, forecasts, t_codes, grad32 = one_step(cfg, torch.float32)
    _, loss64, f64, _, grad64 = one_step(cfg, torch.float64)
    shapes = {part: {k: list(v.shape) for k, v in getattr(net, part).state_dict().items()}
              for part in ('Es', 'Et', 'decoder', 't_resnet')}
    after2 = two_real_steps(cfg)
    gnames = sorted(grad32)
    anames = sorted(after2)
    out = dict(
        cfg=json.dumps({k: v for k, v in cfg.items()}), shapes=json.dumps(shapes), np_seed=NP_SEED,
        loss32=loss32, loss64=loss64,
        forecast_shape=np.array(forecasts.shape), forecast_sub=subsample(forecasts).astype(np.float32),
        forecast_sub64=subsample(f64), forecast_sum=np.array([float(forecasts.double().sum()),
                                                               float(forecasts.double().abs().sum())]),
        t_codes=t_codes.detach().numpy(),
        grad_names=np.array(gnames), grad32=np.stack([grad32[k] for k in gnames]),
        grad64=np.stack([grad64[k] for k in gnames]),
        after2_names=np.array(anames), after2=np.stack([after2[k] for k in anames]),
    )
    out_dir = os.path.join(HERE, subdir) if subdir else HERE
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, cfg['name'] + '.npz')
    np.savez_compressed(path, **out)
    print(f'{cfg["name"]}: losses {loss32}  ->  {path} ({os.path.getsize(path) / 1024:.0f} KiB)')


# flag combinations beyond the BASELINE configurations (host-logic / oracle coverage on the CPU: tests/golden/extra/)
EXTRA = [
    ('mnist', '--architecture vgg --nt_pred 3', 'mnist-small-vgg64'),                     # VGG on 64x64 frames (vgg32=False)
    ('wave', '--mixing concat --code_size_s 20 --n_blocks 1', 'wave-small-concat'),       # MLP encoder/decoder, concat mixing
    ('chairs', '--architecture dcgan --n_blocks 3 --offset 0 --mixing mul --code_size_s 10', 'chairs-small-dcgan-mul'),
    ('taxibj', '--skipco --nt_pred 2', 'taxibj-small-skipco'),                             # VGG skip connections at 32x32
]


if __name__ == '__main__':
    torch.set_num_threads(8)
    if sys.argv[1:] == ['--extra']:
        for n, extra, tag in EXTRA:
            generate(n, small=True, extra=extra, tag=tag, subdir='extra')
        sys.exit(0)
    which = sys.argv[1:] or ['mnist', 'wave', 'taxibj', 'sst', 'chairs']
    for n in which:
        generate(n, small=True)
    if not sys.argv[1:]:
        # variants that exercise the remaining flags of the contract
        generate('mnist', small=True, extra='--mixing mul --code_size_s 6 --n_blocks 2 --offset 0', tag='mnist-small-mul')
        generate('mnist', small=True, extra='--skipco', tag='mnist-small-skipco')
        generate('mnist', small=True, extra='--no_s', tag='mnist-small-no_s')

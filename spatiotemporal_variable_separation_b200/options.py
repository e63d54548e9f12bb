"""Command-line contract of the training program.

Flag names, defaults and choices follow the reference parser one for one
(/root/reference/var_sep/options.py:26-134) so that README command lines —
including the ``--gain_res`` prefix abbreviation of ``--gain_resnet`` that the
SST recipe relies on (README.md:86) — parse to the same values.
The table below is the single source; the parser is generated from it.
"""
import argparse

DATASETS = ['mnist', 'chairs', 'taxibj', 'wave', 'wave_partial', 'sst']
ARCH_TYPES = ['dcgan', 'vgg', 'resnet', 'mlp', 'encoderSST']
DECODER_ARCH_TYPES = ['dcgan', 'vgg', 'mlp', 'decoderSST']
INITIALIZATIONS = ['orthogonal', 'kaiming', 'normal']
MIXING = ['concat', 'mul']

# (group, flag, kwargs)
_FLAGS = [
    (None, 'xp_dir', dict(type=str, required=True, help='directory where models are saved')),
    (None, 'chkpt_interval', dict(type=int, default=None, help='save every N epochs')),
    ('amp', 'torch_amp', dict(action='store_true', help='bf16 tensor-core path (reference: torch.cuda.amp)')),
    ('amp', 'apex_amp', dict(action='store_true', help='accepted for compatibility; same as --torch_amp')),
    ('Distributed', 'device', dict(type=int, default=None, help='GPU index')),
    ('Distributed', 'num_workers', dict(type=int, default=4, help='data-loading processes')),
    ('Model', 'nt_cond', dict(type=int, default=5)),
    ('Model', 'nt_pred', dict(type=int, default=10)),
    ('Model', 'code_size_s', dict(type=int, default=128)),
    ('Model', 'code_size_t', dict(type=int, default=20)),
    ('Model', 'mixing', dict(type=str, default='concat', choices=MIXING)),
    ('Model', 'architecture', dict(type=str, default='dcgan', choices=ARCH_TYPES)),
    ('Model', 'decoder_architecture', dict(type=str, default=None, choices=DECODER_ARCH_TYPES)),
    ('Model', 'skipco', dict(action='store_true')),
    ('Model', 'res_hidden_size', dict(type=int, default=512)),
    ('Model', 'n_blocks', dict(type=int, default=1)),
    ('Model', 'enc_hidden_size', dict(type=int, default=64)),
    ('Model', 'dec_hidden_size', dict(type=int, default=64)),
    ('Model', 'enc_n_layers', dict(type=int, default=3)),
    ('Model', 'dec_n_layers', dict(type=int, default=3)),
    ('Model', 'init_encoder', dict(type=str, default='normal', choices=INITIALIZATIONS)),
    ('Model', 'gain_encoder', dict(type=float, default=0.02)),
    ('Model', 'init_resnet', dict(type=str, default='orthogonal', choices=INITIALIZATIONS)),
    ('Model', 'gain_resnet', dict(type=float, default=1.41)),
    ('Model', 'no_s', dict(action='store_true')),
    ('Model', 'offset', dict(type=int, default=5)),
    ('Optimization', 'lamb_ae', dict(type=float, default=10)),
    ('Optimization', 'lamb_s', dict(type=float, default=45)),
    ('Optimization', 'lamb_t', dict(type=float, default=0.001)),
    ('Optimization', 'lamb_pred', dict(type=float, default=45)),
    ('Optimization', 'batch_size', dict(type=int, default=128)),
    ('Optimization', 'lr', dict(type=float, default=4e-4)),
    ('Optimization', 'beta1', dict(type=float, default=0.9)),
    ('Optimization', 'beta2', dict(type=float, default=0.99)),
    ('Optimization', 'epochs', dict(type=int, default=200)),
    ('Optimization', 'scheduler', dict(action='store_true')),
    ('Optimization', 'scheduler_decay', dict(type=float, default=0.5)),
    ('Optimization', 'scheduler_milestones', dict(type=int, nargs='+', default=[300, 400, 500, 600, 700])),
    ('Dataset', 'data', dict(type=str, default='mnist', choices=DATASETS)),
    ('Dataset', 'data_dir', dict(type=str, required=True)),
    (None, 'downsample', dict(type=int, default=2)),
    (None, 'n_wave_points', dict(type=int, default=100)),
    (None, 'zones', dict(type=int, default=list(range(1, 30)), nargs='+')),
    (None, 'n_object', dict(type=int, default=2)),
]


def build_parser():
    p = argparse.ArgumentParser(prog='PDE-Driven Spatiotemporal Disentanglement (training, B200)',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    groups = {}
    for group, flag, kw in _FLAGS:
        if group is None:
            target = p
        elif group == 'amp':            # mutually exclusive, as options.py:39-43
            if group not in groups:
                groups[group] = p.add_argument_group('Mixed-precision training').add_mutually_exclusive_group()
            target = groups[group]
        else:
            if group not in groups:
                groups[group] = p.add_argument_group(group)
            target = groups[group]
        target.add_argument('--' + flag, **kw)
    return p


parser = build_parser()

# main.py:70-102 — what each dataset implies for the networks
DATA_GEOMETRY = {
    'mnist': ('sigmoid', [1, 64, 64]),
    'chairs': ('sigmoid', [3, 64, 64]),
    'taxibj': (None, [2, 32, 32]),
    'sst': (None, [1, 64, 64]),
    'wave': ('sigmoid', [1, 64, 64]),
}


def config_from_args(args):
    """Namespace -> plain dict used by the model builders (adds last_activation/shape;
    applies the --no_s rewrites of main.py:122-127)."""
    cfg = dict(vars(args))
    if cfg['data'] == 'wave_partial':
        cfg['last_activation'], cfg['shape'] = 'sigmoid', [1, cfg['n_wave_points']]
    else:
        cfg['last_activation'], cfg['shape'] = DATA_GEOMETRY[cfg['data']]
        cfg['shape'] = list(cfg['shape'])
    if cfg['no_s']:
        assert not cfg['skipco']
        cfg['code_size_s'] = cfg['code_size_t']
        cfg['mixing'] = 'mul'
    return cfg

"""The five BASELINE.json configurations, as the README command lines that define them.

Each preset is parsed by our own ``options.parser`` (so the flag contract is
exercised every time a preset is built) and turned into the plain ``cfg`` dict
the builders take.  ``small`` variants shrink only the knobs the reference
exposes as flags (batch, filter counts, horizon); they are the parity-test
cases the CPU oracle finishes in seconds.
"""
import shlex

from .options import parser, config_from_args

# /root/reference/README.md:71-95 (epochs/scheduler flags dropped: they do not touch the step)
README_FLAGS = {
    'mnist': '--data mnist --beta1 0.5',
    'chairs': '--data chairs --gain_resnet 0.71 --code_size_t 10 --architecture resnet '
              '--decoder_architecture dcgan --lamb_ae 1 --lamb_s 1',
    'taxibj': '--data taxibj --nt_cond 4 --nt_pred 4 --lr 4e-5 --batch_size 100 --offset 4 '
              '--gain_resnet 0.71 --architecture vgg --lamb_ae 45 --lamb_s 0.0001',
    'sst': '--data sst --nt_cond 4 --nt_pred 6 --code_size_t 64 --code_size_s 196 --gain_res 0.2 --offset 0 '
           '--gain_resnet 0.71 --architecture encoderSST --decoder_architecture decoderSST --lamb_ae 1 '
           '--lamb_s 100 --lamb_t 5e-6 --skipco --n_blocks 2',
    'wave': '--data wave --nt_cond 5 --nt_pred 20 --batch_size 128 --code_size_t 32 --code_size_s 32 '
            '--gain_resnet 0.71 --offset 5 --n_blocks 3 --mixing mul --architecture mlp '
            '--enc_hidden_size 1200 --dec_hidden_size 1200 --dec_n_layers 4 --lamb_ae 1',
}

# flag overrides for the seconds-scale parity cases
SMALL_FLAGS = {
    'mnist': '--batch_size 4 --enc_hidden_size 8 --dec_hidden_size 8 --res_hidden_size 32 '
             '--code_size_s 12 --code_size_t 6 --nt_pred 4',
    'chairs': '--batch_size 3 --dec_hidden_size 8 --res_hidden_size 32 --code_size_s 12 --nt_pred 3 '
              '--nt_cond 3 --offset 3',
    'taxibj': '--batch_size 4 --enc_hidden_size 8 --dec_hidden_size 8 --res_hidden_size 32 --code_size_s 12 '
              '--code_size_t 6',
    'sst': '--batch_size 2 --nt_pred 3 --nt_cond 2 --code_size_t 16 --code_size_s 24 --res_hidden_size 32',
    'wave': '--batch_size 8 --enc_hidden_size 96 --dec_hidden_size 96 --res_hidden_size 64 --nt_pred 6',
}


def preset(name, small=False, extra=''):
    """cfg dict for BASELINE config ``name``; ``extra`` is an additional flag string."""
    argv = ['--xp_dir', '.', '--data_dir', '.'] + shlex.split(README_FLAGS[name])
    if small:
        argv += shlex.split(SMALL_FLAGS[name])
    argv += shlex.split(extra)
    cfg = config_from_args(parser.parse_args(argv))
    cfg['name'] = name + ('-small' if small else '')
    return cfg


def input_kind(cfg):
    """Value distribution of synthetic inputs per dataset."""
    return {'mnist': 'blobs', 'wave': 'blobs', 'chairs': 'uniform', 'taxibj': 'uniform', 'sst': 'normal'}[cfg['data']]

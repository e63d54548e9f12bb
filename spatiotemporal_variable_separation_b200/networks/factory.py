"""String -> network dispatch with the reference's signatures and assertions
(counterpart of /root/reference/var_sep/networks/factory.py:25-87)."""
import numpy as np

from .conv import (DCGAN64Encoder, VGG64Encoder, DCGAN64Decoder, VGG64Decoder, ResNet18, EncoderSST, DecoderSST,
                   DecoderSST_Skip)
from .mlp_encdec import MLPEncoder, MLPDecoder
from .resnet import MLPResnet, ConvResnet
from .utils import init_net


def _flat_size(shape, nt_cond):
    return int(nt_cond * np.prod(np.array(shape)))


# architecture name -> (allowed image sizes or None, constructor from the factory arguments)
_ENCODERS = {
    'dcgan': ((64,), lambda nc, dim, a: DCGAN64Encoder(nc * a['nt_cond'], a['output_size'], a['hidden_size'])),
    'vgg': ((32, 64), lambda nc, dim, a: VGG64Encoder(nc * a['nt_cond'], a['output_size'], a['hidden_size'], vgg32=dim == 32)),
    'resnet': (None, lambda nc, dim, a: ResNet18(a['output_size'], nc * a['nt_cond'])),
    'encoderSST': (None, lambda nc, dim, a: EncoderSST(nc * a['nt_cond'], a['output_size'])),
    'mlp': (None, lambda nc, dim, a: MLPEncoder(_flat_size(a['shape'], a['nt_cond']), a['hidden_size'], a['output_size'],
                                                a['n_layers'])),
}
_DECODERS = {
    'dcgan': ((64,), lambda nc, dim, a: DCGAN64Decoder(nc, a['input_size'], a['hidden_size'], a['skipco'], a['last_activation'],
                                                       a['mixing'])),
    'vgg': ((32, 64), lambda nc, dim, a: VGG64Decoder(nc, a['input_size'], a['hidden_size'], a['skipco'], a['last_activation'],
                                                      a['mixing'], vgg32=dim == 32)),
    'mlp': (None, lambda nc, dim, a: MLPDecoder(a['input_size'], a['hidden_size'], a['shape'], a['n_layers'],
                                                a['last_activation'], a['mixing'])),
    'decoderSST': (None, lambda nc, dim, a: (DecoderSST_Skip if a['skipco'] else DecoderSST)(a['input_size'], nc,
                                                                                             a['last_activation'])),
}
_SKIPCO_DECODERS = ('dcgan', 'vgg', 'decoderSST')


def _build(table, kind, nn_type, shape, args):
    if nn_type not in table:
        raise ValueError(f'unknown {kind} architecture `{nn_type}`')
    sizes, make = table[nn_type]
    nc, dim = shape[0], shape[-1]
    assert sizes is None or dim in sizes                                  # factory.py:29,33,60,64
    return make(nc, dim, args)


def get_encoder(nn_type, shape, output_size, hidden_size, n_layers, nt_cond, init_type, init_gain):
    """factory.py:25-44."""
    encoder = _build(_ENCODERS, 'encoder', nn_type, shape,
                     dict(shape=shape, output_size=output_size, hidden_size=hidden_size, n_layers=n_layers, nt_cond=nt_cond))
    init_net(encoder, init_type=init_type, init_gain=init_gain)
    return encoder


def get_decoder(nn_type, shape, code_size_t, code_size_s, last_activation, hidden_size, n_layers, mixing, skipco,
                init_type, init_gain):
    """factory.py:47-76: skip connections only for the architectures that consume them, `mul` mixing needs equal code
    sizes, the SST decoder concatenates feature maps."""
    assert not skipco or nn_type in _SKIPCO_DECODERS                       # factory.py:49
    if mixing == 'mul':
        assert code_size_t == code_size_s                                  # factory.py:52
    input_size = code_size_t if mixing == 'mul' else code_size_t + code_size_s
    assert nn_type != 'decoderSST' or mixing == 'concat'                    # factory.py:68
    decoder = _build(_DECODERS, 'decoder', nn_type, shape,
                     dict(shape=shape, input_size=input_size, hidden_size=hidden_size, n_layers=n_layers, mixing=mixing,
                          skipco=skipco, last_activation=last_activation))
    init_net(decoder, init_type=init_type, init_gain=init_gain)
    return decoder


def get_resnet(latent_size, n_blocks, hidden_size, init_type, gain_res, fully_conv=False):
    """factory.py:79-87: the latent time-stepper, convolutional for feature-map codes (SST)."""
    resnet = ConvResnet(latent_size, n_blocks=n_blocks, nf=hidden_size) if fully_conv else MLPResnet(latent_size, n_blocks, hidden_size)
    init_net(resnet, init_type=init_type, init_gain=gain_res)
    return resnet


def build_model(cfg, device=None):
    """The network construction block of main.py:117-140 as a function of a cfg dict."""
    from .model import SeparableNetwork
    from .utils import ConstantS
    c = cfg
    if not c['no_s']:
        Es = get_encoder(c['architecture'], c['shape'], c['code_size_s'], c['enc_hidden_size'], c['enc_n_layers'],
                         c['nt_cond'], c['init_encoder'], c['gain_encoder'])
    else:
        Es = ConstantS(return_value=1, code_size=c['code_size_s'])
    Et = get_encoder(c['architecture'], c['shape'], c['code_size_t'], c['enc_hidden_size'], c['enc_n_layers'],
                     c['nt_cond'], c['init_encoder'], c['gain_encoder'])
    decoder = get_decoder(c['decoder_architecture'] or c['architecture'], c['shape'], c['code_size_t'],
                          c['code_size_s'], c['last_activation'], c['dec_hidden_size'], c['dec_n_layers'], c['mixing'],
                          c['skipco'], c['init_encoder'], c['gain_encoder'])
    t_resnet = get_resnet(c['code_size_t'], c['n_blocks'], c['res_hidden_size'], c['init_resnet'], c['gain_resnet'],
                          c['architecture'] == 'encoderSST')
    net = SeparableNetwork(Es, Et, t_resnet, decoder, c['nt_cond'], c['skipco'])
    return net.to(device) if device is not None else net

"""String -> network dispatch with the reference's signatures and assertions
(counterpart of /root/reference/var_sep/networks/factory.py:25-87)."""
import numpy as np

from .conv import (DCGAN64Encoder, VGG64Encoder, DCGAN64Decoder, VGG64Decoder, ResNet18, EncoderSST, DecoderSST,
                   DecoderSST_Skip)
from .mlp_encdec import MLPEncoder, MLPDecoder
from .resnet import MLPResnet, ConvResnet
from .utils import init_net


def get_encoder(nn_type, shape, output_size, hidden_size, n_layers, nt_cond, init_type, init_gain):
    nc, dim = shape[0], shape[-1]
    if nn_type == 'dcgan':
        assert dim == 64
        encoder = DCGAN64Encoder(nc * nt_cond, output_size, hidden_size)
    elif nn_type == 'vgg':
        assert dim in [32, 64]
        encoder = VGG64Encoder(nc * nt_cond, output_size, hidden_size, vgg32=dim == 32)
    elif nn_type == 'resnet':
        encoder = ResNet18(output_size, nc * nt_cond)
    elif nn_type == 'encoderSST':
        encoder = EncoderSST(nc * nt_cond, output_size)
    elif nn_type == 'mlp':
        encoder = MLPEncoder(int(nt_cond * np.prod(np.array(shape))), hidden_size, output_size, n_layers)
    else:
        raise ValueError(f'unknown encoder architecture `{nn_type}`')
    init_net(encoder, init_type=init_type, init_gain=init_gain)
    return encoder


def get_decoder(nn_type, shape, code_size_t, code_size_s, last_activation, hidden_size, n_layers, mixing, skipco,
                init_type, init_gain):
    assert not skipco or nn_type in ['dcgan', 'vgg', 'decoderSST']
    if mixing == 'mul':
        assert code_size_t == code_size_s
        input_size = code_size_t
    else:
        input_size = code_size_t + code_size_s
    nc, dim = shape[0], shape[-1]
    if nn_type == 'dcgan':
        assert dim == 64
        decoder = DCGAN64Decoder(nc, input_size, hidden_size, skipco, last_activation, mixing)
    elif nn_type == 'vgg':
        assert dim in [32, 64]
        decoder = VGG64Decoder(nc, input_size, hidden_size, skipco, last_activation, mixing, vgg32=dim == 32)
    elif nn_type == 'mlp':
        decoder = MLPDecoder(input_size, hidden_size, shape, n_layers, last_activation, mixing)
    elif nn_type == 'decoderSST':
        assert mixing == 'concat'
        decoder = (DecoderSST_Skip if skipco else DecoderSST)(input_size, nc, last_activation)
    else:
        raise ValueError(f'unknown decoder architecture `{nn_type}`')
    init_net(decoder, init_type=init_type, init_gain=init_gain)
    return decoder


def get_resnet(latent_size, n_blocks, hidden_size, init_type, gain_res, fully_conv=False):
    if fully_conv:
        resnet = ConvResnet(latent_size, n_blocks=n_blocks, nf=hidden_size)
    else:
        resnet = MLPResnet(latent_size, n_blocks, hidden_size)
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

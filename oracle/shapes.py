"""State-dict layouts (key -> shape) of every reference network.  TEST INFRASTRUCTURE ONLY.

Restates the constructors of /root/reference/var_sep/networks/{conv,mlp_encdec,resnet}.py
as shape tables, in ``state_dict()`` order.  ``tests/test_oracle_golden.py`` checks
these tables against the listing dumped from the real reference modules.
"""
import numpy as np


class _T(dict):
    def conv(self, key, cin, cout, k):
        self[key + '.weight'] = (cout, cin, k, k)
        self[key + '.bias'] = (cout,)

    def convT(self, key, cin, cout, k):
        self[key + '.weight'] = (cin, cout, k, k)
        self[key + '.bias'] = (cout,)

    def bn(self, key, c):
        self[key + '.weight'] = (c,)
        self[key + '.bias'] = (c,)
        self[key + '.running_mean'] = (c,)
        self[key + '.running_var'] = (c,)
        self[key + '.num_batches_tracked'] = ()

    def block(self, key, cin, cout, k, bn=True, transposed=False):
        (self.convT if transposed else self.conv)(key + '.0', cin, cout, k)
        if bn:
            self.bn(key + '.1', cout)

    def linear(self, key, nin, nout):
        self[key + '.weight'] = (nout, nin)
        self[key + '.bias'] = (nout,)

    def mlp(self, pre, ninp, nhid, nout, nlayers):
        """mlp.py:45-71."""
        for il in range(nlayers):
            self.linear(f'{pre}module.{il}.{0 if il == 0 else 1}',
                        ninp if il == 0 else nhid, nout if il == nlayers - 1 else nhid)


def encoder_shapes(nn_type, shape, output_size, hidden_size, n_layers, nt_cond):
    """factory.py:25-44."""
    t = _T()
    nc, nf, nh = shape[0] * nt_cond, hidden_size, output_size
    if nn_type == 'dcgan':                                       # conv.py:118-124
        t.block('conv.0', nc, nf, 4, bn=False)
        for i in (1, 2, 3):
            t.block(f'conv.{i}', nf * 2 ** (i - 1), nf * 2 ** i, 4)
        t.linear('last_op.1', nf * 8 * 16, nh)
    elif nn_type == 'vgg':                                       # conv.py:146-171
        cin = nc
        for s, nblk in enumerate((2, 2, 3, 3)):
            for j in range(nblk):
                t.block(f'conv.{s}.{j + (s > 0)}', cin, nf * 2 ** s, 3)
                cin = nf * 2 ** s
        t.block('last_op.1', nf * 8, nh, 4)
    elif nn_type == 'resnet':                                    # conv.py:512-527
        t.conv('conv1', nc, 64, 5)
        t.bn('bn1', 64)
        cin = 64
        for li, planes in ((1, 64), (2, 128), (3, 256), (4, 512)):
            for b in (0, 1):
                pre = f'layer{li}.{b}.'
                t.conv(pre + 'conv1', cin if b == 0 else planes, planes, 3)
                t.bn(pre + 'bn1', planes)
                t.conv(pre + 'conv2', planes, planes, 3)
                t.bn(pre + 'bn2', planes)
                if b == 0 and li > 1:
                    t.conv(pre + 'downsample.0', cin, planes, 1)
                    t.bn(pre + 'downsample.1', planes)
            cin = planes
        t.conv('conv_out', 512, nh, 3)
        t.bn('bn_out', nh)
    elif nn_type == 'encoderSST':                                # conv.py:326-343
        t.block('conv1.0', nc, 64, 3)
        t.block('conv1.1', 64, 64, 3)
        t.block('conv2.1', 64, 128, 3)
        t.block('conv2.2', 128, 128, 3)
        t.block('conv3.1', 128, 256, 3)
        t.block('conv3.2', 256, 256, 3)
        t.block('conv3.3', 256, 256, 3)
        t.block('conv4.0', 256, 512, 3)
        t.block('conv4.1', 512, nh, 3)
        t.block('conv4.2', nh, nh, 3, bn=False)
    elif nn_type == 'mlp':                                       # mlp_encdec.py:26-28
        t.mlp('mlp.', int(nt_cond * np.prod(shape)), hidden_size, output_size, n_layers)
    else:
        raise ValueError(nn_type)
    return dict(t)


def decoder_shapes(nn_type, shape, code_size_t, code_size_s, hidden_size, n_layers, mixing, skipco):
    """factory.py:47-76."""
    t = _T()
    ny = code_size_t if mixing == 'mul' else code_size_t + code_size_s
    nc, nf, coef = shape[0], hidden_size, 2 if skipco else 1
    if nn_type == 'dcgan':                                       # conv.py:256-264
        t.block('first_upconv', ny, nf * 8, 4, transposed=True)
        t.block('conv.0', nf * 8 * coef, nf * 4, 4, transposed=True)
        t.block('conv.1', nf * 4 * coef, nf * 2, 4, transposed=True)
        t.block('conv.2', nf * 2 * coef, nf, 4, transposed=True)
        t.convT('conv.3', nf * coef, nc, 4)
    elif nn_type == 'vgg':                                       # conv.py:293-320
        t.block('first_upconv.0', ny, nf * 8, 4, transposed=True)
        t.block('conv.0.0', nf * 8 * coef, nf * 8, 3)
        t.block('conv.0.1', nf * 8, nf * 8, 3)
        t.block('conv.0.2', nf * 8, nf * 4, 3)
        t.block('conv.1.0', nf * 4 * coef, nf * 4, 3)
        t.block('conv.1.1', nf * 4, nf * 4, 3)
        t.block('conv.1.2', nf * 4, nf * 2, 3)
        t.block('conv.2.0', nf * 2 * coef, nf * 2, 3)
        t.block('conv.2.1', nf * 2, nf, 3)
        t.block('conv.3.0', nf * coef, nf, 3)
        t.convT('conv.3.1', nf, nc, 3)
    elif nn_type == 'mlp':                                       # mlp_encdec.py:36-41
        t.mlp('mlp.', ny, hidden_size, int(np.prod(shape)), n_layers)
    elif nn_type == 'decoderSST':
        if skipco:                                               # conv.py:362-383
            chans = {'conv1': (ny, 256, 256, 128), 'conv2': (384, 128, 64, 64),
                     'conv3': (192, 128, 64, 64), 'conv4': (128, 64, 64, nc)}
        else:                                                    # conv.py:402-417
            chans = {'conv1': (ny, 256, 256, 128), 'conv2': (128, 128, 128, 64), 'conv3': (64, 64, nc)}
        for name, cs in chans.items():
            for j in range(len(cs) - 1):
                t.block(f'{name}.{j}', cs[j], cs[j + 1], 3)
    else:
        raise ValueError(nn_type)
    return dict(t)


def resnet_shapes(latent_size, n_blocks, hidden_size, fully_conv):
    """factory.py:79-87; resnet.py:22-88."""
    t = _T()
    for j in range(n_blocks):
        if fully_conv:
            pre = f'resblock_modules.{j}.conv.'
            t.block(pre + '0', latent_size, hidden_size, 3)
            t.block(pre + '1', hidden_size, hidden_size, 3)
            t.block(pre + '2', hidden_size, latent_size, 3)
        else:
            t.mlp(f'blocks.{j}.mlp.', latent_size, hidden_size, latent_size, 3)
    return dict(t)


def model_shapes(cfg):
    """main.py:120-138 — the four state-dict layouts of one configuration."""
    c = cfg
    code_s = c['code_size_t'] if c.get('no_s') else c['code_size_s']
    mixing = 'mul' if c.get('no_s') else c['mixing']
    if c.get('no_s'):
        es = {}
    else:
        es = encoder_shapes(c['architecture'], c['shape'], code_s, c['enc_hidden_size'], c['enc_n_layers'],
                            c['nt_cond'])
    et = encoder_shapes(c['architecture'], c['shape'], c['code_size_t'], c['enc_hidden_size'],
                        c['enc_n_layers'], c['nt_cond'])
    dec = decoder_shapes(c.get('decoder_architecture') or c['architecture'], c['shape'], c['code_size_t'],
                         code_s, c['dec_hidden_size'], c['dec_n_layers'], mixing, c['skipco'])
    res = resnet_shapes(c['code_size_t'], c['n_blocks'], c['res_hidden_size'],
                        c['architecture'] == 'encoderSST')
    return {'Es': es, 'Et': et, 'decoder': dec, 't_resnet': res}

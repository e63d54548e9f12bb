"""Functional CPU restatement of the reference networks.  TEST INFRASTRUCTURE ONLY.

Every network is a pure function ``f(P, prefix, inputs, train)`` over a flat
dict ``P`` of tensors keyed exactly like the reference ``state_dict`` (so
weights move between the reference, the oracle and the CUDA modules without
renaming).  Nothing here subclasses ``nn.Module``; autograd on the CPU supplies
the gradients.

Citations are into /root/reference/var_sep/networks/.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default (conv.py:56-57)
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default


# --------------------------------------------------------------------------- #
# primitive layers
# --------------------------------------------------------------------------- #
def act(x, name):
    """utils.py:50-72 (activation_factory).  LeakyReLU slope is 0.2."""
    if name in (None, 'none', 'identity'):
        return x
    if name == 'relu':
        return F.relu(x)
    if name == 'leaky_relu':
        return F.leaky_relu(x, 0.2)
    if name == 'elu':
        return F.elu(x)
    if name == 'sigmoid':
        return torch.sigmoid(x)
    if name == 'tanh':
        return torch.tanh(x)
    raise ValueError(f'Activation function `{name}` not yet implemented')


def conv(P, key, x, stride=1, pad=0):
    return F.conv2d(x, P[key + '.weight'], P[key + '.bias'], stride, pad)


def convT(P, key, x, stride=1, pad=0):
    return F.conv_transpose2d(x, P[key + '.weight'], P[key + '.bias'], stride, pad)


def bnorm(P, key, x, train):
    """Train mode: batch statistics of *this call*, running stats EMA with the
    unbiased variance, ``num_batches_tracked += 1``."""
    if train:
        P[key + '.num_batches_tracked'] += 1
    return F.batch_norm(x, P[key + '.running_mean'], P[key + '.running_var'],
                        P[key + '.weight'], P[key + '.bias'], train, BN_MOMENTUM, BN_EPS)


def block(P, key, x, train, a='leaky_relu', stride=1, pad=1, bn=True, transposed=False):
    """conv.py:41-60 make_conv_block: ``key.0`` conv, ``key.1`` BN, activation."""
    op = convT if transposed else conv
    h = op(P, key + '.0', x, stride, pad)
    if bn:
        h = bnorm(P, key + '.1', h, train)
    return act(h, a)


def fold_time(x):
    """conv.py:90 — [B,T,C,H,W] -> [B,T*C,H,W]."""
    return x.reshape(x.size(0), -1, x.size(3), x.size(4))


def mlp(P, pre, x):
    """mlp.py:24-75 — pre-activation MLP: layer 0 is Linear, layer i>0 is ReLU->Linear."""
    il = 0
    while True:
        key = f'{pre}module.{il}.{0 if il == 0 else 1}'
        if key + '.weight' not in P:
            break
        if il > 0:
            x = F.relu(x)
        x = F.linear(x, P[key + '.weight'], P[key + '.bias'])
        il += 1
    return x


def mix(z1, z2, mixing):
    """conv.py:220-223 / mlp_encdec.py:44-47."""
    if mixing == 'concat':
        return torch.cat([z1, z2], dim=1)
    return z1 * z2


# --------------------------------------------------------------------------- #
# encoders   (return code, or (code, skips deepest-first) if return_skip)
# --------------------------------------------------------------------------- #
def enc_dcgan(P, x, train, return_skip=False):
    """conv.py:81-99,102-124."""
    h = fold_time(x)
    skips = []
    h = block(P, 'conv.0', h, train, stride=2, pad=1, bn=False)
    skips.append(h)
    for i in (1, 2, 3):
        h = block(P, f'conv.{i}', h, train, stride=2, pad=1)
        skips.append(h)
    nh = P['last_op.1.weight'].shape[0]
    h = F.linear(h.flatten(1), P['last_op.1.weight'], P['last_op.1.bias']).view(-1, nh)
    return (h, skips[::-1]) if return_skip else h


def enc_vgg(P, x, train, return_skip=False):
    """conv.py:127-171.  Stage s>0 starts with MaxPool2 (module index 0), so the
    conv blocks of stages 1..3 sit at indices 1.. ; last_op = [pool|identity, conv4x4+BN]."""
    h = fold_time(x)
    vgg32 = h.shape[-1] == 32
    skips = []
    for s, nblk in enumerate((2, 2, 3, 3)):
        if s > 0:
            h = F.max_pool2d(h, 2, 2, 0)
        for j in range(nblk):
            h = block(P, f'conv.{s}.{j + (s > 0)}', h, train)
        skips.append(h)
    if not vgg32:
        h = F.max_pool2d(h, 2, 2, 0)
    nh = P['last_op.1.0.weight'].shape[0]
    h = block(P, 'last_op.1', h, train, a='none', stride=1, pad=0).view(-1, nh)
    return (h, skips[::-1]) if return_skip else h


def enc_sst(P, x, train, return_skip=False):
    """conv.py:323-356."""
    h = fold_time(x)
    h1 = block(P, 'conv1.1', block(P, 'conv1.0', h, train), train)
    h = F.max_pool2d(h1, 2, 2, 0)
    h2 = block(P, 'conv2.2', block(P, 'conv2.1', h, train), train)
    h = F.max_pool2d(h2, 2, 2, 0)
    for j in (1, 2, 3):
        h = block(P, f'conv3.{j}', h, train)
    h3 = h
    h = block(P, 'conv4.0', h3, train)
    h = block(P, 'conv4.1', h, train)
    h4 = block(P, 'conv4.2', h, train, a='none', bn=False)
    return (h4, [h3, h2, h1]) if return_skip else h4


def _basic_block(P, pre, x, train, stride):
    """conv.py:439-468."""
    out = F.relu(bnorm(P, pre + 'bn1', conv(P, pre + 'conv1', x, stride, 1), train))
    out = bnorm(P, pre + 'bn2', conv(P, pre + 'conv2', out, 1, 1), train)
    if pre + 'downsample.0.weight' in P:
        res = bnorm(P, pre + 'downsample.1', conv(P, pre + 'downsample.0', x, stride, 0), train)
    else:
        res = x
    return F.relu(out + res)


def enc_resnet18(P, x, train, return_skip=False):
    """conv.py:510-564.  ``bn_out`` exists in the state dict but is never applied."""
    h = fold_time(x)
    h = F.relu(bnorm(P, 'bn1', conv(P, 'conv1', h, 2, 3), train))
    h = F.max_pool2d(h, 3, 2, 1)
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        h = _basic_block(P, f'layer{li}.0.', h, train, stride)
        h = _basic_block(P, f'layer{li}.1.', h, train, 1)
    h = conv(P, 'conv_out', h, 1, 0)
    return h.reshape(len(h), -1)


def enc_mlp(P, x, train, return_skip=False):
    """mlp_encdec.py:25-32."""
    return mlp(P, 'mlp.', x.reshape(len(x), -1))


def enc_constant(P, x, train, return_skip=False):
    """utils.py:21-29 ConstantS."""
    return torch.ones(len(x), P['__code_size__']).to(x) * P.get('__return_value__', 1)


ENCODERS = {'dcgan': enc_dcgan, 'vgg': enc_vgg, 'resnet': enc_resnet18, 'encoderSST': enc_sst,
            'mlp': enc_mlp, 'constant': enc_constant}


# --------------------------------------------------------------------------- #
# decoders
# --------------------------------------------------------------------------- #
def dec_dcgan(P, z1, z2, skip, train, mixing, last_activation):
    """conv.py:207-230,233-264."""
    z = mix(z1, z2, mixing)
    h = block(P, 'first_upconv', z.view(*z.shape, 1, 1), train, stride=1, pad=0, transposed=True)
    for i in range(4):
        if skip is not None:
            h = torch.cat([h, skip[i]], 1)
        if i < 3:
            h = block(P, f'conv.{i}', h, train, stride=2, pad=1, transposed=True)
        else:
            h = convT(P, 'conv.3', h, 2, 1)
    return act(h, last_activation)


def dec_vgg(P, z1, z2, skip, train, mixing, last_activation, vgg32):
    """conv.py:267-320."""
    z = mix(z1, z2, mixing)
    h = block(P, 'first_upconv.0', z.view(*z.shape, 1, 1), train, stride=1, pad=0, transposed=True)
    if not vgg32:
        h = F.interpolate(h, scale_factor=2, mode='nearest')
    for s, nblk in enumerate((3, 3, 2, 1)):
        if skip is not None:
            h = torch.cat([h, skip[s]], 1)
        for j in range(nblk):
            h = block(P, f'conv.{s}.{j}', h, train)
        if s < 3:
            h = F.interpolate(h, scale_factor=2, mode='nearest')
        else:
            h = convT(P, 'conv.3.1', h, 1, 1)
    return act(h, last_activation)


def dec_sst_skip(P, z1, z2, skip, train, mixing, last_activation):
    """conv.py:359-396."""
    h3, h2, h1 = skip
    out = torch.cat([z1, z2], dim=1)
    for j in range(3):
        out = block(P, f'conv1.{j}', out, train)
    out = torch.cat([h3, out], dim=1)
    for j in range(3):
        out = block(P, f'conv2.{j}', out, train)
    out = F.interpolate(out, scale_factor=2, mode='nearest')
    out = torch.cat([h2, out], dim=1)
    for j in range(3):
        out = block(P, f'conv3.{j}', out, train)
    out = F.interpolate(out, scale_factor=2, mode='nearest')
    out = torch.cat([h1, out], dim=1)
    for j in range(3):
        out = block(P, f'conv4.{j}', out, train)
    return act(out, last_activation)


def dec_sst(P, z1, z2, skip, train, mixing, last_activation):
    """conv.py:399-426."""
    x = torch.cat([z1, z2], dim=1)
    for j in range(3):
        x = block(P, f'conv1.{j}', x, train)
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    for j in range(3):
        x = block(P, f'conv2.{j}', x, train)
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    for j in range(2):
        x = block(P, f'conv3.{j}', x, train)
    return act(x, last_activation)


def dec_mlp(P, z1, z2, skip, train, mixing, last_activation, shape):
    """mlp_encdec.py:35-50."""
    x = act(mlp(P, 'mlp.', mix(z1, z2, mixing)), last_activation)
    return x.view([-1] + list(shape))


# --------------------------------------------------------------------------- #
# latent time-steppers    (return x_next, [residual per block])
# --------------------------------------------------------------------------- #
def step_mlp(P, x, train):
    """resnet.py:22-50: x <- x + MLP3(x), n_blocks times; no runtime gain."""
    res = []
    j = 0
    while f'blocks.{j}.mlp.module.0.0.weight' in P:
        r = mlp(P, f'blocks.{j}.mlp.', x)
        x = x + r
        res.append(r)
        j += 1
    return x, res


def step_conv(P, x, train):
    """resnet.py:53-88: three 3x3 conv+BN (LeakyReLU on the first two), identity shortcut."""
    res = []
    i = 0
    while f'resblock_modules.{i}.conv.0.0.weight' in P:
        pre = f'resblock_modules.{i}.conv.'
        r = block(P, pre + '0', x, train)
        r = block(P, pre + '1', r, train)
        r = block(P, pre + '2', r, train, a='none')
        x = x + r
        res.append(r)
        i += 1
    return x, res


# --------------------------------------------------------------------------- #
# model composition
# --------------------------------------------------------------------------- #
class Net:
    """Plain container (not an nn.Module): four state dicts + the flags that
    select the functions above.  ``cfg`` uses the option names of options.py."""

    def __init__(self, cfg, Es, Et, decoder, t_resnet):
        self.cfg = cfg
        self.P = {'Es': Es, 'Et': Et, 'decoder': decoder, 't_resnet': t_resnet}
        self.train = True

    # --- the three module calls -------------------------------------------------
    def Es(self, x, return_skip=False):
        arch = 'constant' if self.cfg.get('no_s') else self.cfg['architecture']
        return ENCODERS[arch](self.P['Es'], x, self.train, return_skip)

    def Et(self, x):
        return ENCODERS[self.cfg['architecture']](self.P['Et'], x, self.train)

    def decoder(self, s, t, skip=None):
        c = self.cfg
        arch = c.get('decoder_architecture') or c['architecture']
        P = self.P['decoder']
        la, mx = c['last_activation'], c['mixing']
        assert (skip is None) == (not c['skipco'])            # conv.py:218
        if arch == 'dcgan':
            return dec_dcgan(P, s, t, skip, self.train, mx, la)
        if arch == 'vgg':
            return dec_vgg(P, s, t, skip, self.train, mx, la, c['shape'][-1] == 32)
        if arch == 'mlp':
            return dec_mlp(P, s, t, skip, self.train, mx, la, c['shape'])
        if arch == 'decoderSST':
            f = dec_sst_skip if c['skipco'] else dec_sst
            return f(P, s, t, skip, self.train, mx, la)
        raise ValueError(arch)

    def t_resnet(self, x):
        f = step_conv if self.cfg['architecture'] == 'encoderSST' else step_mlp
        return f(self.P['t_resnet'], x, self.train)

    # --- model.py:52-89 ---------------------------------------------------------
    def get_forecast(self, cond, n_forecast, init_t_code=None, init_s_code=None):
        skipco = self.cfg['skipco']
        s_code = self.Es(cond, return_skip=skipco) if init_s_code is None else init_s_code
        if skipco:
            s_code, s_skip = s_code
        else:
            s_skip = None
        t_code = self.Et(cond) if init_t_code is None else init_t_code
        t_codes, forecasts, t_residuals = [t_code], [self.decoder(s_code, t_code, s_skip)], []
        for _ in range(1, n_forecast):
            t_code, t_res = self.t_resnet(t_code)
            t_codes.append(t_code)
            t_residuals.append(t_res)
            forecasts.append(self.decoder(s_code, t_code, s_skip))
        return torch.stack(forecasts, 1), torch.stack(t_codes, 1), s_code, t_residuals

    # --- helpers ------------------------------------------------------------------
    def parameters(self):
        """(name, tensor) in the order of ``SeparableNetwork.parameters()``:
        Es, Et, decoder, t_resnet (model.py:34-37), buffers excluded."""
        out = []
        for part in ('Es', 'Et', 'decoder', 't_resnet'):
            for k, v in self.P[part].items():
                if isinstance(v, torch.Tensor) and v.is_floating_point() and \
                        not k.endswith(('running_mean', 'running_var')):
                    out.append((f'{part}.{k}', v))
        return out

    def requires_grad_(self, flag=True):
        for _, p in self.parameters():
            p.requires_grad_(flag)
        return self

"""Convolutional encoders / decoders (DCGAN, VGG, SST, ResNet18).

Counterpart of /root/reference/var_sep/networks/conv.py with the same class names, constructor
arguments, ``forward`` signatures and ``state_dict`` layout.  ``torch.nn`` layers are used only as
*parameter containers* (so keys, shapes and ``init_net`` behave identically); the arithmetic is done
by the sm_100a kernels through ``ops.conv_block`` on NHWC tensors of the compute dtype.

Every network has two entry points:
  * ``forward(...)``  – the reference's public signature, fp32 NCHW in / out;
  * ``encode`` / ``decode`` – the internal NHWC form with a ``groups`` argument: the batch holds
    ``groups`` consecutive reference *calls* and train-mode BatchNorm statistics are taken per
    call, which lets one launch stand for e.g. all 16 decoder calls of a step.
"""
import torch.nn as nn

from .. import ops
from .utils import activation_factory, activation_name


class ConvBlock(nn.Sequential):
    """conv -> [BatchNorm2d] -> [activation]; children indices as make_conv_block (conv.py:41-60)."""

    def __init__(self, conv, activation, bn=True):
        mods = [conv]
        if bn:
            mods.append(nn.BatchNorm2d(conv.out_channels))
        if activation != 'none':
            mods.append(activation_factory(activation))
        super().__init__(*mods)
        self.kind = 'convT' if isinstance(conv, nn.ConvTranspose2d) else 'conv'
        self.act = activation_name(activation)
        self.has_bn = bn

    def forward(self, x, groups=1):
        return ops.conv_block(x, self[0], self[1] if self.has_bn else None, self.act, self.kind, groups)


def make_conv_block(conv, activation, bn=True):
    return ConvBlock(conv, activation, bn)


class _Pool(nn.MaxPool2d):
    def forward(self, x, groups=1):
        return ops.maxpool(x, self.kernel_size, self.stride, self.padding)


class _Up(nn.Upsample):
    def forward(self, x, groups=1):
        return ops.upsample2(x)


class _Id(nn.Identity):
    def forward(self, x, groups=1):
        return x


class _Stage(nn.Sequential):
    """Sequential whose children all take (x, groups)."""

    def forward(self, x, groups=1):
        for m in self:
            x = m(x, groups)
        return x


def _fold_time(x):
    return x.reshape(x.size(0), -1, x.size(3), x.size(4))          # conv.py:90


# ================================================================================================
# encoders
# ================================================================================================
class BaseEncoder(nn.Module):
    """conv.py:63-99."""

    def __init__(self, nh):
        super().__init__()
        self.nh = nh

    def encode(self, h, groups=1, return_skip=False):
        skips = []
        for layer in self.conv:
            h = layer(h, groups)
            skips.append(h)
        h = self._last(h, groups)
        if return_skip:
            return h, skips[::-1]
        return h

    def forward(self, x, return_skip=False):
        out = self.encode(ops.to_internal(_fold_time(x)), 1, return_skip)
        if return_skip:
            return ops.to_external(out[0]).view(-1, self.nh), [ops.to_external(s) for s in out[1]]
        return ops.to_external(out).view(-1, self.nh)


class DCGAN64Encoder(BaseEncoder):
    """conv.py:102-124: 4 x [Conv k4 s2 p1 (+BN) + LeakyReLU] then Flatten + Linear."""

    def __init__(self, nc, nh, nf):
        super().__init__(nh)
        chans = [nc, nf, nf * 2, nf * 4, nf * 8]
        self.conv = nn.ModuleList([ConvBlock(nn.Conv2d(chans[i], chans[i + 1], 4, 2, 1), 'leaky_relu', bn=i > 0)
                                   for i in range(4)])
        self.last_op = nn.Sequential(nn.Flatten(), nn.Linear(nf * 8 * 4 * 4, nh))
        self._flat = (nh, nf * 8, 4, 4)

    def _last(self, h, groups):
        # Flatten (c,h,w order) + Linear == a 4x4 valid convolution with the same weight memory
        return ops.conv_block(h, self.last_op[1], None, None, 'conv', groups, wshape=self._flat)


def _vgg_stage(cin, cout, n, pool):
    mods = [_Pool(kernel_size=2, stride=2, padding=0)] if pool else []
    for j in range(n):
        mods.append(ConvBlock(nn.Conv2d(cin if j == 0 else cout, cout, 3, 1, 1), 'leaky_relu'))
    return _Stage(*mods)


class VGG64Encoder(BaseEncoder):
    """conv.py:127-171."""

    def __init__(self, nc, nh, nf, vgg32=False):
        super().__init__(nh)
        self.conv = nn.ModuleList([
            _vgg_stage(nc, nf, 2, False), _vgg_stage(nf, nf * 2, 2, True),
            _vgg_stage(nf * 2, nf * 4, 3, True), _vgg_stage(nf * 4, nf * 8, 3, True)])
        self.last_op = _Stage(_Pool(kernel_size=2, stride=2, padding=0) if not vgg32 else _Id(),
                              ConvBlock(nn.Conv2d(nf * 8, nh, 4, 1, 0), 'none'))

    def _last(self, h, groups):
        return self.last_op(h, groups)


class EncoderSST(nn.Module):
    """conv.py:323-356: the code is a feature map [B, out_c, 16, 16]."""

    def __init__(self, in_c, out_c):
        super().__init__()
        self.conv1 = _vgg_stage(in_c, 64, 2, False)
        self.conv2 = _vgg_stage(64, 128, 2, True)
        self.conv3 = _vgg_stage(128, 256, 3, True)
        self.conv4 = _Stage(ConvBlock(nn.Conv2d(256, 512, 3, 1, 1), 'leaky_relu'),
                            ConvBlock(nn.Conv2d(512, out_c, 3, 1, 1), 'leaky_relu'),
                            ConvBlock(nn.Conv2d(out_c, out_c, 3, 1, 1), 'none', bn=False))

    def encode(self, h, groups=1, return_skip=False):
        h1 = self.conv1(h, groups)
        h2 = self.conv2(h1, groups)
        h3 = self.conv3(h2, groups)
        h4 = self.conv4(h3, groups)
        if return_skip:
            return h4, [h3, h2, h1]
        return h4

    def forward(self, x, return_skip=False):
        out = self.encode(ops.to_internal(_fold_time(x)), 1, return_skip)
        if return_skip:
            return ops.to_external(out[0]), [ops.to_external(s) for s in out[1]]
        return ops.to_external(out)


class BasicBlock(nn.Module):
    """conv.py:439-468."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x, groups=1):
        out = ops.conv_block(x, self.conv1, self.bn1, 'relu', 'conv', groups)
        out = ops.conv_block(out, self.conv2, self.bn2, None, 'conv', groups)
        res = x
        if self.downsample is not None:
            res = ops.conv_block(x, self.downsample[0], self.downsample[1], None, 'conv', groups)
        return ops.add_act(out, res, 'relu')


class ResNet18(nn.Module):
    """conv.py:510-564 (``bn_out`` is constructed but never applied, as upstream)."""

    def __init__(self, pose_dim, nc=3, out_f=None):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(nc, 64, kernel_size=5, stride=2, padding=3)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = _Pool(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, 2)
        self.layer2 = self._make_layer(128, 2, stride=2)
        self.layer3 = self._make_layer(256, 2, stride=2)
        self.layer4 = self._make_layer(512, 2, stride=2)
        self.conv_out = nn.Conv2d(512, pose_dim, kernel_size=3)
        self.bn_out = nn.BatchNorm2d(pose_dim)
        self.out_function = activation_factory(out_f)
        self._out_act = activation_name(out_f)

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride),
                                       nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        layers += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return _Stage(*layers)

    def encode(self, h, groups=1, return_skip=False):
        h = ops.conv_block(h, self.conv1, self.bn1, 'relu', 'conv', groups)
        h = self.maxpool(h)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            h = layer(h, groups)
        return ops.conv_block(h, self.conv_out, None, self._out_act, 'conv', groups)

    def forward(self, x, return_skip=False):
        h = self.encode(ops.to_internal(_fold_time(x)))
        return ops.to_external(h).reshape(len(x), -1)


# ================================================================================================
# decoders
# ================================================================================================
def _mix(z1, z2, mixing):
    """cat / mul of the S and T codes (conv.py:220-223); z1 may hold one call's S for all groups."""
    if mixing == 'concat':
        return ops.concat_channels(z1, z2)
    return ops.mul_bcast(z1, z2)


def _skips_internal(skip):
    return None if skip is None else [ops.to_internal(s) for s in skip]


class BaseDecoder(nn.Module):
    """conv.py:174-230."""

    def __init__(self, ny, skip, last_activation, mixing):
        super().__init__()
        self.ny = ny
        self.skip = skip
        self.mixing = mixing
        self.last_activation = activation_factory(last_activation)
        self._last_act = activation_name(last_activation)

    def decode(self, z1, z2, skip=None, groups=1):
        assert skip is None and not self.skip or self.skip and skip is not None
        h = self.first_upconv(_mix(z1, z2, self.mixing), groups)
        for i, layer in enumerate(self.conv):
            if skip is not None:
                h = ops.concat_channels(h, skip[i])
            h = layer(h, groups)
        return h

    def decode_external(self, z1, z2, skip=None, groups=1):
        return ops.to_external(self.decode(z1, z2, skip, groups))

    def forward(self, z1, z2, skip=None, groups=1):
        return self.decode_external(ops.to_internal(z1), ops.to_internal(z2), _skips_internal(skip), groups)


class _LastConvT(nn.ConvTranspose2d):
    """Bias-only transposed convolution with the decoder's output activation fused in its epilogue."""
    _act = None

    def forward(self, x, groups=1):
        return ops.conv_block(x, self, None, self._act, 'convT', groups)


class DCGAN64Decoder(BaseDecoder):
    """conv.py:233-264."""

    def __init__(self, nc, ny, nf, skip, last_activation, mixing):
        super().__init__(ny, skip, last_activation, mixing)
        coef = 2 if skip else 1
        self.first_upconv = ConvBlock(nn.ConvTranspose2d(ny, nf * 8, 4, 1, 0), 'leaky_relu')
        last = _LastConvT(nf * coef, nc, 4, 2, 1)
        last._act = self._last_act
        self.conv = nn.ModuleList([
            ConvBlock(nn.ConvTranspose2d(nf * 8 * coef, nf * 4, 4, 2, 1), 'leaky_relu'),
            ConvBlock(nn.ConvTranspose2d(nf * 4 * coef, nf * 2, 4, 2, 1), 'leaky_relu'),
            ConvBlock(nn.ConvTranspose2d(nf * 2 * coef, nf, 4, 2, 1), 'leaky_relu'),
            last])

    def decode(self, z1, z2, skip=None, groups=1):
        if skip is not None:
            return super().decode(z1, z2, skip, groups)
        # without skip connections the last BatchNorm block feeds the thin output convolution directly: one fused
        # operator (ops.decoder_tail) that never materialises the normalised 2*nf... -> nf tensor
        h = self.first_upconv(_mix(z1, z2, self.mixing), groups)
        for layer in self.conv[:2]:
            h = layer(h, groups)
        return ops.decoder_tail(h, self.conv[2], self.conv[3], groups)


class VGG64Decoder(BaseDecoder):
    """conv.py:267-320."""

    def __init__(self, nc, ny, nf, skip, last_activation, mixing, vgg32=False):
        super().__init__(ny, skip, last_activation, mixing)
        coef = 2 if skip else 1

        def cb(cin, cout):
            return ConvBlock(nn.Conv2d(cin, cout, 3, 1, 1), 'leaky_relu')

        self.first_upconv = _Stage(ConvBlock(nn.ConvTranspose2d(ny, nf * 8, 4, 1, 0), 'leaky_relu'),
                                   _Up(scale_factor=2, mode='nearest') if not vgg32 else _Id())
        last = _LastConvT(nf, nc, 3, 1, 1)
        last._act = self._last_act
        self.conv = nn.ModuleList([
            _Stage(cb(nf * 8 * coef, nf * 8), cb(nf * 8, nf * 8), cb(nf * 8, nf * 4), _Up(scale_factor=2, mode='nearest')),
            _Stage(cb(nf * 4 * coef, nf * 4), cb(nf * 4, nf * 4), cb(nf * 4, nf * 2), _Up(scale_factor=2, mode='nearest')),
            _Stage(cb(nf * 2 * coef, nf * 2), cb(nf * 2, nf), _Up(scale_factor=2, mode='nearest')),
            _Stage(cb(nf * coef, nf), last)])


def _sst_stage(chans, last_act='leaky_relu', up=False):
    mods = [ConvBlock(nn.Conv2d(chans[j], chans[j + 1], 3, 1, 1), 'leaky_relu') for j in range(len(chans) - 1)]
    if up:
        mods.append(_Up(scale_factor=2, mode='nearest'))
    return _Stage(*mods)


class DecoderSST_Skip(nn.Module):
    """conv.py:359-396."""

    def __init__(self, in_c, out_c, out_f):
        super().__init__()
        self.conv1 = _sst_stage((in_c, 256, 256, 128))
        self.conv2 = _sst_stage((256 + 128, 128, 64, 64), up=True)
        self.conv3 = _sst_stage((128 + 64, 128, 64, 64), up=True)
        self.conv4 = _sst_stage((64 * 2, 64, 64, out_c))
        self.out_f = activation_factory(out_f)
        self._out_act = activation_name(out_f)

    def decode(self, s_code, t_code, skip, groups=1):
        h3, h2, h1 = skip
        out = self.conv1(ops.concat_channels(s_code, t_code), groups)
        out = self.conv2(ops.concat_channels(h3, out), groups)
        out = self.conv3(ops.concat_channels(h2, out), groups)
        out = self.conv4(ops.concat_channels(h1, out), groups)
        return _standalone_act(out, self._out_act)

    def decode_external(self, s_code, t_code, skip, groups=1):
        return ops.to_external(self.decode(s_code, t_code, skip, groups))

    def forward(self, s_code, t_code, skip, groups=1):
        return self.decode_external(ops.to_internal(s_code), ops.to_internal(t_code), _skips_internal(skip), groups)


class DecoderSST(nn.Module):
    """conv.py:399-426."""

    def __init__(self, in_c, out_c, out_f):
        super().__init__()
        self.conv1 = _sst_stage((in_c, 256, 256, 128), up=True)
        self.conv2 = _sst_stage((128, 128, 128, 64), up=True)
        self.conv3 = _sst_stage((64, 64, out_c))
        self.out_f = activation_factory(out_f)
        self._out_act = activation_name(out_f)

    def decode(self, s_code, t_code, skip=None, groups=1):
        x = self.conv1(ops.concat_channels(s_code, t_code), groups)
        x = self.conv3(self.conv2(x, groups), groups)
        return _standalone_act(x, self._out_act)

    def decode_external(self, s_code, t_code, skip=None, groups=1):
        return ops.to_external(self.decode(s_code, t_code, None, groups))

    def forward(self, s_code, t_code, skip=None, groups=1):
        return self.decode_external(ops.to_internal(s_code), ops.to_internal(t_code), None, groups)


def _standalone_act(x, act):
    """An output activation that follows a BN+LeakyReLU block (SST decoders; identity in every
    shipped configuration, main.py:86-89)."""
    if act is None:
        return x
    raise NotImplementedError('a non-identity output activation after a BatchNorm block is not used by any '
                              'reference configuration (sst: last_activation=None)')

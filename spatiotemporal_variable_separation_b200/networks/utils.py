"""Activation lookup, ConstantS and weight initialisation.

Counterpart of /root/reference/var_sep/networks/utils.py.  Initialisation stays host-side PyTorch
(it is not on the timed path); activations are never standalone modules at run time — they are
fused into the epilogue of the producing kernel — but the modules are still instantiated so the
container hierarchy (and hence every ``state_dict`` key) matches the reference.
"""
import torch
import torch.nn as nn

_ACTIVATIONS = {
    'relu': lambda: nn.ReLU(inplace=True),
    'leaky_relu': lambda: nn.LeakyReLU(0.2, inplace=True),
    'elu': lambda: nn.ELU(inplace=True),
    'sigmoid': nn.Sigmoid,
    'tanh': nn.Tanh,
    None: nn.Identity,
    'identity': nn.Identity,
}


def activation_factory(name):
    """utils.py:50-72 — same names, same error for unknown ones."""
    try:
        return _ACTIVATIONS[name]()
    except KeyError:
        raise ValueError(f'Activation function `{name}` not yet implemented') from None


def activation_name(name):
    """Canonical name handed to the kernels (None for identity)."""
    if name in (None, 'none', 'identity'):
        return None
    if name not in _ACTIVATIONS:
        raise ValueError(f'Activation function `{name}` not yet implemented')
    return name


class ConstantS(nn.Module):
    """--no_s: the content code is constant (utils.py:21-29)."""

    def __init__(self, return_value=1, code_size=1):
        super().__init__()
        self.code_size = code_size
        self.return_value = return_value

    def forward(self, x, return_skip=False):
        return torch.full((len(x), self.code_size), float(self.return_value), device=x.device, dtype=torch.float32)

    def encode(self, h, groups=1, return_skip=False):
        from .. import ops
        return torch.full((h.shape[0], 1, 1, self.code_size), float(self.return_value), device=h.device,
                          dtype=ops.compute_dtype())


_WEIGHT_INIT = {
    'normal': lambda w, gain: nn.init.normal_(w, 0.0, gain),
    'xavier': lambda w, gain: nn.init.xavier_normal_(w, gain=gain),
    'kaiming': lambda w, gain: nn.init.kaiming_normal_(w, a=0, mode='fan_in'),
    'orthogonal': lambda w, gain: nn.init.orthogonal_(w, gain=gain),
}


def init_net(net, init_type='normal', init_gain=0.02):
    """utils.py:75-109: conv / linear weights by ``init_type``, biases 0, BatchNorm gamma ~ N(1, gain), beta 0."""
    if init_type not in _WEIGHT_INIT:
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
    for m in net.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            _WEIGHT_INIT[init_type](m.weight.data, init_gain)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif isinstance(m, nn.BatchNorm2d):
            if m.weight is not None:
                nn.init.normal_(m.weight.data, 1.0, init_gain)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)

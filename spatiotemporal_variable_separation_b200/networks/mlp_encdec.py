"""MLP encoder / decoder (counterpart of /root/reference/var_sep/networks/mlp_encdec.py)."""
import numpy as np
import torch.nn as nn

from .. import ops
from .mlp import MLP
from .utils import activation_factory, activation_name


class MLPEncoder(nn.Module):
    def __init__(self, input_size, hidden_size, output_size, nlayers):
        super().__init__()
        self.mlp = MLP(input_size, hidden_size, output_size, nlayers)

    def encode(self, h, groups=1, return_skip=False):
        # h is the time-folded NHWC frame stack; the reference flattens (t,c,h,w), so go back to that order
        N, H, W, Cc = h.shape
        if H * W > 1 and Cc > 1:
            h = h.permute(0, 3, 1, 2).contiguous()
        return self.mlp.run(h.reshape(N, 1, 1, -1), groups)

    def forward(self, x, return_skip=False):
        h = ops.to_internal(x.reshape(len(x), -1))                    # mlp_encdec.py:31
        return ops.to_external(self.mlp.run(h)).view(len(x), -1)


class MLPDecoder(nn.Module):
    def __init__(self, latent_size, hidden_size, output_shape, nlayers, last_activation, mixing):
        super().__init__()
        self.output_shape = output_shape
        self.mixing = mixing
        self.mlp = MLP(latent_size, hidden_size, int(np.prod(np.array(output_shape))), nlayers)
        self.last_activation = activation_factory(last_activation)
        self._last_act = activation_name(last_activation)

    def decode(self, z1, z2, skip=None, groups=1):
        """-> [N,1,1,C*H*W]: one (c,h,w)-ordered frame per row (mlp_encdec.py:43-50)."""
        z = ops.concat_channels(z1, z2) if self.mixing == 'concat' else ops.mul_bcast(z1, z2)
        return self.mlp.run(z, groups, self._last_act)

    def decode_external(self, z1, z2, skip=None, groups=1):
        x = self.decode(z1, z2, None, groups)
        return ops.to_external(x).view([-1] + list(self.output_shape))

    def forward(self, z1, z2, skip=None, groups=1):
        return self.decode_external(ops.to_internal(z1), ops.to_internal(z2), None, groups)

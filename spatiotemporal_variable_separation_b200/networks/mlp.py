"""Pre-activation MLP (counterpart of /root/reference/var_sep/networks/mlp.py).

Layer 0 is a bare Linear, layer i>0 is activation -> Linear (mlp.py:24-41).  At run time the
activation that *precedes* layer i+1 is fused into the epilogue of layer i's GEMM, which is the
same arithmetic because nothing else reads the pre-activation value.
"""
import torch.nn as nn

from .. import ops
from .utils import activation_factory, activation_name


def make_lin_block(ninp, nout, activation):
    modules = []
    if activation != 'none':
        modules.append(activation_factory(activation))
    modules.append(nn.Linear(ninp, nout))
    return nn.Sequential(*modules)


class MLP(nn.Module):
    def __init__(self, ninp, nhid, nout, nlayers, activation='relu'):
        super().__init__()
        assert nhid == 0 or nlayers > 1
        self.module = nn.Sequential(*[
            make_lin_block(ninp if il == 0 else nhid, nout if il == nlayers - 1 else nhid,
                           activation if il > 0 else 'none') for il in range(nlayers)])
        self._act = activation_name(activation)

    def run(self, h, groups=1, last_act=None):
        """h: NHWC [N,1,1,ninp] -> [N,1,1,nout]; ``last_act`` is fused after the last layer."""
        n = len(self.module)
        for il, blk in enumerate(self.module):
            h = ops.conv_block(h, blk[-1], None, self._act if il < n - 1 else last_act, 'conv', groups)
        return h

    def forward(self, x):
        return ops.to_external(self.run(ops.to_internal(x))).view(len(x), -1)

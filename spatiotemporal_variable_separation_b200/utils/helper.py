"""Checkpoint helpers (counterpart of /root/reference/var_sep/utils/helper.py:22-33).

The reference pickles whole nn.Modules; here the four networks are stored as ``state_dict``s
(identical keys and shapes, so they load into the reference modules and vice versa) under the
reference's file names."""
import os

import torch

_FILES = (('Et', 'ov_Et'), ('Es', 'ov_Es'), ('decoder', 'decoder'), ('t_resnet', 't_resnet'))


def save(elem_xp_path, sep_net, epoch_number=None):
    suffix = f'_{epoch_number}' if epoch_number is not None else ''
    for attr, name in _FILES:
        sd = {k: v.detach().cpu().clone() for k, v in getattr(sep_net, attr).state_dict().items()}
        torch.save(sd, os.path.join(elem_xp_path, f'{name}{suffix}.pt'))


def load(elem_xp_path, sep_net, epoch_number=None):
    suffix = f'_{epoch_number}' if epoch_number is not None else ''
    for attr, name in _FILES:
        sd = torch.load(os.path.join(elem_xp_path, f'{name}{suffix}.pt'), map_location='cpu')
        getattr(sep_net, attr).load_state_dict(sd)
    return sep_net

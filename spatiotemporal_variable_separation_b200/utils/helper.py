"""Checkpoint helpers (counterpart of /root/reference/var_sep/utils/helper.py:22-33,54-78).

The reference pickles whole nn.Modules under ``ov_Et.pt / ov_Es.pt / decoder.pt / t_resnet.pt`` (``save``) and
unpickles them in ``test/utils.py:8-16``.  Here the four networks are stored as ``state_dict``s under the same file
names (identical keys and shapes, so they load into the reference's modules with ``load_state_dict`` and vice versa),
and ``load`` accepts BOTH formats: a file written by the reference (a pickled module; needs the reference package
importable for unpickling) is reduced to its ``state_dict``.  ``save_training_state`` / ``load_training_state`` add what
the reference never saves: the optimizer moments and step count (in ``torch.optim.Adam``'s own format) and the
learning-rate schedule, so that training can resume bit-exactly."""
import json
import os

import torch

_FILES = (('Et', 'ov_Et'), ('Es', 'ov_Es'), ('decoder', 'decoder'), ('t_resnet', 't_resnet'))


def save(elem_xp_path, sep_net, epoch_number=None):
    suffix = f'_{epoch_number}' if epoch_number is not None else ''
    for attr, name in _FILES:
        sd = {k: v.detach().cpu().clone() for k, v in getattr(sep_net, attr).state_dict().items()}
        torch.save(sd, os.path.join(elem_xp_path, f'{name}{suffix}.pt'))


def _state_dict_of(obj, path):
    if isinstance(obj, torch.nn.Module):                 # written by the reference: torch.save(module)
        return obj.state_dict()
    if isinstance(obj, dict):
        return obj
    raise TypeError(f'{path}: expected a state_dict or a pickled nn.Module, got {type(obj).__name__}')


def load(elem_xp_path, sep_net, epoch_number=None):
    suffix = f'_{epoch_number}' if epoch_number is not None else ''
    for attr, name in _FILES:
        path = os.path.join(elem_xp_path, f'{name}{suffix}.pt')
        try:
            obj = torch.load(path, map_location='cpu', weights_only=True)
        except Exception:
            # a module pickled by the reference (torch < 2.6 default); its classes must be importable (var_sep.networks.*)
            obj = torch.load(path, map_location='cpu', weights_only=False)
        getattr(sep_net, attr).load_state_dict(_state_dict_of(obj, path))
    return sep_net


def save_training_state(elem_xp_path, optimizer, scheduler=None, epoch=None):
    state = {'optimizer': optimizer.state_dict(), 'scheduler': scheduler.state_dict() if scheduler is not None else None,
             'epoch': epoch}
    torch.save(state, os.path.join(elem_xp_path, 'training_state.pt'))


def load_training_state(elem_xp_path, optimizer, scheduler=None):
    state = torch.load(os.path.join(elem_xp_path, 'training_state.pt'), map_location='cpu', weights_only=False)
    optimizer.load_state_dict(state['optimizer'])
    if scheduler is not None and state.get('scheduler') is not None:
        scheduler.load_state_dict(state['scheduler'])
    return state.get('epoch')


class DotDict(dict):
    """Dictionary whose entries can also be read, set and deleted as attributes; a missing key reads as None
    (same behaviour as the class of that name in helper.py:54-60)."""

    def __getattr__(self, key):
        return self.get(key)

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]


def load_json(path):
    with open(path, 'r') as f:
        return DotDict(json.load(f))


def load_yaml(path):
    import yaml
    with open(path, 'r') as f:
        return DotDict(yaml.safe_load(f))

"""Checkpoint helpers (counterpart of /root/reference/var_sep/utils/helper.py:22-33,54-78).

``save`` writes what the reference writes: the four networks as whole pickled ``nn.Module``s under
``ov_Et.pt / ov_Es.pt / decoder.pt / t_resnet.pt`` (helper.py:22-33), so the reference's evaluation scripts, which do
``torch.load(path).to(device)`` (test/utils.py:8-16), obtain working modules — of THIS package's classes, i.e. they run
on the CUDA kernels; the package must be importable where they are unpickled.  The pickles hold clean CPU copies
(no optimizer-arena views, no packed-weight caches).  ``save_state_dicts`` / ``load_state_dicts`` move plain
``state_dict``s (identical keys and shapes on both sides) under DISTINCT file names (``*.state_dict.pt``) for exchange
with the reference's own classes.

``load`` accepts a pickled module (this package's classes, or the reference's when ``var_sep`` is importable) or a
``state_dict``.  It unpickles with ``weights_only=True`` and an allow-list of exactly those module classes; the
unrestricted unpickler (arbitrary code execution on a hostile file) is used only on request
(``allow_pickled_modules=True``) and only after the restricted one refused the file.

``save_training_state`` / ``load_training_state`` add what the reference never saves: the optimizer moments and step
count (in ``torch.optim.Adam``'s own format) and the learning-rate schedule, so that training can resume bit-exactly."""
import copy
import inspect
import json
import os
import pickle

import torch

_FILES = (('Et', 'ov_Et'), ('Es', 'ov_Es'), ('decoder', 'decoder'), ('t_resnet', 't_resnet'))


def _clean_cpu_copy(module):
    """Deep copy on the CPU: ``Parameter.__deepcopy__`` clones the data (compact storage instead of a view of the
    optimizer's flat arena) and drops the per-parameter attributes (gradient views, packed-weight caches)."""
    return copy.deepcopy(module).cpu()


def save(elem_xp_path, sep_net, epoch_number=None):
    """helper.py:22-33 — whole modules, the reference's file names."""
    suffix = f'_{epoch_number}' if epoch_number is not None else ''
    for attr, name in _FILES:
        torch.save(_clean_cpu_copy(getattr(sep_net, attr)), os.path.join(elem_xp_path, f'{name}{suffix}.pt'))


def save_state_dicts(elem_xp_path, sep_net, epoch_number=None):
    """Plain ``state_dict``s (``<name>.state_dict.pt``): load into the reference's modules with ``load_state_dict``."""
    suffix = f'_{epoch_number}' if epoch_number is not None else ''
    for attr, name in _FILES:
        sd = {k: v.detach().cpu().clone() for k, v in getattr(sep_net, attr).state_dict().items()}
        torch.save(sd, os.path.join(elem_xp_path, f'{name}{suffix}.state_dict.pt'))


def save_params(elem_xp_path, args):
    """main.py:105-106 — the experiment's flags as ``params.json`` (what ``test/utils.load_model`` rebuilds from)."""
    d = dict(args) if isinstance(args, dict) else dict(vars(args))
    with open(os.path.join(elem_xp_path, 'params.json'), 'w') as f:
        json.dump({k: v for k, v in d.items() if isinstance(v, (int, float, str, bool, list, tuple, type(None)))},
                  f, indent=4, sort_keys=True)


def _module_allow_list():
    """Classes a checkpoint written by ``save`` (ours or the reference's) may contain: ``torch.nn`` layers, this
    package's network classes and, when importable, the reference's."""
    import importlib
    allow = {getattr, setattr, set, frozenset}
    packages = ['torch.nn.modules.' + m for m in ('activation', 'batchnorm', 'container', 'conv', 'flatten', 'linear',
                                                  'pooling', 'upsampling', 'padding', 'dropout', 'normalization')]
    here = __name__.rsplit('.', 2)[0]
    packages += [f'{here}.networks.{m}' for m in ('conv', 'mlp', 'mlp_encdec', 'resnet', 'utils', 'model')]
    packages += [f'var_sep.networks.{m}' for m in ('conv', 'mlp', 'mlp_encdec', 'resnet', 'utils', 'model')]
    for name in packages:
        try:
            mod = importlib.import_module(name)
        except Exception:                     # the reference package is optional
            continue
        for _, obj in inspect.getmembers(mod, inspect.isclass):
            if issubclass(obj, torch.nn.Module):
                allow.add(obj)
    return list(allow)


def _state_dict_of(obj, path):
    if isinstance(obj, torch.nn.Module):                 # torch.save(module): the reference's format and ours
        return obj.state_dict()
    if isinstance(obj, dict):
        return obj
    raise TypeError(f'{path}: expected a state_dict or a pickled nn.Module, got {type(obj).__name__}')


def _restricted_load(path, allow_pickled_modules=False):
    try:
        with torch.serialization.safe_globals(_module_allow_list()):
            return torch.load(path, map_location='cpu', weights_only=True)
    except pickle.UnpicklingError as e:
        if not allow_pickled_modules:
            raise pickle.UnpicklingError(
                f'{path}: refused by the restricted unpickler ({str(e).splitlines()[0]}).  If the file comes from a '
                'trusted source and pickles classes outside torch.nn / this package / var_sep.networks, pass '
                'allow_pickled_modules=True (arbitrary code execution on a hostile file).') from e
        return torch.load(path, map_location='cpu', weights_only=False)


def load(elem_xp_path, sep_net, epoch_number=None, allow_pickled_modules=False):
    suffix = f'_{epoch_number}' if epoch_number is not None else ''
    for attr, name in _FILES:
        path = os.path.join(elem_xp_path, f'{name}{suffix}.pt')
        if not os.path.exists(path) and os.path.exists(path[:-3] + '.state_dict.pt'):
            path = path[:-3] + '.state_dict.pt'
        obj = _restricted_load(path, allow_pickled_modules)
        getattr(sep_net, attr).load_state_dict(_state_dict_of(obj, path))
    _after_external_write(sep_net)
    return sep_net


def _after_external_write(sep_net):
    """Parameters were overwritten in place: packed / padded / folded copies derived from them are stale."""
    from .. import ops
    ops.invalidate_params(list(sep_net.parameters()))


def save_training_state(elem_xp_path, optimizer, scheduler=None, epoch=None):
    state = {'optimizer': optimizer.state_dict(), 'scheduler': scheduler.state_dict() if scheduler is not None else None,
             'epoch': epoch}
    torch.save(state, os.path.join(elem_xp_path, 'training_state.pt'))


def load_training_state(elem_xp_path, optimizer, scheduler=None):
    # tensors and primitives only: the restricted unpickler suffices
    state = torch.load(os.path.join(elem_xp_path, 'training_state.pt'), map_location='cpu', weights_only=True)
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

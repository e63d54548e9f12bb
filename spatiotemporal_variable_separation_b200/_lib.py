"""ctypes binding of libvarsep_sm100a.so (the C ABI declared in include/varsep.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.  PyTorch supplies device memory and the
current CUDA stream only.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvarsep_sm100a.so')

VS_F32, VS_BF16 = 0, 1
ACT = {None: 0, 'none': 0, 'identity': 0, 'relu': 1, 'leaky_relu': 2, 'elu': 3, 'sigmoid': 4, 'tanh': 5}
DIRECT, TRANSPOSED = 0, 1
FLAG_FORCE_SIMT = 1
VS_PACK_BLOCK_ELEMS = 4096          # include/varsep.h


class Geom(C.Structure):
    """vs_conv_geom"""
    _fields_ = [(n, C.c_int32) for n in ('dtype', 'N', 'H', 'W', 'C', 'P', 'Q', 'K', 'R', 'S', 'stride', 'pad',
                                          'groups', 'act', 'flags')]


_p, _i32, _i64, _f, _d = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
_PG = C.POINTER(Geom)
_PD = C.POINTER(C.c_double)
_PP = C.POINTER(C.c_void_p)

# name -> argtypes (all return int status unless listed in _RET)
SIGNATURES = {
    'vs_abi_version': [],
    'vs_last_error': [],
    'vs_launch_count': [],
    'vs_pack_weight': [_p, _p, _i32, _i32, _i32, _i32, _i32, _p],
    'vs_pack_weights_multi': [_p, _i32, _i32, _p],
    'vs_conv_forward': [_PG, _i32, _p, _p, _p, _p, _p, _p],
    'vs_conv_forward_path': [_PG, _i32],
    'vs_conv_forward_variant': [_PG, _i32],
    'vs_conv_wgrad': [_PG, _p, _p, _p, _p],
    'vs_colsum': [_p, _i32, _i64, _i32, _p, _p],
    'vs_bn_finalize': [_p, _i32, _i32, _i64, _f, _f, _p, _p, _p, _p, _p, _p],
    'vs_bn_finalize_act_forward': [_p, _i32, _i32, _i64, _f, _f, _p, _p, _p, _p, _p, _p, _p, _i32, _i64, _p, _p, _i32, _p],
    'vs_bn_eval_stats': [_p, _p, _i32, _f, _p, _p, _p],
    'vs_bn_act_forward': [_p, _p, _i32, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _p],
    'vs_bn_act_backward_reduce': [_p, _p, _i32, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p],
    'vs_bn_act_backward_apply': [_p, _p, _p, _i32, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _p, _i32, _p, _p, _p],
    'vs_act_backward': [_p, _p, _p, _i32, _i64, _i32, _p],
    'vs_add_act': [_p, _p, _p, _i32, _i64, _i32, _p],
    'vs_copy_channels': [_p, _i32, _i64, _p, _i32, _i32, _i64, _i32, _p],
    'vs_slice_channels_reduce': [_p, _i32, _i32, _i64, _p, _i32, _i64, _i32, _p],
    'vs_mul_bcast': [_p, _i64, _p, _p, _i64, _i32, _i32, _p],
    'vs_mul_bcast_backward': [_p, _p, _i64, _p, _p, _p, _i64, _i32, _i32, _p],
    'vs_maxpool_forward': [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    'vs_maxpool_backward': [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p],
    'vs_upsample2_forward': [_p, _p, _i32, _i32, _i32, _i32, _i32, _p],
    'vs_upsample2_backward': [_p, _p, _i32, _i32, _i32, _i32, _i32, _p],
    'vs_frames_to_nhwc': [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p],
    'vs_nhwc_to_nchw': [_p, _i32, _p, _i32, _i32, _i32, _i32, _p],
    'vs_nchw_to_nhwc': [_p, _p, _i32, _i32, _i32, _i32, _i32, _p],
    'vs_sqdiff_sum': [_p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _p, _p],
    'vs_sqdiff_backward': [_p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _f, _p, _p, _f, _p, _i32, _p],
    'vs_loss_combine': [_p, _PD, _PD, _i32, _p, _p],
    'vs_ssim_mse_planes': [_p, _p, _i64, _i32, _i32, _p, _i32, _f, _f, _p, _p, _p, _p],
    'vs_moving_sequences': [_p, _i32, _i32, _i32, _p, _i32, _i32, _i32, _i32, _p, _p],
    'vs_peer_barrier': [_PP, _i32, _i32, _p, _p],
    'vs_peer_allreduce': [_PP, _i32, _i32, _i64, _i32, _p],
    'vs_peer_allreduce_bf16': [_PP, _i32, _i32, _i64, _i32, _p],
    'vs_grad_compress': [_p, _p, _i64, _p],
    'vs_grad_expand': [_p, _p, _i64, _p],
    'vs_adam_step': [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i32, _p, _p, _p],
    'vs_tail_eligible': [_PG],
    'vs_tail_forward': [_PG, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p],
    'vs_tail_wgrad': [_PG, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p],
    'vs_tail_bn_backward': [_PG, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _i32, _i32, _p, _p, _p, _p, _p],
    'vs_latent_rollout_forward': [_p, _PP, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p],
    'vs_latent_rollout_backward': [_p, _PP, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p],
}
_RET = {'vs_last_error': C.c_char_p, 'vs_launch_count': C.c_int64}

_lib = None


def load():
    """Load the library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: build it with '
                               '`python -m spatiotemporal_variable_separation_b200.csrc.build` '
                               '(there is no CPU or PyTorch fallback)')
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = _RET.get(name, C.c_int)
        _lib = lib
    return _lib


class _CurrentStream:
    """Marker returned by ``stream()``: resolved inside ``call`` to the current stream OF THE DEVICE THE TENSOR ARGUMENTS
    LIVE ON (which need not be the process's current device: ``train(device='cuda:1')`` without ``set_device``)."""


_STREAM = _CurrentStream()


def call(name, *args):
    """Invoke an entry point.  torch.Tensor arguments are passed as their device pointer (views keep
    their offset); everything else goes through ctypes unchanged.  The launch goes to the device of the tensor
    arguments — all of them must live on the same one — on that device's current stream."""
    lib = load()
    dev = None
    for a in args:
        if isinstance(a, torch.Tensor):
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise RuntimeError(f'{name}: tensor arguments on different devices ({dev} and {a.device})')
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            conv.append(C.c_void_p(a.data_ptr()))
        elif a is _STREAM:
            conv.append(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        else:
            conv.append(a)
    if dev is not None and dev.type == 'cuda' and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            rc = getattr(lib, name)(*conv)
    else:
        rc = getattr(lib, name)(*conv)
    if rc != 0:
        raise RuntimeError(f'{name} failed: {lib.vs_last_error().decode()}')


def launch_count():
    return int(load().vs_launch_count())


def stream():
    """The stream argument of an entry point (see ``call``)."""
    return _STREAM


def pointer_array(tensors):
    """Host array of device pointers (for entry points that take `const float* const*`)."""
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def ptr(t):
    """Marks an argument as a device pointer (converted in ``call``); None stays NULL."""
    return t


def dtype_code(t):
    if t.dtype == torch.float32:
        return VS_F32
    if t.dtype == torch.bfloat16:
        return VS_BF16
    raise TypeError(f'unsupported dtype {t.dtype}')


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('this operator runs on a CUDA device only (B200, sm_100a); '
                               'got a CPU tensor and there is no CPU fallback')

"""Deterministic, name-keyed weights and inputs.  TEST INFRASTRUCTURE ONLY.

Weights are a pure function of (tensor name, shape, salt), independent of the
construction order of any module, so the reference (in ``gen_golden.py``), the
oracle and the CUDA modules can each materialise the *same* parameters
without shipping tens of MB of fixtures.  Values come from a splitmix-style
integer hash turned into uniform/normal variates with numpy only — no torch
RNG, hence no dependence on the torch version or device.
"""
import zlib

import numpy as np
import torch


def _uniform(name, n, salt=0):
    """n doubles in (0,1), from a counter-based hash seeded by crc32(name)."""
    seed = np.uint64(zlib.crc32(name.encode()) + 0x9E3779B97F4A7C15 * (salt + 1) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over='ignore'):
        z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + seed
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)


def normal(name, shape, salt=0):
    n = int(np.prod(shape)) if len(shape) else 1
    m = (n + 1) // 2
    u1, u2 = _uniform(name + '#a', m, salt), _uniform(name + '#b', m, salt)
    r = np.sqrt(-2.0 * np.log(u1))
    z = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)])[:n]
    return z.reshape(shape)


def uniform(name, shape, salt=0):
    n = int(np.prod(shape)) if len(shape) else 1
    return _uniform(name, n, salt).reshape(shape)


def tensor_for(name, shape, salt=0, dtype=torch.float32):
    """Value policy by parameter kind (well-conditioned, every path exercised):
    weights ~ N(0, 1/max(fan_in, fan_out)); BN gamma ~ 1 + 0.1 N; biases/beta ~ 0.1 N;
    running_mean ~ 0.1 N; running_var ~ 1 + 0.2 U; num_batches_tracked = 0."""
    shape = tuple(shape)
    if name.endswith('num_batches_tracked'):
        return torch.zeros((), dtype=torch.int64)
    if name.endswith('running_var'):
        v = 1.0 + 0.2 * uniform(name, shape, salt)
    elif name.endswith('running_mean'):
        v = 0.1 * normal(name, shape, salt)
    elif len(shape) >= 2:
        # max(fan_in, fan_out): same scale whichever of dim 0/1 is the input (Conv vs ConvTranspose)
        fan = max(int(np.prod(shape[1:])), shape[0] * int(np.prod(shape[2:])))
        v = normal(name, shape, salt) / np.sqrt(fan)
    elif name.endswith('.weight'):           # 1-D weight = BatchNorm gamma
        v = 1.0 + 0.1 * normal(name, shape, salt)
    else:                                    # biases, BN beta
        v = 0.1 * normal(name, shape, salt)
    return torch.from_numpy(np.ascontiguousarray(v)).to(dtype)


def fill_state(shapes, prefix, salt=0, dtype=torch.float32):
    """{key: shape} -> {key: tensor}; ``prefix`` (e.g. 'Es.') separates the four nets."""
    return {k: tensor_for(prefix + k, s, salt, dtype) for k, s in shapes.items()}


@torch.no_grad()
def fill_module_(module, prefix, salt=0):
    """Overwrite every entry of ``module.state_dict()`` in place with the same values."""
    for k, v in module.state_dict().items():
        v.copy_(tensor_for(prefix + k, tuple(v.shape), salt).to(v.dtype))
    return module


def frames(name, batch, nt, shape, salt=0, kind='blobs'):
    """Synthetic sequences [B, nt, C, H, W] in [0,1] (float32).

    'blobs': two Gaussian bumps per sequence moving with constant velocity and
    bouncing on the walls (Moving-MNIST-shaped statistics: sparse, in [0,1]);
    'uniform': U[0,1]; 'normal': N(0,1) (SST anomalies)."""
    C, H, W = shape
    if kind == 'uniform':
        return torch.from_numpy(uniform(name, (batch, nt, C, H, W), salt)).float()
    if kind == 'normal':
        return torch.from_numpy(normal(name, (batch, nt, C, H, W), salt)).float()
    u = uniform(name, (batch, 2, 4), salt)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.zeros((batch, nt, C, H, W))
    for b in range(batch):
        for o in range(2):
            py, px = u[b, o, 0] * (H - 1), u[b, o, 1] * (W - 1)
            vy, vx = (u[b, o, 2] - 0.5) * 8, (u[b, o, 3] - 0.5) * 8
            for t in range(nt):
                out[b, t] += np.exp(-((yy - py) ** 2 + (xx - px) ** 2) / (2 * (0.07 * H) ** 2))[None]
                py, px = py + vy, px + vx
                if py < 0 or py > H - 1:
                    vy, py = -vy, min(max(py, 0), H - 1)
                if px < 0 or px > W - 1:
                    vx, px = -vx, min(max(px, 0), W - 1)
    return torch.from_numpy(np.clip(out, 0, 1)).float()

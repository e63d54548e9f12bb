"""Tensor summaries shared by the golden generator and the tests.

A tensor is summarised by its L2 norm and its dot product with a fixed
name-keyed probe vector: the pair pins magnitude *and* element order (a
transposed or permuted tensor keeps the norm but not the probe dot) in 16 bytes.
"""
import numpy as np
import torch

from oracle import detfill


def probe_dot(name, t):
    t = t.detach().double().cpu().reshape(-1)
    p = torch.from_numpy(detfill.normal('probe:' + name, (t.numel(),)))
    return float((t * p).sum())


def summarize(name, t):
    return np.array([float(t.detach().double().norm()), probe_dot(name, t)])


def subsample(t, n=4096):
    """Deterministic strided subsample of a tensor (flat), at most ~n elements."""
    f = t.detach().double().cpu().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].numpy().copy()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

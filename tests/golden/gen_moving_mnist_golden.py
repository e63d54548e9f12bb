"""Golden Moving-MNIST training samples from the UNMODIFIED reference generator (build container only).

    python tests/golden/gen_moving_mnist_golden.py

Instantiates /root/reference/var_sep/data/moving_mnist.py::MovingMNIST directly on a bank of procedural uint8 glyphs
(MNIST itself cannot be downloaded here; the generator never looks at what the glyphs depict), deterministic variant,
seeds numpy's global RNG and records consecutive ``__getitem__`` results.  Stored: the glyph bank, the seed, and the
frames as uint8 (frame * 255 is integral by construction).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from var_sep.data.moving_mnist import MovingMNIST  # noqa: E402

SEED, N_SAMPLES, N_GLYPHS = 20260, 24, 40
rng = np.random.RandomState(7)
yy, xx = np.mgrid[0:28, 0:28]
glyphs = []
for _ in range(N_GLYPHS):
    img = np.zeros((28, 28))
    for _ in range(3):
        cy, cx, s = rng.uniform(6, 22), rng.uniform(6, 22), rng.uniform(1.5, 4.0)
        img = np.maximum(img, np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s)))
    glyphs.append((255 * img).astype(np.uint8))
ds = MovingMNIST(glyphs, 64, 5, 15, 4, True, 2, True)
np.random.seed(SEED)
frames = []
for i in range(N_SAMPLES):
    cond, target = ds[i]
    x = np.concatenate([cond.numpy(), target.numpy()], 0) * 255
    assert np.abs(x - np.round(x)).max() < 1e-3
    frames.append(np.round(x).astype(np.uint8))
# the stochastic variant, for the oracle's speed re-draw branch (:229-231)
ds_s = MovingMNIST(glyphs, 64, 5, 15, 4, False, 2, True)
np.random.seed(SEED + 1)
frames_s = []
for i in range(8):
    cond, target = ds_s[i]
    frames_s.append(np.round(np.concatenate([cond.numpy(), target.numpy()], 0) * 255).astype(np.uint8))
np.savez_compressed(os.path.join(HERE, 'data', 'moving_mnist.npz'), glyphs=np.stack(glyphs), seed=SEED, frames=np.stack(frames),
                    frames_stochastic=np.stack(frames_s))
print('wrote', N_SAMPLES, 'deterministic +', len(frames_s), 'stochastic samples,', os.path.getsize(os.path.join(HERE, 'data', 'moving_mnist.npz')), 'bytes')

"""CPU restatement of the reference's Moving-MNIST training sample generator.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/var_sep/data/moving_mnist.py: ``sample`` = ``__getitem__`` (train branch, :112-130),
``trajectory`` = ``_compute_trajectory`` (:131-170) and ``collide`` = ``_process_collision`` (:172-253) with its two
intersection helpers (:255-301), in the reference's floating-point arithmetic, draw order and tolerance (eps = 1e-8,
:69).  Pinned to sequences recorded from the unmodified reference (tests/golden/data/moving_mnist.npz, written by
tests/golden/gen_moving_mnist_golden.py).  The CUDA generator (csrc/sequences.cu) is checked against this file.
"""
import numpy as np

EPS = 1e-8


def _cross_x(a, b, x_lim, y_lo, y_hi):
    """:255-277 — where the line y = a x + b meets the vertical border x = x_lim, and whether inside the frame."""
    y = a * x_lim + b
    return (y >= y_lo - EPS) and (y <= y_hi + EPS), (x_lim, y)


def _cross_y(a, b, y_lim, x_lo, x_hi):
    """:279-301 — same for a horizontal border."""
    x = (y_lim - b) / a
    return (x >= x_lo - EPS) and (x <= x_hi + EPS), (x, y_lim)


def collide(sx, sy, dx, dy, x_max, y_max, rng=None, max_speed=None):
    """:172-253.  ``rng`` is None for the deterministic data set; otherwise a new speed is drawn at every bounce
    (:229-231) from ``rng.randint(-max_speed, max_speed + 1)``."""
    x_min = y_min = 0
    left, up = sx < x_min - EPS, sy < y_min - EPS
    right, down = sx > x_max + EPS, sy > y_max + EPS
    while left or right or up or down:
        if dx == 0:
            cx, cy = (sx, y_min) if up else (sx, y_max)
        elif dy == 0:
            cx, cy = (x_min, sy) if left else (x_max, sy)
        else:
            a = dy / dx
            b = sy - a * sx
            if left:
                left, pt = _cross_x(a, b, x_min, y_min, y_max)
                if left:
                    cx, cy = pt
            if right:
                right, pt = _cross_x(a, b, x_max, y_min, y_max)
                if right:
                    cx, cy = pt
            if up:
                up, pt = _cross_y(a, b, y_min, x_min, x_max)
                if up:
                    cx, cy = pt
            if down:
                down, pt = _cross_y(a, b, y_max, x_min, x_max)
                if down:
                    cx, cy = pt
        frac = ((sx - cx) / dx) if dx != 0 else ((sy - cy) / dy)      # share of the time step left after the bounce
        if rng is not None:
            dx = rng.randint(-max_speed, max_speed + 1)
            dy = rng.randint(-max_speed, max_speed + 1)
        if left:
            dx = abs(dx)
        if right:
            dx = -abs(dx)
        if up:
            dy = abs(dy)
        if down:
            dy = -abs(dy)
        sx, sy = cx + dx * frac, cy + dy * frac
        left, up = sx < x_min - EPS, sy < y_min - EPS
        right, down = sx > x_max + EPS, sy > y_max + EPS
    return sx, sy, dx, dy


def trajectory(sx, sy, dx, dy, seq_len, x_max, y_max):
    """:156-170 for given initial conditions (deterministic): rounded positions of every frame."""
    out = []
    for _ in range(seq_len):
        sx, sy, dx, dy = collide(sx, sy, dx, dy, x_max, y_max)
        out.append((int(round(sx)), int(round(sy))))
        sy += dy
        sx += dx
    return out


def draw_objects(rng, n_glyphs, batch, num_digits, frame_size, gh, gw, max_speed):
    """The host draws of ``batch`` consecutive ``__getitem__`` calls in the reference's order (:119-120, :147-150):
    per sample and digit  glyph index, sx, sy, dx, dy.  -> int32 [batch, num_digits, 5]"""
    objs = np.empty((batch, num_digits, 5), dtype=np.int32)
    for b in range(batch):
        for n in range(num_digits):
            objs[b, n, 0] = rng.randint(n_glyphs)
            objs[b, n, 1] = rng.randint(0, frame_size - gh + 1)
            objs[b, n, 2] = rng.randint(0, frame_size - gw + 1)
            objs[b, n, 3] = rng.randint(-max_speed, max_speed + 1)
            objs[b, n, 4] = rng.randint(-max_speed, max_speed + 1)
    return objs


def render(glyphs, objs, seq_len, frame_size):
    """:116-129 — float32 [B, seq_len, 1, F, F] in [0, 1]."""
    B, n_obj, _ = objs.shape
    gh, gw = glyphs.shape[1:]
    out = np.zeros((B, seq_len, 1, frame_size, frame_size), dtype=np.float32)
    for b in range(B):
        for n in range(n_obj):
            g, sx, sy, dx, dy = (int(v) for v in objs[b, n])
            for t, (px, py) in enumerate(trajectory(sx, sy, dx, dy, seq_len, frame_size - gh, frame_size - gw)):
                out[b, t, 0, px:px + gh, py:py + gw] += glyphs[g]
    out[out > 255] = 255
    return out / 255


def sample(rng, glyphs, nt_cond, seq_len, frame_size, max_speed, num_digits, deterministic=True):
    """One ``__getitem__`` (:112-130) with every draw taken from ``rng`` in the reference's order — including, for the
    stochastic data set, the speed re-draws inside the collision loop.  -> (cond, target) float32 arrays."""
    gh, gw = glyphs.shape[1:]
    x = np.zeros((seq_len, 1, frame_size, frame_size), dtype=np.float32)
    for _ in range(num_digits):
        img = glyphs[rng.randint(len(glyphs))]
        x_max, y_max = frame_size - gh, frame_size - gw
        sx, sy = rng.randint(0, x_max + 1), rng.randint(0, y_max + 1)
        dx, dy = rng.randint(-max_speed, max_speed + 1), rng.randint(-max_speed, max_speed + 1)
        for t in range(seq_len):
            sx, sy, dx, dy = collide(sx, sy, dx, dy, x_max, y_max, None if deterministic else rng, max_speed)
            px, py = int(round(sx)), int(round(sy))
            x[t, 0, px:px + gh, py:py + gw] += img
            sy += dy
            sx += dx
    x[x > 255] = 255
    x = x / 255
    return x[:nt_cond], x[nt_cond:]

"""the Moving-MNIST training generator.

  * the oracle restatement (oracle/moving_mnist.py) reproduces, bit for bit, samples recorded from the unmodified
    reference generator for both the deterministic and the stochastic data set (tests/golden/data/moving_mnist.npz);
  * the host sampler of data.MovingSequences consumes numpy's RNG stream exactly like the reference (vectorised draws);
  * the CUDA kernel renders exactly what the reference renders (gpu), including objects that bounce several times,
    objects at rest, glyphs touching the borders and overlapping glyphs that saturate at 255.
"""
import os

import numpy as np
import pytest
import torch

from oracle import moving_mnist as omm
from spatiotemporal_variable_separation_b200 import data as vs_data

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'data', 'moving_mnist.npz'))
GLYPHS, SEED = G['glyphs'], int(G['seed'])


def test_oracle_reproduces_reference_samples_deterministic_and_stochastic():
    rng = np.random.RandomState(SEED)
    for want in G['frames']:
        cond, target = omm.sample(rng, GLYPHS, 5, 15, 64, 4, 2, deterministic=True)
        got = np.round(np.concatenate([cond, target], 0) * 255).astype(np.uint8)
        assert np.array_equal(got, want)
    rng = np.random.RandomState(SEED + 1)
    for want in G['frames_stochastic']:
        cond, target = omm.sample(rng, GLYPHS, 5, 15, 64, 4, 2, deterministic=False)
        assert np.array_equal(np.round(np.concatenate([cond, target], 0) * 255).astype(np.uint8), want)


def test_host_draws_follow_the_reference_rng_stream_and_render_matches():
    ms = vs_data.MovingSequences(GLYPHS, batch_size=len(G['frames']), device='cpu')
    np.random.seed(SEED)
    objs = ms.draw()
    rng = np.random.RandomState(SEED)
    assert np.array_equal(objs, omm.draw_objects(rng, len(GLYPHS), len(G['frames']), 2, 64, 28, 28, 4))
    frames = omm.render(GLYPHS, objs, 15, 64)
    assert np.array_equal(np.round(frames[:, :, 0] * 255).astype(np.uint8), G['frames'][:, :, 0])
    # the RNG state afterwards is where the reference leaves it
    assert np.random.randint(1 << 30) == rng.randint(1 << 30)


def test_triangle_wave_equals_the_collision_loop_for_every_start_and_speed():
    """csrc/sequences.cu folds s + d*t into [0, x_max]; the reference iterates its collision loop (incl. corners)."""
    x_max, T = 36, 40
    for s in (0, 1, 17, 35, 36):
        for d in range(-4, 5):
            ref = [p[0] for p in omm.trajectory(s, s, d, -d, T, x_max, x_max)]
            period = 2 * x_max
            fold = [(lambda m: period - m if m > x_max else m)((s + d * t) % period) for t in range(T)]
            assert ref == fold, (s, d)


@pytest.mark.gpu
def test_cuda_generator_matches_reference_bit_for_bit():
    ms = vs_data.MovingSequences(GLYPHS, batch_size=len(G['frames']), device='cuda')
    np.random.seed(SEED)
    cond, target = ms.batch()
    got = torch.cat([cond, target], 1).cpu().numpy()
    assert got.shape == (len(G['frames']), 15, 1, 64, 64)
    assert np.array_equal(np.round(got * 255).astype(np.uint8)[:, :, 0], G['frames'][:, :, 0])
    assert np.array_equal(got, G['frames'][:, :, None, 0].astype(np.float32).reshape(got.shape) / np.float32(255))
    # edge cases: corners, rest, long horizons (many bounces), saturation where the two glyphs overlap
    objs = np.array([[[0, 0, 0, -4, -4], [1, 36, 36, 4, 4]], [[2, 10, 20, 0, 0], [2, 10, 20, 0, 0]],
                     [[3, 36, 0, 4, -3], [4, 0, 36, -1, 2]]], dtype=np.int32)
    ms.seq_len = 95
    got = ms.render(objs).cpu().numpy()
    want = omm.render(GLYPHS, objs, 95, 64)
    assert np.array_equal(got, want)
    assert got[1].max() == 1.0                                    # two copies of one glyph saturate


@pytest.mark.gpu
def test_generator_feeds_the_training_loop():
    """One epoch of train() straight from the device-side generator (no host frames at all)."""
    import tempfile
    from spatiotemporal_variable_separation_b200 import configs, ops, train as vs_train
    from spatiotemporal_variable_separation_b200.networks.factory import build_model
    from spatiotemporal_variable_separation_b200.optim import FusedAdam
    cfg = configs.preset('mnist', small=True)
    net = build_model(cfg, 'cuda')
    opt = FusedAdam(net.parameters(), cfg['lr'], (cfg['beta1'], cfg['beta2']))
    before = opt.flat_p.clone()
    loader = vs_data.MovingSequences(GLYPHS, nt_cond=cfg['nt_cond'], seq_len=cfg['nt_cond'] + cfg['nt_pred'],
                                     batch_size=cfg['batch_size'], batches_per_epoch=4, device='cuda')
    np.random.seed(3)
    with tempfile.TemporaryDirectory() as d:
        vs_train.train(d, loader, 'cuda', net, opt, None, False, False, 1, cfg['lamb_ae'], cfg['lamb_s'], cfg['lamb_t'],
                       cfg['lamb_pred'], cfg['offset'], cfg['nt_cond'], cfg['nt_pred'], cfg['no_s'], cfg['skipco'], None, False)
        assert os.path.exists(os.path.join(d, 'decoder.pt'))
    assert int(opt.step_dev) == 4 and torch.isfinite(opt.flat_p).all() and not torch.equal(before, opt.flat_p)
    assert ops.compute_dtype() == torch.float32


@pytest.mark.gpu
def test_main_entry_point_writes_a_reloadable_experiment(tmp_path):
    """main.py counterpart: params.json + checkpoints of a (tiny) run, reloaded by test/utils.load_model."""
    from spatiotemporal_variable_separation_b200 import configs, main as vs_main
    from spatiotemporal_variable_separation_b200.test.utils import load_model
    import shlex
    argv = ['--xp_dir', str(tmp_path), '--data_dir', str(tmp_path), '--device', '0', '--epochs', '1', '--synthetic_batches', '3'] + \
        shlex.split(configs.README_FLAGS['mnist']) + shlex.split(configs.SMALL_FLAGS['mnist'])
    vs_main.main(argv)
    assert os.path.exists(tmp_path / 'params.json') and os.path.exists(tmp_path / 'ov_Es.pt')
    net = load_model({'xp_dir': str(tmp_path), 'device': 'cuda'})
    ms = vs_data.MovingSequences(GLYPHS, batch_size=4, device='cuda')
    cond, _ = ms.batch()
    with torch.no_grad():
        f = net.get_forecast(cond, 6)[0]
    assert f.shape == (4, 6, 1, 64, 64) and torch.isfinite(f).all()

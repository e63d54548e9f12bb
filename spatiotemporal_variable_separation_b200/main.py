"""Training entry point (counterpart of /root/reference/var_sep/main.py:50-175 for the hot path).

    python -m spatiotemporal_variable_separation_b200.main --xp_dir runs/mnist --data_dir . --data mnist --beta1 0.5 \
           --device 0 --torch_amp --epochs 1 [--synthetic_batches 200]

Same flags as the reference (options.py); the differences are what this package is about:
  * the experiment's flags are written to ``xp_dir/params.json`` exactly as main.py:105-106 does, so that
    ``test/utils.load_model`` (ours or the reference's) can rebuild the networks of a run trained here;
  * the networks are this package's modules on a B200, the optimizer is ``FusedAdam`` (+ this package's ``MultiStepLR``),
    the loop is ``train.train`` (CUDA-graph replay per step, batches prefetched to the device);
  * data: the dataset classes of the reference are out of scope (they need h5py / netCDF4 / torchdiffeq / downloads, none
    of which exist here).  ``--data mnist`` trains on the device-side Moving-MNIST generator (``data.MovingSequences``:
    the reference's bounce semantics and RNG stream) over a glyph bank — ``<data_dir>/glyphs.npy`` (uint8 [G, 28, 28], e.g.
    the MNIST training digits) if present, procedural strokes otherwise; every other ``--data`` value trains on synthetic
    batches of that data set's shape and value range (``data.synthetic_batch``).
Launch with torchrun for data-parallel training (one process per GPU; ``--device`` is then ignored in favour of LOCAL_RANK).
"""
import os

import numpy as np
import torch

from . import data as vs_data
from .networks.factory import build_model
from .optim import FusedAdam, MultiStepLR
from .options import config_from_args, parser
from .train import train
from .utils import helper


class _SyntheticLoader:
    """``batches`` device-resident batches of one configuration's shape per epoch."""

    def __init__(self, cfg, device, batches, seed=0):
        self.cfg, self.device, self.batches, self.seed = cfg, device, batches, seed

    def __len__(self):
        return self.batches

    def __iter__(self):
        nc = self.cfg['nt_cond']
        for i in range(self.batches):
            full = vs_data.synthetic_batch(self.cfg, device=self.device, seed=self.seed + i)
            yield full[:, :nc], full[:, nc:]


def main(argv=None):
    if not any(a.dest == 'synthetic_batches' for a in parser._actions):
        parser.add_argument('--synthetic_batches', type=int, default=None,
                            help='batches per epoch of the synthetic loaders (default: 200000 samples, as the reference)')
    args = parser.parse_args(argv)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0')) if world > 1 else (args.device if args.device is not None else 0)
    if not torch.cuda.is_available():
        raise RuntimeError('this package runs on a CUDA device (B200, sm_100a) only; there is no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    reducer = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    seed = np.random.randint(0, 10000) if world == 1 else 1234       # (every rank must draw the same t_random)
    torch.manual_seed(seed)
    np.random.seed(seed)
    cfg = config_from_args(args)
    os.makedirs(args.xp_dir, exist_ok=True)
    if int(os.environ.get('RANK', '0')) == 0:
        helper.save_params(args.xp_dir, {k: v for k, v in vars(args).items() if k != 'synthetic_batches'})
    n_batches = args.synthetic_batches
    if args.data == 'mnist':
        path = os.path.join(args.data_dir, 'glyphs.npy')
        glyphs = np.load(path) if os.path.exists(path) else vs_data.procedural_glyphs(1024)
        loader = vs_data.MovingSequences(glyphs, 64, args.nt_cond, args.nt_cond + args.nt_pred, 4, args.n_object,
                                         args.batch_size, n_batches, device)
    else:
        loader = _SyntheticLoader(cfg, device, n_batches or 200000 // args.batch_size, seed)
    sep_net = build_model(cfg, device)
    if world > 1:
        from .parallel import GradReducer, broadcast_model
        broadcast_model(sep_net)
    optimizer = FusedAdam(sep_net.parameters(), lr=args.lr, betas=(args.beta1, args.beta2))
    if world > 1:
        reducer = GradReducer(sep_net, optimizer)
    scheduler = MultiStepLR(optimizer, args.scheduler_milestones, gamma=args.scheduler_decay) if args.scheduler else None
    train(args.xp_dir, loader, device, sep_net, optimizer, scheduler, args.apex_amp, args.torch_amp, args.epochs,
          args.lamb_ae, args.lamb_s, args.lamb_t, args.lamb_pred, args.offset, args.nt_cond, args.nt_pred, args.no_s,
          args.skipco, args.chkpt_interval, args.architecture == 'encoderSST', reducer=reducer)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

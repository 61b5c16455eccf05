"""Throughput of the AutoAttack-compatible evaluation (BASELINE config 5): APGD-CE + APGD-T (9 targets), 100
iterations, l-inf 4/255, ConvNeXt-L-CvSt at 320x320, random-init weights, synthetic points, one GPU's share.

    python profiles/aa_bench.py [--arch convnext_large] [--res 320] [--n 100] [--bs 100] [--iters 100] [--targets 9]

Labels = the model's own predictions, so every point starts robust.  A random-init network is not robust: at
4/255 APGD-CE breaks every point and APGD-T has nothing left to do; `--eps 1e-7` gives the other extreme (nothing is
ever broken: all 1 + targets runs execute on every point, the worst case of the protocol).  The robustness-independent
figure is `model_evaluations_per_sec` (images pushed through forward [+ input-gradient backward] per second)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import revisiting_at_b200  # noqa: E402,F401
from revisiting_at_b200 import _abi, autoattack, convnext  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--arch', default='convnext_large')
ap.add_argument('--res', type=int, default=320)
ap.add_argument('--n', '--points', dest='n', type=int, default=100)
ap.add_argument('--bs', type=int, default=100)
ap.add_argument('--iters', type=int, default=100)
ap.add_argument('--targets', type=int, default=9)
ap.add_argument('--norm', default='Linf')
ap.add_argument('--eps', type=float, default=None, help='override the radius (e.g. 1e-7: no point is ever broken, so all 1 + targets runs execute on every point -- the worst case of the protocol)')
a = ap.parse_args()
# under torchrun (config 5: "batch-sharded over 8 x B200"): one rank per GPU, the library splits the N points into
# contiguous shards, attacks them without communication and all-gathers flags + adversarial points once at the end
world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', '1'), ('RANK', '0'), ('LOCAL_RANK', '0')))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
torch.backends.cudnn.benchmark = True
m = convnext.build(a.arch, normalize=True, seed=0).to(dev).eval()


class LazyPoints:
    """[N,3,R,R] synthetic points generated per 100-point block on demand (5000 x 3 x 320 x 320 fp32 = 6 GB: every rank
    only ever touches its own shard); supports the `x[lo:hi]` / `.shape` the evaluation uses."""
    def __init__(self, n, res):
        self.shape = (n, 3, res, res)

    def block(self, b):
        g = torch.Generator().manual_seed(1000 + b)
        return torch.rand(100, 3, self.shape[2], self.shape[3], generator=g)

    def __getitem__(self, sl):
        lo, hi, _ = sl.indices(self.shape[0])
        parts = [self.block(b)[max(lo - 100 * b, 0):min(hi - 100 * b, 100)] for b in range(lo // 100, (hi + 99) // 100)]
        return torch.cat(parts) if parts else torch.empty(0, *self.shape[1:])


if world > 1:
    per = (a.n + world - 1) // world                                  # the library's contiguous shards (autoattack._shard)
    lo, hi = min(rank * per, a.n), min((rank + 1) * per, a.n)
    xs = LazyPoints(a.n, a.res)[lo:hi]
    with torch.no_grad():
        ys = torch.cat([m(xs[i:i + a.bs].to(dev)).float().max(1)[1].cpu() for i in range(0, hi - lo, a.bs)])

    class Sharded:                                                    # x_orig / y_orig views that serve exactly this rank's slice
        def __init__(self, full_shape, lo, data):
            self.shape, self.lo, self.data = full_shape, lo, data

        def __getitem__(self, sl):
            l, h, _ = sl.indices(self.shape[0])
            assert l >= self.lo and h - self.lo <= self.data.shape[0], 'evaluation reached outside its shard'
            return self.data[l - self.lo:h - self.lo]
    x, y = Sharded((a.n, 3, a.res, a.res), lo, xs), Sharded((a.n,), lo, ys)
else:
    g = torch.Generator().manual_seed(0)
    x = torch.rand(a.n, 3, a.res, a.res, generator=g).to(dev)
    with torch.no_grad():
        y = torch.cat([m(x[i:i + a.bs]).float().max(1)[1] for i in range(0, a.n, a.bs)])
eps = {'Linf': 4 / 255., 'L2': 2., 'L1': 75.}[a.norm] if a.eps is None else a.eps
seen = {'n': 0}
m.register_forward_hook(lambda mod, inp, out: seen.__setitem__('n', seen['n'] + inp[0].shape[0]))
adv = autoattack.AutoAttack(m, norm=a.norm, eps=eps, version='standard', seed=0, verbose=False, device=dev)
adv.attacks_to_run = ['apgd-ce', 'apgd-t']
adv.apgd.n_iter = adv.apgd_targeted.n_iter = a.iters
adv.apgd.n_iter_orig = adv.apgd_targeted.n_iter_orig = a.iters
adv.apgd_targeted.n_target_classes = a.targets
# warm-up: a short evaluation (cuDNN/cuBLAS heuristics, lazy module loads)
w = autoattack.AutoAttack(m, norm=a.norm, eps=eps, version='standard', seed=0, verbose=False, device=dev)
w.attacks_to_run = ['apgd-ce', 'apgd-t']
w.apgd.n_iter = w.apgd_targeted.n_iter = 3
w.apgd.n_iter_orig = w.apgd_targeted.n_iter_orig = 3
w.apgd_targeted.n_target_classes = 1
if world > 1:
    w.run_standard_evaluation(xs[:a.bs], ys[:a.bs], bs=a.bs, shard=False)
    dist.barrier()
else:
    w.run_standard_evaluation(x[:a.bs], y[:a.bs], bs=a.bs)
torch.cuda.synchronize()
n0 = _abi.LAUNCHES['count']
seen['n'] = 0
t0 = time.time()
x_adv = adv.run_standard_evaluation(x, y, bs=a.bs)
torch.cuda.synchronize()
dt = time.time() - t0
if world > 1:
    tt = torch.tensor([dt, float(seen['n']), float(_abi.LAUNCHES['count'] - n0)], device=dev, dtype=torch.float64)
    mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = tt.clone(); dist.all_reduce(sm)
    dt, seen['n'] = mx[0].item(), int(sm[1].item())
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
    x = x_adv                                        # max_abs_delta below is only meaningful for the single-process run
evals = seen['n']                      # images pushed through the model (forward; all but ~1 % also backward)
print(json.dumps({'metric': 'aa_eval_points_per_sec', 'value': a.n / dt, 'unit': 'points/s', 'seconds': dt,
                  'config': {'workload': f'AutoAttack standard [apgd-ce, apgd-t x{a.targets}] {a.iters} iterations, '
                                         f'{a.norm} eps={eps:.5f}, {a.arch}-CvSt at {a.res}x{a.res}, {a.n} points, bs {a.bs} '
                                         f'(BASELINE.json configs[4], {world} GPU' + ('s, points sharded by rank, one all-gather at the end)' if world > 1 else ')')},
                  'n_gpus': world,
                  'model_evaluations': evals, 'model_evaluations_per_sec': evals / dt,
                  'robust_accuracy': adv.results, 'gpu_launches': _abi.LAUNCHES['count'] - n0,
                  'max_abs_delta': (x_adv - x).abs().max().item()}))

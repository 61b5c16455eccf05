"""CPU, world_size 2, gloo: the N>1 path of the attack and of the adversarial train step.

The attack shards by sample with no data-path collective: per-rank results concatenated
must equal the single-process result on the full batch.  The outer step's only collective is DDP's
gradient all-reduce: after one step both ranks hold identical parameters, equal to a single-process step
on the full batch with the mean loss.  The attack runs through the host-compiled kernel bodies here
(tests/hostcheck) because the product kernels need a GPU."""
import os
import socket
import sys
from functools import partial

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)


def _data():
    g = torch.Generator().manual_seed(77)
    return torch.rand(8, 3, 16, 16, generator=g), torch.randint(0, 10, (8,), generator=g)


def _shard_worker(rank, world, port, out):
    _setup(rank, world, port)
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import attack
    from hostcheck.backend import HostBackend
    from oracle.small_cnn import SmallCNN
    torch.manual_seed(0)
    model = SmallCNN().eval()
    x, y = _data()
    lo, hi = rank * 4, rank * 4 + 4
    res = {}
    for norm, eps in (('Linf', 8 / 255.), ('L2', 0.5), ('L1', 12.)):
        mine = attack.run_apgd(HostBackend(), model, x[lo:hi], y[lo:hi], norm, eps, n_iter=4)
        parts = [[torch.empty_like(t) if t.dtype != torch.bool else torch.empty(t.shape, dtype=torch.uint8)
                  for _ in range(world)] for t in mine]
        for t, ps in zip(mine, parts):
            dist.all_gather(ps, t if t.dtype != torch.bool else t.to(torch.uint8))
        if rank == 0:
            full = attack.run_apgd(HostBackend(), model, x, y, norm, eps, n_iter=4)
            res[norm] = all(torch.equal(torch.cat(ps), f if f.dtype != torch.bool else f.to(torch.uint8))
                            for ps, f in zip(parts, full))
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_attack_shards_without_communication(tmp_path):
    out = str(tmp_path / 'res.pt')
    mp.spawn(_shard_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res == {'Linf': True, 'L2': True, 'L1': True}, res


def _ddp_worker(rank, world, port, out, mode='flat'):
    # 'flat': one all-reduce over a flat buffer; 'flat3': the same buffer in gradient-arrival order, reduced in 3 ranges from
    # post-accumulate hooks; 'torch': DistributedDataParallel
    os.environ['B200AT_DDP'] = 'flat' if mode.startswith('flat') else mode
    os.environ['B200AT_FLAT_BUCKETS'] = '3' if mode == 'flat3' else '1'
    _setup(rank, world, port)
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import attack
    from revisiting_at_b200.train_step import AdvTrainStep
    from hostcheck.backend import HostBackend
    from oracle.small_cnn import SmallCNN
    perturb = partial(attack.run_apgd, HostBackend(), norm='Linf', eps=8 / 255., n_iter=2)
    torch.manual_seed(rank)                          # ranks start from DIFFERENT weights: the constructor must broadcast rank 0's
    model = SmallCNN()
    if rank == 0:
        torch.manual_seed(0)
        model = SmallCNN()
    x, y = _data()
    step = AdvTrainStep(model, distributed=True, device=torch.device('cpu'), autocast_dtype=torch.float32,
                        channels_last=False, perturb=perturb, lr=1e-2)
    lo, hi = rank * 4, rank * 4 + 4
    for _ in range(3 if mode == 'flat3' else 1):     # flat3: step 1 learns the arrival order, steps 2-3 reduce from the hooks
        loss = step(x[lo:hi], y[lo:hi])
    if mode == 'flat3':
        fr = step.flat_reduce
        assert len(fr.buckets) == 3 and not fr.learning and sorted(fr.index_of.values()) == list(range(len(fr.params)))
    params = torch.cat([p.detach().reshape(-1) for p in step.raw.parameters()])
    gathered = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather(gathered, params)
    if rank == 0:
        torch.manual_seed(0)
        single = AdvTrainStep(SmallCNN(), distributed=False, device=torch.device('cpu'),
                              autocast_dtype=torch.float32, channels_last=False, perturb=perturb, lr=1e-2)
        for _ in range(3 if mode == 'flat3' else 1):
            single(x, y)
        ref = torch.cat([p.detach().reshape(-1) for p in single.raw.parameters()])
        torch.save({'ranks_equal': torch.equal(gathered[0], gathered[1]),
                    'max_diff_vs_single': (gathered[0] - ref).abs().max().item(),
                    'loss_finite': bool(torch.isfinite(loss))}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['flat', 'flat3', 'torch'])
def test_ddp_step_matches_single_process(tmp_path, mode):
    out = str(tmp_path / 'res.pt')
    mp.spawn(_ddp_worker, args=(2, _free_port(), out, mode), nprocs=2, join=True)
    res = torch.load(out)
    assert res['ranks_equal'] and res['loss_finite'], res
    assert res['max_diff_vs_single'] < (3e-4 if mode == 'flat3' else 1e-4), res     # AdamW amplifies fp32 sum-order noise of the mean gradient (3 steps for flat3)


def _aa_worker(rank, world, port, out):
    _setup(rank, world, port)
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import autoattack as aa
    from hostcheck.backend import HostBackend
    from oracle.small_cnn import SensitiveNet
    torch.manual_seed(0)
    model = SensitiveNet().eval()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(10, 3, 16, 16, generator=g)
    with torch.no_grad():
        y = model(x).max(1)[1]

    def evaluate(xs, ys, shard):
        adv = aa.AutoAttack(model, norm='Linf', eps=8 / 255., version='standard', seed=7, verbose=False, device='cpu',
                            backend=HostBackend(4))
        adv.attacks_to_run = ['apgd-ce', 'apgd-t']
        adv.apgd.n_iter = adv.apgd_targeted.n_iter = 6
        adv.apgd_targeted.n_target_classes = 2
        adv.apgd.rng_device = adv.apgd_targeted.rng_device = 'cpu'
        return adv.run_standard_evaluation(xs, ys, bs=5, shard=shard), adv.results

    sharded, res = evaluate(x, y, True)                 # every rank gets the full, gathered result
    if rank == 0:
        parts = [evaluate(x[lo:lo + 5], y[lo:lo + 5], False)[0] for lo in (0, 5)]   # the shards, one process
        with torch.no_grad():
            rob = (model(sharded).max(1)[1] == y).float().mean().item()
        torch.save({'equal': torch.equal(sharded, torch.cat(parts)), 'results': res, 'robust': rob}, out)
    dist.destroy_process_group()


def test_autoattack_evaluation_shards_by_rank(tmp_path):
    """run_standard_evaluation under a process group: contiguous shards per rank, one all-gather at the end; equal
    to evaluating the two shards in one process, and the reported robust accuracy is the global one."""
    out = str(tmp_path / 'res.pt')
    mp.spawn(_aa_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res['equal'], res
    assert abs(res['results']['apgd-t'] - res['robust']) < 1e-6 and res['results']['clean'] == 1.0, res

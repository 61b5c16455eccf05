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


def _ddp_worker(rank, world, port, out):
    _setup(rank, world, port)
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import attack
    from revisiting_at_b200.train_step import AdvTrainStep
    from hostcheck.backend import HostBackend
    from oracle.small_cnn import SmallCNN
    perturb = partial(attack.run_apgd, HostBackend(), norm='Linf', eps=8 / 255., n_iter=2)
    torch.manual_seed(0)
    model = SmallCNN()
    x, y = _data()
    step = AdvTrainStep(model, distributed=True, device=torch.device('cpu'), autocast_dtype=torch.float32,
                        channels_last=False, perturb=perturb, lr=1e-2)
    lo, hi = rank * 4, rank * 4 + 4
    loss = step(x[lo:hi], y[lo:hi])
    params = torch.cat([p.detach().reshape(-1) for p in step.raw.parameters()])
    gathered = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather(gathered, params)
    if rank == 0:
        torch.manual_seed(0)
        single = AdvTrainStep(SmallCNN(), distributed=False, device=torch.device('cpu'),
                              autocast_dtype=torch.float32, channels_last=False, perturb=perturb, lr=1e-2)
        single(x, y)
        ref = torch.cat([p.detach().reshape(-1) for p in single.raw.parameters()])
        torch.save({'ranks_equal': torch.equal(gathered[0], gathered[1]),
                    'max_diff_vs_single': (gathered[0] - ref).abs().max().item(),
                    'loss_finite': bool(torch.isfinite(loss))}, out)
    dist.destroy_process_group()


def test_ddp_step_matches_single_process(tmp_path):
    out = str(tmp_path / 'res.pt')
    mp.spawn(_ddp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res['ranks_equal'] and res['loss_finite'], res
    assert res['max_diff_vs_single'] < 1e-4, res     # AdamW amplifies fp32 sum-order noise of the mean gradient

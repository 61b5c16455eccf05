"""`main.py` driven on the GPU through the reference's command line (/root/reference/main.py:331-345,
:696-756, :856-887): one tiny epoch of APGD adversarial training with EMA on synthetic batches, the checkpoint files it
writes, and a second trainer resumed from them -- whose EMA shadow must start from the LOADED weights (the reference
loads the checkpoint before it creates ModelEmaV2)."""
import importlib.util
import os
import shlex

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

ARGS = ("--data.num_workers=1 --data.in_memory=1 --data.train_dataset=synthetic:32 --data.val_dataset=synthetic "
        "--logging.folder={folder} --logging.log_level 1 --adv.attack apgd --adv.n_iter 2 --adv.norm Linf "
        "--training.distributed 0 --training.batch_size 8 --validation.batch_size 8 --lr.lr 1e-3 --logging.save_freq 1 "
        "--resolution.min_res 64 --resolution.max_res 64 --data.seed {seed} --data.augmentations 0 "
        "--model.add_normalization 1 --model.not_original 1 --model.model_ema 1 --model.pretrained 0 "
        "--model.arch convnext_tiny --training.epochs 1 --training.label_smoothing 0.1")


def _driver():
    from revisiting_at_b200 import fastargs_compat
    fastargs_compat.set_current_config(fastargs_compat.Config())
    spec = importlib.util.spec_from_file_location('_b200at_main_gpu', os.path.join(ROOT, 'main.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_main_trains_saves_and_resumes_with_ema(cuda_dev, tmp_path):
    from revisiting_at_b200 import checkpoint
    main = _driver()
    main.make_config(shlex.split(ARGS.format(folder=tmp_path / 'a', seed=0)), quiet=True)
    tr = main.ImageNetTrainer(gpu=0)
    before = {k: v.detach().clone() for k, v in tr.step.raw.state_dict().items()}
    tr.train()
    torch.cuda.synchronize()
    folder = tr.log_folder
    for name in ('weights_0.pt', 'weights_ema_0.pt', 'full_model_0.pth', 'params.json', 'log'):
        assert (folder / name).exists(), name
    w = torch.load(folder / 'weights_0.pt', map_location='cpu')
    assert 'module.base_model.model.stem.stem.0.weight' in w and 'module.base_model.normalize.mean' in w   # main.py:739
    moved = [k for k, v in tr.step.raw.state_dict().items() if v.is_floating_point() and not torch.equal(v, before[k])]
    assert len(moved) > 100                                            # the optimiser stepped
    assert all(bool(torch.isfinite(v).all()) for v in w.values())

    # resume in a trainer whose own initialisation differs (other seed): model AND EMA shadow equal the checkpoint
    main2 = _driver()
    main2.make_config(shlex.split(ARGS.format(folder=tmp_path / 'b', seed=5) +
                                  f" --model.ckpt_path {folder / 'weights_0.pt'}"), quiet=True)
    tr2 = main2.ImageNetTrainer(gpu=0)
    sd2 = tr2.step.raw.state_dict()
    ema2 = tr2.step.ema.state_dict()
    new, missing, unused = checkpoint.adapt_state_dict(w, sd2.keys())
    assert not missing and not unused
    for k, v in new.items():
        assert torch.equal(sd2[k].cpu(), v), k
        assert torch.equal(ema2[k].cpu(), v), ('EMA shadow was not seeded from the loaded checkpoint', k)

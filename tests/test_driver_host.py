"""Host logic of the reference-compatible driver: the fastargs command line of run_train.sh,
learning-rate schedules, weight-decay groups, checkpoint key formats, Mixup targets.  CPU only (no kernels)."""
import importlib.util
import math
import os
import shlex
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import revisiting_at_b200  # noqa: E402,F401
from revisiting_at_b200 import checkpoint, convnext, fastargs_compat, mixup, vit  # noqa: E402
from revisiting_at_b200.train_step import DeviceEma, WrappedModel  # noqa: E402
from oracle import convnext_oracle, ref_loader  # noqa: E402

# the argument line of /root/reference/run_train.sh:10-18 with its placeholders filled in
RUN_TRAIN_SH = """--data.num_workers=12 --data.in_memory=1
        --data.train_dataset=synthetic:64 --data.val_dataset=synthetic
        --logging.folder=/tmp/b200at_runs --logging.log_level 2
    --adv.attack apgd --adv.n_iter 2 --adv.norm Linf --training.distributed 1 --training.batch_size 80 --lr.lr 1e-3 --logging.save_freq 2
    --resolution.min_res 224 --resolution.max_res 224 --data.seed 0 --data.augmentations 1 --model.add_normalization 0
     --model.not_original 1 --model.model_ema 1 --lr.lr_peak_epoch 20
    --training.label_smoothing 0.1 --logging.addendum='additional_text'
    --dist.world_size 8 --training.distributed 1 --model.pretrained 0 --model.arch convnext_base --training.epochs 300"""


@pytest.fixture()
def driver():
    """main.py imported on a fresh config object (its Section declarations run at import)."""
    fastargs_compat.set_current_config(fastargs_compat.Config())
    spec = importlib.util.spec_from_file_location('_b200at_main', os.path.join(ROOT, 'main.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_run_train_sh_command_line_parses_under_the_reference_keys(driver):
    cfg = driver.make_config(shlex.split(RUN_TRAIN_SH), quiet=True)
    assert cfg['adv.attack'] == 'apgd' and cfg['adv.n_iter'] == 2 and cfg['adv.norm'] == 'Linf'
    assert cfg['adv.eps'] == 4. / 255. and cfg['adv.alpha'] == 1.          # defaults of main.py:175-184
    assert cfg['training.batch_size'] == 80 and cfg['training.distributed'] == 1 and cfg['dist.world_size'] == 8
    assert cfg['model.arch'] == 'convnext_base' and cfg['model.not_original'] == 1 and cfg['model.model_ema'] == 1.
    assert cfg['lr.lr'] == 1e-3 and cfg['lr.lr_peak_epoch'] == 20 and cfg['training.epochs'] == 300
    assert cfg['training.label_smoothing'] == 0.1 and cfg['logging.addendum'] == 'additional_text'
    assert cfg['data.train_dataset'] == 'synthetic:64' and cfg['data.num_workers'] == 12
    assert cfg[('resolution', 'min_res')] == 224
    assert cfg.get().adv.n_iter == 2
    # every key of the reference's sections is declared (main.py:106-189)
    want = {'model': 9, 'resolution': 4, 'data': 6, 'lr': 5, 'logging': 4, 'validation': 4, 'training': 10, 'dist': 3,
            'adv': 8, 'misc': 2}
    have = {}
    for path in cfg.entries:
        have[path[0]] = have.get(path[0], 0) + 1
    assert have == want


def test_config_errors(driver):
    cfg = fastargs_compat.get_current_config()
    import argparse
    p = argparse.ArgumentParser()
    cfg.augment_argparse(p)
    cfg.collect_argparse_args(p, ['--lr.lr_schedule_type', 'triangle', '--adv.n_iter', 'two'])
    errs = cfg.validate(mode='errordict')
    assert 'lr.lr_schedule_type' in errs and 'adv.n_iter' in errs
    assert {'data.train_dataset', 'data.val_dataset', 'data.num_workers', 'data.in_memory'} <= set(errs)   # required


def test_yaml_config_file_and_cli_precedence(driver, tmp_path):
    f = tmp_path / 'rn.yaml'
    f.write_text('adv:\n  attack: fgsm\n  alpha: 1.25\ndata:\n  train_dataset: synthetic\n  val_dataset: synthetic\n'
                 '  num_workers: 1\n  in_memory: 1\n')
    cfg = driver.make_config(['--config-file', str(f), '--adv.alpha', '2.0'], quiet=True)
    assert cfg['adv.attack'] == 'fgsm' and cfg['adv.alpha'] == 2.0


def test_param_decorator_injects_and_caller_wins(driver):
    driver.make_config(shlex.split(RUN_TRAIN_SH), quiet=True)

    @fastargs_compat.param('adv.n_iter')
    @fastargs_compat.param('adv.eps', alias='radius')
    def f(x, n_iter, radius):
        return x, n_iter, radius
    assert f(7) == (7, 2, 4. / 255.)
    assert f(7, n_iter=5) == (7, 5, 4. / 255.)


def test_lr_schedules(driver):
    driver.make_config(shlex.split(RUN_TRAIN_SH), quiet=True)             # lr 1e-3, peak 20, epochs 300
    lr, peak, n = 1e-3, 20, 300
    assert driver.get_cosine_lr(0) == pytest.approx(1e-7)
    assert driver.get_cosine_lr(10) == pytest.approx(1e-7 + (lr - 1e-7) * 0.5)
    assert driver.get_cosine_lr(20) == pytest.approx(lr)
    for e in (21, 160, 299, 300):
        want = 5e-6 + .5 * (lr - 5e-6) * (1 + math.cos(math.pi * (e - peak) / (n - peak)))
        assert driver.get_cosine_lr(e) == pytest.approx(want, rel=1e-12)
    assert driver.get_cosine_lr(300) == pytest.approx(5e-6)
    assert driver.get_cyclic_lr(20) == pytest.approx(lr) and driver.get_cyclic_lr(300) == 0.
    assert driver.get_cyclic_lr(160) == pytest.approx(lr * (300 - 160) / 280)
    assert driver.get_step_lr(0) == lr and driver.get_step_lr(65) == pytest.approx(lr * 0.1 ** 2) and driver.get_step_lr(300) == 0


def test_weight_decay_groups_follow_the_reference_rule(driver):
    m = WrappedModel(convnext.Normalized(convnext.ConvNeXtCvSt('convnext_tiny')), None)
    g0, g1 = driver.weight_decay_groups(m.named_parameters(), 'convnext_tiny', 0.05)
    names = {id(p): k for k, p in m.named_parameters()}
    no_decay = {names[id(p)] for p in g0['params']}
    decay = {names[id(p)] for p in g1['params']}
    assert g0['weight_decay'] == 0. and g1['weight_decay'] == 0.05
    assert all(k.endswith('.bias') for k in no_decay)                     # 'bn' never occurs, '.bias' does
    assert 'base_model.model.stages.0.blocks.0.gamma' in decay            # layer scale and LN weights ARE decayed
    assert 'base_model.model.stages.0.blocks.0.norm.weight' in decay
    assert len(no_decay) + len(decay) == len(names)
    v = vit.build(normalize=False)
    g0, g1 = driver.weight_decay_groups(v.named_parameters(), 'vit_s', 0.05)
    assert all(p.ndim <= 1 for p in g0['params']) and all(p.ndim > 1 for p in g1['params'])


@pytest.mark.parametrize('arch', ['convnext_tiny', 'convnext_base'])
def test_vendored_names_map_onto_timm_names(arch):
    table = convnext_oracle.vendored_key_map(arch)                        # independent table: timm -> vendored
    for timm_name, vendored in table.items():
        assert checkpoint.vendored_to_timm(vendored) == timm_name
        assert checkpoint.vendored_to_timm(timm_name) == timm_name


def _filled(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape, generator=g))
    return model


def test_reference_checkpoint_round_trip(tmp_path):
    src = WrappedModel(convnext.Normalized(_filled(convnext.ConvNeXtCvSt('convnext_tiny'), 1)), None)
    opt = torch.optim.AdamW(src.parameters(), lr=1e-3)
    ema = DeviceEma(src)
    checkpoint.save_checkpoint(str(tmp_path), 0, src, opt, ema.state_dict(), epochs=1)
    w = torch.load(tmp_path / 'weights_0.pt')
    assert 'module.base_model.model.stem.stem.0.weight' in w and 'module.base_model.normalize.mean' in w   # main.py:739 names
    full = torch.load(tmp_path / 'full_model_0.pth')
    assert set(full) == {'model_state_dict', 'optimizer_state_dict', 'loss_scaler_state_dict', 'epoch', 'state_dict_ema'}
    want = dict(src.base_model.model.state_dict())
    # into every wrapper combination the reference's three-way retry (main.py:856-872) covers
    targets = [convnext.ConvNeXtCvSt('convnext_tiny'), convnext.Normalized(convnext.ConvNeXtCvSt('convnext_tiny')),
               WrappedModel(convnext.Normalized(convnext.ConvNeXtCvSt('convnext_tiny')), None)]
    for tgt, src_file in zip(targets, ('weights_0.pt', 'full_model_0.pth', 'weights_ema_0.pt')):
        missing, unused = checkpoint.load_checkpoint(tgt, str(tmp_path / src_file))
        assert not missing and not unused
        got = {checkpoint._core(k): v for k, v in tgt.state_dict().items() if 'normalize.' not in k}
        assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)
    with pytest.raises(KeyError):
        checkpoint.load_checkpoint(convnext.ConvNeXtCvSt('convnext_small'), str(tmp_path / 'weights_0.pt'))
    bad = {k: v for k, v in w.items()}
    bad['module.base_model.model.head.fc.weight'] = torch.zeros(10, 768)
    with pytest.raises(ValueError):
        checkpoint.load_checkpoint(convnext.ConvNeXtCvSt('convnext_tiny'), bad)


@pytest.mark.skipif(not ref_loader.available(), reason='reference checkout not mounted')
def test_state_dict_of_the_reference_vendored_convnext_loads():
    ref, _ = ref_loader.convnext_t_cvst()
    tgt = convnext.Normalized(convnext.ConvNeXtCvSt('convnext_tiny'))
    missing, unused = checkpoint.load_checkpoint(tgt, {'module.' + k: v for k, v in ref.state_dict().items()})
    assert not missing and not unused
    table = convnext_oracle.vendored_key_map('convnext_tiny')
    rsd, tsd = ref.state_dict(), tgt.model.state_dict()
    assert all(torch.equal(tsd[t], rsd[v]) for t, v in table.items()) and len(table) == len(tsd)


def test_ema_state_dict_uses_model_names():
    m = convnext.Normalized(_filled(convnext.ConvNeXtCvSt('convnext_tiny'), 3))
    ema = DeviceEma(m, decay=0.5)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(2.)
    ema.update()
    sd, cur = ema.state_dict(), m.state_dict()
    assert list(sd) == list(cur)
    k = 'model.stages.1.blocks.0.mlp.fc1.weight'
    assert torch.allclose(sd[k], cur[k] - 1.)                             # halfway between old (cur-2) and new
    assert torch.equal(sd['normalize.mean'], cur['normalize.mean'])


def test_mixup_batch_mode_targets_and_images():
    mx = mixup.Mixup(label_smoothing=0.1, num_classes=10)
    x = torch.rand(8, 3, 16, 16)
    y = torch.arange(8) % 10
    seen = set()
    for seed in range(12):
        np.random.seed(seed)
        xm, ym = mx(x, y)
        np.random.seed(seed)                                              # replay the draws independently
        assert np.random.rand() < 1.0
        cut = np.random.rand() < 0.5
        lam = float(np.random.beta(1.0, 1.0) if cut else np.random.beta(0.8, 0.8))
        if cut:
            r = np.sqrt(1. - lam)
            ch, cw = int(16 * r), int(16 * r)
            cy, cx = np.random.randint(0, 16), np.random.randint(0, 16)
            yl, yh = np.clip(cy - ch // 2, 0, 16), np.clip(cy + ch // 2, 0, 16)
            xl, xh = np.clip(cx - cw // 2, 0, 16), np.clip(cx + cw // 2, 0, 16)
            lam = 1. - (yh - yl) * (xh - xl) / 256.
            want = x.clone()
            want[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
        else:
            want = x * lam + x.flip(0) * (1. - lam)
        seen.add(bool(cut))
        assert torch.equal(xm, want)
        one = torch.full((8, 10), 0.01)
        one[torch.arange(8), y] = 0.91
        assert torch.allclose(ym, one * lam + one.flip(0) * (1. - lam), atol=1e-7)
        assert torch.allclose(ym.sum(1), torch.ones(8), atol=1e-6) and ym.dtype == torch.float32
    assert seen == {True, False}
    with pytest.raises(AssertionError):
        mx(x[:7], y[:7])


def test_interpolate_pos_encoding():
    torch.manual_seed(0)
    pe = torch.randn(1, 197, 384)
    assert checkpoint.interpolate_pos_encoding(pe, 224) is pe
    out = checkpoint.interpolate_pos_encoding(pe, 320)
    assert out.shape == (1, 401, 384) and torch.equal(out[:, 0], pe[:, 0])
    if ref_loader.available():
        _, ua = ref_loader.convnext_t_cvst()
        assert torch.equal(out, ua.interpolate_pos_encoding(pe, new_img_size=320, patch_size=16))


def _aa_eval():
    spec = importlib.util.spec_from_file_location('_b200at_aa_eval', os.path.join(ROOT, 'AA_eval.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_aa_eval_accepts_the_runner_command_line(tmp_path):
    """the argument string runner_aa_eval.py:13-16 builds (with its `--not-orig` abbreviation and `--a100` flag)"""
    aa = _aa_eval()
    job = '--model_in {} --mod {} --not-orig {} --a100 {} --full_aa {} --l_norms {} --batch_size {}'.format(
        str(tmp_path), 'convnext_base', 1, 1, 0, 'Linf', 100)
    a = aa.get_args_parser(shlex.split(job))
    assert a.model_in == [str(tmp_path)] and a.mod == 'convnext_base' and a.not_original == 1 and a.a100 == 1
    assert a.full_aa == 0 and a.l_norms == 'Linf' and a.batch_size == 100 and a.img_size == 224 and a.n_ex == 5000
    assert aa.eps_dict['imagenet'] == {'Linf': 4. / 255., 'L2': 2., 'L1': 75.}
    # run folder -> weights_20.pt when present (AA_eval.py:124), else the latest weights_N.pt; EMA files are not picked
    for n in (2, 11, 7):
        torch.save({}, tmp_path / f'weights_{n}.pt')
    torch.save({}, tmp_path / 'weights_ema_30.pt')
    assert aa.resolve_checkpoint(str(tmp_path)) == (str(tmp_path / 'weights_11.pt'), str(tmp_path))
    torch.save({}, tmp_path / 'weights_20.pt')
    assert aa.resolve_checkpoint(str(tmp_path))[0] == str(tmp_path / 'weights_20.pt')
    assert aa.resolve_checkpoint(str(tmp_path / 'weights_7.pt')) == (str(tmp_path / 'weights_7.pt'), str(tmp_path))
    assert aa.resolve_checkpoint('random') == (None, None)
    x, y = aa.load_points('synthetic', 6, 32)
    assert x.shape == (6, 3, 32, 32) and y.shape == (6,) and 0. <= float(x.min()) and float(x.max()) < 1.
    x2, _ = aa.load_points('synthetic', 6, 32)
    assert torch.equal(x, x2)                                              # the subset is fixed


def test_vit_checkpoint_round_trip_keeps_timm_names():
    """ViT keys (`pos_embed`, `norm.weight`, `head.weight`) must NOT be taken for vendored ConvNeXt spellings"""
    src = WrappedModel(vit.build(normalize=True, seed=1), None)
    sd = checkpoint.reference_state_dict(src)
    assert 'module.base_model.model.pos_embed' in sd and 'module.base_model.model.norm.weight' in sd
    tgt = vit.build(normalize=False, seed=2)
    missing, unused = checkpoint.load_checkpoint(tgt, sd)
    assert not missing and not unused
    a = {checkpoint._core(k): v for k, v in src.state_dict().items() if 'normalize' not in k}
    b = {checkpoint._core(k): v for k, v in tgt.state_dict().items()}
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in b)

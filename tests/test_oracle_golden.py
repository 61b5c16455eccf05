"""CPU: the oracle restatement reproduces the reference's own outputs (fixtures made by
oracle/make_goldens.py from the unmodified /root/reference functions)."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, same
from oracle import apgd_oracle as ao
from oracle.scripted_model import ScriptedModel


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize('name', golden_names('scripted_'))
def test_scripted_matches_reference(name):
    g = golden(name)
    norm, eps, n_iter = str(g['norm']), float(g['eps']), int(g['n_iter'])
    model = ScriptedModel(_t(g['logits']), _t(g['grads']))
    x_best, acc, loss_best, x_best_adv = ao.apgd_train_oracle(
        model, _t(g['x']), _t(g['y']), norm, eps, n_iter=n_iter, loss=str(g['loss']),
        mixup=(object() if bool(g['soft']) else None), is_train=bool(g['is_train']))
    seen = torch.stack(model.seen)
    # identical gradients in -> identical iterates out, for every call the model saw
    assert same(seen, _t(g['x_calls'])), f'{name}: iterate trajectory differs'
    assert same(x_best, _t(g['x_best']))
    assert same(x_best_adv, _t(g['x_best_adv']))
    assert same(loss_best, _t(g['loss_best']))
    assert same(acc, _t(g['acc']))


def test_schedule_matches_survey_appendix_a3():
    # SURVEY.md A.3 (probe-verified against the reference loop)
    s = ao.checkpoint_schedule('Linf', 100)
    assert [i for i, k in enumerate(s) if k] == [21, 40, 56, 69, 79, 86, 92, 98]
    assert [k for k in s if k] == [22, 19, 16, 13, 10, 7, 6, 6]
    assert ao.checkpoint_schedule('Linf', 2) == [1, 1]
    s10 = ao.checkpoint_schedule('L2', 10)
    assert s10[0] == 0 and s10[1] == 2 and all(k == 1 for k in s10[2:])
    assert ao.checkpoint_schedule('L1', 100) == [4 if (i + 1) % 4 == 0 else 0 for i in range(100)]


def test_error_behaviour():
    class M:
        training = True
    x = torch.rand(2, 3, 4, 4)
    y = torch.zeros(2, dtype=torch.long)
    with pytest.raises(AssertionError):
        ao.apgd_train_oracle(M(), x, y, 'Linf', 0.1)
    M.training = False
    with pytest.raises(KeyError):
        ao.apgd_train_oracle(M(), x, y, 'Linf', 0.1, loss='nope')
    with pytest.raises(TypeError):
        ao.apgd_train_oracle(M(), x, y, 'Linf', 0.1, use_rs=True)


@pytest.mark.parametrize('name', golden_names('cnn_'))
def test_cnn_loop_matches_reference(name):
    from oracle.small_cnn import from_fixture
    g = golden(name)
    model = from_fixture(g)
    out = ao.apgd_train_oracle(model, _t(g['x']), _t(g['y']), str(g['norm']), float(g['eps']), n_iter=int(g['n_iter']))
    for got, key in zip(out, ('x_best', 'acc', 'loss_best', 'x_best_adv')):
        assert same(got, _t(g[key])), f'{name}:{key}'


def test_fgsm_matches_reference():
    from oracle.small_cnn import from_fixture
    g = golden('fgsm_cnn')
    model = from_fixture(g)
    x, y, eps = _t(g['x']), _t(g['y']), float(g['eps'])
    for tag, kw in (('plain', dict(use_rs=False)), ('rs', dict(use_rs=True, alpha=1.25, noise_level=1.)),
                    ('rs_skip', dict(use_rs=True, alpha=1.0, noise_level=0.5, skip_projection=True))):
        out = ao.fgsm_train_oracle(model, x, y, eps, noise=_t(g['noise_' + tag]), **kw)
        assert same(out, _t(g['out_' + tag])), tag


def test_convnext_oracle_matches_vendored_reference_model():
    from oracle import convnext_oracle as co
    g = golden('convnext_t_cvst')
    m = co.build('convnext_tiny', normalize=False, seed=0)
    csum = float(sum(v.double().abs().sum() for v in m.state_dict().values()))
    assert abs(csum - float(g['weight_abs_sum'])) < 1e-6 * csum, 'seed-0 init drifted from the fixture'
    with torch.no_grad():
        logits = m(_t(g['x']))
    assert torch.allclose(logits, _t(g['logits']), atol=1e-5, rtol=0)
    out = ao.apgd_train_oracle(m, _t(g['x']), _t(g['y']), 'Linf', 4 / 255., n_iter=2)
    assert (out[0] - _t(g['x_best'])).abs().max() <= 1e-6
    assert same(out[1], _t(g['acc']))
    assert torch.allclose(out[2], _t(g['loss_best']), atol=1e-5)

"""CPU: the oracle restatement reproduces the reference's own outputs.

Two layers of pinning:
* committed fixtures (`tests/golden/*.npz`, made by `oracle/make_goldens.py` from the unmodified
  /root/reference functions): elementwise arithmetic (l-inf, every decision) is compared bit for bit;
  anything downstream of a floating-point *reduction* (l2 norms, the l1 cumsum, CPU convolutions) at
  the north-star tolerance of 1e-6, because torch's CPU reduction order depends on the host's vector
  ISA (fixtures written on an AVX-512 host are 1 ulp away on an AVX2 one);
* live (`test_live_*`, only where /root/reference is mounted): reference and oracle run side by side
  on this host, where both see the same torch kernels -> bit for bit for all three norms.
"""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, same
from oracle import apgd_oracle as ao
from oracle import ref_loader
from oracle.scripted_model import ScriptedModel

TOL = 1e-6          # north_star: final x_adv within 1e-6 absolute in fp32
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not mounted')


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(a, b, tol=TOL):
    return a.shape == b.shape and float((a - b).abs().max()) <= tol


def _run_scripted(g):
    model = ScriptedModel(_t(g['logits']), _t(g['grads']))
    out = ao.apgd_train_oracle(
        model, _t(g['x']), _t(g['y']), str(g['norm']), float(g['eps']), n_iter=int(g['n_iter']),
        loss=str(g['loss']), mixup=(object() if bool(g['soft']) else None), is_train=bool(g['is_train']))
    return torch.stack(model.seen), out


@pytest.mark.parametrize('name', golden_names('scripted_'))
def test_scripted_matches_reference(name):
    g = golden(name)
    seen, (x_best, acc, loss_best, x_best_adv) = _run_scripted(g)
    # identical gradients in -> identical iterates out, for every call the model saw
    eq = same if str(g['norm']) == 'Linf' else _close
    assert eq(seen, _t(g['x_calls'])), f'{name}: iterate trajectory differs'
    assert eq(x_best, _t(g['x_best']))
    assert eq(x_best_adv, _t(g['x_best_adv']))
    assert same(loss_best, _t(g['loss_best']))
    assert same(acc, _t(g['acc']))


@needs_ref
def test_live_scripted_bit_exact():
    from oracle import make_goldens as mg
    ref = ref_loader.attack_module()
    for name, (args, kw) in mg.SCRIPTED_CASES.items():
        g = mg.compute_scripted(ref, *args, **kw)
        seen, out = _run_scripted(g)
        assert same(seen, _t(g['x_calls'])), name
        for got, key in zip(out, ('x_best', 'acc', 'loss_best', 'x_best_adv')):
            assert same(got, _t(g[key])), f'{name}:{key}'


@needs_ref
def test_live_cnn_loop_bit_exact():
    from oracle import make_goldens as mg
    from oracle.small_cnn import from_fixture
    ref = ref_loader.attack_module()
    for name, args in mg.CNN_CASES.items():
        g = mg.compute_cnn(ref, *args)
        out = ao.apgd_train_oracle(from_fixture(g), _t(g['x']), _t(g['y']), str(g['norm']), float(g['eps']),
                                   n_iter=int(g['n_iter']))
        for got, key in zip(out, ('x_best', 'acc', 'loss_best', 'x_best_adv')):
            assert same(got, _t(g[key])), f'{name}:{key}'


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
    """Fixture written on another host: its conv kernels differ in the last ulp, and a ~0 gradient that
    flips sign moves that pixel by a whole step -> >= 99.9 % of pixels within 1e-6, masks equal."""
    from oracle.small_cnn import from_fixture
    g = golden(name)
    model = from_fixture(g)
    out = ao.apgd_train_oracle(model, _t(g['x']), _t(g['y']), str(g['norm']), float(g['eps']), n_iter=int(g['n_iter']))
    for got, key in ((out[0], 'x_best'), (out[3], 'x_best_adv')):
        frac = ((got - _t(g[key])).abs() <= TOL).float().mean().item()
        assert frac >= 0.999, f'{name}:{key} {frac}'
    assert same(out[1], _t(g['acc']))
    assert torch.allclose(out[2], _t(g['loss_best']), atol=2e-5, rtol=0)


def test_fgsm_matches_reference():
    from oracle.small_cnn import from_fixture
    g = golden('fgsm_cnn')
    model = from_fixture(g)
    x, y, eps = _t(g['x']), _t(g['y']), float(g['eps'])
    for tag, kw in (('plain', dict(use_rs=False)), ('rs', dict(use_rs=True, alpha=1.25, noise_level=1.)),
                    ('rs_skip', dict(use_rs=True, alpha=1.0, noise_level=0.5, skip_projection=True))):
        out = ao.fgsm_train_oracle(model, x, y, eps, noise=_t(g['noise_' + tag]), **kw)
        assert ((out - _t(g['out_' + tag])).abs() <= TOL).float().mean().item() >= 0.999, tag


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
    assert ((out[0] - _t(g['x_best'])).abs() <= TOL).float().mean().item() >= 0.999
    assert same(out[1], _t(g['acc']))
    assert torch.allclose(out[2], _t(g['loss_best']), atol=1e-5)

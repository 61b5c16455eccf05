"""CPU: the AutoAttack-compatible evaluation protocol (`revisiting_at_b200.autoattack`) driven through the
host-compiled kernel bodies agrees with the CPU restatement of autoattack-0.1's published algorithm
(oracle/autoattack_oracle.py; parity unpinned: the package is not under /root/reference) on a small real model:
same random starts (one CPU generator), same restarts / target classes / batching."""
import pytest
import torch

from hostcheck.backend import HostBackend
from oracle import autoattack_oracle as ao
from oracle.small_cnn import SensitiveNet
import revisiting_at_b200  # noqa: F401
from revisiting_at_b200 import autoattack as aa


def _setup(seed=0, B=12, C=10, hw=16):
    torch.manual_seed(seed)
    model = SensitiveNet(C, hw).eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, 3, hw, hw, generator=g)
    with torch.no_grad():
        y = model(x).max(1)[1]
    y[::5] = (y[::5] + 1) % C                       # some points are misclassified from the start
    return model, x, y


def _frac_close(a, b, tol=1e-6):
    return ((a - b).abs() <= tol).float().mean().item()


@pytest.mark.parametrize('norm,eps', [('Linf', 8 / 255.), ('L2', 0.5)])
def test_apgd_ce_perturb_matches_oracle(norm, eps):
    model, x, y = _setup()
    att = aa.APGDAttack(model, n_iter=20, norm=norm, n_restarts=2, eps=eps, seed=3, loss='ce', backend=HostBackend(4))
    att.rng_device = 'cpu'
    adv = att.perturb(x, y)
    ref = ao.apgd_perturb(model, x, y, norm, eps, n_iter=20, n_restarts=2, loss='ce', seed=3)
    # a ~0 gradient component can change sign between two summation orders and move that pixel by a step
    assert _frac_close(adv, ref) >= 0.999, _frac_close(adv, ref)
    with torch.no_grad():
        assert torch.equal(model(adv).max(1)[1] == y, model(ref).max(1)[1] == y)
    if norm == 'Linf':
        assert (adv - x).abs().max() <= eps + 1e-6
    else:
        assert ((adv - x) ** 2).flatten(1).sum(1).sqrt().max() <= eps * (1 + 1e-4)
    assert adv.min() >= 0 and adv.max() <= 1


def test_apgd_dlr_loss_runs_through_perturb():
    model, x, y = _setup(1)
    att = aa.APGDAttack(model, n_iter=10, norm='Linf', eps=8 / 255., seed=1, loss='dlr', backend=HostBackend(4))
    att.rng_device = 'cpu'
    adv = att.perturb(x, y)
    ref = ao.apgd_perturb(model, x, y, 'Linf', 8 / 255., n_iter=10, n_restarts=1, loss='dlr', seed=1)
    assert _frac_close(adv, ref) >= 0.999


def test_apgd_targeted_matches_oracle():
    model, x, y = _setup(2)
    att = aa.APGDAttack_targeted(model, n_iter=10, norm='Linf', eps=8 / 255., seed=5, n_target_classes=3,
                                 backend=HostBackend(4))
    att.rng_device = 'cpu'
    adv = att.perturb(x, y)
    ref = ao.apgd_targeted_perturb(model, x, y, 'Linf', 8 / 255., n_iter=10, n_restarts=1, n_target_classes=3, seed=5)
    assert _frac_close(adv, ref) >= 0.999, _frac_close(adv, ref)
    assert att.y_target is None


def test_l1_largereps_matches_oracle():
    model, x, y = _setup(3, B=6)
    att = aa.APGDAttack(model, n_iter=10, norm='L1', n_restarts=1, eps=6., seed=2, loss='ce', use_largereps=True,
                        backend=HostBackend(4))
    att.rng_device = 'cpu'
    assert att._schedule() == ao.largereps_schedule(6., 10)
    adv = att.perturb(x, y)
    ref = ao.apgd_perturb(model, x, y, 'L1', 6., n_iter=10, n_restarts=1, loss='ce', seed=2, largereps=True)
    assert (adv - x).abs().flatten(1).sum(1).max() <= 6. * (1 + 1e-4)
    assert adv.min() >= 0 and adv.max() <= 1
    with torch.no_grad():
        assert torch.equal(model(adv).max(1)[1] == y, model(ref).max(1)[1] == y)
    assert _frac_close(adv, ref, 1e-5) >= 0.99, _frac_close(adv, ref, 1e-5)


def test_run_standard_evaluation_matches_oracle():
    model, x, y = _setup(0, B=14)
    adv = aa.AutoAttack(model, norm='Linf', eps=8 / 255., version='standard', seed=7, verbose=False, device='cpu',
                        backend=HostBackend(4))
    assert adv.attacks_to_run == ['apgd-ce', 'apgd-t', 'fab-t', 'square']
    assert adv.apgd.n_restarts == 1 and adv.apgd_targeted.n_target_classes == 9
    with pytest.raises(NotImplementedError):
        adv.run_standard_evaluation(x, y, bs=5)                       # fab-t / square are outside the hot path
    adv.attacks_to_run = ['apgd-ce', 'apgd-t']                        # AA_eval.py:233-234
    adv.apgd.n_iter = adv.apgd_targeted.n_iter = 8
    adv.apgd_targeted.n_target_classes = 3
    adv.apgd.rng_device = adv.apgd_targeted.rng_device = 'cpu'
    x_adv, y_adv = adv.run_standard_evaluation(x, y, bs=5, return_labels=True)
    # oracle with the same reduced budget
    N = x.shape[0]
    robust = torch.zeros(N, dtype=torch.bool)
    with torch.no_grad():
        robust[:] = model(x).max(1)[1] == y
    ref = x.clone()
    for name in ('apgd-ce', 'apgd-t'):
        idcs = robust.nonzero().squeeze(1)
        for s in range(0, idcs.numel(), 5):
            bi = idcs[s:s + 5]
            if name == 'apgd-ce':
                a = ao.apgd_perturb(model, x[bi], y[bi], 'Linf', 8 / 255., 8, 1, 'ce', 7)
            else:
                a = ao.apgd_targeted_perturb(model, x[bi], y[bi], 'Linf', 8 / 255., 8, 1, 3, 7)
            with torch.no_grad():
                fb = ~(model(a).max(1)[1] == y[bi])
            robust[bi[fb]] = False
            ref[bi[fb]] = a[fb]
    with torch.no_grad():
        got_flags = model(x_adv).max(1)[1] == y
    assert torch.equal(got_flags, robust)
    assert _frac_close(x_adv, ref) >= 0.999
    assert abs(adv.results['apgd-t'] - robust.sum().item() / N) < 1e-9
    assert adv.results['apgd-t'] < adv.results['apgd-ce'] < adv.results['clean']     # the compaction paths ran
    assert torch.equal(y_adv, model(x_adv).max(1)[1])


def test_constructor_contract():
    model, x, y = _setup()
    with pytest.raises(ValueError):
        aa.AutoAttack(model, attacks_to_run=['apgd-ce'], version='standard', verbose=False)
    with pytest.raises(NotImplementedError):
        aa.APGDAttack(model, eps=0.1, eot_iter=2)
    import autoattack
    assert autoattack.AutoAttack is aa.AutoAttack

"""GPU: the AutoAttack-compatible APGD-CE / APGD-T evaluation (BASELINE config 5) on the CUDA kernels against the CPU
restatement of autoattack-0.1's published algorithm (oracle/autoattack_oracle.py; parity unpinned: the package is
not under /root/reference; the targeted DLR loss is pinned by autopgd_train_clean.py:106-111)."""
import pytest
import torch

from oracle import autoattack_oracle as ao
from oracle.small_cnn import SensitiveNet

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_dlr_targeted_kernel_vs_reference_formula(cuda_dev, dtype):
    """-(z_y - z_t) / (z_(1) - (z_(3) + z_(4)) / 2 + 1e-12) and its gradient (autopgd_train_clean.py:106-111)"""
    import revisiting_at_b200  # noqa: F401
    from revisiting_at_b200 import _abi
    g = torch.Generator().manual_seed(12)
    if dtype == torch.float32:
        B, C = 37, 1000
        z = (torch.randn(B, C, generator=g) * 3).to(cuda_dev)
    else:
        # bf16 logits: distinct, exactly representable values (k/8, k < 200), so that the order statistics have no
        # ties -- with ties the (sub)gradient depends on the sort's arbitrary tie order, in the reference too
        B, C = 37, 200
        z = torch.stack([torch.randperm(C, generator=g) for _ in range(B)]).float().div(8).sub(12).to(cuda_dev).to(dtype)
    y = torch.randint(0, C, (B,), generator=g).to(cuda_dev)
    y[:10] = z[:10].float().argmax(1)
    yt = torch.randint(0, C, (B,), generator=g).to(cuda_dev)
    yt[10:20] = z[10:20].float().sort(1)[1][:, -3]           # the target coincides with a sorted position
    zf = z.float().clone().requires_grad_(True)
    want = ao.dlr_targeted_rows(zf, y, yt)
    (dwant,) = torch.autograd.grad(want.sum(), zf)
    dl, lo = torch.empty_like(z), torch.empty(B, device=cuda_dev)
    st = torch.zeros(_abi.ST_ROWS, B, device=cuda_dev)
    _abi.loss_bookkeep(z, y, dl, lo, st, torch.zeros(1, B, device=cuda_dev), -1, 1, 0, 'Linf', 'dlr-targeted', 0.1,
                       0.01, 10, y_target=yt)
    torch.cuda.synchronize()
    assert torch.allclose(lo, want.detach(), atol=1e-6, rtol=1e-6)
    tol = 1e-6 if dtype == torch.float32 else 4e-3
    assert (dl.float() - dwant).abs().max() <= tol * max(1.0, dwant.abs().max().item())
    assert torch.equal(st[_abi.ST_PRED].view(torch.int32) != 0, z.float().max(1)[1] == y)


def _setup(seed, B, hw=16, C=10):
    torch.manual_seed(seed)
    model = SensitiveNet(C, hw).eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, 3, hw, hw, generator=g)
    with torch.no_grad():
        y = model(x).max(1)[1]
    y[::5] = (y[::5] + 1) % C
    return model, x, y


@pytest.mark.parametrize('norm,eps', [('Linf', 8 / 255.), ('L2', 0.5), ('L1', 6.)])
def test_standard_evaluation_matches_oracle(cuda_dev, norm, eps):
    """AutoAttack(version='standard') with attacks_to_run = [apgd-ce, apgd-t] (AA_eval.py:226-239), reduced budget:
    >= 99.9 % of the pixels of x_adv within 1e-6 of the oracle's (a ~0 gradient may change sign between two conv
    implementations and moves that pixel by a full step), robust flags equal on >= 99.9 % (here: all) samples."""
    import autoattack                                              # the drop-in shim package
    model, x, y = _setup(0, 20)
    gm = SensitiveNet(10, 16).eval()
    gm.load_state_dict(model.state_dict())
    gm = gm.to(cuda_dev)
    adv = autoattack.AutoAttack(gm, norm=norm, eps=eps, version='standard', seed=7, verbose=False, device=cuda_dev)
    adv.attacks_to_run = ['apgd-ce', 'apgd-t']
    n_iter = 10 if norm == 'L1' else 12
    adv.apgd.n_iter = adv.apgd_targeted.n_iter = n_iter
    adv.apgd.n_iter_orig = adv.apgd_targeted.n_iter_orig = n_iter
    adv.apgd.n_restarts = 1
    adv.apgd_targeted.n_target_classes = 3
    adv.apgd.rng_device = adv.apgd_targeted.rng_device = 'cpu'
    x_adv = adv.run_standard_evaluation(x.to(cuda_dev), y.to(cuda_dev), bs=8).cpu()

    robust = torch.zeros(20, dtype=torch.bool)
    with torch.no_grad():
        robust[:] = model(x).max(1)[1] == y
    ref = x.clone()
    l1 = norm == 'L1'
    for name in ('apgd-ce', 'apgd-t'):
        idcs = robust.nonzero().squeeze(1)
        for s in range(0, idcs.numel(), 8):
            bi = idcs[s:s + 8]
            if name == 'apgd-ce':
                a = ao.apgd_perturb(model, x[bi], y[bi], norm, eps, n_iter, 1, 'ce', 7, largereps=l1)
            else:
                a = ao.apgd_targeted_perturb(model, x[bi], y[bi], norm, eps, n_iter, 1, 3, 7, largereps=l1)
            with torch.no_grad():
                fb = ~(model(a).max(1)[1] == y[bi])
            robust[bi[fb]] = False
            ref[bi[fb]] = a[fb]
    with torch.no_grad():
        got = model(x_adv).max(1)[1] == y
    agree = (got == robust).float().mean().item()
    tol = 1e-6 if norm != 'L1' else 1e-5
    frac = ((x_adv - ref).abs() <= tol).float().mean().item()
    assert agree >= 0.999, agree
    assert frac >= (0.999 if norm != 'L1' else 0.99), frac
    d = (x_adv - x).flatten(1)
    if norm == 'Linf':
        assert d.abs().max() <= eps + 1e-6
    elif norm == 'L2':
        assert (d ** 2).sum(1).sqrt().max() <= eps * (1 + 1e-4)
    else:
        assert d.abs().sum(1).max() <= eps * (1 + 1e-4)
    assert x_adv.min() >= 0 and x_adv.max() <= 1


def test_evaluation_on_convnext_engine(cuda_dev):
    """the whole path on the B200 engine (ConvNeXt-T-CvSt, bf16 kernels): runs, stays in the ball, never increases the
    robust count, and every reported adversarial point is misclassified by the engine"""
    import autoattack
    from revisiting_at_b200 import convnext
    m = convnext.build('convnext_tiny', normalize=True, seed=0).to(cuda_dev).eval()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(6, 3, 64, 64, generator=g).to(cuda_dev)
    with torch.no_grad():
        y = m(x).float().max(1)[1]
    eps = 4 / 255.
    adv = autoattack.AutoAttack(m, norm='Linf', eps=eps, version='standard', seed=1, verbose=False, device=cuda_dev)
    adv.attacks_to_run = ['apgd-ce', 'apgd-t']
    adv.apgd.n_iter = adv.apgd_targeted.n_iter = 10
    adv.apgd_targeted.n_target_classes = 2
    x_adv, y_adv = adv.run_standard_evaluation(x, y, bs=4, return_labels=True)
    assert (x_adv - x).abs().max() <= eps + 1e-6 and x_adv.min() >= 0 and x_adv.max() <= 1
    changed = (x_adv != x).flatten(1).any(1)
    with torch.no_grad():
        pred = m(x_adv).float().max(1)[1]
    assert bool((pred[changed] != y[changed]).all())
    assert adv.results['clean'] == 1.0 and adv.results['apgd-t'] <= adv.results['apgd-ce'] <= 1.0


def test_l1_projection_export(cuda_dev):
    """`L1_projection` (exported next to apgd_train, autopgd_train_clean.py:24-91) vs its CPU restatement"""
    import autopgd_train_clean as product
    from oracle.apgd_oracle import l1_projection_rows
    g = torch.Generator().manual_seed(3)
    x = torch.rand(5, 3, 12, 12, generator=g)
    d = torch.randn(5, 3, 12, 12, generator=g) * 0.3
    d[0] *= 0.001                                    # already inside the ball: only the box matters
    want = l1_projection_rows(x, d, 4.0)
    got = product.L1_projection(x.to(cuda_dev), d.to(cuda_dev), 4.0).cpu()
    assert (got - want).abs().max() <= 2e-6, (got - want).abs().max()
    z = x + d + got
    assert (z - x).abs().flatten(1).sum(1).max() <= 4.0 * (1 + 1e-5) and z.min() >= -1e-6 and z.max() <= 1 + 1e-6

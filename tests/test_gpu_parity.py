"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the product entry
points (`autopgd_train_clean.apgd_train`, `fgsm_train.fgsm_train`) or the raw C ABI; the oracle and the
golden fixtures are only the checkers.  Nothing here reads /root/reference.

Bars (BASELINE.json north_star):
  * identical gradients in (scripted model)  -> iterates, x_best, x_best_adv, acc bit-exact
    (numeric equality; signed zeros are outside the contract, see oracle/apgd_oracle.py);
  * real model in fp32 -> |x_best - ref| <= 1e-6 on >= 99.9 % of pixels (a gradient that is ~0 can change
    sign between two conv implementations, which moves that pixel by a full step), masks agree.
"""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, same
from oracle import apgd_oracle as ao
from oracle.scripted_model import ScriptedModel

pytestmark = pytest.mark.gpu


def _t(a, dev=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(dev) if dev is not None else t


@pytest.fixture(scope='module')
def product(cuda_dev):
    import autopgd_train_clean
    return autopgd_train_clean


def _scripted_names():
    from revisiting_at_b200 import _abi
    return [n for n in golden_names('scripted_')]


@pytest.mark.parametrize('name', golden_names('scripted_'))
def test_scripted_bit_exact(product, cuda_dev, name):
    g = golden(name)
    norm = str(g['norm'])
    from revisiting_at_b200 import _abi
    if norm != 'Linf' and not hasattr(_abi, norm.lower() + '_step'):
        pytest.skip(f'{norm} kernels not built yet')
    model = ScriptedModel(_t(g['logits'], cuda_dev), _t(g['grads'], cuda_dev))
    out = product.apgd_train(model, _t(g['x'], cuda_dev), _t(g['y'], cuda_dev), norm, float(g['eps']),
                             n_iter=int(g['n_iter']), loss=str(g['loss']),
                             mixup=(object() if bool(g['soft']) else None), is_train=bool(g['is_train']))
    torch.cuda.synchronize()
    seen = torch.stack(model.seen).cpu()
    x_best, acc, loss_best, x_best_adv = [o.cpu() for o in out]
    assert out[0].dtype == torch.float32 and out[1].dtype == torch.bool and not out[0].requires_grad
    if norm == 'Linf':
        assert same(seen, _t(g['x_calls'])), 'iterate trajectory differs'
        assert same(x_best, _t(g['x_best']))
        assert same(x_best_adv, _t(g['x_best_adv']))
    else:  # dependent per-sample reductions: tolerance 1e-6 absolute (north_star)
        assert (seen - _t(g['x_calls'])).abs().max() <= 1e-6
        assert (x_best - _t(g['x_best'])).abs().max() <= 1e-6
        assert (x_best_adv - _t(g['x_best_adv'])).abs().max() <= 1e-6
    assert same(acc, _t(g['acc']))
    assert torch.allclose(loss_best, _t(g['loss_best']), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize('name', golden_names('cnn_'))
def test_cnn_loop_fp32(product, cuda_dev, name):
    from oracle.small_cnn import from_fixture
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden(name)
    from revisiting_at_b200 import _abi
    if str(g['norm']) != 'Linf' and not hasattr(_abi, str(g['norm']).lower() + '_step'):
        pytest.skip('kernels not built yet')
    model = from_fixture(g).to(cuda_dev)
    out = product.apgd_train(model, _t(g['x'], cuda_dev), _t(g['y'], cuda_dev), str(g['norm']), float(g['eps']),
                             n_iter=int(g['n_iter']))
    x_best, acc, loss_best, x_best_adv = [o.cpu() for o in out]
    frac = ((x_best - _t(g['x_best'])).abs() <= 1e-6).float().mean().item()
    assert frac >= 0.999, f'{name}: only {frac:.5f} of pixels within 1e-6'
    assert same(acc, _t(g['acc']))
    assert torch.allclose(loss_best, _t(g['loss_best']), atol=1e-3)
    assert all(p.grad is None for p in model.parameters())


@pytest.mark.parametrize('name', golden_names('scripted_linf'))
def test_scripted_bit_exact_copying_kernels(cuda_dev, name):
    """same fixtures forced through the copying kernels (b200at_linf_step + flush), the path of long attacks"""
    from revisiting_at_b200 import attack
    g = golden(name)
    model = ScriptedModel(_t(g['logits'], cuda_dev), _t(g['grads'], cuda_dev))
    out = attack.run_apgd(attack._CUDA, model, _t(g['x'], cuda_dev), _t(g['y'], cuda_dev), 'Linf', float(g['eps']),
                          n_iter=int(g['n_iter']), loss=str(g['loss']),
                          mixup=(object() if bool(g['soft']) else None), log_slots=0)
    assert same(torch.stack(model.seen), _t(g['x_calls']))
    assert same(out[0], _t(g['x_best'])) and same(out[3], _t(g['x_best_adv'])) and same(out[1], _t(g['acc']))


def test_raw_log_kernels_vs_host_bodies(cuda_dev):
    """b200at_linf_step_log / b200at_gather_best with every slot combination against the host build."""
    from hostcheck.backend import HostBackend
    from revisiting_at_b200 import _abi
    hb = HostBackend(4)
    B, shape, eps = 27, (3, 32, 32), 4 / 255.
    g = torch.Generator().manual_seed(9)
    x = torch.rand(B, *shape, generator=g)
    xs = [(x + (torch.rand(B, *shape, generator=g) * 2 - 1) * eps).clamp(0, 1) for _ in range(3)]
    gs = [torch.randn(B, *shape, generator=g) * 1e-3 for _ in range(2)]
    st = torch.zeros(_abi.ST_ROWS, B)
    st[_abi.ST_STEP] = torch.tensor([2 * eps, eps, eps / 2] * 9)
    idx = torch.arange(B, dtype=torch.int32)
    for row, v in ((_abi.ST_IDX_CUR, idx % 3), (_abi.ST_IDX_OLD, (idx // 3) % 3), (_abi.ST_GIDX_CUR, (idx // 9) % 2),
                   (_abi.ST_IDX_BEST, (idx + 1) % 3), (_abi.ST_IDX_BEST_ADV, (idx // 2) % 3)):
        st[row] = v.view(torch.float32)
    want, wb, wa = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    hb.linf_step_log(x, xs, gs + [gs[0]], want, st, eps, 0.75)
    hb.gather_best(xs, wb, wa, st)
    d = lambda t: t.to(cuda_dev)
    got, gb_, ga = torch.empty_like(x, device=cuda_dev), torch.empty_like(x, device=cuda_dev), torch.empty_like(x, device=cuda_dev)
    _abi.linf_step_log(d(x), [d(t) for t in xs], [d(t) for t in gs], got, d(st), eps, 0.75)
    _abi.gather_best([d(t) for t in xs], gb_, ga, d(st))
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), want) and torch.equal(gb_.cpu(), wb) and torch.equal(ga.cpu(), wa)


def test_raw_linf_kernel_vs_host_bodies_all_flag_combinations(cuda_dev):
    """Raw C ABI, 16 x 3x224x224, every pending-flag combination, against the host build of the bodies."""
    from hostcheck.backend import HostBackend
    from revisiting_at_b200 import _abi
    hb = HostBackend(4)
    B, shape = 16, (3, 224, 224)
    g = torch.Generator().manual_seed(3)
    eps = 4 / 255.
    x = torch.rand(B, *shape, generator=g)
    xa = (x + (torch.rand(B, *shape, generator=g) * 2 - 1) * eps).clamp(0, 1)
    xo = (xa + (torch.rand(B, *shape, generator=g) * 2 - 1) * eps).clamp(0, 1)
    gr = torch.randn(B, *shape, generator=g) * 1e-3
    gr[torch.rand(B, *shape, generator=g) < 0.1] = 0.
    xb, gb, xba = torch.rand(B, *shape, generator=g), torch.randn(B, *shape, generator=g), torch.rand(B, *shape, generator=g)
    st = torch.zeros(_abi.ST_ROWS, B)
    st[_abi.ST_STEP] = torch.tensor([2 * eps, eps, eps / 2, eps / 4] * 4)
    st[_abi.ST_FLAGS] = torch.tensor(list(range(8)) * 2, dtype=torch.int32).view(torch.float32)
    for a in (1.0, 0.75):
        h = [t.clone() for t in (x, xa, xo, gr, xb, gb, xba)]
        d = [t.clone().to(cuda_dev) for t in (x, xa, xo, gr, xb, gb, xba)]
        hb.linf_step(h[0], h[1], h[2], h[2], h[3], h[4], h[5], h[6], st, eps, a)
        _abi.linf_step(d[0], d[1], d[2], d[2], d[3], d[4], d[5], d[6], st.to(cuda_dev), eps, a)
        torch.cuda.synchronize()
        for name, th, td in zip(('x', 'x_adv', 'x_new', 'grad', 'x_best', 'grad_best', 'x_best_adv'), h, d):
            assert torch.equal(th, td.cpu()), f'a={a}: {name} differs'


def test_full_size_linf_step_matches_eager_and_invariants(cuda_dev):
    """BASELINE config size (128 x 3x224x224): fused kernel == eager torch restatement bit for bit,
    and the result stays inside the fp32 eps-ball and [0,1]."""
    from revisiting_at_b200 import _abi
    B, shape, eps = 128, (3, 224, 224), 4 / 255.
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.rand(B, *shape, generator=g, device=cuda_dev)
    xa = (x + (torch.rand(B, *shape, generator=g, device=cuda_dev) * 2 - 1) * eps).clamp(0, 1)
    xo = (xa + (torch.rand(B, *shape, generator=g, device=cuda_dev) * 2 - 1) * eps).clamp(0, 1)
    gr = torch.randn(B, *shape, generator=g, device=cuda_dev) * 1e-3
    gr[torch.rand(B, *shape, generator=g, device=cuda_dev) < 0.1] = 0.
    step = torch.tensor([2 * eps, eps, eps / 2, eps / 4] * 32, device=cuda_dev)
    st = torch.zeros(_abi.ST_ROWS, B, device=cuda_dev)
    st[_abi.ST_STEP] = step
    want = ao.linf_update(x, xa, xo, gr, step, eps, 0.75)
    xb, gb, xba = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    new = torch.empty_like(x)
    _abi.linf_step(x, xa, xo, new, gr, xb, gb, xba, st, eps, 0.75)
    torch.cuda.synchronize()
    assert torch.equal(new, want)
    eps32 = torch.tensor(eps, dtype=torch.float32, device=cuda_dev)
    assert bool((new >= x - eps32).all()) and bool((new <= x + eps32).all())
    assert bool((new >= 0).all()) and bool((new <= 1).all()) and not bool(torch.isnan(new).any())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize('soft', [False, True])
def test_loss_kernel_vs_torch(cuda_dev, dtype, soft):
    from revisiting_at_b200 import _abi
    B, C = 37, 1000
    g = torch.Generator().manual_seed(11)
    z = (torch.randn(B, C, generator=g) * 3).to(dtype).to(cuda_dev)
    yh = torch.randint(0, C, (B,), generator=g).to(cuda_dev)
    if soft:
        oh = torch.nn.functional.one_hot(yh, C).float() * 0.9 + 0.1 / C
        y = 0.3 * oh + 0.7 * oh.flip(0)
    else:
        y = yh
    zf = z.float().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(zf, y, reduction='none')
    (dwant,) = torch.autograd.grad(want.sum(), zf)
    dl = torch.empty_like(z)
    lo = torch.empty(B, device=cuda_dev)
    st = torch.zeros(_abi.ST_ROWS, B, device=cuda_dev)
    ls = torch.zeros(1, B, device=cuda_dev)
    _abi.loss_bookkeep(z, y, dl, lo, st, ls, -1, 1, 0, 'Linf', 'ce', 0.1, 0.01, 10)
    torch.cuda.synchronize()
    assert torch.allclose(lo, want.detach(), atol=2e-5, rtol=2e-6)
    tol = {torch.float32: 1e-6, torch.bfloat16: 4e-3, torch.float16: 5e-4}[dtype]   # half an ulp of |dz| <= 1
    assert (dl.float() - dwant).abs().max() <= tol
    label = y.max(1)[1] if soft else y
    assert torch.equal(st[_abi.ST_PRED].view(torch.int32) != 0, z.float().max(1)[1] == label)
    assert torch.equal(st[_abi.ST_LOSS_BEST], lo)
    assert bool((st[_abi.ST_FLAGS].view(torch.int32) == 3).all())


def test_dlr_kernel_vs_oracle(cuda_dev):
    from revisiting_at_b200 import _abi
    B, C = 33, 1000
    g = torch.Generator().manual_seed(12)
    z = (torch.randn(B, C, generator=g) * 3).to(cuda_dev)
    y = torch.randint(0, C, (B,), generator=g).to(cuda_dev)
    y[:10] = z[:10].argmax(1)          # both branches of the "other" selection
    zf = z.clone().requires_grad_(True)
    want = ao.dlr_rows(zf, y)
    (dwant,) = torch.autograd.grad(want.sum(), zf)
    dl, lo = torch.empty_like(z), torch.empty(B, device=cuda_dev)
    st = torch.zeros(_abi.ST_ROWS, B, device=cuda_dev)
    _abi.loss_bookkeep(z, y, dl, lo, st, torch.zeros(1, B, device=cuda_dev), -1, 1, 0, 'Linf', 'dlr', 0.1, 0.01, 10)
    torch.cuda.synchronize()
    assert torch.allclose(lo, want.detach(), atol=1e-6, rtol=1e-6)
    assert torch.allclose(dl, dwant, atol=1e-6, rtol=1e-5)


def test_bookkeeping_kernel_vs_host_state_machine(cuda_dev):
    """Random loss/pred sequences: the device state block equals the host build of the same function
    after every call (acc, loss_best, flags, step halving, checkpoint memory)."""
    from hostcheck.backend import HostBackend
    from revisiting_at_b200 import _abi, attack
    hb = HostBackend()
    B, C, n_iter = 64, 10, 25
    sched = attack.checkpoint_schedule('Linf', n_iter)
    g = torch.Generator().manual_seed(21)
    y = torch.randint(0, C, (B,), generator=g)
    st_h = torch.zeros(_abi.ST_ROWS, B)
    st_h[_abi.ST_STEP] = 0.03
    st_h[_abi.ST_REDUCED_LAST] = 1.
    st_d = st_h.clone().to(cuda_dev)
    ls_h = torch.zeros(n_iter, B)
    ls_d = ls_h.clone().to(cuda_dev)
    for it in range(-1, n_iter):
        z = torch.randn(B, C, generator=g) * 2
        k = sched[it] if it >= 0 else 0
        hb.loss_bookkeep(z, y, None, None, st_h, ls_h, it, n_iter, k, 'Linf', 'ce', 0.03, 0.003, 100)
        _abi.loss_bookkeep(z.to(cuda_dev), y.to(cuda_dev), None, None, st_d, ls_d, it, n_iter, k, 'Linf', 'ce',
                           0.03, 0.003, 100)
        torch.cuda.synchronize()
        for row in (_abi.ST_ACC, _abi.ST_FLAGS, _abi.ST_PRED):
            assert torch.equal(st_h[row].view(torch.int32), st_d[row].cpu().view(torch.int32)), (it, row)
        assert torch.equal(st_h[_abi.ST_STEP], st_d[_abi.ST_STEP].cpu()), it
        assert torch.equal(st_h[_abi.ST_REDUCED_LAST], st_d[_abi.ST_REDUCED_LAST].cpu()), it
        assert torch.allclose(st_h[_abi.ST_LOSS_BEST], st_d[_abi.ST_LOSS_BEST].cpu(), atol=1e-5)


def test_fgsm_matches_reference_fixture(cuda_dev):
    import fgsm_train as product_fgsm
    from revisiting_at_b200 import fgsm
    from oracle.small_cnn import from_fixture
    torch.backends.cudnn.allow_tf32 = False
    g = golden('fgsm_cnn')
    model = from_fixture(g).to(cuda_dev)
    x, y, eps = _t(g['x'], cuda_dev), _t(g['y'], cuda_dev), float(g['eps'])
    for tag, kw in (('plain', dict(use_rs=False)), ('rs', dict(use_rs=True, alpha=1.25, noise_level=1.)),
                    ('rs_skip', dict(use_rs=True, alpha=1.0, noise_level=0.5, skip_projection=True))):
        out = fgsm.run_fgsm(fgsm.CudaFgsmBackend(), model, x, y, eps, noise=_t(g['noise_' + tag], cuda_dev), **kw)
        frac = ((out.cpu() - _t(g['out_' + tag])).abs() <= 1e-6).float().mean().item()
        assert frac >= 0.999, (tag, frac)
    out = product_fgsm.fgsm_train(model, x, y, eps, use_rs=True)
    assert out.shape == x.shape and bool(((out - x).abs() <= eps + 1e-6).all())


def test_edge_shapes(product, cuda_dev):
    """ragged n (scalar path), B=1, n_iter=0 and 1, channels_last input, non-contiguous input."""
    from oracle.small_cnn import SmallCNN
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    cnn = SmallCNN().eval()
    for shape, n_iter in (((1, 3, 7, 5), 3), ((3, 3, 9, 9), 0), ((2, 3, 16, 16), 1), ((5, 3, 10, 6), 4)):
        g = torch.Generator().manual_seed(sum(shape))
        x = torch.rand(*shape, generator=g)
        y = torch.randint(0, 10, (shape[0],), generator=g)
        ref = ao.apgd_train_oracle(cnn, x, y, 'Linf', 8 / 255., n_iter=n_iter)
        got = product.apgd_train(cnn.to(cuda_dev), x.to(cuda_dev), y.to(cuda_dev), 'Linf', 8 / 255., n_iter=n_iter)
        cnn.cpu()
        assert ((got[0].cpu() - ref[0]).abs() <= 1e-6).float().mean() >= 0.995, shape
        assert same(got[1], ref[1])
    x = torch.rand(4, 3, 16, 16)
    y = torch.randint(0, 10, (4,))
    ref = ao.apgd_train_oracle(cnn, x, y, 'Linf', 8 / 255., n_iter=2)
    cnn.to(cuda_dev)
    xcl = x.to(cuda_dev).contiguous(memory_format=torch.channels_last)
    got = product.apgd_train(cnn, xcl, y.to(cuda_dev), 'Linf', 8 / 255., n_iter=2)
    assert ((got[0].cpu() - ref[0]).abs() <= 1e-6).float().mean() >= 0.995
    xnc = torch.rand(4, 3, 16, 32)[..., ::2].to(cuda_dev)
    got = product.apgd_train(cnn, xnc, y.to(cuda_dev), 'Linf', 8 / 255., n_iter=2)
    assert got[0].shape == xnc.shape
    before = xnc.clone()
    product.apgd_train(cnn, xnc, y.to(cuda_dev), 'Linf', 8 / 255., n_iter=2)
    assert torch.equal(before, xnc), 'x must not be mutated'

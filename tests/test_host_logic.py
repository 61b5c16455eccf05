"""CPU: the product's host-side state machine (`attack.run_apgd`) driven through the host-compiled
kernel bodies reproduces the reference's golden trajectories bit-exactly (l-inf path)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden, golden_names, same
from oracle.scripted_model import ScriptedModel
from hostcheck.backend import HostBackend
import revisiting_at_b200
from revisiting_at_b200 import attack


def _t(a):
    return torch.from_numpy(np.asarray(a))


LINF = [n for n in golden_names('scripted_linf')]


@pytest.mark.parametrize('log_slots', [0, 8, 32])     # copying kernels / iterate log (n_iter <= 7) / log for all
@pytest.mark.parametrize('vec', [4, 1])
@pytest.mark.parametrize('name', LINF)
def test_linf_host_path_bit_exact(name, vec, log_slots, monkeypatch):
    g = golden(name)
    if log_slots == 32:
        if int(g['n_iter']) + 1 > 8:
            pytest.skip('more iterates than log slots')
        log_slots = 8
    model = ScriptedModel(_t(g['logits']), _t(g['grads']))
    out = attack.run_apgd(HostBackend(vec), model, _t(g['x']), _t(g['y']), str(g['norm']), float(g['eps']),
                          n_iter=int(g['n_iter']), loss=str(g['loss']),
                          mixup=(object() if bool(g['soft']) else None), log_slots=log_slots)
    assert same(torch.stack(model.seen), _t(g['x_calls'])), 'iterate trajectory differs'
    for got, key in zip(out, ('x_best', 'acc', 'loss_best', 'x_best_adv')):
        assert same(got, _t(g[key])), key


@pytest.mark.parametrize('vec', [4, 1])
@pytest.mark.parametrize('name', golden_names('scripted_l2'))
def test_l2_host_path_within_tolerance(name, vec):
    g = golden(name)
    model = ScriptedModel(_t(g['logits']), _t(g['grads']))
    out = attack.run_apgd(HostBackend(vec), model, _t(g['x']), _t(g['y']), 'L2', float(g['eps']),
                          n_iter=int(g['n_iter']))
    assert (torch.stack(model.seen) - _t(g['x_calls'])).abs().max() <= 1e-6
    assert (out[0] - _t(g['x_best'])).abs().max() <= 1e-6
    assert (out[3] - _t(g['x_best_adv'])).abs().max() <= 1e-6
    assert same(out[1], _t(g['acc'])) and same(out[2], _t(g['loss_best']))


@pytest.mark.parametrize('name', golden_names('scripted_l1'))
def test_l1_host_path_within_tolerance(name):
    g = golden(name)
    model = ScriptedModel(_t(g['logits']), _t(g['grads']))
    out = attack.run_apgd(HostBackend(4), model, _t(g['x']), _t(g['y']), 'L1', float(g['eps']),
                          n_iter=int(g['n_iter']), is_train=bool(g['is_train']))
    err = (torch.stack(model.seen) - _t(g['x_calls'])).abs().max().item()
    assert err <= 1e-6, err
    assert (out[0] - _t(g['x_best'])).abs().max() <= 1e-6
    assert (out[3] - _t(g['x_best_adv'])).abs().max() <= 1e-6
    assert same(out[1], _t(g['acc'])) and same(out[2], _t(g['loss_best']))


def test_schedule_equals_oracle():
    from oracle.apgd_oracle import checkpoint_schedule as ref
    for norm in ('Linf', 'L2', 'L1'):
        for n in (0, 1, 2, 3, 5, 10, 17, 50, 100):
            assert attack.checkpoint_schedule(norm, n) == ref(norm, n)


def test_product_path_refuses_cpu_tensors():
    class M:
        training = False
    with pytest.raises(RuntimeError):
        attack.apgd_train(M(), torch.rand(2, 3, 4, 4), torch.zeros(2, dtype=torch.long), 'Linf', 0.1)


def test_error_behaviour_matches_reference():
    class M:
        training = True
    x, y = torch.rand(2, 3, 4, 4), torch.zeros(2, dtype=torch.long)
    with pytest.raises(AssertionError):
        attack.run_apgd(HostBackend(), M(), x, y, 'Linf', 0.1)
    M.training = False
    with pytest.raises(KeyError):
        attack.run_apgd(HostBackend(), M(), x, y, 'Linf', 0.1, loss='nope')
    with pytest.raises(TypeError):
        attack.run_apgd(HostBackend(), M(), x, y, 'Linf', 0.1, use_rs=True)


@pytest.mark.parametrize('vec', [4, 1])
def test_fgsm_host_path_bit_exact(vec):
    from oracle.small_cnn import from_fixture
    from revisiting_at_b200 import fgsm
    g = golden('fgsm_cnn')
    model = from_fixture(g)
    x, y, eps = _t(g['x']), _t(g['y']), float(g['eps'])
    for tag, kw in (('plain', dict(use_rs=False)), ('rs', dict(use_rs=True, alpha=1.25, noise_level=1.)),
                    ('rs_skip', dict(use_rs=True, alpha=1.0, noise_level=0.5, skip_projection=True))):
        out = fgsm.run_fgsm(HostBackend(vec), model, x, y, eps, noise=_t(g['noise_' + tag]), **kw)
        assert same(out, _t(g['out_' + tag])), tag


def test_derived_parameter_copies_follow_a_fused_optimizer_step():
    """torch's fused optimisers update parameters in place without moving `_version`; the kernel-side copies
    (bf16 / transposed / folded weights) must still be rebuilt after `optimizer.step()`."""
    from revisiting_at_b200 import ops
    p = torch.nn.Parameter(torch.randn(8, 8))
    first = ops._derived(p, 'unit_test_copy', lambda w: w.clone())
    assert ops._derived(p, 'unit_test_copy', lambda w: w.clone()) is first          # cached
    try:
        opt = torch.optim.AdamW([p], lr=0.1, fused=True)
    except RuntimeError:
        pytest.skip('no fused optimiser on this device')
    p.grad = torch.ones_like(p)
    v = p._version
    opt.step()
    again = ops._derived(p, 'unit_test_copy', lambda w: w.clone())
    assert torch.equal(again, p.detach()) and not torch.equal(again, first), (v, p._version)
    with torch.no_grad():
        p.mul_(2.)                                                                    # plain in-place ops: version check
    assert torch.equal(ops._derived(p, 'unit_test_copy', lambda w: w.clone()), p.detach())


def test_conv3x3s2_weight_layout_is_the_implicit_gemm_of_the_convolution():
    """b200at_conv3x3s2_fwd's contract (include/b200at_model.h): with wk = ops._conv3x3s2_wk(w), the convolution equals the
    GEMM of the per-tap strided input rows (channels padded to 64 per tap) with wk^T -- checked here on the CPU against
    F.conv2d of the same bf16 operands (utils_architecture.py:205-211, Conv2d(k=3, s=2, p=1))."""
    import torch.nn.functional as F
    from revisiting_at_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, H, W, Ci, Co = 2, 12, 16, 24, 32
    x = torch.randn(B, H, W, Ci, generator=g).to(torch.bfloat16)
    w = torch.randn(Co, Ci, 3, 3, generator=g) * 0.1
    wk = ops._conv3x3s2_wk(w)
    assert wk.shape == (Co, 576) and wk.dtype == torch.bfloat16
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))                       # zero ring = the TMA fill
    rows = torch.zeros(B, H // 2, W // 2, 9, 64)
    for kh in range(3):
        for kw in range(3):
            rows[..., kh * 3 + kw, :Ci] = xp[:, :, kh:kh + H:2, kw:kw + W:2].permute(0, 2, 3, 1)[:, :H // 2, :W // 2]
    got = rows.reshape(-1, 576) @ wk.float().t()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), None, stride=2, padding=1)
    assert torch.allclose(got.view(B, H // 2, W // 2, Co), ref.permute(0, 2, 3, 1), atol=1e-4, rtol=1e-4)
    assert float(wk.view(Co, 9, 64)[:, :, Ci:].abs().max()) == 0.0                # the padded channels multiply zero-filled input


def test_gelu_polynomial_constants_in_the_kernel_header():
    """csrc/b200at_gelu.cuh: Phi(-|v|) = 2^P7(-|v|) with the coefficients B200AT_GELU_P0..P7 (profiles/fit_gelu_poly.py).
    The header's constants, evaluated the way the kernel does (fp32 Horner with fused multiply-adds, input clamped at
    -6.5), must give GELU (models/convnext.py:43 nn.GELU(), exact erf form) to 2e-7 absolute over [-12, 12]."""
    import re
    import numpy as np
    from math import erf, sqrt
    src = open(os.path.join(os.path.dirname(__file__), '..', 'revisiting-at_b200', 'csrc', 'b200at_gelu.cuh')).read()
    d = [float(re.search(r'#define B200AT_GELU_P%d \(([-+0-9.e]+)f\)' % k, src).group(1)) for k in range(8)]
    v = np.linspace(-12, 12, 96001)
    nax = -np.abs(v)
    t = np.maximum(nax, -6.5).astype(np.float32)
    acc = np.full(t.shape, np.float32(d[7]), dtype=np.float32)
    for k in range(6, -1, -1):
        acc = (acc.astype(np.float64) * t.astype(np.float64) + np.float64(np.float32(d[k]))).astype(np.float32)
    got = nax * np.exp2(acc.astype(np.float64)) + np.maximum(v, 0)
    ref = np.array([0.5 * x * (1 + erf(x / sqrt(2))) for x in v])
    assert float(np.abs(got - ref).max()) < 2e-7

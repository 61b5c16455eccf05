"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- build container only.

    python oracle/make_goldens.py

Every vector below is an input/output pair of `/root/reference/autopgd_train_clean.py:apgd_train`
(or `fgsm_train.py:fgsm_train`, `models/convnext.py` + `utils_architecture.py:ConvBlock1`) executed
here on CPU in fp32.  The reference has no tests or golden vectors of its own, so
these are the pins for the oracle and for the CUDA path.  Fixtures are small on purpose (about 5 MB in total).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, convnext_oracle            # noqa: E402
from oracle.scripted_model import ScriptedModel           # noqa: E402
from oracle.small_cnn import SmallCNN                     # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def scripted_inputs(seed, B, shape, C, n_calls, soft, in_box=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, *shape, generator=g)
    flat = x.view(-1)
    # edge pixels: exact 0/1, just outside the box (clamped on entry), denormal-ish.  The l1 fixtures keep
    # x inside [0,1]: L1_projection (:24-91) assumes an in-box centre, and its sparsity counts
    # (`L0_norm(x_best - x)`, :353) flip on one-ulp ties that out-of-box / 1e-9 pixels provoke.
    flat[0:6] = torch.tensor([0., 1., 0., 1., 0.5, 0.25]) if in_box else torch.tensor([0., 1., -0.25, 1.5, 1e-9, 1. - 1e-7])
    y = torch.randint(0, C, (B,), generator=g)
    logits = torch.randn(n_calls, B, C, generator=g) * 2.
    # true-class logit boosted on a random half of the calls so that predictions flip both ways
    boost = (torch.rand(n_calls, B, generator=g) < 0.5).float() * 5.
    logits.scatter_add_(2, y.view(1, B, 1).expand(n_calls, B, 1), boost.unsqueeze(-1))
    # sample 0: loss rises every call (never oscillates); sample 1: loss falls every call
    # sample 2 is always correctly classified with a large margin, sample 3 never
    for k in range(n_calls):
        logits[k, 0, y[0]] = 3. - 0.7 * k
        if B > 3:
            logits[k, 1, y[1]] = -3. + 0.7 * k
            logits[k, 2, y[2]] = 25.
            logits[k, 3, y[3]] = -25.
    grads = torch.randn(n_calls, B, *shape, generator=g) * 1e-3
    zero = torch.rand(n_calls, B, *shape, generator=g) < 0.1
    grads[zero] = 0.
    grads.view(n_calls, -1)[:, 7] = -0.0
    if soft:
        lam = 0.3
        oh = torch.nn.functional.one_hot(y, C).float() * 0.9 + 0.1 / C
        y = lam * oh + (1 - lam) * oh.flip(0)
    return x, y, logits, grads


def compute_scripted(ref, norm, eps, n_iter, seed, soft=False, loss='ce', B=8, shape=(3, 12, 12), C=10,
                     is_train=True):
    """One scripted-model run of the reference; returns the fixture dict (also used live by the tests)."""
    x, y, logits, grads = scripted_inputs(seed, B, shape, C, n_iter + 1, soft, in_box=(norm == 'L1'))
    model = ScriptedModel(logits, grads)
    out = ref.apgd_train(model, x, y, norm=norm, eps=eps, n_iter=n_iter, loss=loss,
                         mixup=(object() if soft else None), is_train=is_train)
    x_best, acc, loss_best, x_best_adv = out
    return dict(
        norm=norm, eps=np.float64(eps), n_iter=n_iter, soft=soft, loss=loss, is_train=is_train,
        x=x.numpy(), y=y.numpy(), logits=logits.numpy(), grads=grads.numpy(),
        x_calls=torch.stack(model.seen).numpy(),
        x_best=x_best.numpy(), acc=acc.numpy(), loss_best=loss_best.numpy(), x_best_adv=x_best_adv.numpy())


def run_scripted(ref, name, norm, eps, n_iter, seed, **kw):
    d = compute_scripted(ref, norm, eps, n_iter, seed, **kw)
    np.savez_compressed(os.path.join(OUT, f'scripted_{name}.npz'), **d)
    print(f'scripted_{name}: acc={d["acc"].mean():.2f} loss_best={d["loss_best"].mean():.4f}')


def compute_cnn(ref, norm, eps, n_iter, seed):
    torch.manual_seed(seed)
    model = SmallCNN().eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(8, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (8,), generator=g)
    x_best, acc, loss_best, x_best_adv = ref.apgd_train(model, x, y, norm=norm, eps=eps, n_iter=n_iter)
    sd = {'w_' + k.replace('.', '_'): v.numpy() for k, v in model.state_dict().items()}
    return dict(norm=norm, eps=np.float64(eps), n_iter=n_iter,
                x=x.numpy(), y=y.numpy(), x_best=x_best.numpy(), acc=acc.numpy(),
                loss_best=loss_best.numpy(), x_best_adv=x_best_adv.numpy(), **sd)


def run_cnn(ref, name, norm, eps, n_iter, seed):
    d = compute_cnn(ref, norm, eps, n_iter, seed)
    np.savez_compressed(os.path.join(OUT, f'cnn_{name}.npz'), **d)
    print(f'cnn_{name}: acc={d["acc"].mean():.2f} loss_best={d["loss_best"].mean():.4f}')


def run_convnext(ref):
    """ConvNeXt-T-CvSt, oracle seed-0 weights copied into the reference's vendored model."""
    m_ref, _ = ref_loader.convnext_t_cvst()
    m_or = convnext_oracle.build('convnext_tiny', normalize=False, seed=0)
    km = convnext_oracle.vendored_key_map()
    m_ref.load_state_dict({km[k]: v for k, v in m_or.state_dict().items()})
    csum = float(sum(v.double().abs().sum() for v in m_or.state_dict().values()))
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(4, 3, 64, 64, generator=g)
    y = torch.randint(0, 1000, (4,), generator=g)
    with torch.no_grad():
        logits = m_ref(x)
    x_best, acc, loss_best, x_best_adv = ref.apgd_train(m_ref, x, y, norm='Linf', eps=4 / 255., n_iter=2)
    np.savez_compressed(os.path.join(OUT, 'convnext_t_cvst.npz'), weight_abs_sum=csum, x=x.numpy(), y=y.numpy(),
                        logits=logits.numpy(), x_best=x_best.numpy(), acc=acc.numpy(),
                        loss_best=loss_best.numpy(), x_best_adv=x_best_adv.numpy())
    print('convnext_t_cvst: loss_best', loss_best.tolist())


def run_fgsm():
    for n in ('robustbench', 'autoattack'):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.path.insert(0, ref_loader.REF)
    try:
        fg = ref_loader._load('_ref_fgsm_train', os.path.join(ref_loader.REF, 'fgsm_train.py'))
    finally:
        sys.path.remove(ref_loader.REF)
    torch.manual_seed(5)
    model = SmallCNN().eval()
    g = torch.Generator().manual_seed(6)
    x = torch.rand(8, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (8,), generator=g)
    outs = {}
    for tag, kw in (('plain', dict(use_rs=False)), ('rs', dict(use_rs=True, alpha=1.25, noise_level=1.)),
                    ('rs_skip', dict(use_rs=True, alpha=1.0, noise_level=0.5, skip_projection=True))):
        torch.manual_seed(99)
        noise = torch.rand_like(x)
        torch.manual_seed(99)
        outs['out_' + tag] = fg.fgsm_train(model, x.clone(), y, eps=4 / 255., **kw).detach().numpy()
        outs['noise_' + tag] = noise.numpy()
    sd = {'w_' + k.replace('.', '_'): v.numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, 'fgsm_cnn.npz'), eps=np.float64(4 / 255.), x=x.numpy(), y=y.numpy(),
                        **outs, **sd)
    print('fgsm_cnn done')


# name -> ((norm, eps, n_iter), kwargs of compute_scripted); the tests replay the same table live
SCRIPTED_CASES = {}
for _n in (1, 2, 10):
    SCRIPTED_CASES[f'linf_n{_n}'] = (('Linf', 4 / 255., _n), dict(seed=10 + _n))
    SCRIPTED_CASES[f'l2_n{_n}'] = (('L2', 0.5, _n), dict(seed=20 + _n))
    SCRIPTED_CASES[f'l1_n{_n}'] = (('L1', 12., _n), dict(seed=30 + _n))
SCRIPTED_CASES.update({
    'linf_n25': (('Linf', 8 / 255., 25), dict(seed=41, B=5)),
    'l1_n25': (('L1', 12., 25), dict(seed=42, B=5)),
    'l1_n10_eval': (('L1', 12., 10), dict(seed=43, is_train=False)),
    'linf_n2_soft': (('Linf', 4 / 255., 2), dict(seed=51, soft=True)),
    'linf_n10_soft': (('Linf', 4 / 255., 10), dict(seed=52, soft=True)),
    'linf_n10_dlr': (('Linf', 4 / 255., 10), dict(seed=53, loss='dlr')),
    'linf_n2_b1': (('Linf', 4 / 255., 2), dict(seed=54, B=1)),
})
CNN_CASES = {f'{_norm.lower()}_n5': (_norm, _eps, 5, 7) for _norm, _eps in (('Linf', 4 / 255.), ('L2', 0.5), ('L1', 12.))}


def main():
    assert ref_loader.available(), 'reference not mounted'
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.attack_module()
    for name, (args, kw) in SCRIPTED_CASES.items():
        run_scripted(ref, name, *args, **kw)
    for name, args in CNN_CASES.items():
        run_cnn(ref, name, *args)
    run_convnext(ref)
    run_fgsm()


if __name__ == '__main__':
    main()

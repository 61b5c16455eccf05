"""CPU oracle for the AutoAttack APGD-CE / APGD-T evaluation -- TEST INFRASTRUCTURE ONLY.  **Parity unpinned.**

The reference evaluates with `autoattack.AutoAttack(..., version='standard')`, `attacks_to_run = ['apgd-ce',
'apgd-t']` (AA_eval.py:226-239).  That arithmetic lives in the un-vendored pip dependency autoattack-0.1
(README.md:15; fra31/auto-attack), which is absent from /root/reference and from this image, so no vector of the
reference pins it.  What IS pinned: the per-iteration update, `check_oscillation`, `L1_projection` and both DLR
losses are shared with the reference's own fork `autopgd_train_clean.py` (apgd_train :123-371, dlr_loss :99-104,
dlr_loss_targeted :106-111), whose restatement in `apgd_oracle.py` is checked bit for bit against golden vectors.
This file restates the package's published protocol around those pieces (autoattack/autopgd_base.py
`APGDAttack.attack_single_run / perturb / decr_eps_pgd`, `APGDAttack_targeted.perturb`; autoattack/autoattack.py
`AutoAttack.run_standard_evaluation`) in the package's own style (boolean-mask indexing, `nonzero`), on CPU
tensors.  Differences kept from the package relative to `apgd_train`: random start, the gradient is also taken
on the last iteration, hard labels only, top-k 0.2 for l1.

Random numbers: the package draws its start on the CPU (`torch.rand(x.shape).to(device)`); here every draw comes
from one `torch.Generator` passed by the caller, in the order start(run 1), start(run 2), ... so that a test can
give the product the same stream.
"""
import math

import torch

from .apgd_oracle import (_bc, ce_rows, checkpoint_schedule, dlr_rows, l0_norm_rows, l1_projection_rows, l1_update,
                          l2_update, linf_update)


def dlr_targeted_rows(logits, y, y_target):
    """autopgd_train_clean.py:106-111"""
    zs, _ = logits.sort(dim=1)
    u = torch.arange(logits.shape[0])
    return -(logits[u, y] - logits[u, y_target]) / (zs[:, -1] - .5 * (zs[:, -3] + zs[:, -4]) + 1e-12)


def random_start(x, norm, eps, gen):
    B = x.shape[0]
    if norm == 'Linf':
        t = 2 * torch.rand(x.shape, generator=gen) - 1
        return x + eps * torch.ones_like(x) * (t / (_bc(t.abs().reshape(B, -1).max(1)[0], x) + 1e-12))
    t = torch.randn(x.shape, generator=gen)
    if norm == 'L2':
        return x + eps * torch.ones_like(x) * (t / (_bc((t ** 2).reshape(B, -1).sum(-1).sqrt(), x) + 1e-12))
    return x + t + l1_projection_rows(x, t, eps)


def single_run(model, x, y, norm, eps, n_iter, loss='ce', x_init=None, y_target=None, gen=None, restart_state=False):
    """`attack_single_run`: (x_best, acc, loss_best, x_best_adv)"""
    B, n_fts = x.shape[0], math.prod(x.shape[1:])
    crit = {'ce': lambda z: ce_rows(z, y), 'dlr': lambda z: dlr_rows(z, y),
            'dlr-targeted': lambda z: dlr_targeted_rows(z, y, y_target)}[loss]
    x_adv = (random_start(x, norm, eps, gen) if x_init is None else x_init.clone()).clamp(0., 1.)
    x_best, x_best_adv = x_adv.clone(), x_adv.clone()
    loss_steps = torch.zeros(n_iter, B)

    def evaluate(xa):
        xa = xa.detach().requires_grad_()
        with torch.enable_grad():
            logits = model(xa)
            li = crit(logits)
            g = torch.autograd.grad(li.sum(), [xa])[0].detach()
        return logits.detach(), li.detach(), g

    logits, li, grad = evaluate(x_adv)
    grad_best = grad.clone()
    acc = logits.max(1)[1] == y
    loss_best = li.clone()
    alpha = 2. if norm in ('Linf', 'L2') else 1.
    step = alpha * eps * torch.ones(B)
    x_old = x_adv.clone()
    sched = checkpoint_schedule(norm, n_iter)
    if norm == 'L1':
        if x_init is None or not restart_state:
            topk, sp_old = .2 * torch.ones(B), n_fts * torch.ones(B)
        else:
            sp_old = l0_norm_rows(x_adv - x)
            topk = sp_old / n_fts / 1.5
    loss_best_last, reduced_last = loss_best.clone(), torch.ones_like(loss_best)
    for i in range(n_iter):
        a = 0.75 if i > 0 else 1.0
        if norm == 'Linf':
            x_new = linf_update(x, x_adv, x_old, grad, step, eps, a)
        elif norm == 'L2':
            x_new = l2_update(x, x_adv, x_old, grad, step, eps, a)
        else:
            x_new = l1_update(x, x_adv, grad, step, eps, topk)
        x_old, x_adv = x_adv, x_new
        logits, li, grad = evaluate(x_adv)
        pred = logits.max(1)[1] == y
        acc = torch.min(acc, pred)
        ind_pred = (pred == 0).nonzero().squeeze(1)
        x_best_adv[ind_pred] = x_adv[ind_pred] + 0.
        loss_steps[i] = li
        ind = (li > loss_best).nonzero().squeeze(1)
        x_best[ind] = x_adv[ind].clone()
        grad_best[ind] = grad[ind].clone()
        loss_best[ind] = li[ind] + 0
        k = sched[i]
        if k > 0:
            if norm in ('Linf', 'L2'):
                ups = torch.zeros(B)
                for c in range(k):
                    ups += (loss_steps[(i - c) % n_iter] > loss_steps[(i - c - 1) % n_iter]).float()
                fl = torch.max((ups <= k * .75 * torch.ones_like(ups)).float(),
                               (1. - reduced_last) * (loss_best_last >= loss_best).float())
                reduced_last, loss_best_last = fl.clone(), loss_best.clone()
                idx = (fl > 0).nonzero().squeeze(1)
                step[idx] /= 2.0
            else:
                sp = l0_norm_rows(x_best - x)
                idx = ((sp / sp_old) < .95).nonzero().squeeze(1)
                topk = sp / n_fts / 1.5
                new = step / 1.5
                new[idx] = alpha * eps
                step = new.clamp(alpha * eps / 10., alpha * eps)
                sp_old = sp.clone()
            x_adv = x_adv.clone()
            x_adv[idx] = x_best[idx].clone()
            grad[idx] = grad_best[idx].clone()
    return x_best, acc, loss_best, x_best_adv


def largereps_schedule(eps, n_iter):
    epss = [3. * eps, 2. * eps, 1. * eps]
    iters = [math.ceil(c) for c in (.3 * n_iter, .3 * n_iter, .4 * n_iter)]
    iters[-1] = n_iter - sum(iters[:-1])
    return epss, iters


def decr_eps_pgd(model, x, y, eps, n_iter, loss, y_target, gen):
    epss, iters = largereps_schedule(eps, n_iter)
    x_init = x + torch.randn(x.shape, generator=gen)
    x_init = x_init + l1_projection_rows(x, x_init - x, 1. * float(epss[0]))
    for e, it in zip(epss, iters):
        x_init = x_init + l1_projection_rows(x, x_init - x, 1. * e)
        x_init, acc, lb, x_adv = single_run(model, x, y, 'L1', e, it, loss, x_init=x_init, y_target=y_target,
                                            restart_state=True)
    return x_init, acc, lb, x_adv


def _run(model, x, y, norm, eps, n_iter, loss, y_target, gen, largereps):
    if largereps:
        return decr_eps_pgd(model, x, y, eps, n_iter, loss, y_target, gen)
    return single_run(model, x, y, norm, eps, n_iter, loss, y_target=y_target, gen=gen)


def apgd_perturb(model, x, y, norm, eps, n_iter=100, n_restarts=1, loss='ce', seed=0, largereps=False):
    """`APGDAttack.perturb`"""
    with torch.no_grad():
        y_pred = model(x).max(1)[1]
    adv, acc = x.clone(), y_pred == y
    gen = torch.Generator().manual_seed(seed)
    for _ in range(n_restarts):
        ind = acc.nonzero().squeeze(1)
        if ind.numel() == 0:
            break
        _, acc_c, _, adv_c = _run(model, x[ind].clone(), y[ind].clone(), norm, eps, n_iter, loss, None, gen, largereps)
        bad = (acc_c == 0).nonzero().squeeze(1)
        acc[ind[bad]] = False
        adv[ind[bad]] = adv_c[bad].clone()
    return adv


def apgd_targeted_perturb(model, x, y, norm, eps, n_iter=100, n_restarts=1, n_target_classes=9, seed=0,
                          largereps=False):
    """`APGDAttack_targeted.perturb`"""
    with torch.no_grad():
        y_pred = model(x).max(1)[1]
    adv, acc = x.clone(), y_pred == y
    gen = torch.Generator().manual_seed(seed)
    for target_class in range(2, n_target_classes + 2):
        for _ in range(n_restarts):
            ind = acc.nonzero().squeeze(1)
            if ind.numel() == 0:
                break
            xf, yf = x[ind].clone(), y[ind].clone()
            with torch.no_grad():
                y_target = model(xf).sort(dim=1)[1][:, -target_class]
            _, acc_c, _, adv_c = _run(model, xf, yf, norm, eps, n_iter, 'dlr-targeted', y_target, gen, largereps)
            bad = (acc_c == 0).nonzero().squeeze(1)
            acc[ind[bad]] = False
            adv[ind[bad]] = adv_c[bad].clone()
    return adv


def run_standard_evaluation(model, x_orig, y_orig, norm, eps, bs=250, attacks=('apgd-ce', 'apgd-t'), seed=0, n_iter=100):
    """`AutoAttack(version='standard').run_standard_evaluation` restricted to APGD-CE / APGD-T: (x_adv, robust flags)"""
    N = x_orig.shape[0]
    robust = torch.zeros(N, dtype=torch.bool)
    with torch.no_grad():
        for s in range(0, N, bs):
            robust[s:s + bs] = model(x_orig[s:s + bs]).max(1)[1] == y_orig[s:s + bs]
    x_adv = x_orig.clone()
    l1 = norm == 'L1'
    for name in attacks:
        idcs = robust.nonzero().squeeze(1)
        for s in range(0, idcs.numel(), bs):
            bi = idcs[s:s + bs]
            x, y = x_orig[bi].clone(), y_orig[bi].clone()
            if name == 'apgd-ce':
                adv = apgd_perturb(model, x, y, norm, eps, n_iter, 5 if l1 else 1, 'ce', seed, l1)
            else:
                adv = apgd_targeted_perturb(model, x, y, norm, eps, n_iter, 1, 5 if l1 else 9, seed, l1)
            with torch.no_grad():
                false_batch = ~(model(adv).max(1)[1] == y)
            robust[bi[false_batch]] = False
            x_adv[bi[false_batch]] = adv[false_batch]
    return x_adv, robust

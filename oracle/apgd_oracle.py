"""CPU oracle for the APGD training attack -- TEST INFRASTRUCTURE ONLY.

This file is a restatement (not a copy) of the algorithm in the reference
`autopgd_train_clean.py` (`apgd_train`, :123-371) and `fgsm_train.py`
(`fgsm_train`, :72-98), written as a *sync-free* per-sample state machine in
plain torch-CPU ops: every data-dependent branch of the reference
(`nonzero`, boolean-mask indexing, `if flag.sum() > 0`) is replaced by a
per-sample predicate + `torch.where`.  That formulation is the semantic spec
the CUDA kernels in `revisiting-at_b200/csrc/` implement.

Who may import this: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs.  The product path
(`revisiting-at_b200/`) must never import it.

Pinning: `oracle/make_goldens.py` runs the *unmodified* reference function
(imported from /root/reference in the build container) on scripted-model and
small-CNN inputs and commits inputs+outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this oracle against those vectors on
every CPU run, and its `test_live_*` cases compare with the reference side by
side when /root/reference is mounted.  Parity status: PINNED for the attack
(l-inf / l2 / l1, hard and soft labels, ce and dlr).

Arithmetic contract (reference line numbers in brackets):
  * every op is individually rounded to fp32, no FMA contraction [:214-226];
  * eps enters tensor ops as float32(eps); step = float32(alpha*eps)
    (double product rounded once) [:169-170];
  * sign(NaN) = sign(+-0) = 0; max/min/clamp propagate NaN;
  * signed zeros are NOT part of the contract (torch's own vectorised and
    scalar CPU paths disagree on max(-0., +0.)); comparisons are numeric.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# per-sample reductions  [autopgd_train_clean.py:8-21]
# --------------------------------------------------------------------------
def _rows(t: torch.Tensor) -> torch.Tensor:
    return t.reshape(t.shape[0], -1)


def l1_norm_rows(t):
    return _rows(t.abs()).sum(-1)


def l2_norm_rows(t):
    return _rows(t * t).sum(-1).sqrt()


def l0_norm_rows(t):
    return _rows(t != 0.).sum(-1)


def _bc(v: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """[B] -> [B,1,1,...] broadcastable against `like`."""
    return v.reshape(-1, *([1] * (like.dim() - 1)))


# --------------------------------------------------------------------------
# losses  [autopgd_train_clean.py:94-114]
# --------------------------------------------------------------------------
def ce_rows(logits, y):
    """Per-sample CE; y is int64 [B] or a soft target [B,C] (:113)."""
    return F.cross_entropy(logits, y, reduction='none')


def dlr_rows(logits, y):
    """DLR loss (:99-104): -(z_y - max_{i!=y} z_i) / (z_(1) - z_(3) + 1e-12)."""
    zs, order = logits.sort(dim=1)
    top_is_y = (order[:, -1] == y).to(logits.dtype)
    zy = logits.gather(1, y.view(-1, 1)).squeeze(1)
    other = zs[:, -2] * top_is_y + zs[:, -1] * (1. - top_is_y)
    return -(zy - other) / (zs[:, -1] - zs[:, -3] + 1e-12)


LOSSES = {'ce': ce_rows, 'dlr': dlr_rows}


# --------------------------------------------------------------------------
# l1 projection  [autopgd_train_clean.py:24-91]
# --------------------------------------------------------------------------
def l1_projection_rows(x2: torch.Tensor, y2: torch.Tensor, eps1: float) -> torch.Tensor:
    """delta such that x2+y2+delta lies in {||.-x2||_1 <= eps1} ∩ [0,1]^n.

    Follows the reference's sorted-breakpoint scheme so that rounding of the
    cumulative sums matches it:  with a_i = box excess, b_i = |y_i|, find the
    water level alpha with  sum_i clamp(alpha, a_i, b_i) = ||y||_1 - eps1.
    """
    B = x2.shape[0]
    x = x2.detach().clone().float().reshape(B, -1)
    y = y2.detach().clone().float().reshape(B, -1)
    n = x.shape[1]
    sgn = y.sign()
    u = torch.min(1 - x - y, x + y)
    u = torch.min(torch.zeros_like(y), u)          # -a_i  (:37-39)
    l = -y.abs()                                   # -b_i  (:40)
    d = u.clone()

    bp, src = torch.sort(-torch.cat((u, l), 1), dim=1)          # 2n breakpoints (:43)
    bp_next = torch.cat((bp[:, 1:], torch.zeros(B, 1)), 1)      # (:44)
    active = (2 * (src < n).float() - 1).cumsum(dim=1)          # slope after each bp (:46-47)
    base = -u.sum(dim=1)                                        # g(0) (:49)
    c = eps1 - y.abs().sum(dim=1)                               # (:51)
    need = base + c < 0                                         # (:52)
    g_at_next = base.unsqueeze(-1) + torch.cumsum((bp_next - bp) * active, dim=1)   # (:55)

    rows = need.nonzero().squeeze(1)
    if rows.numel() > 0:
        lo = torch.zeros(rows.numel())
        hi = torch.full_like(lo, 2 * n - 1)
        # the reference evaluates ceil(log2(float32(2n))) in fp32 (:67)
        rounds = int(torch.ceil(torch.log2(torch.tensor(2 * n).float())).item())
        for _ in range(rounds):                                  # (:71-85)
            mid = torch.floor((lo + hi) / 2.)
            below = g_at_next[rows, mid.long()] + c[rows] < 0
            lo = torch.where(below, mid, lo)
            hi = torch.where(below, hi, mid)
        j = lo.long()
        alpha = (-g_at_next[rows, j] - c[rows]) / active[rows, j + 1] + bp_next[rows, j]   # (:88)
        d[rows] = -torch.min(torch.max(-u[rows], alpha.unsqueeze(-1)), -l[rows])           # (:89)
    return (sgn * d).reshape(x2.shape)


# --------------------------------------------------------------------------
# checkpoint schedule (data independent)  [:153-161, 327-349, 364]
# --------------------------------------------------------------------------
def checkpoint_schedule(norm: str, n_iter: int):
    """List of length n_iter: k used at iteration i if i is a checkpoint, else 0."""
    out = [0] * n_iter
    if norm in ('Linf', 'L2'):
        k = max(int(0.22 * n_iter), 1)
        k_min = max(int(0.06 * n_iter), 1)
        dec = max(int(0.03 * n_iter), 1)
    else:
        k = max(int(.04 * n_iter), 1)
        k_min, dec = k, 0
    cnt = 0
    for i in range(n_iter):
        cnt += 1
        if cnt == k:
            out[i] = k
            cnt = 0
            k = max(k - dec, k_min)
    return out


# --------------------------------------------------------------------------
# single-iterate updates
# --------------------------------------------------------------------------
def linf_update(x, x_adv, x_old, grad, step, eps, a):
    """One l-inf APGD move with momentum (:214-226). `step` is [B]."""
    eps32 = torch.tensor(eps, dtype=x.dtype)
    lo, hi = x - eps32, x + eps32

    def proj(v):
        return torch.clamp(torch.min(torch.max(v, lo), hi), 0.0, 1.0)

    g2 = x_adv - x_old
    z = proj(x_adv + _bc(step, x) * torch.sign(grad))
    return proj(x_adv + (z - x_adv) * a + g2 * (1 - a))


def l2_update(x, x_adv, x_old, grad, step, eps, a):
    """One l2 APGD move with momentum (:228-237)."""
    def ball(v):
        dv = v - x
        nv = _bc(l2_norm_rows(dv), x)
        return torch.clamp(x + dv / (nv + 1e-12) * torch.min(eps * torch.ones_like(x), nv), 0.0, 1.0)

    g2 = x_adv - x_old
    z = ball(x_adv + _bc(step, x) * grad / (_bc(l2_norm_rows(grad), x) + 1e-12))
    return ball(x_adv + (z - x_adv) * a + g2 * (1 - a))


def l1_update(x, x_adv, grad, step, eps, topk):
    """One l1 APGD move: sparse sign step + projection, no momentum (:239-250)."""
    B = x.shape[0]
    n_fts = math.prod(x.shape[1:])
    mags = _rows(grad.abs()).sort(-1)[0]
    pos = torch.clamp((1. - topk) * n_fts, min=0, max=n_fts - 1).long()
    thr = _bc(mags[torch.arange(B), pos], x)
    sparse = grad * (grad.abs() >= thr).float()
    s = sparse.sign()
    nnz = _bc(_rows(s.abs()).sum(dim=-1), x)
    moved = x_adv + _bc(step, x) * s / (nnz + 1e-10)
    du = moved - x
    return x + du + l1_projection_rows(x, du, eps)


# --------------------------------------------------------------------------
# the attack
# --------------------------------------------------------------------------
def apgd_train_oracle(model: Callable, x, y, norm, eps, n_iter=10, use_rs=False, loss='ce',
                      verbose=False, mixup=None, is_train=True, trace: Optional[dict] = None):
    """Same contract as the reference `apgd_train` (:123-371):
    returns (x_best, acc, loss_best, x_best_adv).  `trace`, if given, receives
    the per-call iterates / flags for step-level parity tests."""
    assert not model.training                                   # (:125)
    if use_rs:
        raise TypeError("exceptions must derive from BaseException")   # `raise NotImplemented` (:137)
    crit = LOSSES[loss] if loss in LOSSES else _unsupported(loss)
    B = x.shape[0]
    n_fts = math.prod(x.shape[1:])
    sched = checkpoint_schedule(norm, n_iter)

    if norm in ('Linf', 'L2'):
        alpha = 2.
    elif norm == 'L1':
        alpha = 1.
        topk = (.05 if is_train else .2) * torch.ones(B)
        sp_old = n_fts * torch.ones(B)
    else:
        raise UnboundLocalError(f"norm {norm!r}: no step rule")   # reference dies at `alpha * eps` (:169)
    step = (alpha * eps * torch.ones(B, dtype=x.dtype))

    x_adv = x.detach().clone().clamp(0., 1.)
    x_best = x_adv.clone()
    x_best_adv = x_adv.clone()
    loss_steps = torch.zeros(n_iter, B)

    def evaluate(xa, need_grad):
        xa = xa.detach().requires_grad_(need_grad)
        with torch.enable_grad():
            logits = model(xa)
            li = crit(logits, y)
            g = torch.autograd.grad(li.sum(), [xa])[0].detach() if need_grad else None
        return logits.detach(), li.detach(), g

    def correct(logits):
        label = y.max(1)[1] if mixup is not None else y
        return logits.max(1)[1] == label

    logits, li, grad = evaluate(x_adv, True)
    grad_best = grad.clone()
    acc = correct(logits)
    loss_best = li.clone()
    loss_best_last = loss_best.clone()
    reduced_last = torch.ones_like(loss_best)
    x_old = x_adv.clone()
    if trace is not None:
        trace.update(x_calls=[x_adv.clone()], flags=[], steps=[step.clone()])

    for i in range(n_iter):
        a = 0.75 if i > 0 else 1.0
        if norm == 'Linf':
            x_new = linf_update(x, x_adv, x_old, grad, step, eps, a)
        elif norm == 'L2':
            x_new = l2_update(x, x_adv, x_old, grad, step, eps, a)
        else:
            x_new = l1_update(x, x_adv, grad, step, eps, topk)
        x_old, x_adv = x_adv, x_new

        last = i == n_iter - 1
        logits, li, g_new = evaluate(x_adv, not last)           # last backward skipped (:281-283)
        if not last:
            grad = g_new
        pred = correct(logits)
        acc = acc & pred                                        # (:296)
        x_best_adv = torch.where(_bc(~pred, x), x_adv, x_best_adv)   # (:304)

        loss_steps[i] = li
        better = li > loss_best                                 # strict (:321)
        x_best = torch.where(_bc(better, x), x_adv, x_best)
        grad_best = torch.where(_bc(better, x), grad, grad_best)     # stale grad on last iter (:323)
        loss_best = torch.where(better, li, loss_best)

        k = sched[i]
        flag = torch.zeros(B, dtype=torch.bool)
        if k > 0:
            if norm in ('Linf', 'L2'):
                ups = torch.zeros(B)
                for c in range(k):                              # rows wrap like Python indexing (:119)
                    ups += (loss_steps[(i - c) % n_iter] > loss_steps[(i - c - 1) % n_iter]).float()
                osc = (ups <= k * 0.75 * torch.ones_like(ups)).float()
                stalled = (1. - reduced_last) * (loss_best_last >= loss_best).float()
                fl = torch.max(osc, stalled)
                reduced_last = fl.clone()
                loss_best_last = loss_best.clone()
                flag = fl > 0
                step = torch.where(flag, step / 2.0, step)
            else:                                               # l1 sparsity adaptation (:351-364)
                sp = l0_norm_rows(x_best - x)
                flag = (sp / sp_old) < .95
                topk = sp / n_fts / 1.5
                step = torch.where(flag, torch.full_like(step, alpha * eps), step / 1.5)
                step = step.clamp(alpha * eps / 10., alpha * eps)
                sp_old = sp.clone()
            x_adv = torch.where(_bc(flag, x), x_best, x_adv)
            grad = torch.where(_bc(flag, x), grad_best, grad)
        if trace is not None:
            trace['x_calls'].append(x_new.clone())
            trace['flags'].append(dict(pred=pred.clone(), better=better.clone(), flag=flag.clone()))
            trace['steps'].append(step.clone())

    return x_best, acc, loss_best, x_best_adv


def _unsupported(loss):
    if loss in ('softloss', 'dlr-targeted'):
        # present in the reference table (:113-114) but not usable through apgd_train:
        # 'dlr-targeted' needs a third argument, 'softloss' returns a scalar.
        raise TypeError(f"loss {loss!r} cannot be driven through apgd_train")
    raise KeyError(loss)


def fgsm_train_oracle(model, x, y, eps, loss='ce', alpha=1.25, use_rs=False, noise_level=1.,
                      skip_projection=False, noise: Optional[torch.Tensor] = None):
    """Restatement of `fgsm_train` (fgsm_train.py:72-98).  `noise` (U[0,1), same shape
    as x) makes the random start reproducible; default draws `torch.rand_like(x)`."""
    assert not model.training
    if use_rs:
        t = torch.rand_like(x) if noise is None else noise
        x_adv = x + (2. * t - 1.) * eps * noise_level
        if not skip_projection:
            x_adv = x_adv.clamp(0., 1.)
    else:
        x_adv = x.clone()
    if loss != 'ce':
        raise KeyError(loss)                                    # fgsm_train.py:12 rebinds the table to 'ce' only
    xa = x_adv.detach().requires_grad_(True)
    li = ce_rows(model(xa), y)
    grad = torch.autograd.grad(li.sum(), xa)[0].detach()
    x_adv = xa.detach() + alpha * eps * grad.sign()
    if not skip_projection:
        x_adv = x + (x_adv - x).clamp(-eps, eps)
        x_adv = x_adv.clamp(0., 1.)
    return x_adv

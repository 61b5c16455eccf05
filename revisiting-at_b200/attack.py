"""Device-resident APGD state machine behind the reference's `apgd_train` signature.

Reference: /root/reference/autopgd_train_clean.py:123-371.  Same positional/keyword arguments, same
return tuple `(x_best, acc, loss_best, x_best_adv)`, same preconditions and error behaviour.

What is different is everything underneath (SURVEY.md §3.2 -> DESIGN.md):
  * no host synchronisation inside the loop: the three data-dependent host branches of the
    reference (`nonzero` at :304/:321, `if fl.sum() > 0` at :340) are per-sample flags in a
    device-resident state block, produced by one loss+bookkeeping kernel per forward and consumed
    by the next image pass;
  * one HBM pass per iteration over the image-sized tensors (the ~36 eager launches of
    :213-226,:304,:321-324,:345-346 collapse into `b200at_linf_step`);
  * the checkpoint schedule (:153-158,:327-349) is data independent and precomputed on the host.

`run_apgd` is written against a small backend object so the CPU test-suite can drive the same
host logic through the host-compiled kernel bodies (tests/hostcheck); the product entry point
`apgd_train` always binds the CUDA library and refuses non-CUDA inputs.
"""
from __future__ import annotations

import math
import time

import torch

from . import _abi
from .ops import input_grad_only

import os

_SUPPORTED_LOSS = ('ce', 'dlr')
# l-inf attacks with n_iter + 1 <= LOG_SLOTS keep every iterate/gradient in its own buffer and replace all
# masked image copies by per-sample slot indices (b200at_linf_step_log); longer attacks use the copying kernel.
LOG_SLOTS = int(os.environ.get('B200AT_APGD_LOG_SLOTS', str(_abi.LOG_MAX_SLOTS)))


def checkpoint_schedule(norm: str, n_iter: int):
    """k at iteration i if i is a checkpoint else 0 (autopgd_train_clean.py:153-161,327-349,364)."""
    out = [0] * n_iter
    if norm in ('Linf', 'L2'):
        k = max(int(0.22 * n_iter), 1)
        k_min = max(int(0.06 * n_iter), 1)
        dec = max(int(0.03 * n_iter), 1)
    else:
        k = max(int(.04 * n_iter), 1)
        k_min, dec = k, 0
    since = 0
    for i in range(n_iter):
        since += 1
        if since == k:
            out[i], since = k, 0
            k = max(k - dec, k_min)
    return out


class CudaBackend:
    """Product backend: hand-written sm_100a kernels through the C ABI."""
    name = 'cuda'

    def check_input(self, x):
        if not x.is_cuda:
            raise _abi.B200atError('apgd_train: x must live on a CUDA device; this build has no CPU path')
        _abi.lib()

    init = staticmethod(_abi.apgd_init)
    linf_step = staticmethod(_abi.linf_step)
    linf_step_log = staticmethod(_abi.linf_step_log)
    gather_best = staticmethod(_abi.gather_best)
    flush_best = staticmethod(_abi.flush_best)
    loss_bookkeep = staticmethod(_abi.loss_bookkeep)

    def l2_step(self, *a, **k):
        return _abi.l2_step(*a, **k)

    def l1_step(self, *a, **k):
        return _abi.l1_step(*a, **k)


_CUDA = CudaBackend()


def _dense_like_input(x):
    """One dense layout per call (SURVEY.md §7 'tensor layouts'): keep NCHW-contiguous or
    channels_last inputs as they are, densify anything else."""
    if x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        return x
    return x.contiguous()


def run_apgd(be, model, x, y, norm, eps, n_iter=10, use_rs=False, loss='ce', verbose=False, mixup=None,
             is_train=True, log_slots=None, x_init=None, y_target=None, l1_restart_state=False):
    """`x_init`, `y_target`, `l1_restart_state`: the extras of AutoAttack's `APGDAttack.attack_single_run`
    (start point instead of clamp(x); targeted DLR; l1 top-k / sparsity initialised from the start point) --
    reached through `autoattack.py`, never through `apgd_train`, whose signature stays the reference's."""
    assert not model.training                                     # autopgd_train_clean.py:125
    if use_rs:
        raise TypeError('exceptions must derive from BaseException')   # `raise NotImplemented` (:137)
    if loss == 'dlr-targeted' and y_target is not None:
        pass                                                      # autoattack.APGDAttack_targeted
    elif loss not in _SUPPORTED_LOSS:
        if loss in ('softloss', 'dlr-targeted'):
            # in the reference's table (:113-114) but not drivable through apgd_train's call sites
            raise TypeError(f'loss {loss!r} cannot be driven through apgd_train')
        raise KeyError(loss)                                      # criterion_dict[loss] (:149)
    if norm not in ('Linf', 'L2', 'L1'):
        raise UnboundLocalError(f"norm {norm!r}: the reference defines no step rule (:153-169)")
    be.check_input(x)
    t_total = time.time()

    x = _dense_like_input(x.detach())
    if x.dtype != torch.float32:
        raise _abi.B200atError(f'apgd_train: x must be fp32 (got {x.dtype})')
    B = x.shape[0]
    n_fts = math.prod(x.shape[1:])
    dev = x.device
    alpha = 2. if norm in ('Linf', 'L2') else 1.
    step_full = alpha * eps                                       # double product, rounded once (:169)
    step_min = alpha * eps / 10.                                  # l1 adasp_minstep (:166,358)
    topk0 = (.05 if is_train else .2) if norm == 'L1' else 0.
    sched = checkpoint_schedule(norm, n_iter)
    soft = mixup is not None
    if loss in ('dlr', 'dlr-targeted') and (soft or y.dim() != 1):
        raise _abi.B200atError(f"loss {loss!r} needs hard labels")
    tkw = {'y_target': y_target} if loss == 'dlr-targeted' else {}

    log_slots = LOG_SLOTS if log_slots is None else log_slots
    use_log = (norm == 'Linf' and 1 <= n_iter and n_iter + 1 <= min(log_slots, _abi.LOG_MAX_SLOTS)
               and x_init is None)
    grad_buf = None
    state = torch.empty(_abi.ST_ROWS, B, device=dev, dtype=torch.float32)
    loss_steps = torch.zeros(max(n_iter, 1), B, device=dev, dtype=torch.float32)
    scratch = None
    times = {'fp': 0., 'bp': 0.}

    def evaluate(x_cur, it, need_grad, own_grad=False):
        nonlocal grad_buf
        xin = x_cur.detach()
        if need_grad:
            xin.requires_grad_()
        t0 = time.time()
        with (torch.enable_grad() if need_grad else torch.no_grad()), input_grad_only():
            logits = model(xin)
        times['fp'] += time.time() - t0
        lg = logits.detach()
        if not lg.is_contiguous():
            lg = lg.contiguous()
        dl = torch.empty_like(lg) if need_grad else None
        be.loss_bookkeep(lg, y, dl, None, state, loss_steps, it, n_iter, sched[it] if it >= 0 else 0, norm, loss,
                         step_full, step_min, n_fts, **tkw)
        if not need_grad:
            return None
        t0 = time.time()
        (g,) = torch.autograd.grad(logits, [xin], grad_outputs=dl.view_as(logits))   # input-grad only (:185,:283)
        times['bp'] += time.time() - t0
        if g.dtype != torch.float32 or g.shape != x.shape or g.stride() != x.stride():
            # backward handed back another layout (e.g. channels_last from a cuDNN dgrad): one dense copy
            if grad_buf is None or own_grad:
                grad_buf = torch.empty_like(x)
            grad_buf.copy_(g)
            g = grad_buf
        return g

    def finish(x_best, x_best_adv):
        acc = state[_abi.ST_ACC].view(torch.int32) != 0
        loss_best = state[_abi.ST_LOSS_BEST].clone()
        if verbose:
            times['total'] = time.time() - t_total
            print(' '.join(f'{k}={v:.5f} s' for k, v in times.items()))
        return x_best, acc, loss_best, x_best_adv

    if use_log:
        xs = [torch.empty_like(x) for _ in range(n_iter + 1)]       # slot k = iterate k
        be.init(x, xs[0], state, step_full, topk0)
        gs = [evaluate(xs[0], -1, True, own_grad=True)]             # slot k = gradient at iterate k
        for i in range(n_iter):
            be.linf_step_log(x, xs, gs, xs[i + 1], state, eps, 0.75 if i > 0 else 1.0)
            g = evaluate(xs[i + 1], i, i < n_iter - 1, own_grad=True)
            if g is not None:
                gs.append(g)
            if verbose:
                _report(state, i, norm, n_fts)
        x_best, x_best_adv = torch.empty_like(x), torch.empty_like(x)
        be.gather_best(xs, x_best, x_best_adv, state)
        return finish(x_best, x_best_adv)

    buf_a, buf_b = torch.empty_like(x), torch.empty_like(x)
    x_best, x_best_adv, grad_best = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    be.init(x, buf_a, state, step_full, topk0)
    if x_init is not None:
        # autoattack `attack_single_run`: x_adv = x_init.clamp(0, 1) as the first iterate (random start / the
        # previous stage's x_best of the l1 large-eps schedule)
        torch.clamp(x_init.detach().to(torch.float32), 0., 1., out=buf_a)
        if norm == 'L1':
            nnz = (buf_a != x).reshape(B, -1).sum(-1)
            state[_abi.ST_SP_ADV] = nnz.to(torch.int32).view(torch.float32)    # L0(x_adv - x) of the start point
            if l1_restart_state:
                # topk = L0(x_adv - x) / n_fts / 1.5 ; sp_old = L0(x_adv - x)   (autopgd_base.py, x_init branch)
                sp = nnz.to(torch.float32)
                state[_abi.ST_TOPK] = sp / n_fts / 1.5
                state[_abi.ST_SP_OLD] = sp
    grad = evaluate(buf_a, -1, True)
    cur, old, first = buf_a, buf_a, True
    for i in range(n_iter):
        a = 0.75 if i > 0 else 1.0
        new = buf_b if first else old                             # in place over the previous iterate
        if norm == 'Linf':
            be.linf_step(x, cur, old, new, grad, x_best, grad_best, x_best_adv, state, eps, a)
        elif norm == 'L2':
            scratch = be.l2_step(x, cur, old, new, grad, x_best, grad_best, x_best_adv, state, eps, a, scratch)
        else:
            scratch = be.l1_step(x, cur, new, grad, x_best, grad_best, x_best_adv, state, eps, scratch)
        old, cur, first = cur, new, False
        g = evaluate(cur, i, i < n_iter - 1)                      # last backward skipped (:281-283)
        if g is not None:
            grad = g
        if verbose:
            _report(state, i, norm, n_fts)
    be.flush_best(cur, x_best, x_best_adv, state)
    return finish(x_best, x_best_adv)


def _report(state, i, norm, n_fts):
    """verbose line of autopgd_train_clean.py:306-311 (reads device state => synchronises)."""
    s = state.cpu()
    acc = (s[_abi.ST_ACC].view(torch.int32) != 0).float().mean().item()
    extra = ' - topk: {:.2f}'.format(s[_abi.ST_TOPK].mean().item() * n_fts) if norm == 'L1' else ''
    print('iteration: {} - best loss: {:.6f} curr loss {:.6f} - robust accuracy: {:.2%} - step size: {:.5f}{}'.format(
        i, s[_abi.ST_LOSS_BEST].sum().item(), s[_abi.ST_LOSS_CUR].sum().item(), acc,
        s[_abi.ST_STEP].mean().item(), extra))


def apgd_train(model, x, y, norm, eps, n_iter=10, use_rs=False, loss='ce',
               verbose=False, mixup=None, is_train=True):
    """Drop-in for the reference `apgd_train` (autopgd_train_clean.py:123-124)."""
    return run_apgd(_CUDA, model, x, y, norm, eps, n_iter=n_iter, use_rs=use_rs, loss=loss, verbose=verbose,
                    mixup=mixup, is_train=is_train)

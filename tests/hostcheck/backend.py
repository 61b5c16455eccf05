"""TEST INFRASTRUCTURE ONLY: drives `attack.run_apgd` through the host-compiled kernel bodies
(tests/hostcheck/hostcheck.cpp) so the host logic + flag protocol + per-element arithmetic can be
checked on a machine without a GPU.  The loss/dlogits part (a warp kernel on the GPU) is replaced
by torch-CPU here; `b200at_bookkeep_sample` itself is the shared code."""
import ctypes
import os
import sys
from ctypes import c_float, c_int, c_int64, c_void_p

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as _build  # noqa: E402

NORMS = {'Linf': 0, 'L2': 1, 'L1': 2}


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


class HostBackend:
    name = 'hostcheck'

    def __init__(self, vec=4):
        self.L = ctypes.CDLL(_build.build())
        self.vec = vec
        P, Fl, I, I64 = c_void_p, c_float, c_int, c_int64
        self.L.hc_init.argtypes = [P, P, P, I64, I64, Fl, Fl, I]
        self.L.hc_linf_step.argtypes = [P] * 9 + [I64, I64, Fl, Fl, I]
        self.L.hc_flush.argtypes = [P, P, P, P, I64, I64, I]
        self.L.hc_bookkeep.argtypes = [P, P, I64, P, P, I, I, I, I, Fl, Fl, I64, I]
        self.L.hc_linf_step_log.argtypes = [P, P, P, I, P, P, I64, I64, Fl, Fl, I]
        self.L.hc_linf_step_log.restype = None
        self.L.hc_gather_best.argtypes = [P, I, P, P, P, I64, I64, I]
        self.L.hc_gather_best.restype = None
        self.L.hc_l2_step.argtypes = [P] * 10 + [I64, I64, Fl, Fl, I]
        self.L.hc_l2_step.restype = None
        self.L.hc_l1_step.argtypes = [P] * 8 + [I64, I64, Fl]
        self.L.hc_l1_step.restype = None
        self.L.hc_fgsm_start.argtypes = [P, P, P, I64, Fl, Fl, I, I]
        self.L.hc_fgsm_step.argtypes = [P, P, P, P, I64, Fl, Fl, I, I]
        for f in ('hc_init', 'hc_linf_step', 'hc_flush', 'hc_bookkeep', 'hc_fgsm_start', 'hc_fgsm_step'):
            getattr(self.L, f).restype = None

    def _vec(self, n):
        return self.vec if n % self.vec == 0 else 1

    def check_input(self, x):
        assert not x.is_cuda

    def init(self, x, x_adv, state, step0, topk0):
        B, n = x.shape[0], x[0].numel()
        self.L.hc_init(_p(x), _p(x_adv), _p(state), B, n, step0, topk0, self._vec(n))

    def linf_step(self, x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, state, eps, a):
        B, n = x.shape[0], x[0].numel()
        assert grad.is_contiguous()
        self.L.hc_linf_step(_p(x), _p(x_adv), _p(x_old), _p(x_new), _p(grad), _p(x_best), _p(grad_best),
                            _p(x_best_adv), _p(state), B, n, eps, a, self._vec(n))

    def l2_step(self, x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, state, eps, a, scratch):
        B, n = x.shape[0], x[0].numel()
        if scratch is None:
            scratch = torch.zeros(B, 3)
        self.L.hc_l2_step(_p(x), _p(x_adv), _p(x_old), _p(x_new), _p(grad.contiguous()), _p(x_best), _p(grad_best),
                          _p(x_best_adv), _p(state), _p(scratch), B, n, eps, a, self._vec(n))
        return scratch

    def l1_step(self, x, x_adv, x_new, grad, x_best, grad_best, x_best_adv, state, eps, scratch):
        B, n = x.shape[0], x[0].numel()
        assert x_new.data_ptr() != x_adv.data_ptr()
        self.L.hc_l1_step(_p(x), _p(x_adv), _p(x_new), _p(grad.contiguous()), _p(x_best), _p(grad_best),
                          _p(x_best_adv), _p(state), B, n, eps)
        return scratch

    @staticmethod
    def _ptrs(ts):
        return (c_void_p * len(ts))(*[t.data_ptr() for t in ts])

    def linf_step_log(self, x, x_slots, g_slots, x_new, state, eps, a):
        B, n = x.shape[0], x[0].numel()
        assert all(t.is_contiguous() for t in list(x_slots) + list(g_slots))
        self.L.hc_linf_step_log(_p(x), self._ptrs(x_slots), self._ptrs(g_slots), len(x_slots), _p(x_new), _p(state),
                                B, n, eps, a, self._vec(n))

    def gather_best(self, x_slots, x_best, x_best_adv, state):
        B, n = x_best.shape[0], x_best[0].numel()
        self.L.hc_gather_best(self._ptrs(x_slots), len(x_slots), _p(x_best), _p(x_best_adv), _p(state), B, n,
                              self._vec(n))

    def flush_best(self, x_adv, x_best, x_best_adv, state):
        B, n = x_adv.shape[0], x_adv[0].numel()
        self.L.hc_flush(_p(x_adv), _p(x_best), _p(x_best_adv), _p(state), B, n, self._vec(n))

    def l1_projection(self, x, d, eps):
        from oracle.apgd_oracle import l1_projection_rows
        return l1_projection_rows(x, d, eps)

    def loss_bookkeep(self, logits, y, dlogits, loss_out, state, loss_steps, it, n_iter, ckpt_k, norm, loss,
                      step_full, step_min, n_fts, y_target=None):
        z = logits.detach().float().requires_grad_(True)
        if loss == 'ce':
            li = F.cross_entropy(z, y, reduction='none')
        elif loss == 'dlr-targeted':
            from oracle.autoattack_oracle import dlr_targeted_rows
            li = dlr_targeted_rows(z, y, y_target)
        else:
            from oracle.apgd_oracle import dlr_rows
            li = dlr_rows(z, y)
        if dlogits is not None:
            dlogits.copy_(torch.autograd.grad(li.sum(), z)[0])
        label = y.max(1)[1] if y.dim() == 2 else y
        pred = (z.detach().max(1)[1] == label).to(torch.int32).contiguous()
        li = li.detach().contiguous()
        self.L.hc_bookkeep(_p(state), _p(loss_steps), logits.shape[0], _p(li), _p(pred), it, n_iter, ckpt_k,
                           NORMS[norm], step_full, step_min, n_fts, int(dlogits is not None))

    # fgsm backend surface
    def start(self, x, noise, x_adv, eps, noise_level, skip):
        self.L.hc_fgsm_start(_p(x), _p(noise.contiguous()), _p(x_adv), x.numel(), eps, noise_level, int(skip),
                             self._vec(x.numel()))

    def step(self, x, x_adv, grad, out, eps, step, skip):
        self.L.hc_fgsm_step(_p(x), _p(x_adv), _p(grad.contiguous()), _p(out), x.numel(), eps, step, int(skip),
                            self._vec(x.numel()))

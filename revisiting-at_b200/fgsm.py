"""`fgsm_train` behind the reference's signature (/root/reference/fgsm_train.py:72-98):
one-step l-inf attack with optional random start; returns the bare x_adv tensor (main.py:836-842
binds it with partial(eps, use_rs=True, alpha, noise_level, skip_projection))."""
import torch

from . import _abi
from .ops import input_grad_only


class CudaFgsmBackend:
    def check_input(self, x):
        if not x.is_cuda:
            raise _abi.B200atError('fgsm_train: x must live on a CUDA device; this build has no CPU path')
        _abi.lib()

    start = staticmethod(lambda *a: _abi.fgsm_start(*a))
    step = staticmethod(lambda *a: _abi.fgsm_step(*a))
    loss_bookkeep = staticmethod(lambda *a: _abi.loss_bookkeep(*a))


def run_fgsm(be, model, x, y, eps, loss='ce', alpha=1.25, use_rs=False, noise_level=1., skip_projection=False,
             noise=None):
    assert not model.training                                  # fgsm_train.py:74
    if loss != 'ce':
        raise KeyError(loss)                                   # fgsm_train.py:12 rebinds the table to 'ce' only
    be.check_input(x)
    x = x.detach()
    x = x if (x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))) \
        else x.contiguous()
    if x.dtype != torch.float32:
        raise _abi.B200atError(f'fgsm_train: x must be fp32 (got {x.dtype})')
    B = x.shape[0]
    x_adv = torch.empty_like(x)
    if use_rs:
        t = torch.rand_like(x) if noise is None else noise     # same RNG call as the reference (:80)
        be.start(x, t, x_adv, eps, noise_level, skip_projection)
    else:
        x_adv.copy_(x)
    xin = x_adv.requires_grad_()
    with torch.enable_grad(), input_grad_only():
        logits = model(xin)
    lg = logits.detach().contiguous()
    dl = torch.empty_like(lg)
    state = torch.zeros(_abi.ST_ROWS, B, device=x.device, dtype=torch.float32)
    loss_steps = torch.zeros(1, B, device=x.device, dtype=torch.float32)
    be.loss_bookkeep(lg, y, dl, None, state, loss_steps, -1, 1, 0, 'Linf', 'ce', 0., 0., x[0].numel())
    (g,) = torch.autograd.grad(logits, [xin], grad_outputs=dl.view_as(logits))
    if g.dtype != torch.float32 or g.stride() != x.stride():
        g = torch.empty_like(x).copy_(g)
    out = torch.empty_like(x)
    be.step(x, x_adv.detach(), g, out, eps, alpha * eps, skip_projection)
    return out


def fgsm_train(model, x, y, eps, loss='ce', alpha=1.25, use_rs=False,
               noise_level=1., skip_projection=False):
    """Drop-in for the reference `fgsm_train` (fgsm_train.py:72-73)."""
    return run_fgsm(CudaFgsmBackend(), model, x, y, eps, loss=loss, alpha=alpha, use_rs=use_rs,
                    noise_level=noise_level, skip_projection=skip_projection)

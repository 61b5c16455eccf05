"""Scripted model fixture -- TEST INFRASTRUCTURE ONLY.

`apgd_train` accepts any callable with `.training == False`.  This stub returns
prescribed logits on call k and makes autograd return a prescribed dL/dx, so
the reference (autopgd_train_clean.py:179-185, :273-283) and the CUDA path see
IDENTICAL losses and gradients; what is left is exactly the per-step update,
projection, clamp, best-tracking and step-halving arithmetic.
"""
import torch


class _Scripted(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, logits, grad):
        ctx.save_for_backward(grad)
        return logits.clone()

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad.clone(), None, None


class ScriptedModel:
    """logits_seq: [n_calls,B,C]; grad_seq: [n_calls,B,*img]. Records every input."""
    training = False

    def __init__(self, logits_seq: torch.Tensor, grad_seq: torch.Tensor):
        self.logits_seq, self.grad_seq = logits_seq, grad_seq
        self.seen = []

    def __call__(self, x):
        k = len(self.seen)
        self.seen.append(x.detach().clone())
        return _Scripted.apply(x, self.logits_seq[k], self.grad_seq[k])

    def eval(self):
        return self

    def train(self, mode=True):
        raise RuntimeError("scripted model is eval-only")

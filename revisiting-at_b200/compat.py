"""Small public helpers the reference module exports next to `apgd_train`
(autopgd_train_clean.py:8-21, :94-121).  They are API surface, not the hot path: thin torch
expressions on whatever device the caller's tensors live on.  `L1_projection` is the one with real
arithmetic and forwards to the CUDA kernel."""
import torch
import torch.nn.functional as F


def L1_norm(x, keepdim=False):
    z = x.abs().reshape(x.shape[0], -1).sum(-1)
    return z.view(-1, *[1] * (x.dim() - 1)) if keepdim else z


def L2_norm(x, keepdim=False):
    z = (x ** 2).reshape(x.shape[0], -1).sum(-1).sqrt()
    return z.view(-1, *[1] * (x.dim() - 1)) if keepdim else z


def L0_norm(x):
    return (x != 0.).reshape(x.shape[0], -1).sum(-1)


def softloss(x, target):
    return torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1).mean()


def dlr_loss(x, y, reduction='none'):
    top3 = x.topk(3, dim=1).values
    zy = x.gather(1, y.view(-1, 1)).squeeze(1)
    is_top = (x.argmax(1) == y).to(x.dtype)
    other = top3[:, 1] * is_top + top3[:, 0] * (1. - is_top)
    return -(zy - other) / (top3[:, 0] - top3[:, 2] + 1e-12)


def dlr_loss_targeted(x, y, y_target):
    top4 = x.topk(4, dim=1).values
    zy = x.gather(1, y.view(-1, 1)).squeeze(1)
    zt = x.gather(1, y_target.view(-1, 1)).squeeze(1)
    return -(zy - zt) / (top4[:, 0] - .5 * (top4[:, 2] + top4[:, 3]) + 1e-12)


criterion_dict = {'ce': lambda x, y: F.cross_entropy(x, y, reduction='none'), 'softloss': softloss,
                  'dlr': dlr_loss, 'dlr-targeted': dlr_loss_targeted}


def check_oscillation(x, j, k, y5, k3=0.75):
    """[k-window oscillation flag per sample] (autopgd_train_clean.py:116-121)."""
    rows = [(j - c) % x.shape[0] for c in range(k)]
    prev = [(j - c - 1) % x.shape[0] for c in range(k)]
    ups = (x[rows] > x[prev]).float().sum(0)
    return (ups <= k * k3).float()


def L1_projection(x2, y2, eps1):
    """Projection of x2+y2 onto the l1 ball of radius eps1 around x2 intersected with [0,1]^n
    (autopgd_train_clean.py:24-91); returns the correction delta.  CUDA kernel."""
    from . import _abi
    return _abi.l1_projection(x2, y2, eps1)

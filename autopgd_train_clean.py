"""Drop-in module for the reference's `autopgd_train_clean.py` (same module name, same public names).

`main.py:63` does `from autopgd_train_clean import apgd_train`; `fgsm_train.py:5` imports
`criterion_dict`.  Everything here forwards to the B200-native implementation in
`revisiting-at_b200/` (hand-written sm_100a kernels behind include/b200at.h).  CUDA tensors only.
"""
import revisiting_at_b200  # noqa: F401  (registers the package alias)
from revisiting_at_b200.attack import apgd_train, checkpoint_schedule  # noqa: F401
from revisiting_at_b200.compat import (  # noqa: F401
    L0_norm, L1_norm, L2_norm, L1_projection, check_oscillation, criterion_dict, dlr_loss, dlr_loss_targeted,
    softloss)

__all__ = ['apgd_train', 'criterion_dict', 'L1_projection', 'check_oscillation', 'L0_norm', 'L1_norm', 'L2_norm']

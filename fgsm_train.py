"""Drop-in module for the reference's `fgsm_train.py` (`main.py:64`: `from fgsm_train import fgsm_train`)."""
import revisiting_at_b200  # noqa: F401
from revisiting_at_b200.fgsm import fgsm_train  # noqa: F401
from revisiting_at_b200.compat import criterion_dict  # noqa: F401

__all__ = ['fgsm_train', 'criterion_dict']

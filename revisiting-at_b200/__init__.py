"""revisiting-at_b200 -- B200-native APGD adversarial-example hot path.

Host side (Python/PyTorch plumbing) of the sm_100a kernels in `csrc/`, reached through the C ABI
declared in `include/b200at.h`.  Mirrors the reference's interface for this path:

    attack.apgd_train   <- /root/reference/autopgd_train_clean.py:123  (same signature / returns)
    fgsm.fgsm_train     <- /root/reference/fgsm_train.py:72

There is no CPU fallback: without the built CUDA library every entry point raises.
"""
__all__ = ['attack', 'fgsm', '_abi']

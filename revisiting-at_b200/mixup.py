"""Mixup / CutMix on the device, as the reference's train loop applies it to the already-uploaded batch
(main.py:967-968: `images, target = self.mixup_fn(images, target)`; constructed at main.py:599-607 with the
settings of parserr.py:17-40: mixup 0.8, cutmix 1.0, prob 1.0, switch 0.5, mode 'batch', label smoothing from
`training.label_smoothing`, 1000 classes).

The arithmetic is timm 0.8's `timm.data.mixup.Mixup` in 'batch' mode -- an un-vendored dependency (SURVEY §8c,
**parity unpinned**), restated from its published algorithm: one lambda ~ Beta(alpha, alpha) per batch drawn with
numpy's global generator, partner = the batch flipped along dim 0, CutMix box centred uniformly with side
sqrt(1-lambda) of the image and lambda corrected to the clipped box area; targets are the lambda-mix of the
smoothed one-hot rows (on = 1 - s + s/K, off = s/K).  It produces the soft `[B, K]` fp32 targets that switch
`apgd_train` to its soft-label cross-entropy (autopgd_train_clean.py:194-197).  Out-of-place on `x` (timm writes
in place into the loader's batch; a resident synthetic batch must survive).
"""
import numpy as np
import torch


def smoothed_one_hot(target, num_classes, smoothing):
    off = smoothing / num_classes
    on = 1. - smoothing + off
    out = torch.full((target.shape[0], num_classes), off, device=target.device, dtype=torch.float32)
    return out.scatter_(1, target.view(-1, 1).long(), on)


def cutmix_box(height, width, lam, rng=np.random):
    """(yl, yh, xl, xh) and the area-corrected lambda."""
    ratio = np.sqrt(1. - lam)
    cut_h, cut_w = int(height * ratio), int(width * ratio)
    cy = int(rng.randint(0, height))
    cx = int(rng.randint(0, width))
    yl, yh = int(np.clip(cy - cut_h // 2, 0, height)), int(np.clip(cy + cut_h // 2, 0, height))
    xl, xh = int(np.clip(cx - cut_w // 2, 0, width)), int(np.clip(cx + cut_w // 2, 0, width))
    return (yl, yh, xl, xh), 1. - (yh - yl) * (xh - xl) / float(height * width)


class Mixup:
    def __init__(self, mixup_alpha=0.8, cutmix_alpha=1.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode='batch',
                 correct_lam=True, label_smoothing=0.1, num_classes=1000):
        if mode != 'batch' or cutmix_minmax is not None:
            raise ValueError("only the reference's configuration is built: mode='batch', cutmix_minmax=None (parserr.py:27-32)")
        self.mixup_alpha, self.cutmix_alpha = mixup_alpha, cutmix_alpha
        self.mix_prob, self.switch_prob = prob, switch_prob
        self.correct_lam = correct_lam
        self.label_smoothing, self.num_classes = label_smoothing, num_classes
        self.mixup_enabled = True

    def draw(self):
        """(lambda, use_cutmix) for this batch; same order of numpy draws as timm's `_params_per_batch`."""
        lam, use_cutmix = 1., False
        if self.mixup_enabled and np.random.rand() < self.mix_prob:
            if self.mixup_alpha > 0. and self.cutmix_alpha > 0.:
                use_cutmix = np.random.rand() < self.switch_prob
                a = self.cutmix_alpha if use_cutmix else self.mixup_alpha
            elif self.mixup_alpha > 0.:
                a = self.mixup_alpha
            elif self.cutmix_alpha > 0.:
                use_cutmix, a = True, self.cutmix_alpha
            else:
                raise ValueError('one of mixup_alpha > 0, cutmix_alpha > 0 is required')
            lam = float(np.random.beta(a, a))
        return lam, use_cutmix

    def __call__(self, x, target):
        if x.shape[0] % 2:
            raise AssertionError('Batch size should be even when using this')
        lam, use_cutmix = self.draw()
        if lam != 1.:
            if use_cutmix:
                (yl, yh, xl, xh), lam_box = cutmix_box(x.shape[-2], x.shape[-1], lam)
                if self.correct_lam:
                    lam = lam_box
                x = x.clone()
                x[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
            else:
                x = x * lam + x.flip(0) * (1. - lam)
        y1 = smoothed_one_hot(target, self.num_classes, self.label_smoothing)
        return x, y1 * lam + y1.flip(0) * (1. - lam)

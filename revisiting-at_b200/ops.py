"""Layer functions of the ConvNeXt-CvSt forward/backward (reference math: models/convnext.py:37-50,
utils_architecture.py:57-81).  ROUND-1 STATE: these run on torch's library kernels (cuDNN / cuBLAS /
ATen) under the caller's autocast; they are the measured baseline that the hand-written NHWC kernels
(depthwise 7x7, LayerNorm, GELU, layer-scale, tcgen05 GEMM) replace one by one.  Inputs/outputs are
NCHW-shaped tensors in channels_last memory."""
import torch
import torch.nn.functional as F


def _ln_channels_first(x, w, b, eps=1e-6):
    # per-pixel LayerNorm over C on an NCHW-shaped tensor == layer_norm on the NHWC view
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps).permute(0, 3, 1, 2)


def stem_layer(x, cw, cb, lw, lb, stride, mean=None, std=None):
    if mean is not None:
        x = (x - mean) / std
    x = F.conv2d(x.contiguous(memory_format=torch.channels_last), cw, cb, stride=stride, padding=1)
    return F.gelu(_ln_channels_first(x, lw, lb))


def convnext_block(x, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, gamma):
    h = F.conv2d(x, dw_w, dw_b, padding=3, groups=x.shape[1]).permute(0, 2, 3, 1)
    h = F.layer_norm(h, (h.shape[-1],), ln_w, ln_b, 1e-6)
    h = F.linear(F.gelu(F.linear(h, w1, b1)), w2, b2) * gamma
    return x + h.permute(0, 3, 1, 2)


def downsample(x, ln_w, ln_b, cw, cb):
    return F.conv2d(_ln_channels_first(x, ln_w, ln_b), cw, cb, stride=2)


def head(x, ln_w, ln_b, fw, fb):
    return F.linear(F.layer_norm(x.mean((-2, -1)), (x.shape[1],), ln_w, ln_b, 1e-6), fw, fb)

"""ConvNeXt-{T,S,B,L}-CvSt for the attack / adversarial train step.

Architecture the reference builds through timm + its conv stems (utils_architecture.py:174-217,
:241-269; block math = models/convnext.py:37-50).  Parameter names follow timm 0.8
(`stem.stem.N`, `stages.S.downsample.N`, `stages.S.blocks.M.{conv_dw,norm,mlp.fc1,mlp.fc2,gamma}`,
`head.norm`, `head.fc`) so the reference's checkpoints load; `normalize_model` wrapping
(`normalize.mean/std`, `model.*`, utils_architecture.py:86-117) is `with_normalizer=True`.

Every layer goes through `ops.py`, which is where the hand-written sm_100a kernels plug in
(NHWC depthwise conv, LayerNorm, GELU, layer-scale); this module is only shape plumbing.  Activations
are NHWC bf16 from the first stem layer to the head, whatever the caller's autocast state.
"""
import torch
import torch.nn as nn

from . import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

ARCHS = {
    # depths, dims, stem widths, stem strides
    'convnext_tiny': ((3, 3, 9, 3), (96, 192, 384, 768), (48, 96), (2, 2)),
    'convnext_small': ((3, 3, 27, 3), (96, 192, 384, 768), (48, 96), (2, 2)),
    'convnext_base': ((3, 3, 27, 3), (128, 256, 512, 1024), (64, 96, 128), (2, 2, 1)),
    'convnext_large': ((3, 3, 27, 3), (192, 384, 768, 1536), (96, 144, 192), (2, 2, 1)),
}


class _LN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _Conv(nn.Module):
    def __init__(self, cin, cout, k, groups=1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
        self.bias = nn.Parameter(torch.zeros(cout))


class _Linear(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.zeros(cout))


class _Mlp(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1 = _Linear(c, 4 * c)
        self.fc2 = _Linear(4 * c, c)


class _Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_dw = _Conv(c, c, 7, groups=c)
        self.norm = _LN(c)
        self.mlp = _Mlp(c)
        self.gamma = nn.Parameter(1e-6 * torch.ones(c))

    def forward(self, x):
        return ops.convnext_block(x, self.conv_dw.weight, self.conv_dw.bias, self.norm.weight, self.norm.bias,
                                  self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias,
                                  self.gamma)


class _Stage(nn.Module):
    def __init__(self, cin, cout, depth, first):
        super().__init__()
        self.downsample = nn.Identity() if first else nn.ModuleList([_LN(cin), _Conv(cin, cout, 2)])
        self.blocks = nn.ModuleList([_Block(cout) for _ in range(depth)])

    def forward(self, x):
        if not isinstance(self.downsample, nn.Identity):
            ln, conv = self.downsample
            x = ops.downsample(x, ln.weight, ln.bias, conv.weight, conv.bias)
        for b in self.blocks:
            x = b(x)
        return x


class _Stem(nn.Module):
    """ConvBlock1 / ConvBlock3: [conv3x3, LN (channels-first), GELU] * n (utils_architecture.py:174-217)."""
    def __init__(self, widths, strides):
        super().__init__()
        mods, cin = [], 3
        for w in widths:
            mods += [_Conv(cin, w, 3), _LN(w), nn.Identity()]   # slot 3j+2 is the GELU (no parameters)
            cin = w
        self.stem = nn.ModuleList(mods)
        self.strides = tuple(strides)

    def forward(self, x, mean=None, std=None):
        for j, s in enumerate(self.strides):
            conv, ln = self.stem[3 * j], self.stem[3 * j + 1]
            x = ops.stem_layer(x, conv.weight, conv.bias, ln.weight, ln.bias, s, j == 0,
                               mean if j == 0 else None, std if j == 0 else None)
        return x


class _Head(nn.Module):
    def __init__(self, c, n_cls):
        super().__init__()
        self.norm = _LN(c)
        self.fc = _Linear(c, n_cls)

    def forward(self, x):
        return ops.head(x, self.norm.weight, self.norm.bias, self.fc.weight, self.fc.bias)


class ConvNeXtCvSt(nn.Module):
    def __init__(self, arch='convnext_tiny', n_cls=1000):
        super().__init__()
        depths, dims, widths, strides = ARCHS[arch]
        self.arch = arch
        self.stem = _Stem(widths, strides)
        self.stages = nn.ModuleList([_Stage(dims[max(i - 1, 0)], dims[i], depths[i], i == 0) for i in range(4)])
        self.head = _Head(dims[-1], n_cls)
        self.reset_parameters()

    def reset_parameters(self):
        """timm init (trunc-normal .02, zero bias) for the backbone; torch default conv init for the
        stem, which the reference swaps in after timm's init (utils_architecture.py:243-244)."""
        for m in self.modules():
            if isinstance(m, (_Conv, _Linear)):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.zeros_(m.bias)
        for m in self.stem.modules():
            if isinstance(m, _Conv):
                nn.init.kaiming_uniform_(m.weight, a=5 ** 0.5)
                fan_in = m.weight[0].numel()
                nn.init.uniform_(m.bias, -1 / fan_in ** 0.5, 1 / fan_in ** 0.5)

    def forward(self, x, mean=None, std=None):
        x = self.stem(x, mean, std)
        for st in self.stages:
            x = st(x)
        return self.head(x)


class Normalized(nn.Module):
    """`normalize_model` (utils_architecture.py:111-117); the (x-mean)/std is folded into the first stem layer."""
    def __init__(self, model):
        super().__init__()
        self.normalize = nn.Module()
        self.normalize.register_buffer('mean', torch.as_tensor(IMAGENET_MEAN).view(1, 3, 1, 1))
        self.normalize.register_buffer('std', torch.as_tensor(IMAGENET_STD).view(1, 3, 1, 1))
        self.model = model

    def forward(self, x):
        return self.model(x, self.normalize.mean, self.normalize.std)


def get_new_model(modelname, pretrained=False, not_original=True, updated=False):
    """The ConvNeXt branches of the reference factory (utils_architecture.py:225-269) with the CvSt stem."""
    if modelname not in ARCHS:
        raise ValueError(f'{modelname!r}: only the ConvNeXt-CvSt family is built on this path')
    if pretrained:
        raise RuntimeError('no network: load a checkpoint with load_state_dict instead')
    if not not_original:
        raise ValueError('the patch stem (not_original=False) is outside the hot path')
    return ConvNeXtCvSt(modelname)


def build(arch='convnext_tiny', normalize=True, seed=0):
    torch.manual_seed(seed)
    m = ConvNeXtCvSt(arch)
    return Normalized(m) if normalize else m

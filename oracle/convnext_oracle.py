"""CPU model oracle: ConvNeXt-{T,S,B,L}-CvSt -- TEST INFRASTRUCTURE ONLY.

Plain torch restatement of the architecture the reference builds through timm
(`utils_architecture.py:241-269`) with the conv stems of
`utils_architecture.py:174-217`; block math follows the vendored
`models/convnext.py:37-50` (dw7x7 -> LN(C, eps 1e-6) -> Linear C->4C -> GELU(erf)
-> Linear 4C->C -> gamma -> +residual), downsample = channels-first LN + 2x2 s2
conv (`models/convnext.py:79-82`), head = mean-pool -> LN -> Linear (:113-117).
Parameter names follow timm 0.8 (`stem.stem.N`, `stages.S.downsample.N`,
`stages.S.blocks.M.{conv_dw,norm,mlp.fc1,mlp.fc2,gamma}`, `head.norm`,
`head.fc`) -- the names the reference's checkpoints carry.

Pinned by `tests/test_model_oracle.py`: same logits as the vendored
`/root/reference/models/convnext.py` (timm stubbed) + `ConvBlock1` under a key
mapping, and by the committed logits fixture `tests/golden/convnext_t_cvst.npz`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

VARIANTS = {
    # name: (depths, dims, stem widths (channels after each 3x3 conv), stride of each stem conv)
    'convnext_tiny': ((3, 3, 9, 3), (96, 192, 384, 768), (48, 96), (2, 2)),
    'convnext_small': ((3, 3, 27, 3), (96, 192, 384, 768), (48, 96), (2, 2)),
    'convnext_base': ((3, 3, 27, 3), (128, 256, 512, 1024), (64, 96, 128), (2, 2, 1)),
    'convnext_large': ((3, 3, 27, 3), (192, 384, 768, 1536), (96, 144, 192), (2, 2, 1)),
}


class LNChannelsFirst(nn.Module):
    """utils_architecture.py:57-81, data_format='channels_first' (biased variance, eps inside sqrt)."""
    def __init__(self, c, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.eps = eps

    def forward(self, x):
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        xh = (x - mu) / torch.sqrt(var + self.eps)
        return self.weight[:, None, None] * xh + self.bias[:, None, None]


class ConvStem(nn.Module):
    """ConvBlock1 / ConvBlock3 (utils_architecture.py:174-217): [3x3 conv, LN, GELU] * n."""
    def __init__(self, widths, strides):
        super().__init__()
        layers, cin = [], 3
        for w, s in zip(widths, strides):
            layers += [nn.Conv2d(cin, w, 3, stride=s, padding=1), LNChannelsFirst(w), nn.GELU()]
            cin = w
        self.stem = nn.Sequential(*layers)

    def forward(self, x):
        return self.stem(x)


class Mlp(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c, 4 * c)
        self.fc2 = nn.Linear(4 * c, c)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_dw = nn.Conv2d(c, c, 7, padding=3, groups=c)
        self.norm = nn.LayerNorm(c, eps=1e-6)
        self.mlp = Mlp(c)
        self.gamma = nn.Parameter(1e-6 * torch.ones(c))

    def forward(self, x):
        h = self.conv_dw(x).permute(0, 2, 3, 1)
        h = self.mlp(self.norm(h)) * self.gamma
        return x + h.permute(0, 3, 1, 2)


class Stage(nn.Module):
    def __init__(self, cin, cout, depth, first):
        super().__init__()
        if first:
            self.downsample = nn.Identity()
        else:
            self.downsample = nn.Sequential(LNChannelsFirst(cin), nn.Conv2d(cin, cout, 2, stride=2))
        self.blocks = nn.Sequential(*[Block(cout) for _ in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class Head(nn.Module):
    def __init__(self, c, n_cls):
        super().__init__()
        self.norm = nn.LayerNorm(c, eps=1e-6)
        self.fc = nn.Linear(c, n_cls)

    def forward(self, x):
        return self.fc(self.norm(x.mean((-2, -1))))


class ConvNeXtCvSt(nn.Module):
    def __init__(self, arch='convnext_tiny', n_cls=1000):
        super().__init__()
        depths, dims, widths, strides = VARIANTS[arch]
        assert widths[-1] == dims[0]
        self.stem = ConvStem(widths, strides)
        self.stages = nn.Sequential(*[
            Stage(dims[max(i - 1, 0)], dims[i], depths[i], first=(i == 0)) for i in range(4)])
        self.head = Head(dims[-1], n_cls)
        self.apply(self._init)
        # the reference swaps the stem in AFTER timm's init (utils_architecture.py:243-244),
        # so the stem keeps torch's default conv init
        for m in self.stem.modules():
            if isinstance(m, nn.Conv2d):
                m.reset_parameters()

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=.02)
            nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.head(self.stages(self.stem(x)))


class Normalized(nn.Sequential):
    """normalize_model (utils_architecture.py:86-117): keys `normalize.mean/std`, `model.*`."""
    def __init__(self, model):
        super().__init__()
        self.normalize = _Normalizer()
        self.model = model


class _Normalizer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('mean', torch.as_tensor(IMAGENET_MEAN).view(1, 3, 1, 1))
        self.register_buffer('std', torch.as_tensor(IMAGENET_STD).view(1, 3, 1, 1))

    def forward(self, x):
        return (x - self.mean) / self.std


def build(arch='convnext_tiny', normalize=True, seed=0):
    torch.manual_seed(seed)
    m = ConvNeXtCvSt(arch)
    if normalize:
        m = Normalized(m)
    return m.eval()


def vendored_key_map(arch='convnext_tiny'):
    """timm-style key -> key in the reference's vendored models/convnext.py + ConvBlock stem."""
    depths = VARIANTS[arch][0]
    n_stem = len(VARIANTS[arch][2])
    out = {}
    for j in range(n_stem):
        for p in ('weight', 'bias'):
            out[f'stem.stem.{3 * j}.{p}'] = f'downsample_layers.0.stem.{3 * j}.{p}'
            out[f'stem.stem.{3 * j + 1}.{p}'] = f'downsample_layers.0.stem.{3 * j + 1}.{p}'
    for s in range(4):
        if s > 0:
            for j in (0, 1):
                for p in ('weight', 'bias'):
                    out[f'stages.{s}.downsample.{j}.{p}'] = f'downsample_layers.{s}.{j}.{p}'
        for b in range(depths[s]):
            t, v = f'stages.{s}.blocks.{b}.', f'stages.{s}.{b}.'
            out[t + 'gamma'] = v + 'gamma'
            for p in ('weight', 'bias'):
                out[t + 'conv_dw.' + p] = v + 'dwconv.' + p
                out[t + 'norm.' + p] = v + 'norm.' + p
                out[t + 'mlp.fc1.' + p] = v + 'pwconv1.' + p
                out[t + 'mlp.fc2.' + p] = v + 'pwconv2.' + p
    for p in ('weight', 'bias'):
        out['head.norm.' + p] = 'norm.' + p
        out['head.fc.' + p] = 'head.' + p
    return out

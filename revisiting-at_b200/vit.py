"""ViT-S-CvSt (and DeiT-S-CvSt, same architecture) for the attack / adversarial train step -- BASELINE config 3.

What the reference builds at utils_architecture.py:271-284: timm `vit_small_patch16_224` (D=384, depth 12,
6 heads of 64, MLP x4, qkv bias, LN eps 1e-6, class token, pos_embed [1,197,384]) with `patch_embed.proj`
replaced by `ConvBlock(48, end_siz=8)` (utils_architecture.py:120-144: four 3x3 stride-2 convs each followed by a
channels-first LN and GELU, then a 1x1 conv to 384 channels at 14x14).  Parameter names follow timm so the
reference's checkpoints load.  timm itself is un-vendored: the transformer math is restated from its
published source (see oracle/vit_oracle.py; parity unpinned).

Every layer goes through `ops.py`: fused first stem stage, LN(+GELU) kernels, tcgen05 GEMMs with bias /
residual epilogues, the hand-written attention kernel, bias+GELU kernels; tokens are [B,197,384] bf16.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .convnext import Normalized, _Conv, _LN, _Linear

BF16 = torch.bfloat16


class _ConvBlock(nn.Module):
    """ConvBlock(siz, end_siz): `stem` = Sequential([conv3x3 s2, LN, GELU] x 4, conv1x1)."""
    def __init__(self, siz=48, end_siz=8):
        super().__init__()
        mods, cin = [], 3
        for m in (1, 2, 4, 8):
            mods += [_Conv(cin, siz * m, 3), _LN(siz * m), nn.Identity()]
            cin = siz * m
        mods.append(_Conv(cin, siz * end_siz, 1))
        self.stem = nn.ModuleList(mods)

    def forward(self, x, mean=None, std=None):
        for j in range(4):
            conv, ln = self.stem[3 * j], self.stem[3 * j + 1]
            x = ops.stem_layer(x, conv.weight, conv.bias, ln.weight, ln.bias, 2, j == 0,
                               mean if j == 0 else None, std if j == 0 else None)
        last = self.stem[12]
        return ops.linear_rows(x, last.weight, last.bias)               # NHWC [B,14,14,384]


class _PatchEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = _ConvBlock(48, 8)


class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = _Linear(dim, 3 * dim)
        self.proj = _Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = _Linear(dim, 4 * dim)
        self.fc2 = _Linear(4 * dim, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.norm1 = _LN(dim)
        self.attn = _Attn(dim)
        self.norm2 = _LN(dim)
        self.mlp = _Mlp(dim)

    def forward(self, x):
        return ops.vit_block(x, self.norm1.weight, self.norm1.bias, self.attn.qkv.weight, self.attn.qkv.bias,
                             self.attn.proj.weight, self.attn.proj.bias, self.norm2.weight, self.norm2.bias,
                             self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias, self.heads)


class ViTCvSt(nn.Module):
    def __init__(self, dim=384, depth=12, heads=6, n_cls=1000, n_tokens=197):
        super().__init__()
        if dim // heads != 64:
            raise ValueError('the attention kernel is built for a head dimension of 64')
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_tokens, dim))
        self.patch_embed = _PatchEmbed()
        self.blocks = nn.ModuleList([_Block(dim, heads) for _ in range(depth)])
        self.norm = _LN(dim)
        self.head = _Linear(dim, n_cls)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for name, p in self.named_parameters():
            if p.ndim == 2 and not name.startswith('patch_embed'):
                nn.init.trunc_normal_(p, std=.02)
        for m in self.patch_embed.modules():
            if isinstance(m, _Conv):
                nn.init.kaiming_uniform_(m.weight, a=5 ** 0.5)
                fan_in = m.weight[0].numel()
                nn.init.uniform_(m.bias, -1 / fan_in ** 0.5, 1 / fan_in ** 0.5)

    def forward(self, x, mean=None, std=None):
        x = self.patch_embed.proj(x, mean, std)                         # [B,14,14,D] bf16
        B, D = x.shape[0], x.shape[-1]
        x = x.reshape(B, -1, D)
        cls, pos = ops._cast(self.cls_token), ops._cast(self.pos_embed)
        x = torch.cat((cls.expand(B, -1, -1), x), dim=1) + pos          # timm `_pos_embed`
        for b in self.blocks:
            x = b(x)
        c = x[:, 0].float()                                             # LN is per token: only the class row is needed
        c = F.layer_norm(c, (D,), self.norm.weight.float(), self.norm.bias.float(), 1e-6)
        return F.linear(c.to(BF16), ops._cast(self.head.weight), ops._cast(self.head.bias))


def build(normalize=True, seed=0, **kw):
    torch.manual_seed(seed)
    m = ViTCvSt(**kw)
    return Normalized(m) if normalize else m

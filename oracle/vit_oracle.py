"""CPU model oracle: ViT-S-CvSt -- TEST INFRASTRUCTURE ONLY.  **Parity unpinned.**

The reference builds this model as timm's `vit_small_patch16_224` with `patch_embed.proj` replaced by its own
conv stem `ConvBlock(48, end_siz=8)` (utils_architecture.py:120-144, call site :271-275).  The conv stem is under
/root/reference and is restated from it; the transformer itself lives in the un-vendored dependency
timm-0.8.0.dev0 (README.md:15), which is absent from /root/reference and from this image, so its published
algorithm (`timm/models/vision_transformer.py`: VisionTransformer / Block / Attention / Mlp) is restated here:

    tokens = cat(cls_token, flatten(proj(x))) + pos_embed                  [B, 197, 384]
    12 x { x += proj(softmax(q k^T / sqrt(64)) v),  (q,k,v) = qkv(LN(x)) ;  x += fc2(GELU(fc1(LN(x)))) }
    logits = head(LN(x)[:, 0])                                              LN eps 1e-6, qkv bias, no layer scale

Nothing in the reference (no tests, no golden vectors, no source) pins the transformer part; the CUDA engine is
compared against this restatement only.  Parameter names follow timm (`cls_token`, `pos_embed`,
`patch_embed.proj.stem.N`, `blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}`, `norm`, `head`), the
names the reference's ViT checkpoints carry.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .convnext_oracle import IMAGENET_MEAN, IMAGENET_STD, LNChannelsFirst


class ConvBlock(nn.Module):
    """utils_architecture.py:120-144: 4 x [conv3x3 s2, channels-first LN, GELU] then a 1x1 conv."""
    def __init__(self, siz=48, end_siz=8):
        super().__init__()
        layers, cin = [], 3
        for m in (1, 2, 4, 8):
            layers += [nn.Conv2d(cin, siz * m, 3, stride=2, padding=1), LNChannelsFirst(siz * m), nn.GELU()]
            cin = siz * m
        layers.append(nn.Conv2d(cin, siz * end_siz, 1))
        self.stem = nn.Sequential(*layers)

    def forward(self, x):
        return self.stem(x)


class PatchEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = ConvBlock(48, 8)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class ViTCvStOracle(nn.Module):
    def __init__(self, dim=384, depth=12, heads=6, n_cls=1000, n_tokens=197):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_tokens, dim))
        self.patch_embed = PatchEmbed()
        self.blocks = nn.ModuleList([Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, n_cls)
        init_vit_(self)

    def forward(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        for b in self.blocks:
            x = b(x)
        return self.head(self.norm(x)[:, 0])


def init_vit_(m):
    """timm init: trunc-normal .02 for pos_embed and every Linear weight, normal 1e-6 for the class token,
    zero biases; the conv stem keeps torch's default init (it is swapped in after timm's init)."""
    nn.init.trunc_normal_(m.pos_embed, std=.02)
    nn.init.normal_(m.cls_token, std=1e-6)
    for name, p in m.named_parameters():
        if name.startswith('patch_embed') or p.ndim != 2:
            continue
        nn.init.trunc_normal_(p, std=.02)


class Normalized(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.normalize = nn.Module()
        self.normalize.register_buffer('mean', torch.as_tensor(IMAGENET_MEAN).view(1, 3, 1, 1))
        self.normalize.register_buffer('std', torch.as_tensor(IMAGENET_STD).view(1, 3, 1, 1))
        self.model = model

    def forward(self, x):
        return self.model((x - self.normalize.mean) / self.normalize.std)


def build(normalize=True, seed=0, **kw):
    torch.manual_seed(seed)
    m = ViTCvStOracle(**kw)
    return Normalized(m) if normalize else m

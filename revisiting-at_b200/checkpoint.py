"""Checkpoint formats of the reference, read and written for the engine's models.

What the reference writes (main.py:737-756): `weights_N.pt` = `self.model.state_dict()` of the
DDP(WrappedModel(normalize_model(timm model))) stack, i.e. keys `module.base_model.model.<timm name>` plus
`module.base_model.normalize.{mean,std}`; `weights_ema_N.pt` = the same without `module.` (timm
`get_state_dict` unwraps); `full_model_N.pth` = `{'model_state_dict', 'optimizer_state_dict',
'loss_scaler_state_dict', 'epoch' [, 'state_dict_ema']}`.  What it reads (main.py:856-872, AA_eval.py:184-188):
strip `module.`, then try as-is, then with `base_model.` added, then with `base_model.` removed.

`adapt_state_dict` folds those trials into one deterministic key rewrite against the TARGET's own key set (so
nothing is loaded by trial and exception), and also accepts the parameter names of the reference's vendored
`models/convnext.py` (`downsample_layers.N`, `stages.S.B.dwconv/pwconv1/pwconv2`, `norm`, `head`), which differ
from timm's only by name.  Pure host code: tensors are only renamed, never touched.
"""
import math
import re
from collections import OrderedDict

import torch

_WRAPPERS = ('module.', 'base_model.', 'model.')

_VENDORED = (
    # (pattern on the vendored name, replacement giving the timm name)
    (re.compile(r'^downsample_layers\.0\.'), 'stem.'),
    (re.compile(r'^downsample_layers\.([1-3])\.([01])\.'), r'stages.\1.downsample.\2.'),
    (re.compile(r'^stages\.(\d)\.(\d+)\.dwconv\.'), r'stages.\1.blocks.\2.conv_dw.'),
    (re.compile(r'^stages\.(\d)\.(\d+)\.pwconv1\.'), r'stages.\1.blocks.\2.mlp.fc1.'),
    (re.compile(r'^stages\.(\d)\.(\d+)\.pwconv2\.'), r'stages.\1.blocks.\2.mlp.fc2.'),
    (re.compile(r'^stages\.(\d)\.(\d+)\.(norm\.|gamma$)'), r'stages.\1.blocks.\2.\3'),
    (re.compile(r'^norm\.'), 'head.norm.'),
    (re.compile(r'^head\.(weight|bias)$'), r'head.fc.\1'),
)


def vendored_to_timm(name):
    """models/convnext.py parameter name -> timm 0.8 `ConvNeXt` name (identity for a name that already is timm's)."""
    for pat, rep in _VENDORED:
        new, n = pat.subn(rep, name)
        if n:
            return new
    return name


def _core(name):
    """Key without any wrapper prefix: (`normalize.<buf>` | `<model parameter name>`)."""
    changed = True
    while changed:
        changed = False
        for w in _WRAPPERS:
            if name.startswith(w):
                name, changed = name[len(w):], True
    return name


def unwrap(obj, prefer_ema=False):
    """A `weights_N.pt` dict is the state dict itself; a `full_model_N.pth` dict carries it under
    'model_state_dict' (and the EMA weights under 'state_dict_ema')."""
    if isinstance(obj, dict) and 'model_state_dict' in obj:
        if prefer_ema and obj.get('state_dict_ema') is not None:
            return obj['state_dict_ema']
        return obj['model_state_dict']
    return obj


def adapt_state_dict(ckpt, target_keys):
    """Rewrite the keys of `ckpt` onto `target_keys` (an iterable of the target module's state-dict keys).

    Every target key is matched by its wrapper-free core name; checkpoint keys are reduced the same way, after
    mapping vendored ConvNeXt names to timm's when the plain name is not one of the target's.  Normaliser buffers (`normalize.mean/std`) are dropped when the
    target has no normaliser and left to the target's own constants when the checkpoint has none (they are
    constants of the reference: main.py:190-191).  Returns (new state dict, missing target keys, unused ckpt keys)."""
    by_core = {}
    for k in target_keys:
        by_core.setdefault(_core(k), k)
    out, unused = OrderedDict(), []
    for k, v in ckpt.items():
        c = _core(k)
        tk = by_core.get(c)
        if tk is None and not c.startswith('normalize.'):      # not a timm name of this model: vendored spelling?
            tk = by_core.get(vendored_to_timm(c))
        if tk is None:
            unused.append(k)
        else:
            out[tk] = v
    missing = [k for c, k in by_core.items() if k not in out and not c.startswith('normalize.')]
    unused = [k for k in unused if not _core(k).startswith('normalize.')]
    return out, missing, unused


def load_checkpoint(model, ckpt, strict=True, prefer_ema=False):
    """Load a reference-format checkpoint (path, `weights_N.pt` dict or `full_model_N.pth` dict) into `model`,
    whatever wrappers (`DistributedDataParallel`, `WrappedModel`, `Normalized`) either side carries.
    Raises KeyError on a missing / unexpected parameter when `strict` (the reference's final `load_state_dict`
    attempt raises too: main.py:869-871)."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, '__fspath__'):
        ckpt = torch.load(ckpt, map_location='cpu')
    sd = unwrap(ckpt, prefer_ema)
    own = model.state_dict()
    new, missing, unused = adapt_state_dict(sd, own.keys())
    if strict and (missing or unused):
        raise KeyError(f'checkpoint does not fit the model: missing {missing[:5]}{"..." if len(missing) > 5 else ""}, '
                       f'unexpected {unused[:5]}{"..." if len(unused) > 5 else ""}')
    for k, v in new.items():
        if tuple(v.shape) != tuple(own[k].shape):
            raise ValueError(f'{k}: checkpoint shape {tuple(v.shape)} != model shape {tuple(own[k].shape)}')
    model.load_state_dict(new, strict=False)
    return missing, unused


def reference_state_dict(model, prefix='module.'):
    """State dict under the reference's key names for `weights_N.pt` (main.py:739): `model` is the
    WrappedModel(Normalized(engine)) (or DDP of it); plain engines get the `base_model.model.` prefix added."""
    sd = model.state_dict()
    normalised = any(_core(k).startswith('normalize.') for k in sd)      # `normalize_model` adds the `model.` level
    out = OrderedDict()
    for k, v in sd.items():
        c = _core(k)
        inner = c if (c.startswith('normalize.') or not normalised) else 'model.' + c
        out[prefix + 'base_model.' + inner] = v.detach()
    return out


def save_checkpoint(folder, epoch, model, optimizer=None, ema_state=None, epochs=None):
    """The files main.py:737-756 writes at the end of an epoch: `weights_N.pt`, `weights_ema_N.pt` and, every 5th
    epoch or at the last one, `full_model_N.pth` (bf16 training has no loss scaler: its entry is an empty dict)."""
    import os
    sd = reference_state_dict(model)
    torch.save(sd, os.path.join(folder, f'weights_{epoch}.pt'))
    full = {'model_state_dict': sd, 'optimizer_state_dict': optimizer.state_dict() if optimizer is not None else {},
            'loss_scaler_state_dict': {}, 'epoch': epoch}
    if ema_state is not None:
        torch.save(ema_state, os.path.join(folder, f'weights_ema_{epoch}.pt'))
        full['state_dict_ema'] = ema_state
    if epoch % 5 == 0 or (epochs is not None and epoch == epochs - 1):
        torch.save(full, os.path.join(folder, f'full_model_{epoch}.pth'))


def interpolate_pos_encoding(pos_embed, new_img_size, old_img_size=224, patch_size=16):
    """ViT position embedding for another (square) resolution (utils_architecture.py:22-53): class row kept, the
    sqrt(N) x sqrt(N) grid resampled bicubically with the reference's +0.1 guard on the scale factor."""
    n_old = pos_embed.shape[1] - 1
    side_new = new_img_size // patch_size
    if side_new * side_new == n_old:
        return pos_embed
    side_old = int(math.sqrt(n_old))
    dim = pos_embed.shape[-1]
    grid = pos_embed[:, 1:].reshape(1, side_old, side_old, dim).permute(0, 3, 1, 2)
    s = (side_new + 0.1) / math.sqrt(n_old)
    grid = torch.nn.functional.interpolate(grid, scale_factor=(s, s), mode='bicubic')
    if grid.shape[-2] != side_new or grid.shape[-1] != side_new:
        raise AssertionError(f'interpolated grid {tuple(grid.shape[-2:])} != {side_new}')
    grid = grid.permute(0, 2, 3, 1).reshape(1, side_new * side_new, dim)
    return torch.cat((pos_embed[:, :1], grid), dim=1)

"""Import the UNMODIFIED reference by file path -- TEST / BENCH INFRASTRUCTURE ONLY.

Source: `/root/reference` where it is mounted (the build container), else the copy `oracle/stage_ref.py` staged
under `oracle/_ref/` at build time (git-ignored; it travels to the GPU box with the snapshot).  Used by
`oracle/make_goldens.py`, by the CPU tests that compare the oracle with the reference directly, and by
`bench.py`'s CPU arm (`--impl reference`, `cpu_baseline`: kind "reference").  The `-m gpu` tests and `smoke()`
never need it; the product package never imports it.
"""
import importlib.util
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
REF = os.environ.get('B200AT_REFERENCE', '/root/reference')
if not os.path.isfile(os.path.join(REF, 'autopgd_train_clean.py')) and \
        os.path.isfile(os.path.join(_STAGED, 'autopgd_train_clean.py')):
    REF = _STAGED


def available() -> bool:
    return os.path.isfile(os.path.join(REF, 'autopgd_train_clean.py'))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def attack_module():
    """reference autopgd_train_clean (needs only torch: SURVEY F3)."""
    return _load('_ref_autopgd_train_clean', os.path.join(REF, 'autopgd_train_clean.py'))


def _timm_stub():
    """4-symbol timm stub: enough for models/convnext.py and utils_architecture.py."""
    import torch.nn as nn
    if 'timm' in sys.modules and not getattr(sys.modules['timm'], '_b200at_stub', False):
        return
    def mk(name):
        m = types.ModuleType(name)
        m._b200at_stub = True
        sys.modules[name] = m
        return m
    timm = mk('timm'); models = mk('timm.models'); layers = mk('timm.models.layers')
    registry = mk('timm.models.registry'); cnx = mk('timm.models.convnext'); vit = mk('timm.models.vision_transformer')
    timm.models = models; models.layers = layers; models.registry = registry
    models.convnext = cnx; models.vision_transformer = vit
    layers.trunc_normal_ = nn.init.trunc_normal_
    class DropPath(nn.Identity):
        def __init__(self, p=0.):
            super().__init__()
    layers.DropPath = DropPath
    registry.register_model = lambda f: f
    models.create_model = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('timm stub'))
    cnx._create_convnext = None
    vit.VisionTransformer = object


def convnext_t_cvst():
    """Reference ConvNeXt-T-CvSt from the vendored models/convnext.py + ConvBlock1."""
    _timm_stub()
    cn = _load('_ref_convnext', os.path.join(REF, 'models', 'convnext.py'))
    ua = _load('_ref_utils_architecture', os.path.join(REF, 'utils_architecture.py'))
    m = cn.ConvNeXt(depths=[3, 3, 9, 3], dims=[96, 192, 384, 768])
    m.downsample_layers[0] = ua.ConvBlock1(48, end_siz=8)
    return m.eval(), ua


def convnext_t_cvst_normalized(state_dict=None):
    """`normalize_model(ConvNeXt-T-CvSt)` exactly as main.py:826-828 wraps it (utils_architecture.py:86-117), the
    reference's own modules end to end; optionally loaded with a timm-named state dict of the oracle / engine."""
    m, ua = convnext_t_cvst()
    if state_dict is not None:
        from . import convnext_oracle
        km = convnext_oracle.vendored_key_map()
        m.load_state_dict({km[k]: v for k, v in state_dict.items() if k in km})
    return ua.normalize_model(m, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)).eval()

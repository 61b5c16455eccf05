"""Import the UNMODIFIED reference from /root/reference (build container only).

Used by `oracle/make_goldens.py` and by the CPU tests that compare the oracle
with the reference directly.  /root/reference does not exist on the GPU box:
nothing reachable from `-m gpu` tests, `smoke()` or `bench.py` imports this.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get('B200AT_REFERENCE', '/root/reference')


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

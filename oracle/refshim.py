"""Import the UNMODIFIED reference hot-path modules: from /root/reference in the build container, or
from the staged copy under oracle/_ref/ (written by oracle/stage_ref.py, git-ignored, travels to the
GPU box with the snapshot) -- the reference arm of bench.py (`--impl reference`, cpu_baseline.kind
"reference") runs those files as they are.

Test / baseline infrastructure only: used by tests/golden/make_golden.py, tests/, and bench.py's CPU
legs; never by the product path.

`import mogen` needs mmcv 1.7.2, lmdb, fairseq, librosa, fuzzywuzzy, kornia, none of which are in
the image.  This shim puts minimal stand-ins on sys.modules -- only the names the
hot-path modules touch at import time -- and creates empty `mogen.*` packages whose __path__ points
into the reference tree, so the heavy package __init__ files are skipped while every module the
path needs is the reference's own file, executed as is.
"""
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(_HERE, "_ref")


def _find_ref():
    for c in (os.environ.get("RG_REFERENCE"), "/root/reference", STAGED):
        if c and os.path.isdir(os.path.join(c, "mogen")):
            return c
    return os.environ.get("RG_REFERENCE", "/root/reference")


REF = _find_ref()


def available():
    return os.path.isdir(os.path.join(REF, "mogen"))


class _Registry:
    """Stand-in for mmcv.utils.Registry: register_module / get / build (type= dispatch)."""

    def __init__(self, name, parent=None, build_func=None):
        self.name = name
        self._module_dict = {}
        self.build_func = build_func or _build_from_cfg
        self.parent = parent

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self._module_dict.get(key)

    def build(self, cfg, *a, **kw):
        return self.build_func(cfg, self, *a, **kw)


def _build_from_cfg(cfg, registry, default_args=None):
    if cfg is None:
        return None
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop("type")
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f"{t} is not in the {registry.name} registry")
    return cls(**args)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch.nn as nn
    import transformers  # noqa: F401  (import the real one before stubbing its optional deps)
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_rg_stub", False):
        return
    models = _Registry("model")

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    mmcv = _mod("mmcv", __version__="1.7.2", _rg_stub=True)
    mmcv.utils = _mod("mmcv.utils", Registry=_Registry, build_from_cfg=_build_from_cfg,
                      get_logger=lambda *a, **k: None)
    mmcv.cnn = _mod("mmcv.cnn", MODELS=models)
    mmcv.runner = _mod("mmcv.runner", BaseModule=BaseModule)
    mmcv.parallel = _mod("mmcv.parallel", collate=None)
    _mod("lmdb")
    _mod("fairseq")
    _mod("librosa")
    fw = _mod("fuzzywuzzy")
    from . import fuzz_ratio                                   # restated third-party function, see its header
    fw.fuzz = _mod("fuzzywuzzy.fuzz", partial_ratio=fuzz_ratio.partial_ratio)
    k = _mod("kornia")
    k.filters = _mod("kornia.filters")
    k.filters.kernels = _mod("kornia.filters.kernels", laplacian_1d=lambda *a, **kw: None)
    if "pyarrow" not in sys.modules:
        try:
            import pyarrow  # noqa: F401
        except Exception:
            _mod("pyarrow")
    for pkg in ["mogen", "mogen.models", "mogen.models.utils", "mogen.models.transformers",
                "mogen.models.transformers.rag", "mogen.models.attentions",
                "mogen.models.architectures", "mogen.models.losses"]:
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split("."))]
        m.__package__ = pkg
        sys.modules[pkg] = m


def load():
    """Returns a namespace of the reference modules on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    install_stubs()
    ns = types.SimpleNamespace()
    ns.builder = importlib.import_module("mogen.models.builder")
    ns.gd = importlib.import_module("mogen.models.utils.gaussian_diffusion")
    ns.styl = importlib.import_module("mogen.models.utils.stylization_block")
    ns.attn = importlib.import_module("mogen.models.attentions.efficient_attention")
    ns.mse = importlib.import_module("mogen.models.losses.mse_loss")
    ns.dt = importlib.import_module("mogen.models.transformers.diffusion_transformer")
    ns.rag_utils = importlib.import_module("mogen.models.transformers.rag.utils")
    ns.discourse = importlib.import_module("mogen.models.transformers.rag.discourse_retrieval")
    ns.gesture_type = importlib.import_module("mogen.models.transformers.rag.gesture_type_retrieval")
    ns.rg = importlib.import_module("mogen.models.transformers.raggesture")
    ns.arch = importlib.import_module("mogen.models.architectures.diffusion_architecture")
    import torch
    torch.autograd.set_detect_anomaly(False)  # the reference turns it on globally at import
    return ns

"""ORACLE (test infrastructure, runs in THIS container only — needs /root/reference).

Imports the reference's own `cubercnn` modeling code from /root/reference, unmodified, on top of
oracle/d2lite (the restated detectron2 / pytorch3d / fvcore surface).  Every other third-party
module the reference drags in at import time but never touches on the model path (matplotlib,
pycocotools, iopath, termcolor, seaborn, pytorch3d.renderer, detectron2.engine ...) is replaced by
an inert stub.  Used by tests/golden/make_model_golden.py to produce the fixtures that pin
oracle/cubercnn_oracle and the CUDA path.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
D2LITE = os.path.join(HERE, "d2lite")


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = _AnyMeta(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None,
                                  "__getattr__": lambda self, n: (lambda *a, **k: None)})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Last-resort finder: fabricates inert modules for packages that are absent here."""
    ROOTS = ("detectron2", "pytorch3d", "fvcore", "pycocotools", "matplotlib", "iopath", "termcolor", "seaborn",
             "mpl_toolkits")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make `import cubercnn` resolve to /root/reference and `import detectron2` to d2lite."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REF, "cubercnn")):
        raise RuntimeError("ref_runner needs /root/reference (only present in the build container)")
    for p in (REF, D2LITE):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.meta_path.append(_StubFinder())     # appended: real modules (d2lite, cv2, scipy) win
    _installed = True


def reference_cfg(config_file="configs/cubercnn_DLA34_FPN.yaml", opts=()):
    install()
    from detectron2.config import get_cfg
    from cubercnn.config import get_cfg_defaults
    cfg = get_cfg()
    get_cfg_defaults(cfg)
    cfg.merge_from_file(os.path.join(REF, config_file))
    cfg.merge_from_list(["MODEL.DEVICE", "cpu", "MODEL.WEIGHTS_PRETRAIN", "none", "VIS_PERIOD", 0] + list(opts))
    return cfg


def build_reference_model(cfg, priors=None):
    install()
    from cubercnn.modeling.meta_arch import build_model
    import cubercnn.modeling.backbone  # noqa: F401  (registers builders)
    import cubercnn.modeling.proposal_generator  # noqa: F401
    import cubercnn.modeling.roi_heads  # noqa: F401
    return build_model(cfg, priors=priors)

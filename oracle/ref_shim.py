"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim that lets the *unmodified* reference (abhi1kumar/groomed_nms, mounted read-only at
/root/reference in the build container) run on python 3.12 / torch 2.11 / numpy 2.x on CPU.  It is used
only by oracle/gen_golden.py (to generate tests/golden/*.npz) and by tests that validate the oracle
restatement against the real reference when /root/reference happens to exist.  /root/reference does NOT
exist on the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py may call this.

What the shim does; no reference source is edited or copied:
  1. torch.Tensor.masked_fill_ accepts uint8 masks again (lib/groomed_nms.py:56,73 pass .byte() masks).
  2. stub modules for packages that are absent here and unused on the hot path:
     matplotlib(+pyplot, patches, backends.backend_agg), mpl_toolkits.mplot3d, shapely.geometry, imp,
     visdom, easydict (a 6-line EasyDict), and lib.nms.gpu_nms (Cython/CUDA build, not rebuilt).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GROOMED_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "groomed_nms.py"))


class _Anything(types.ModuleType):
    """Module stub whose every attribute is a harmless callable/namespace."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = _AnyObj()
        setattr(self, name, val)
        return val


class _AnyObj(object):
    def __call__(self, *a, **k):
        return _AnyObj()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _AnyObj()

    def __getitem__(self, k):
        return _AnyObj()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())


class _EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


_installed = False


def install():
    """Idempotently install the shim and put the reference root on sys.path (as its scripts do with cwd)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import torch

    _orig_masked_fill_ = torch.Tensor.masked_fill_

    def masked_fill_(self, mask, value):
        if mask.dtype == torch.uint8:
            mask = mask.bool()
        return _orig_masked_fill_(self, mask, value)

    torch.Tensor.masked_fill_ = masked_fill_

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.backends",
                 "matplotlib.backends.backend_agg", "matplotlib.path", "matplotlib.ticker",
                 "matplotlib.colors", "matplotlib.cm", "matplotlib.lines",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "shapely", "shapely.geometry", "imp", "visdom",
                 "plot", "plot.plotting_params", "plot.common_operations"]:
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    if "easydict" not in sys.modules:
        ed = types.ModuleType("easydict")
        ed.EasyDict = _EasyDict
        sys.modules["easydict"] = ed
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # lib.nms.gpu_nms is a Cython+CUDA extension (sm_35); give importers a stub that fails loudly if called.
    import importlib
    importlib.import_module("lib")  # namespace package rooted at the reference
    if "lib.nms.gpu_nms" not in sys.modules:
        g = types.ModuleType("lib.nms.gpu_nms")

        def gpu_nms(*a, **k):
            raise RuntimeError("reference lib.nms.gpu_nms (Cython/CUDA) is not built under the shim")

        g.gpu_nms = gpu_nms
        sys.modules["lib.nms.gpu_nms"] = g
    _installed = True


def load():
    """Return a namespace with the reference modules on the hot path."""
    install()
    import importlib
    ns = types.SimpleNamespace()
    ns.groomed_nms = importlib.import_module("lib.groomed_nms")
    ns.core = importlib.import_module("lib.core")
    ns.math_3d = importlib.import_module("lib.math_3d")
    ns.aploss = importlib.import_module("lib.loss.aploss")
    ns.nms_others = importlib.import_module("lib.nms_others")
    ns.py_cpu_nms = importlib.import_module("lib.nms.py_cpu_nms")
    return ns

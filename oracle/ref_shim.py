"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim that lets the *unmodified* reference (abhi1kumar/groomed_nms) run on python 3.12 / torch 2.11 /
numpy 2.x, on CPU or on a GPU.  The reference tree is looked up, in this order, at $GROOMED_REFERENCE_ROOT,
`<repo>/baseline/_ref` (the git-ignored byte-for-byte staging made by tools/stage_reference.py: it travels to the GPU
box with the snapshot) and /root/reference (build container only).  Users: oracle/gen_golden*.py (golden vectors),
oracle/ref_harness.py (the reference's loss / model run stock or with groomed_nms_b200.install()), the tests that
compare against the live reference, and `bench.py --impl reference` (the reference's own CPU PyTorch path as the
baseline arm).  The product never imports this.

What the shim does; no reference source is edited or copied:
  1. torch.Tensor.masked_fill_ accepts uint8 masks again (lib/groomed_nms.py:56,73 pass .byte() masks).
  2. stub modules for packages that are absent here and unused on the hot path:
     matplotlib(+pyplot, patches, backends.backend_agg), mpl_toolkits.mplot3d, shapely.geometry, imp,
     visdom, easydict (a 6-line EasyDict), and lib.nms.gpu_nms (Cython/CUDA build, not rebuilt).
  3. (GPU runs) legacy device semantics the reference's loss relies on (written for torch 0.4.1 with
     set_default_tensor_type('torch.cuda.FloatTensor')): numpy conversion of a CUDA tensor copies to the host
     first (lib/loss/rpn_3d.py:470 assigns CUDA tensor slices into numpy arrays), and indexing a CPU tensor with a
     CUDA index tensor moves the index to the host (lib/loss/rpn_3d.py:737: fg_inds_tensor[sorted_index[...]]).
  4. (CPU debugging only, `fake_cuda()`) `.cuda()` / torch.cuda.*Tensor become their CPU counterparts so that the
     CUDA-hard-coded RPN_3D_loss.forward can be stepped through in the GPU-less build container.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _find_root():
    for cand in (os.environ.get("GROOMED_REFERENCE_ROOT"), STAGED_ROOT, "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "lib", "groomed_nms.py")):
            return cand
    return os.environ.get("GROOMED_REFERENCE_ROOT", STAGED_ROOT)


REFERENCE_ROOT = _find_root()


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


def _install_legacy_device_semantics(torch):
    """Item 3 of the module docstring.  No-ops for CPU tensors."""
    _orig_array = torch.Tensor.__array__

    def __array__(self, *a, **k):
        if self.is_cuda:
            return _orig_array(self.detach().cpu(), *a, **k)
        return _orig_array(self, *a, **k)

    torch.Tensor.__array__ = __array__

    def _host_index(self, idx):
        if self.is_cuda:
            return idx
        if isinstance(idx, torch.Tensor):
            return idx.cpu() if idx.is_cuda else idx
        if isinstance(idx, tuple):
            return tuple(i.cpu() if isinstance(i, torch.Tensor) and i.is_cuda else i for i in idx)
        return idx

    _orig_get, _orig_set = torch.Tensor.__getitem__, torch.Tensor.__setitem__

    def __getitem__(self, idx):
        return _orig_get(self, _host_index(self, idx))

    def __setitem__(self, idx, val):
        if not self.is_cuda and isinstance(val, torch.Tensor) and val.is_cuda:
            val = val.cpu()
        return _orig_set(self, _host_index(self, idx), val)

    torch.Tensor.__getitem__ = __getitem__
    torch.Tensor.__setitem__ = __setitem__


def fake_cuda():
    """Item 4 of the module docstring: make the CUDA-hard-coded reference code run on the CPU (debugging and golden
    generation in the GPU-less build container).  Never used on the GPU box."""
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    for name in ("FloatTensor", "DoubleTensor", "LongTensor", "IntTensor", "ByteTensor", "BoolTensor"):
        setattr(torch.cuda, name, getattr(torch, name))


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
    _install_legacy_device_semantics(torch)

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

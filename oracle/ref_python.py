"""TEST INFRASTRUCTURE -- import the reference's own Python package (dev container only).

``/root/reference`` is read-only and does not exist on the GPU box, so this module is used
only by ``oracle/gen_golden.py`` and by the CPU tests that pin ``oracle.port`` (those tests
skip when the reference tree is absent).  Nothing is copied: the package is imported in
place, with

* stub modules for the plotting / phonopy / h5py imports that are not installed here and
  not on the hot path (matplotlib, mpl_toolkits, phonopy, seekpath, h5py), and
* the compiled reference C kernels from ``oracle/_ref`` pre-registered in ``sys.modules``
  under the names the package imports (``dynaphopy.power_spectrum.mem`` ...), because the
  read-only tree has no built extensions.  ``mem`` is the deterministic build
  (``Data[N] := 0``) unless ``shipped_mem=True``.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DP_REFERENCE_ROOT", "/root/reference")
_STUB_PREFIXES = ("matplotlib", "mpl_toolkits", "phonopy", "seekpath", "h5py")
_state = {"done": False}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dynaphopy"))


class _AutoModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _AutoModule(spec.name)

    def exec_module(self, module):
        pass


def load(shipped_mem=False):
    """Returns the imported reference package ``dynaphopy`` (unmodified, in place)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _state["done"]:
        from . import ref
        for name in _STUB_PREFIXES:
            try:
                __import__(name)
            except ImportError:
                if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
                    sys.meta_path.insert(0, _StubFinder())
        sys.modules["dynaphopy.power_spectrum.mem"] = ref.mem_shipped() if shipped_mem else ref.mem_det()
        sys.modules["dynaphopy.power_spectrum.correlation"] = ref.correlation_shipped()
        sys.modules["dynaphopy.displacements"] = ref.displacements()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _state["done"] = True
    import dynaphopy  # noqa: F401  (the reference package)
    import dynaphopy.projection  # noqa: F401
    import dynaphopy.power_spectrum  # noqa: F401
    import dynaphopy.dynamics  # noqa: F401
    import dynaphopy.atoms  # noqa: F401
    return sys.modules["dynaphopy"]

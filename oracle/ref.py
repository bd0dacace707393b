"""TEST INFRASTRUCTURE -- loader for the compiled, unmodified reference C kernels.

``oracle/Makefile`` compiles ``/root/reference/c/{correlation,mem,displacements}.c``
where they lie into CPython extension modules under ``oracle/_ref/`` (git-ignored; the
object code travels to the GPU box, the sources never enter this repository).

Variants (see ``oracle/Makefile``):
  shipped/   as shipped                      (mem is nondeterministic: OOB read, c/mem.c:210)
  det/mem    + ``-include shims/zalloc.h``   canonical ``Data[N] := 0`` semantics
  intended/correlation + ``-include shims/fixcexp.h``  ``exp(i w tau)`` weights
"""
import importlib.machinery
import importlib.util
import os
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_EXT = sysconfig.get_config_var("EXT_SUFFIX")
_cache = {}


def available():
    return os.path.exists(os.path.join(_REF, "det", "mem" + _EXT))


def _load(variant, name):
    key = (variant, name)
    if key not in _cache:
        path = os.path.join(_REF, variant, name + _EXT)
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle ref` in the dev container "
                "(needs /root/reference)")
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _cache[key] = mod
    return _cache[key]


def mem_det():
    """``mem.mem(frequency, velocity, time_step, coefficients=)`` -- c/mem.c:102, deterministic build."""
    return _load("det", "mem")


def mem_shipped():
    return _load("shipped", "mem")


def correlation_shipped():
    """``correlation.correlation_par(frequency, velocity, time_step, step=, integration_method=)`` -- c/correlation.c:101."""
    return _load("shipped", "correlation")


def correlation_intended():
    return _load("intended", "correlation")


def displacements():
    """``displacements.atomic_displacements(trajectory, positions, cell)`` -- c/displacements.c:95."""
    return _load("shipped", "displacements")

"""Build the CUDA shared library in-tree (``dynaphopy_b200/libdynaphopy_b200.so``).

nvcc cross-compiles for sm_100a without a GPU.  The built ``.so`` is git-ignored but travels
to the GPU box with the gpurun snapshot.  ``build_emulator`` compiles the same sources for the
host with ``-DDP_EMUL`` into ``tests/_emul_build/`` -- TEST INFRASTRUCTURE only (see
``tests/emul_include/dp_emul.h``); the package never loads it.
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdynaphopy_b200.so")
EMUL_DIR = os.path.join(ROOT, "tests", "_emul_build")
EMUL_LIB = os.path.join(EMUL_DIR, "libdp_emul.so")
OBJ_DIR = os.path.join(HERE, "_build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + glob.glob(os.path.join(ROOT, "tests", "emul_include", "*.h"))


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in _deps())


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build_cuda(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every .cu for sm_100a (one object per source, in parallel; objects are cached under
    ``dynaphopy_b200/_build`` keyed on the flags) and link them into one shared library.  Returns the library path.
    ``extra_flags`` / ``out``: alternative builds of the same library for experiments (``DYNAPHOPY_B200_LIB``)."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor
    out = LIB if out is None else out
    if not force and not extra_flags and out == LIB and not _stale(LIB):
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libdynaphopy_b200.so")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)
    tag = hashlib.sha1(" ".join(flags).encode()).hexdigest()[:10]
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [d for d in _deps() if not d.endswith(".cu")]
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, "%s.%s.o" % (os.path.basename(src)[:-3], tag))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.check_call(cmd, env=env)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _sources()))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", out],
                          env=env)
    return out


def build_emulator(force=False):
    """Host build of the same kernels on the fiber emulator (tests only)."""
    if not force and not _stale(EMUL_LIB):
        return EMUL_LIB
    os.makedirs(EMUL_DIR, exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    for src in _sources():
        obj = os.path.join(EMUL_DIR, os.path.basename(src) + ".o")
        cmd = [gxx, "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-DDP_EMUL", "-ffp-contract=off", "-w",
               "-I", os.path.join(ROOT, "tests", "emul_include"), "-c", src, "-o", obj]
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([gxx, "-shared", "-o", EMUL_LIB] + objs + ["-lpthread"])
    return EMUL_LIB


if __name__ == "__main__":
    import sys
    print(build_cuda(force=True, verbose="-v" in sys.argv))

"""In-tree build of the engine's CUDA library (``filterpy_b200/_C/libbke.so``) for sm_100a.

``nvcc`` cross-compiles without a GPU.  The ``.so`` is git-ignored but travels to the GPU box
with the repo snapshot, so a box without nvcc in PATH still finds a prebuilt library.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
OBJ_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libbke.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "--extended-lambda",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "bke.h"))
    return hdrs


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + _deps())


def build(force=False, verbose=False, extra_flags=()):
    """Compile every .cu under csrc/ and link libbke.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and %s is missing or stale" % LIB)
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    # one builder at a time: the ranks of a torchrun job may all find the library stale
    import fcntl
    lock = open(os.path.join(OUT_DIR, ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not needs_build():
            return LIB
        return _build_locked(nvcc, force, verbose, extra_flags)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc, force, verbose, extra_flags):
    hdr_t = max(os.path.getmtime(p) for p in _deps())

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", LIB + ".tmp"] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(LIB + ".tmp", LIB)
    return LIB


TORCH_OPS_SRC = os.path.join(HERE, "torch_ops", "bke_torch_ops.cpp")
TORCH_OPS_LIB = os.path.join(OUT_DIR, "libbke_torch_ops.so")


def build_torch_ops(force=False):
    """Compile the torch.ops.bke.* twin of the C-ABI (torch_ops/bke_torch_ops.cpp) against the installed
    PyTorch and link it to libbke.so (rpath $ORIGIN).  g++ only: the file holds no device code."""
    build()
    deps = [TORCH_OPS_SRC, os.path.join(os.path.dirname(HERE), "include", "bke.h")]
    if (not force and os.path.exists(TORCH_OPS_LIB)
            and all(os.path.getmtime(TORCH_OPS_LIB) > os.path.getmtime(d) for d in deps)):
        return TORCH_OPS_LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found and %s is missing or stale" % TORCH_OPS_LIB)
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + d for d in ce.include_paths()] + ["-I" + cuda_inc, TORCH_OPS_SRC, "-o", TORCH_OPS_LIB + ".tmp",
            "-L" + OUT_DIR, "-lbke", "-L" + tlib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda",
            "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib]
    import fcntl
    lock = open(os.path.join(OUT_DIR, ".build_ops.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for the torch.ops twin:\n%s\n%s" % (r.stdout, r.stderr))
        os.replace(TORCH_OPS_LIB + ".tmp", TORCH_OPS_LIB)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()
    return TORCH_OPS_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra_flags=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else []))

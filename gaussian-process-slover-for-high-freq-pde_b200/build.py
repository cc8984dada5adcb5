"""Build libgphm.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python gaussian-process-slover-for-high-freq-pde_b200/build.py [--force]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgphm.so")
SOURCES = ["gram.cu", "dgemm.cu", "factor.cu", "elemwise.cu", "fft.cu", "toeplitz_inv.cu", "toeplitz_fused.cu", "ozaki.cu", "peer.cu", "plan.cu"]
HEADERS = ["common.cuh", "kernfun.cuh", "kernels.h", "fft_core.cuh", os.path.join("..", "..", "include", "gphm.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((s, subprocess.Popen(cmd)))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on " + s)
    if force or procs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    build_ffi_shim(verbose)
    return OUT


def build_ffi_shim(verbose=True):
    """experimental/ffi_shim.cc -> libgphm_ffi.so, ONLY where JAX ships the XLA FFI headers (not in this image; the shim has
    never been compiled - see experimental/README.md)."""
    try:
        import jax
        inc = jax.ffi.include_dir()
    except Exception:
        if verbose:
            print("ffi_shim.cc skipped: no jax / XLA FFI headers in this environment", flush=True)
        return None
    out = os.path.join(HERE, "libgphm_ffi.so")
    src = os.path.join(HERE, "experimental", "ffi_shim.cc")
    if _stale(out, [src, OUT]):
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I" + inc, "-I" + os.path.join(HERE, "..", "include"),
               "-I/usr/local/cuda/include", src, "-L" + HERE, "-lgphm", "-Wl,-rpath," + HERE, "-o", out]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

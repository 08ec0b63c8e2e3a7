"""In-tree build of libe4s_b200.so (hand-written sm_100a CUDA behind the C ABI of include/e4s_b200.h).

    python -m e4s_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to e4s_b200/csrc/_obj/, the shared library to
e4s_b200/libe4s_b200.so (git-ignored, but it travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "libe4s_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
          "-I", INCLUDE]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _compile(src, verbose, objdir=None, extra=()):
    obj = os.path.join(objdir or OBJ, os.path.basename(src)[:-3] + ".o")
    deps = [src] + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    if not _newer(deps, obj):
        return obj, ""
    cmd = [NVCC] + ARCH + CFLAGS + list(extra) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if force:
        for f in glob.glob(os.path.join(OBJ, "*.o")):
            os.remove(f)
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write(log)
    if force or _newer(objs, LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_profile(verbose: bool = False) -> str:
    """Diagnostic twin of the library with the tensor-core kernels' stall counters compiled in (tools/opbench.py --prof):
    libe4s_b200_prof.so, selected with E4S_B200_LIB=<path>."""
    build(verbose=verbose)
    objdir = os.path.join(CSRC, "_obj_prof")
    os.makedirs(objdir, exist_ok=True)
    prof_srcs = [os.path.join(CSRC, "modconv_tcr.cu"), os.path.join(CSRC, "modconv_tch.cu")]
    prof_objs = [_compile(src, verbose, objdir, ["-DE4S_TC_PROFILE"])[0] for src in prof_srcs]
    others = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) if s not in prof_srcs]
    lib = os.path.join(PKG, "libe4s_b200_prof.so")
    if _newer(prof_objs + others, lib):
        r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", lib] + prof_objs + others + ["-lcudart", "-lcuda"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A/B twin of the library with extra -D switches on the tensor-core kernels (diagnostics only): libe4s_b200_<name>.so."""
    build(verbose=verbose)
    objdir = os.path.join(CSRC, "_obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    srcs = [os.path.join(CSRC, "modconv_tcr.cu"), os.path.join(CSRC, "modconv_tch.cu")]
    objs = [_compile(src, verbose, objdir, ["-D" + d for d in defines])[0] for src in srcs]
    others = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) if s not in srcs]
    lib = os.path.join(PKG, f"libe4s_b200_{name}.so")
    r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", lib] + objs + others + ["-lcudart", "-lcuda"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    if "--profile" in sys.argv:
        print(build_profile(verbose="--verbose" in sys.argv))
        sys.exit(0)
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

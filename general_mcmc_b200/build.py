"""In-tree build of libgmcmc.so (CUDA kernels + C ABI) for sm_100a.

nvcc cross-compiles without a GPU; objects go to general_mcmc_b200/build/, the shared library to
general_mcmc_b200/libgmcmc.so (git-ignored, shipped to the GPU box with the gpurun snapshot).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# tuning builds: GMCMC_VARIANT=name GMCMC_DEFINES="-DGM_MINB=5" python build.py -> libgmcmc_name.so
_VARIANT = os.environ.get("GMCMC_VARIANT", "")
BUILD = os.path.join(HERE, "build" + ("_" + _VARIANT if _VARIANT else ""))
LIB = os.path.join(HERE, "libgmcmc%s.so" % ("_" + _VARIANT if _VARIANT else ""))
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = (["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]
          + os.environ.get("GMCMC_DEFINES", "").split())

# per-target K1 translation units, each compiled in both math modes
K1_TARGETS = ["k_rosen", "k_iso", "k_dense", "k_mix", "k_rosen2d", "k_dgauss2d", "k_gauss2d"]
NUTS_TARGETS = ["n_rosen", "n_iso", "n_dense", "n_mix", "n_rosen2d", "n_dgauss2d"]
EXACT_FLAGS = ["-DGM_EXACT=1", "--fmad=false"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _units():
    """(source, object, extra flags) for every translation unit."""
    units = []
    for t in K1_TARGETS:
        units.append((t + ".cu", t + "_fast.o", []))
        units.append((t + ".cu", t + "_exact.o", EXACT_FLAGS))
    for t in NUTS_TARGETS:
        units.append((t + ".cu", t + "_fast.o", []))
        units.append((t + ".cu", t + "_exact.o", EXACT_FLAGS))
    units.append(("mh_fast.cu", "mh_fast.o", []))
    units.append(("mh_exact.cu", "mh_exact.o", ["--fmad=false"]))
    units.append(("mh_int.cu", "mh_int.o", ["--fmad=false"]))
    units.append(("mass_dense.cu", "mass_dense.o", ["--fmad=false"]))
    units.append(("gibbs.cu", "gibbs.o", ["--fmad=false"]))
    for extra in ("dense_tc.cu",):
        if os.path.exists(os.path.join(CSRC, extra)):
            flags = ["--fmad=false"] if extra.endswith("_exact.cu") else []
            units.append((extra, extra[:-3] + ".o", flags))
    for src in ("stats.cu", "export.cu", "dispatch.cu", "runtime.cu"):
        units.append((src, src[:-3] + ".o", []))
    return units


_INC_RE = __import__("re").compile(r'^\s*#\s*include\s+"([^"]+)"', __import__("re").M)


def _deps_mtime(src, _seen=None):
    """Newest mtime of `src` and of every header it includes (transitively) from csrc/ or include/."""
    seen = _seen if _seen is not None else set()
    if src in seen or not os.path.exists(src):
        return 0.0
    seen.add(src)
    m = os.path.getmtime(src)
    with open(src) as f:
        text = f.read()
    for inc in _INC_RE.findall(text):
        for d in (os.path.dirname(src), CSRC, INCLUDE):
            cand = os.path.join(d, inc)
            if os.path.exists(cand):
                m = max(m, _deps_mtime(cand, seen))
                break
    return m


def build(force=False, verbose=False, jobs=None):
    """Compile whatever is stale and link libgmcmc.so.  Returns the library path."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    todo = []
    objs = []
    for src, obj, flags in _units():
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, obj)
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < _deps_mtime(s):
            todo.append([nvcc] + ARCH + COMMON + flags + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if todo:
        with ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 4)) as ex:
            list(ex.map(run, todo))
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        run([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

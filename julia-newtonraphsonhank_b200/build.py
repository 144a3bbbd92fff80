"""Builds libhankb200.so (CUDA, sm_100a only) in-tree: csrc/*.cu -> lib/libhankb200.so.

nvcc cross-compiles without a GPU.  Objects are compiled in parallel (one translation unit per
n_e instantiation and launcher group) and only when their sources changed.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libhankb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _deps(src, seen=None):
    """Transitive #include "..." closure of a source (so that touching one header rebuilds only its users)."""
    import re
    seen = set() if seen is None else seen
    out = []
    try:
        text = open(src).read()
    except OSError:
        return out
    for inc in re.findall(r'^\s*#include\s+"([^"]+)"', text, flags=re.M):
        p = os.path.normpath(os.path.join(os.path.dirname(src), inc))
        if p not in seen and os.path.exists(p):
            seen.add(p)
            out.append(p)
            out += _deps(p, seen)
    return out


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, obj, log):
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
    return obj


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + _deps(s)):
            jobs.append((s, o, o[:-2] + ".log"))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(jobs))) as ex:
            for f in [ex.submit(_compile, *j) for j in jobs]:
                f.result()
                if verbose:
                    print("compiled", f.result())
    if jobs or _newer(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcusolver", "-lcublas", "-ldl",
                                                     "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

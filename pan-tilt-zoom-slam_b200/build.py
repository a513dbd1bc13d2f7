"""
Builds csrc/*.cu into csrc/libptzba.so for sm_100a with nvcc (in-tree, so the .so travels to the GPU box).
Usage: python pan-tilt-zoom-slam_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PTZBA_BUILD_TAG=<tag> (with PTZBA_EXTRA_NVCC_FLAGS) builds a tuning variant beside the default library:
# objects in csrc/build_<tag>/, output csrc/libptzba_<tag>.so, selected at run time with PTZBA_LIBRARY=<path>
TAG = os.environ.get("PTZBA_BUILD_TAG", "")
OBJDIR = os.path.join(CSRC, "build_" + TAG) if TAG else CSRC
OUT = os.path.join(CSRC, "libptzba_%s.so" % TAG if TAG else "libptzba.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("PTZBA_EXTRA_NVCC_FLAGS", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "ptzba.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, force, verbose):
    os.makedirs(OBJDIR, exist_ok=True)
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
    if not force and not _stale(obj, [src] + headers()):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr if verbose else ""


def build(force=False, verbose=False):
    srcs = sources()
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [r[0] for r in res]
    for r in res:
        if r[1]:
            print(r[1])
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

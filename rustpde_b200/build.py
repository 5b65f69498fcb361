"""Build librustpde_b200.so (nvcc, sm_100a) in-tree.

    python -m rustpde_b200.build            # build if sources are newer
    python -m rustpde_b200.build --force

There is exactly one product library and it is CUDA-only.  (The CPU
*emulation* of the same kernels used by the GPU-less tests is built by
tests/cuemu/build_emu.py into tests/cuemu/_build/ and is never loaded by this
package.)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librustpde_b200.so")
SOURCES = ["kernels.cu", "fast_x.cu", "fast_xs.cu", "fast_xw.cu", "fast_y.cu", "fast_p.cu", "fast_pw.cu", "tables.cu", "progbuild.cu", "field.cu", "solver.cu", "navier.cu", "snapshot.cu", "adjoint.cu", "lapack.cu", "capi.cu"]
NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-fvisibility=hidden",
]


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(HERE, "..", "include", "rustpde_b200.h"))
    return out


def _includes(path, seen):
    """Transitive closure of the quoted #includes of `path` inside csrc/."""
    import re

    if path in seen or not os.path.exists(path):
        return
    seen.add(path)
    for m in re.finditer(r'#include\s+"([^"]+)"', open(path).read()):
        _includes(os.path.join(os.path.dirname(path), m.group(1)), seen)
        _includes(os.path.join(CSRC, m.group(1)), seen)


def _stale(src_path, obj, extra=()):
    """True when `obj` is older than the source or any header it includes (per-file incremental build)."""
    if not os.path.exists(obj):
        return True
    deps = set()
    _includes(src_path, deps)
    deps.update(extra)
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and not verbose and not _stale(os.path.join(CSRC, src), obj):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("RUSTPDE_B200_NVCC_EXTRA", "").split() + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

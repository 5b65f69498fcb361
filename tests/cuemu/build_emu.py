"""Build the CPU *emulation* of the CUDA kernels (TEST INFRASTRUCTURE ONLY).

The same sources as the product library are compiled with g++ and -DRP_EMU
against tests/cuemu/cuemu.h, which supplies a cooperative-fiber model of
blocks/threads/__syncthreads/shuffles/mma.  The result,
tests/cuemu/_build/librustpde_b200_emu.so, lets the GPU-less test-suite check
every kernel's arithmetic against the oracle.  Nothing in rustpde_b200/ loads it.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "rustpde_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "librustpde_b200_emu.so")
SOURCES = ["kernels.cu", "fast_x.cu", "fast_xs.cu", "fast_xw.cu", "fast_y.cu", "fast_p.cu", "fast_pw.cu", "tables.cu", "progbuild.cu", "field.cu", "solver.cu", "navier.cu", "snapshot.cu", "adjoint.cu", "lapack.cu", "capi.cu"]
FLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-DRP_EMU", "-I", HERE, "-I", CSRC, "-fvisibility=hidden", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas"]


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
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("cuemu.h", "cuemu_switch.S")]
    deps.append(os.path.join(ROOT, "include", "rustpde_b200.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OUT, src.replace(".cu", ".o"))
        if not force and not _stale(os.path.join(CSRC, src), obj, [os.path.join(HERE, "cuemu.h")]):
            return obj
        cmd = ["g++"] + FLAGS + ["-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for %s:\n%s" % (src, r.stderr[-6000:]))
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    sw = os.path.join(OUT, "cuemu_switch.o")
    subprocess.run(["gcc", "-c", os.path.join(HERE, "cuemu_switch.S"), "-o", sw], check=True)
    r = subprocess.run(["g++", "-shared", "-o", LIB] + objs + [sw, "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""Build libmoldy_b200.so (CUDA C++, sm_100a only) in-tree with nvcc.

The shared object is git-ignored but travels to the GPU box with the repo
snapshot.  There is deliberately no other backend and no CPU fallback.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmoldy_b200.so")
SOURCES = ["mdb_host.cpp", "mdb_cells.cu", "mdb_pair.cu", "mdb_pair_tiled.cu", "mdb_kspace.cu", "mdb_molframe.cu", "mdb_md.cu", "mdb_peer.cu", "mdb_group.cu", "mdb_engine.cu", "moldy_abi.cu"]
HEADERS = ["mdb_internal.h", "mdb_math.cuh", os.path.join(ROOT, "include", "moldy_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False, extra=(), out: str = LIB, tag: str = "") -> str:
    """extra: additional nvcc flags (e.g. -DMDB_TILED_MINB=3) for tuning variants written to `out`."""
    if not force and out == LIB and not _stale():
        return LIB
    objdir = os.path.join(HERE, "build" + tag)
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        objs.append(o)
        cmd = [NVCC, *FLAGS, *extra, "-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    fail = False
    for s, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s}\n{log}\n")
        fail |= p.returncode != 0
    if fail:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", out, *objs, "-lcudart"])
    return out


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))

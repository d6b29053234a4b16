"""In-tree build of libfock_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build() and on first import."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfock_b200.so")
SOURCES = ["capi.cu", "slos.cu", "slos_mu.cu", "slos_thin.cu", "slos_masked.cu", "permanent.cu", "cc2017.cu", "peaks.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "slos_tile.cuh"), os.path.join(os.path.dirname(HERE), "include", "fock_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfock_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    host_cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + (["-ccbin", host_cc] if host_cc else []) + ["-c", src, "-o", obj]
            jobs.append((s, cmd, obj))

    def run(job):
        name, cmd, obj = job
        p = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{p.stdout}\n{p.stderr}")
        if verbose:
            sys.stderr.write(p.stderr)
        return name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + (["-ccbin", host_cc] if host_cc else []) + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

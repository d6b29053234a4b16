"""Builds a tuning variant of libfock_b200.so with extra -D macros into scratch/variants/libfock_<name>.so (not the product
library; load it with FOCK_B200_LIB=<path>).   python tools/build_variant.py regcol3 -DTH6_REGCOL -DTH_MINB=3"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceval_b200 import _build as B

name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "scratch", "variants")
obj_dir = os.path.join(out_dir, "obj_" + name)
os.makedirs(obj_dir, exist_ok=True)
nvcc = B._nvcc()


def cc(src):
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    p = subprocess.run([nvcc] + B.NVCC_FLAGS + defs + ["-ccbin", "/usr/bin/g++", "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
    open(obj + ".log", "w").write(p.stdout + p.stderr)
    assert p.returncode == 0, p.stderr
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, B.SOURCES))
lib = os.path.join(out_dir, f"libfock_{name}.so")
subprocess.check_call([nvcc, "-shared", "-o", lib] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
log = open(os.path.join(obj_dir, "slos_thin.o.log")).read()
i = log.find("slos_thin6_kernelILi16ELi2ELb0")
print(lib, [ln.strip() for ln in log[i:i + 600].split("\n") if "Used" in ln or "spill" in ln][:2])

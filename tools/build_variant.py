#!/usr/bin/env python3
"""Build an experimental variant of libbioem_b200.so: the fused kernel of ONE image edge is
recompiled with extra -D flags and linked with the regular objects.
usage: build_variant.py <N> <out.so> [-DFLAG=..]...   (run with BIOEM_B200_LIB=<out.so>)"""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from bioem_b200 import build as b  # noqa: E402

n, out, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
b.build()
obj = out + f".lik_{n}.o"
cmd = [b.NVCC] + b.CFLAGS + flags + [f"-DBIOEM_N={n}", "-x", "cu", "-c", os.path.join(b.CSRC, "lik_instance.inl"), "-o", obj]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr[-3000:])
for ln in r.stderr.splitlines():
    if "spill" in ln or "Used" in ln:
        print(ln.strip())
# the other image edges are stubbed out: a variant library only has to run the edge under test (and stays small)
stub = out + ".stubs.cpp"
with open(stub, "w") as f:
    f.write("struct CUstream_st;\n")
    for m in b.sizes():
        if str(m) != str(n):
            f.write(f'extern "C" int bioem_lik_launch_{m}(const void *, int, int, CUstream_st *) {{ return 1; /* cudaErrorInvalidValue */ }}\n')
stub_o = out + ".stubs.o"
r = subprocess.run(["/usr/bin/g++", "-O1", "-fPIC", "-c", stub, "-o", stub_o], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr[-3000:])
objs = [os.path.join(b.OBJ, f) for f in os.listdir(b.OBJ) if f.endswith(".o") and not f.startswith("lik_")] + [obj, stub_o]
r = subprocess.run([b.NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", out] + objs, capture_output=True, text=True)
os.remove(stub)
os.remove(stub_o)
if r.returncode:
    sys.exit(r.stderr[-3000:])
os.remove(obj)
print(out)

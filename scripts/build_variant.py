"""Build libfrb200 with extra -D flags into another file (A/B runs of compile-time knobs on the GPU box):
    python scripts/build_variant.py pend32 -DFR_K2_PEND_K32=32 -DFR_K2_STAGES_K32=4
    FRB200_LIB=financial_rag_b200/libfrb200_pend32.so python scripts/sweep_rows.py ..."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from financial_rag_b200 import build as b

name, defs = sys.argv[1], sys.argv[2:]
obj_dir = os.path.join(b.OBJ, "variant_" + name)
os.makedirs(obj_dir, exist_ok=True)
b.build()  # the default objects are reused for the sources a knob does not touch
nvcc = b._nvcc()
objs = []
for src in b.SOURCES:
    text = open(os.path.join(b.CSRC, src)).read()
    touched = any(d.split("=")[0][2:] in text for d in defs)
    obj = os.path.join(obj_dir if touched else b.OBJ, src.replace(".cu", ".o"))
    if touched:
        subprocess.run([nvcc, *b.NVCC_FLAGS, *defs, "-c", os.path.join(b.CSRC, src), "-o", obj], check=True)
    objs.append(obj)
out = os.path.join(b.PKG, f"libfrb200_{name}.so")
subprocess.run([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl"], check=True)
print(out)

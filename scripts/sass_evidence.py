"""Blackwell opcode evidence: per-kernel counts of the tcgen05 / TMA / TMEM SASS mnemonics in libfrb200.so, plus the
register / spill / shared-memory lines of ptxas.  Runs without a GPU:  python scripts/sass_evidence.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "financial_rag_b200", "libfrb200.so")
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "FFMA", "HFMA2", "HMMA", "LDG", "STG", "ATOMG", "REDG", "SHFL", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    for name in OPS:
        if op == name or op.startswith(name + "."):
            counts[cur][name] += 1
            break
print("# SASS opcode counts per kernel of financial_rag_b200/libfrb200.so (cuobjdump -sass, sm_100a)")
print("# UTCHMMA = tcgen05.mma (kind::f16 family), .2CTA = cta_group::2; UTMALDG = TMA tensor load (cp.async.bulk.tensor);")
print("# LDTM = tcgen05.ld (TMEM -> registers); UTCBAR = tcgen05.commit; SYNCS = mbarrier ops")
hdr = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "FFMA", "LDG", "SHFL"]
print(f"{'kernel':100s} " + " ".join(f"{h:>12s}" for h in hdr))
tot = collections.Counter()
others = collections.Counter()
for k, c in counts.items():
    if not any(c[h] for h in hdr):
        continue
    tot.update(c)
    if not (c["UTCHMMA"] or c["UTCHMMA.2CTA"] or c["UTMALDG"] or c["LDTM"]):
        others[re.sub(r"<.*", "", k)] += 1  # CUDA-core kernels: one line per template family below
        continue
    print(f"{k[:100]:100s} " + " ".join(f"{c[h]:12d}" for h in hdr))
print("# CUDA-core kernels (no tensor-core / TMA opcodes), instances per template: " + ", ".join(f"{k} x{v}" for k, v in others.items()))
print(f"{'TOTAL':100s} " + " ".join(f"{tot[h]:12d}" for h in hdr))
print()
print("# ptxas -v (registers, spills, static shared memory) for the tensor-core and streaming scan kernels")
from financial_rag_b200 import build as frbuild  # noqa: E402
for src in ("scan_mma.cu", "scan_mma_small.cu", "scan_stream.cu"):
    cmd = [frbuild._nvcc(), *frbuild.NVCC_FLAGS, "-Xptxas", "-v", "-c", os.path.join(frbuild.CSRC, src), "-o", "/dev/null"]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    fn = None
    for line in err.splitlines():
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
        m = re.search(r"Used (\d+) registers.*", line)
        if m and fn and ("scan_mma" in fn or "scan_stream_kernel<true, 2, 1, 1, false>" in fn):
            print(f"{fn[:110]:110s} {line.strip().replace('ptxas info    : ', '')}")
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and fn and "scan_mma" in fn and (m.group(1) != "0" or m.group(2) != "0"):
            print(f"{'':110s} SPILL: {line.strip()}")

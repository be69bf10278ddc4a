"""SASS evidence for the tensor-core kernels of the shipped library (CPU: cuobjdump only).

    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt

Per kernel family: how many of each Blackwell mnemonic the SASS contains (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTCBAR = tcgen05.commit, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store, SYNCS = mbarrier, HMMA = mma.sync, LDSM = ldmatrix,
LDGSTS = cp.async, FFMA2/FADD2 = packed fp32x2, MUFU.TANH/EX2 .F16 = half2 SFU path, REDUX) and the first occurrence of each."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sst_b200", "libsstb200.so")
FAMILIES = ["sra_chain2_kernel", "win_attn_batch_kernel", "umma_gemm_kernel", "vfe_l1_umma_kernel", "sir_a_kernel", "sir_b_kernel",
            "dw_gemm_kernel", "pos_qk_kernel"]
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "LDSM", "LDGSTS", "FFMA2", "FADD2",
             "MUFU.TANH", "MUFU.EX2", "HFMA2", "REDUX", "ACQBULK", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : ", sass)
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}  ({len(parts) - 1} functions)")
    for fam in FAMILIES:
        fns = [p for p in parts[1:] if fam in p.split("\n", 1)[0]]
        if not fns:
            print(f"\n## {fam}: not in the library")
            continue
        fn = max(fns, key=len)   # the largest instantiation
        name = fn.split("\n", 1)[0].strip()
        lines = [l for l in fn.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        print(f"\n## {fam}: {len(fns)} instantiation(s); largest = {name[:110]}  ({len(lines)} SASS instructions)")
        for m in MNEMONICS:
            hits = [l for l in lines if re.search(r"\b" + re.escape(m), l)]
            if hits:
                first = re.sub(r"\s+", " ", hits[0].split("*/", 1)[1].split("/*")[0]).strip()
                print(f"  {m:10s} x{len(hits):4d}   first: {first[:120]}")


if __name__ == "__main__":
    sys.exit(main())

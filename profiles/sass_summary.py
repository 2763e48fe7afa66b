"""Static SASS summary of the hot kernels of libtmfwm.so (cuobjdump -sass): instruction count, opcode
histogram, and the lines that prove the TMA path (UTMALDG / UTMASTG / UTMAPF, mbarrier SYNCS).
    python profiles/sass_summary.py > profiles/r02_sass_hot_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "thatsmyface_b200", "lib", "libtmfwm.so")
WANT = [("k_embed_tile", "FAST embed, TMA-tiled persistent (default)"), ("k_embed_fastILi8", "FAST embed, per-thread (any alignment)"),
        ("k_extract_fastILi8", "FAST extract, per-thread (default)"), ("k_embed_faithfulILi8ELb0", "FAITHFUL embed, block 8"),
        ("k_extract_faithfulILi8", "FAITHFUL extract, block 8"), ("k_svd8x8ILb0", "SVD tap, values only")]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
print("# cuobjdump -sass", os.path.relpath(LIB, ROOT), "(sm_100a)")
for key, title in WANT:
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if key not in name:
            continue
        ops = collections.Counter()
        proof = []
        n = 0
        for line in f.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
            if not m:
                continue
            ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
            op = ins.split()[0]
            n += 1
            base = op.split(".")[0]
            ops[op if base in ("UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UBLKCP", "FFMA2", "FADD2", "FMUL2", "IDP", "I2IP", "MUFU", "LDS", "STS", "LDG", "STG") else base] += 1
            if base in ("UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP") or (base == "SYNCS" and len(proof) < 12):
                proof.append(ins)
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(f"\n== {title}\n   {demangled[:150]}\n   {n} SASS instructions ({n * 16 / 1024:.1f} KB)")
        print("   " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(40)))
        for p in proof:
            print("     " + p)
        break

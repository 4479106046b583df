#!/usr/bin/env python
"""Static SASS mnemonic counts per GEMM kernel of libhalva_dpa.so (cuobjdump -sass; runs on the CPU box):
    python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "halva_b200", "libhalva_dpa.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
for m in re.finditer(r"Function : (\S+)", sass):
    names[m.group(1)] = None
dem = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(names, dem))
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTCATOMSWS", "HMMA"]
rows = collections.OrderedDict()
for blk in sass.split("Function : ")[1:]:
    mang = blk.split()[0]
    d = names.get(mang, mang)
    m = re.search(r"(dpa_gemm\w*_kernel)<(?:dpa::)?GemmCfg<([^>]*)>, (?:dpa::)?(\w+<[^>]*>)", d)
    if not m:
        continue
    key = f"`{m.group(1).replace('_kernel', '')}<{m.group(2).replace(' ', '').replace('(bool)', '').replace('(int)', '')}>/{m.group(3).replace(' ', '').replace('(bool)', '').replace('(int)', '')}`"
    c = {k: len(re.findall(r"\b" + k + r"\b[\w.]*", blk)) for k in MN}
    c["2CTA"] = len(re.findall(r"UTCHMMA\.2CTA", blk))
    rows[key] = c
print("# SASS mnemonics of libhalva_dpa.so (`cuobjdump -sass`, sm_100a; static counts per GEMM kernel) -- `python tools/sass_mnemonics.py`\n")
print("`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `UTCBAR` = tcgen05.commit, `LDTM` / `STTM` = tcgen05.ld / tcgen05.st, `UTMALDG` = TMA tensor")
print("load, `UTCATOMSWS` = TMEM alloc/dealloc.  No `HMMA` (legacy mma.sync) anywhere.  GemmCfg<BN, DUAL, A major, B major, cta_group, K segments>;")
print("`dpa_gemm_split` = the fp32-input kernels (six split-bf16 K segments, K-chunk promotion: the `STTM` parks the promoted sum in TMEM).\n")
print("| kernel | UTCHMMA (of which .2CTA) | UTCBAR | LDTM | STTM | UTMALDG | UTCATOMSWS | HMMA |\n|---|---:|---:|---:|---:|---:|---:|---:|")
for k in sorted(rows):
    c = rows[k]
    print(f"| {k} | {c['UTCHMMA']} ({c['2CTA']}) | {c['UTCBAR']} | {c['LDTM']} | {c['STTM']} | {c['UTMALDG']} | {c['UTCATOMSWS']} | {c['HMMA']} |")

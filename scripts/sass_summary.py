#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the shipped library (run HERE, no GPU needed):
  python scripts/sass_summary.py > profiles/r02_sass_summary.txt
Counts the wide loads/stores and any TMA / bulk-copy / mbarrier instruction of every kernel in liblmb200.so."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "line_mod_pipeline_b200", "liblmb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
names = [f.split("\n", 1)[0].strip() for f in funcs]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print("SASS of the shipped liblmb200.so (cuobjdump -sass, sm_100a): per kernel, instruction count, wide memory operations and the 14 most frequent opcodes.\n")
tot_tma = 0
for f, name in zip(funcs, dem):
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    if not ops:
        continue
    base = collections.Counter()
    for k, v in ops.items():
        base[k.split(".")[0]] += v
    cnt = lambda p: sum(v for k, v in ops.items() if k.startswith(p))
    tma = cnt(("UTMALDG", "UBLKCP", "LDGSTS", "SYNCS"))
    tot_tma += tma
    print("%s\n  %d instructions; LDG.E.128: %d, LDG.E.64: %d, STG.E.128: %d, LDS.128: %d, UTMALDG/UBLKCP/LDGSTS/SYNCS: %d\n  %s\n" % (
        re.sub(r"\(.*", "", name), sum(ops.values()), cnt("LDG.E.128"), cnt("LDG.E.64"), cnt("STG.E.128"), cnt("LDS.128"), tma,
        ", ".join("%s %d" % kv for kv in base.most_common(14))))
print("TMA / bulk-copy / mbarrier instructions in the whole library: %d — the kernels gather through L1 with LDG.E.128(.CONSTANT) / LDG.E.64 (DESIGN.md section 7)." % tot_tma)

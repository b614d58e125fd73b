#!/usr/bin/env python
"""Static SASS instruction mix of the built library per kernel (cuobjdump -sass): counts of the mnemonics that prove
what the code uses -- FFMA / FFMA2 (packed fp32), LDS / LDS.128, PRMT, UTMALDG (TMA), SYNCS (mbarrier), SHFL, MUFU.
    python tools/sass_mix.py [kernel-regex] > profiles/rN_sass_mix.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "oclcomputervision_b200", "libraisr_b200.so")
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["FFMA2", "FFMA", "FMUL2", "FADD2", "FMUL", "FADD", "HADD2", "PRMT", "LDS.128", "LDS.64", "LDS", "STS", "LDG", "STG", "UTMALDG", "SYNCS", "SHFL", "MUFU",
        "BAR", "IMAD", "ISETP", "BRA"]
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = name if (pat is None or pat.search(name)) else None
        if cur:
            per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    per[cur]["total"] += 1
    for k in KEYS:
        if op == k or op.startswith(k + ".") or (k in ("LDS.128", "LDS.64") and op.startswith(k)):
            if k == "LDS" and (op.startswith("LDS.128") or op.startswith("LDS.64")):
                continue
            if k == "FFMA" and op.startswith("FFMA2"):
                continue
            per[cur][k] += 1
            break
print("static SASS instruction counts per kernel of libraisr_b200.so (sm_100a); columns: " + " ".join(["total"] + KEYS))
for name, c in per.items():
    print("%s\n    %s" % (name[:150], "  ".join("%s=%d" % (k, c[k]) for k in ["total"] + KEYS if c[k])))

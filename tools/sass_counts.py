#!/usr/bin/env python
"""cuobjdump -sass of libasr_b200.so -> per-kernel counts of the opcodes that prove tcgen05 / TMEM / TMA use
(profiles/<tag>_sass.txt).  Runs on the build box (no GPU needed).
    python tools/sass_counts.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "chinese_asr_b200", "libasr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF",
         "SYNCS", "LDGSTS", "MUFU", "FFMA2", "FMNMX", "HMMA", "ACQBULK", "UCGABAR"]
name = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("asr::", "")
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and name:
        counts[name]["_total"] += 1
        op = m.group(1)
        if op in WATCH:
            counts[name][op] += 1
            if ".2CTA" in m.group(2):
                counts[name][op + ".2CTA"] += 1
out = os.path.join(ROOT, "profiles", f"{tag}_sass.txt")
with open(out, "w") as f:
    f.write("# `cuobjdump -sass chinese_asr_b200/libasr_b200.so` (sm_100a), opcode counts per kernel (tools/sass_counts.py).\n"
            "# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st\n"
            "# (TMEM), UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier,\n"
            "# LDGSTS = cp.async, UCGABAR = cluster barrier.\n\n")
    tot = collections.Counter()
    for k, c in counts.items():
        items = ", ".join(f"{op} {n}" for op, n in c.items() if op != "_total")
        f.write(f"{k}: {c['_total']} instructions" + (f"; {items}" if items else "") + "\n")
        tot.update(c)
    f.write("\nlibrary total: " + ", ".join(f"{op} {n}" for op, n in tot.items() if op != "_total") + "\n")
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    f.write("\n# ldd (no cuBLAS / cuDNN / cuFFT):\n" + "".join("#   " + l.strip() + "\n" for l in ldd.splitlines()))
print(open(out).read()[-1800:])

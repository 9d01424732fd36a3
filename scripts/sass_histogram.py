"""SASS opcode histogram of libtta_b200.so (cuobjdump -sass): which kernels issue tcgen05 / TMEM / TMA instructions.
    python scripts/sass_histogram.py profiles/sass_r2.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "multimodal_tta_b200", "libtta_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMACCTL", "SYNCS", "ELECT", "HMMA", "FFMA",
       "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "MUFU", "SHFL"]
total = collections.Counter()
per_kernel = collections.defaultdict(collections.Counter)
inst = collections.Counter()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        b = re.search(r"(conv_tc_kernel|conv_t2s_kernel|wgrad_tc_kernel)", name)
        cur = b.group(1) if b else "other"
        inst[cur] += 1
        continue
    if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        for o in OPS:
            if re.search(r"(?<![A-Z0-9_])" + o + r"(?![A-Z0-9_])", line):
                total[o] += 1
                per_kernel[cur][o] += 1
rows = ["# SASS opcode histogram of libtta_b200.so (cuobjdump -sass, sm_100a), round 2 (final build)", "",
        "Proof of the instruction families the kernels are built on: UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld "
        "(TMEM -> registers), UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (bulk copy), UTCBAR = "
        "tcgen05.commit -> mbarrier, SYNCS = mbarrier ops.  No HMMA (mma.sync / wmma) anywhere.  Written by "
        "`scripts/sass_histogram.py`.", "", "| opcode | count (whole library) |", "|---|---|"]
rows += [f"| {o} | {total[o]} |" for o in OPS]
rows += ["", "## kernels that issue tcgen05 / TMA instructions", "", "| kernel | UTCHMMA | LDTM | UTMALDG | UBLKCP | UTCBAR |",
         "|---|---|---|---|---|---|"]
for k in ("conv_tc_kernel", "conv_t2s_kernel", "wgrad_tc_kernel"):
    c = per_kernel[k]
    rows.append(f"| {k} ({inst[k]} instantiations) | {c['UTCHMMA']} | {c['LDTM']} | {c['UTMALDG']} | {c['UBLKCP']} | {c['UTCBAR']} |")
open(sys.argv[1], "w").write("\n".join(rows) + "\n")
print("\n".join(rows))

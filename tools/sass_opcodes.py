"""Per-kernel counts of the SASS mnemonics that tell a Blackwell-native kernel from a recompiled one (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), HMMA (mma.sync), LDGSTS (cp.async).
usage: python tools/sass_opcodes.py [libcellvit_b200.so] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
lib = sys.argv[1] if len(sys.argv) > 1 else "cellvit_b200/libcellvit_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|HMMA|LDGSTS|LDSM|MUFU)\b")
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        counts[kern] = collections.Counter()
        continue
    if kern:
        m = pat.search(line)
        if m:
            counts[kern][m.group(1)] += 1
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMAPF", "HMMA", "LDGSTS", "LDSM", "MUFU"]
print(f"{'kernel':58s}" + "".join(f"{c:>9s}" for c in cols))
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print(f"{k[:58]:58s}" + "".join(f"{c.get(x, 0):9d}" for x in cols))

"""Opcode census of libfdfd_b200.so (cuobjdump -sass): per kernel the tensor / copy / fp64 instruction counts that show
which hardware path it uses.  Written to profiles/r02_sass_summary.txt by tools/run_profiles.sh.

tcgen05 (UTCMMA / UTCHMMA ...), TMEM (LDTM / STTM) and TMA (UTMALDG / UTMASTG) have no FP64 kind on sm_100a, so the
FP64 tensor path is DMMA.8x8x4 (mma.sync.m8n8k4.f64); LDGSTS = cp.async."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "fdfdpy_b200/libfdfd_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        counts[kern][m.group(1)] += 1
keys = ["DMMA", "LDGSTS", "LDS", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "BAR", "LDG", "STG", "UTMALDG", "UTMASTG", "UTCMMA", "LDTM", "STTM"]
tot = collections.Counter()
print("%-86s %6s " % ("kernel", "instr") + " ".join("%7s" % k for k in keys))
for k, c in counts.items():
    tot.update(c)
    print("%-86s %6d " % (k[:86], sum(c.values())) + " ".join("%7d" % c.get(x, 0) for x in keys))
print("%-86s %6d " % ("TOTAL (%d kernels)" % len(counts), sum(tot.values())) + " ".join("%7d" % tot.get(x, 0) for x in keys))

"""SASS instruction census of libbke.so: python scripts/sass_census.py > profiles/<round>_sass_census.txt
Per kernel: instruction count and the mnemonics that prove (or rule out) TMA / mbarrier / cluster / tensor-core use."""
import collections
import re
import subprocess
import sys
import os

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "filterpy_b200", "_C", "libbke.so")
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "UCGABAR", "UTCMMA", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "HMMA", "DFMA", "DADD", "DMUL",
        "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "LDL", "STL", "MUFU", "ATOM", "RED", "MEMBAR", "FENCE", "NANOSLEEP"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = ins.match(line)
    if m and cur is not None:
        op = m.group(1)
        cur["instr"] += 1
        for k in KEYS:
            if op.startswith(k):
                cur[k] += 1
                break
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# SASS instruction census of filterpy_b200/_C/libbke.so (sm_100a): cuobjdump -sass, per-kernel counts")
print("# TMA: UTMALDG (tensor) / UBLKCP (bulk) / UTMAPF (prefetch); mbarrier: SYNCS; cluster barrier: UCGABAR;")
print("# tensor memory / 5th-gen MMA: UTCHMMA (tcgen05.mma kind::tf32), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc) —")
print("# only in kf_cov_tc_kernel<16|32> (csrc/kf_tc.cu), the one shape class that fills an MMA fragment; no HMMA / mma.sync anywhere")
print()
print("kernels: %d   total: " % len(per) + ", ".join("%s=%d" % (k, tot[k]) for k in ["instr"] + KEYS if tot[k]))
print()
try:
    dem = subprocess.run(["cu++filt"] + list(per.keys()), capture_output=True, text=True).stdout.splitlines()
except Exception:
    dem = list(per.keys())
for (name, c), d in zip(per.items(), dem):
    short = d[:d.rfind(")(") + 1] if ")(" in d else re.sub(r"\((?!bool|int)[^()]*\)$", "", d)      # drop the parameter list, keep template arguments
    short = short.replace("(bool)", "").replace("(int)", "")
    print("%-110s instr=%6d  %s" % (short[:110], c["instr"], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))

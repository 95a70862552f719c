"""SASS opcode census of libshgvqa.so per kernel (cuobjdump -sass): which kernels are Blackwell-native.
    python tools/sass_census.py > profiles/sass_census_rNN.txt        # runs without a GPU
UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add
(cp.async.bulk.tensor, cp.reduce.async.bulk.tensor), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations; HMMA (legacy mma.sync) must be absent."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "shg_vqa_b200", "libshgvqa.so")
sass = subprocess.check_output(["cuobjdump", "-sass", so]).decode()
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "SYNCS", "REDG",
       "ATOMG", "MUFU.EX2", "MUFU.RCP"]
cur, counts = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for o in OPS:
            if re.search(r"\b" + re.escape(o) + r"\b", ln):
                counts[cur][o] += 1
        if "UTCHMMA" in ln and ".2CTA" in ln:
            counts[cur]["UTCHMMA.2CTA"] += 1


def demangle(n):
    try:
        return subprocess.check_output(["c++filt", n]).decode().strip().split("(")[0]
    except Exception:
        return n


tot = collections.Counter()
rows = []
for k, c in counts.items():
    if not sum(c.values()):
        continue
    tot.update(c)
    rows.append("%-72s %s" % (demangle(k)[:72], " ".join("%s=%d" % (o, c[o]) for o in OPS + ["UTCHMMA.2CTA"] if c[o])))
print(__doc__.strip().split("\n", 2)[2] if False else "SASS opcode census of %s (sm_100a), per kernel; see tools/sass_census.py for the legend." % os.path.relpath(so, ROOT))
print("TOTAL: " + " ".join("%s=%d" % (o, tot[o]) for o in OPS + ["UTCHMMA.2CTA"] if tot[o] or o == "HMMA"))
print()
print("\n".join(rows))

"""Print the hottest SASS lines (warp-stall samples) of an ncu --page source --csv dump.
usage: ncu -i rep --page source --csv > x.csv ; python tools/ncu_hot.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isrc, ci = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    try:
        s = float(r[ci])
    except Exception:
        continue
    top = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    data.append((s, r[isrc][:100], " ".join("%s=%d" % (h[6:], v) for v, h in top if v > 0)))
tot = sum(d[0] for d in data)
print("total samples", tot)
for idx, (s, src, st) in enumerate(data):
    pass
for s, src, st in sorted(data, reverse=True)[:n]:
    print("%5.1f%%  %-100s %s" % (100 * s / tot, src, st))

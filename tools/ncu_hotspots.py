"""Stall-sample hot spots per SASS instruction from `ncu -i X.ncu-rep --page source --csv --kernel-id :::N`.

    ncu -i rep --page source --csv --kernel-id :::6 > /tmp/k.csv ; python tools/ncu_hotspots.py /tmp/k.csv [min_pct]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc = hdr.index("Source"); iall = hdr.index("Warp Stall Sampling (All Samples)")
iex = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def I(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(I(r[iall]) for r in data)
print(rows[0][1][:150]); print("total samples", tot, "instructions", len(data))
cum = 0
for n, r in enumerate(data):
    s = I(r[iall]); cum += s
    if s >= tot * minp / 100:
        st = sorted([(I(r[i]), h[6:]) for i, h in stall_cols], reverse=True)[:2]
        print("%5d %-78s %6d %5.1f%% cum %3.0f%% ex %8d %s" % (n, r[isrc].strip()[:78], s, 100 * s / tot, 100 * cum / tot, I(r[iex]), st))

"""Executed warp instructions and stall samples of one kernel, per barrier-delimited SASS section and opcode, from
`ncu -i X.ncu-rep --page source --csv --kernel-id :::N > k.csv`.

    python tools/sass_sections.py k.csv [warps_for_per_warp_figures]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
isrc, iex, iall = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[iex]) for r in data)
tots = sum(int(r[iall]) for r in data)
print(rows[0][1][:120])
print("warp instructions", tot, "SASS lines", len(data), "stall samples", tots)
sec, secs = 0, collections.OrderedDict()
for r in data:
    s = r[isrc].strip()
    f = s.split()
    op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
    d = secs.setdefault(sec, {"ex": 0, "sm": 0, "ops": collections.Counter(), "n": 0})
    d["ex"] += int(r[iex]); d["sm"] += int(r[iall]); d["ops"][op] += int(r[iex]); d["n"] += 1
    if op == "BAR":
        sec += 1
for k, d in secs.items():
    print("section %d: %d SASS lines, %.1f%% of executed, %.1f%% of samples; %s" % (
        k, d["n"], 100.0 * d["ex"] / tot, 100.0 * d["sm"] / max(tots, 1),
        " ".join("%s=%.1f%%" % (o, 100.0 * c / tot) for o, c in d["ops"].most_common(10))))

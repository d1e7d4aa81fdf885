"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` output (scratch tool).
usage: python tools/ncu_stalls.py source.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
num = lambda s: int(float(s)) if s not in ("", "-") else 0
tot = sum(num(r[isamp]) for r in data)
print("total samples", tot, "sass rows", len(data), "warp-instructions", sum(num(r[iex]) for r in data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + num(r[i])
print({k: "%.1f%%" % (100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]})
for r in sorted(data, key=lambda r: -num(r[isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    st = sorted([(num(r[i]), hdr[i][6:]) for i in stall_cols], reverse=True)[:2]
    print(r[ia][-5:], "%5d %4.1f%%" % (num(r[isamp]), 100 * num(r[isamp]) / tot), r[iex].rjust(8), r[isrc][:64].ljust(64), st)

"""Stall samples along the hottest loop of a kernel (buckets of SASS rows) from an `ncu --page source --csv` dump (scratch tool).
usage: python tools/ncu_loop.py source.csv <exec count of the loop rows, comma separated> [bucket]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
data = data[:len(data) // 2] if len(data) > 3000 and data[0][0] == data[len(data) // 2][0] else data
isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
num = lambda s: int(float(s)) if s not in ("", "-") else 0
counts = [int(c) for c in sys.argv[2].split(",")]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 25
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
idx = [i for i, r in enumerate(data) if num(r[iex]) in counts]
loop = data[idx[0]:idx[-1] + 1]
tot = sum(num(r[isamp]) for r in data)
print("rows", len(loop), "loop samples", sum(num(r[isamp]) for r in loop), "of", tot)
for k in range(0, len(loop), B):
    chunk = loop[k:k + B]
    sm = sum(num(r[isamp]) for r in chunk)
    agg = {}
    for r in chunk:
        for i in stall_cols:
            agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + num(r[i])
    top = [(a, b) for a, b in sorted(agg.items(), key=lambda x: -x[1])[:3] if b]
    ops = [(r[isrc].split()[1] if r[isrc].startswith('@') else r[isrc].split()[0]).split('.')[0] for r in chunk if r[isrc].split()]
    key = [o for o in ops if o in ('LDTM', 'STTM', 'LDG', 'STG', 'STS', 'LDS', 'SYNCS', 'MEMBAR', 'FENCE', 'UTCHMMA', 'UTCBAR', 'UBLKCP')]
    print("%4d %5d %-70s %s" % (k, sm, top, ' '.join(key)[:70]))

"""Summarise an .ncu-rep here (no GPU): key counters per launch + the top stall instructions of one kernel (scratch tool).
usage: python tools/ncu_summary.py <rep> [kernel-regex-for-source-page] [launch-skip]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
names = [d[idx["Kernel Name"]].split("(")[0].replace("void <unnamed>::", "")[:34] for d in data]
print(f"{'metric':62s}", *[f"{n:>22s}" for n in names])
for w in want:
    if w in idx:
        vals = [d[idx[w]] for d in data]
        try:
            if all(float(v) < 0.05 for v in vals) and "stalled" in w:
                continue
        except ValueError:
            pass
        short = w.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")
        print(f"{short[:62]:62s}", *[f"{v[:22]:>22s}" for v in vals])
if len(sys.argv) > 2:
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{sys.argv[2]}", "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[h], rows[h + 1:]
    ia, isrc, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    seen, uniq = set(), []
    for k, r in enumerate(data):
        if len(r) > ia and r[ia].isdigit() and r[0] not in seen:
            seen.add(r[0]); uniq.append((int(r[ia]), k, r[isrc].strip(), r[iex]))
    tot = sum(u[0] for u in uniq)
    print("\ntop stall instructions (", tot, "samples,", len(uniq), "instructions )")
    for s, k, text, ex in sorted(uniq, reverse=True)[:32]:
        print(f"{s:6d} {100 * s / tot:5.1f}%  #{k:5d} ex={ex:>8s}  {text[:100]}")

"""Turn the round's ncu captures (gpurun_out/*.ncu-rep, read here without a GPU) into the committed profile summaries:
  python tools/make_profiles.py step  <rep> <out.md> <title> [traffic.json commit batch]   per-launch table (+ DRAM bytes json)
  python tools/make_profiles.py list  <launches.csv> <out.md> <title>                        launch list -> share per kernel
"""
import csv, io, json, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def step(rep, out, title, traffic=None, commit=None, batch=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", ""]
    per_kernel = collections.defaultdict(list)
    for n, d in enumerate(data):
        name = short(d[idx["Kernel Name"]])
        lines += [f"## launch {n}: {name}", "", "| metric | value | unit |", "|---|---:|---|"]
        for k in KEYS:
            if k in idx:
                lines.append(f"| {k} | {d[idx[k]]} | {units[idx[k]]} |")
        rd = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        per_kernel[name.split("<")[0]].append(rd + wr)
        lines.append("")
    open(out, "w").write("\n".join(lines))
    if traffic:
        try:
            js = json.load(open(traffic))
        except Exception:
            js = {}
        js.update({"commit": commit, "batch_per_gpu": int(batch), "source": f"ncu --set full --clock-control none, one capture per launch ({out})",
                   "kernels": {k: sum(v) / len(v) for k, v in per_kernel.items()},
                   "launches_per_kernel": {k: len(v) for k, v in per_kernel.items()}})
        json.dump(js, open(traffic, "w"), indent=1)


def launch_list(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(r[iu], 1e-6)
        t = tot[short(r[ik])[:90]]
        t[0] += 1
        t[1] += ms
    total = sum(t[1] for t in tot.values())
    ours = sum(t[1] for k, t in tot.items() if not k.startswith(("at::", "void at::", "ncclDevKernel", "ncclKernel")) and "at::native" not in k and "cub::" not in k)
    lines = [f"# {title}", "", f"{sum(t[0] for t in tot.values())} launches, {total:.1f} ms in total (per-launch times under ncu are cold-cache and serialised: compare SHARES).",
             "", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
        lines.append(f"| {k} | {t[0]} | {t[1]:.3f} | {100 * t[1] / total:.1f}% |")
    lines += ["", f"Kernels of libroomslam_b200.so: {100 * ours / total:.1f}% of the time; torch plumbing (fills, copies, casts, cat, reductions): {100 * (1 - ours / total):.1f}%."]
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "step":
        step(*sys.argv[2:5], *(sys.argv[5:8] if len(sys.argv) > 5 else []))
    else:
        launch_list(*sys.argv[2:5])

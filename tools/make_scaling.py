"""profiles/r2_scaling.md from the bench lines profiles/r2_bench_n{1,2,4,8}.json (scratch tool)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for n in (1, 2, 4, 8):
    d = None
    for l in open(os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")):
        if l.startswith("{"):
            d = json.loads(l)
    rows.append((n, d))
b = rows[0][1]
out = ["# Round 2 scaling (builder-run, `python bench.py --gpus N --steps 20 --warmup 3` under torchrun; one box, N B200s)", "",
       "BASELINE config 3 as written is STRONG scaling: a global batch of 8192 traces over N GPUs (8192/N per GPU). Weak scaling (8192 per GPU) is timed in the same run.",
       "Efficiency here is only for orientation (the driver computes its own): strong = (N=1 ms) / (N * ms at N); weak = value at N / (N * value at 1).", "",
       "| N | strong traces/s | strong ms/step | strong eff | strong e2e traces/s | weak traces/s | weak ms/step | weak eff | invariance (grad fp32 / bf16 rel L2, heatmap) |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---|"]
for n, d in rows:
    w = d.get("weak") or {}
    ic = d.get("invariance_detail") or {}
    se = b["ms_per_step"] / (n * d["ms_per_step"])
    wv = w.get("value", d["value"]); wm = w.get("ms_per_step", d["ms_per_step"])
    we = wv / (n * b["value"])
    invs = "n/a (N = 1)" if n == 1 else (f"{d.get('invariance')}: {ic.get('grad_rel_l2_fp32', 0):.1e} / {ic.get('grad_rel_l2_bf16', 0):.1e}, "
                                         f"grids bit-identical = {ic.get('heatmap_bit_identical')}")
    out.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {se:.2f} | {d['e2e']['value']:.0f} | {wv:.0f} | {wm:.2f} | {we:.2f} | {invs} |")
out += ["", "Per-kernel CUDA-event times of one strong-scaling step (ms; both launches of a kernel summed):", "",
        "| N | rec_fwd_pair | projection | rec_bwd_pair | blk_wgrad | dgrad | decoder fwd+bwd |", "|---|---:|---:|---:|---:|---:|---:|"]
for n, d in rows:
    k = d["train"]["kernel_ms"]
    g = lambda s: sum(v[0] for kk, v in k.items() if s in kk)
    out.append(f"| {n} | {g('rec_fwd'):.2f} | {g('projection'):.2f} | {g('rec_bwd'):.2f} | {g('wgrad'):.2f} | {g('dgrad'):.2f} | {g('decoder'):.2f} |")
out += ["", "Reading: from N = 4 on the step is the serial chain of 4 x 500 time steps (2.1 us forward, 2.3 us backward each, whatever the batch); the time-parallel GEMMs scale.",
        "Round 1's one-CTA-per-tile kernels needed 12.2 ms at 1024 traces per GPU (strong efficiency 0.24 at N = 8); the first CTA-pair kernels of this round 6.9 ms."]
open(os.path.join(ROOT, "profiles", "r2_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

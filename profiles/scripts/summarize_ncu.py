"""Turns gpurun_out/launches.csv (ncu launch list) and gpurun_out/prof_*.ncu-rep (ncu --set full) into the
tracked summaries under profiles/.   usage: python profiles/scripts/summarize_ncu.py r01 prof_coop [prof_task ...]"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GO = os.path.join(ROOT, "gpurun_out")
TIMED = os.environ.get("FM_TIMED_KERNEL", "fm_search_wide_kernel<2, 5, 1, 256, 8, 0>")     # bench.py's timed kernel (r01 / early r02: fm_search_sparse_kernel<2, 2, 3, 256, 4, 0>)
PR = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(tag):
    path = os.path.join(GO, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 8]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg, total = collections.OrderedDict(), 0.0
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        ns = float(r[ix["Metric Value"]]) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ix["Metric Unit"]], 1)
        a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0]); a[0] += 1; a[1] += ns; total += ns
    with open(os.path.join(PR, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3` (1 GPU, 2 Gbp index, 10 M reads)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 8000` (profiles/scripts/capture_ncu.sh) -- per-launch times are cold-cache and serialised;\n"
                "compare SHARES.  Setup kernels (index build, re-block, probe, read synthesis) are outside bench.py's timed regions.\n\n"
                "| total ms | launches | share | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {t / 1e6:.3f} | {c} | {100 * t / total:.1f}% | `{k[:110]}` |\n")
        f.write(f"\nprofiled launches: {sum(c for c, _ in agg.values())}, total {total / 1e6:.1f} ms\n")
        # the timed kernel by launch shape: whole-batch launches are bench.py's `value` region (one launch = one step, nothing else
        # is launched there), the small grids are the 512 K-read chunks of the end-to-end pipeline
        per = collections.OrderedDict()
        for r in rows[1:]:
            if r[ix["Metric Name"]] == "gpu__time_duration.sum" and TIMED in r[ix["Kernel Name"]]:
                ns = float(r[ix["Metric Value"]]) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ix["Metric Unit"]], 1)
                per.setdefault(r[ix["Grid Size"]], []).append(ns)
        if per:
            f.write(f"\n## The timed kernel, `{TIMED}`, by grid\n\n"
                    "| grid | launches | mean ms per launch | where |\n|---|---:|---:|---|\n")
            def blocks(g): return int(g.strip("()").split(",")[0])
            top = max(blocks(g) for g in per)
            for g, v in sorted(per.items(), key=lambda kv: -blocks(kv[0])):
                where = ("`value` region: one launch over the rank's 10 M reads = one step (warm-up + timed steps); it is the only kernel "
                         "launched there, i.e. 100 % of the step" if blocks(g) == top else
                         "`skewed_text` extra key (4 M reads on the 400 Mbp repeat-rich text)" if blocks(g) * 10 > top * 3 else
                         "end-to-end pipeline (`e2e`), one launch per 512 K-read chunk (or tail chunk)")
                f.write(f"| {g} | {len(v)} | {sum(v) / len(v) / 1e6:.3f} | {where} |\n")


def full(tag, name):
    rep = os.path.join(GO, name + ".ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d.get("Kernel Name")}
        for k in KEYS:
            if k in d:
                rec[k] = f"{d[k]} {units[hdr.index(k)]}".strip()
        stalls = {h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(d[h])
                  for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]}
        rec["warp_stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        out.append(rec)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    json.dump(out, open(os.path.join(PR, f"{tag}_{name}.json"), "w"), indent=1)
    # keep the SASS text, stall samples and execution counts of the first profiled launch only
    keep = ["Source", "Warp Stall Sampling (All Samples)", "# Samples", "Instructions Executed"]
    lines = list(csv.reader(io.StringIO(src)))
    with open(os.path.join(PR, f"{tag}_{name}_source.csv"), "w") as f:
        w = csv.writer(f)
        cols, seen_kernels = None, 0
        for r in lines:
            if r and r[0] == "Kernel Name":
                seen_kernels += 1
                if seen_kernels > 1:
                    break
                w.writerow(r[:2]); continue
            if r and r[0] == "Address":
                cols = [r.index(k) for k in keep if k in r]; w.writerow([r[c] for c in cols]); continue
            if cols and len(r) > max(cols):
                w.writerow([r[c].strip() for c in cols])
    return out


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    for n in sys.argv[2:]:
        for rec in full(tag, n):
            print(json.dumps(rec)[:1500])

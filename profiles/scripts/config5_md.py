"""gpurun_out/config5_sweep.jsonl (config5_sweep.py) -> profiles/<tag>_config5_sweep.{jsonl,md}.   usage: python profiles/scripts/config5_md.py r01"""
import json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out", "config5_sweep.jsonl")
rows = [json.loads(l) for l in open(src)]
shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_config5_sweep.jsonl"))
pts, idxs = {}, {}
for r in rows:
    if r["what"] == "index":
        idxs[r["k"]] = r
    elif r["what"] == "search":
        pts.setdefault((r["k"], r["len"]), {})[r["kernel"]] = r
ok = all(p["all_found"] and p["same_as_task"] for d in pts.values() for p in d.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_config5_sweep.md"), "w") as f:
    f.write(f"# {tag}: BASELINE config 5 on one B200 (2 Gbp index, 10 M reads per point) -- `profiles/scripts/config5_sweep.py`\n\n"
            f"Mq/s per kernel; every point verified (all reads found; Task == Coop == Fused == Sparse bit for bit: {'yes' if ok else 'NO'}).  "
            "G LF/s = reference k-step LF steps per second.\n\n"
            "| k | read length | Task Mq/s | Coop Mq/s | Fused Mq/s | Sparse Mq/s | Sparse G LF/s | Sparse / Coop | mean hits |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for (k, ln), d in pts.items():
        f.write(f"| {k} | {ln} | {d['task']['mq_per_s']:.0f} | {d['coop']['mq_per_s']:.0f} | {d['fused']['mq_per_s']:.0f} | {d['sparse']['mq_per_s']:.0f} | "
                f"{d['sparse']['g_ref_lf_steps_per_s']:.1f} | {d['sparse']['mq_per_s'] / d['coop']['mq_per_s']:.1f}x | {d['sparse']['mean_hits']:.2f} |\n")
    f.write("\nTables: " + "; ".join(f"k={k}: SB96 {i['sb96_gb']:.2f} GB, fused {i['fused_bases']} bases {i['fused_gb']:.1f} GB, sparse {i['sparse_bases']} bases {i['sparse_gb']:.1f} GB"
                                     for k, i in idxs.items()) + ".\n\n"
            "12-bp reads never leave the L2-resident phase (120 hits per read; the sparse kernel answers them from its 10-base start table plus one 2-base step);\n"
            "from 50 bp on every kernel sits on the random-access ceiling and the ratio between them is the ratio of block fetches per read.\n"
            "The fused and sparse tables are composed from either index, so a 1-step index searches as fast as a 2-step one.\n"
            "Odd lengths at k=2 (25, 75, 101 bp) are undefined in the reference; here the plain and fused kernels take the last base from the tail table (one extra\n"
            "fetch) and the sparse kernel starts from a lead table that already holds it (`DESIGN.md` 4c): 101 bp costs what 100 bp costs, 75 bp = 7 fetches.\n")

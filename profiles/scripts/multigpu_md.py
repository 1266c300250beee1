"""Turns gpurun_out / profiles r02_dropin_multigpu.jsonl and r02_config5_n8.jsonl into profiles/r02_dropin_multigpu.json,
profiles/r02_dropin_multigpu.md and profiles/r02_config5_n8.md."""
import json, os
PR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
drop = [json.loads(x) for x in open(os.path.join(PR, "r02_dropin_multigpu.jsonl")) if x.strip()]
runs = [r for r in drop if r.get("what", "").startswith("config 4")]
par = [r for r in drop if r.get("what", "").startswith("parity")]
summary = {"what": "BASELINE config 4 as written through the single-process host driver (csrc/fm_host.c) on physical GPUs: 2 Gbp index FILE written by "
                   "gfmi_b200, 100 M x 100 bp reads in a FASTA file, FMGPU_DEVICES=0..N-1; profiles/scripts/dropin_multigpu.sh",
           "runs": runs, "parity": par, "distinct_res_gpu_md5": sorted({r["res_gpu_md5"] for r in runs})}
json.dump(summary, open(os.path.join(PR, "r02_dropin_multigpu.json"), "w"), indent=1)
with open(os.path.join(PR, "r02_dropin_multigpu.md"), "w") as f:
    f.write("# r02: the single-process host driver on physical GPUs (SURVEY 8 row a-9) -- `profiles/scripts/dropin_multigpu.sh`\n\n"
            "BASELINE config 4 **as written**: the 2 Gbp index file (3.0 GB, written by `bin/gfmi_b200 --synth`, md5-identical to the reference builder's), "
            "100 000 000 reads of 100 bp in a 13.5 GB FASTA file, ONE process, `FMGPU_DEVICES=0..N-1`: `loadIndex -> loadQueries -> initResults -> transferCPUtoGPU "
            "(one H2D + re-block on GPU 0, `cudaMemcpyPeer` to the others, sparse-step table on every replica, reads sharded contiguously, ASCII H2D + 2-bit pack) -> "
            "5 x searchIndexGPU -> transferGPUtoCPU -> saveResults`.  Two callers: `bin/fmIndexSearchGPU_b200` (our rebuild of the reference `main()`) and "
            "`oracle/_ref/fmIndexSearchGPU_refmain` = the reference's OWN `common/searchQueries.c`, unmodified, compiled `-DCUDA` and linked against `libfmindex_b200.so`.\n"
            "8 x B200 box, 32 host cores, 1 TB RAM; raw lines in `r02_dropin_multigpu.jsonl` / `.json`, log in `r02_dropin_multigpu.log`.\n\n"
            "| N | main() | `TIME:` s per iteration | M reads/s from TIME (wall clock, launch + sync of all shards) | M reads/s, kernels (CUDA events, slowest GPU) | per-GPU kernel ms | "
            "index H2D + re-block s (GB/s) | peer copies s (GB/s over NVLink) | table build s per GPU | reads H2D + pack s | results D2H s | `.res.gpu` md5 |\n|---|---|---:|---:|---:|---|---|---|---|---:|---:|---|\n")
    for r in runs:
        pc = ", ".join(f"{a:.4f}" for a in r["peer_copy_s"]) or "-"
        pg = ", ".join(f"{a:.0f}" for a in r["peer_copy_gbs"]) or "-"
        f.write(f"| {r['n_gpus']} | {'reference searchQueries.c' if 'reference' in r['main'] else 'fmIndexSearchGPU_b200'} | {r['TIME_s_per_iteration']:.6f} | {r['mqueries_per_s_from_TIME']:.0f} | "
                f"{r['mqueries_per_s_kernels_max_over_gpus']:.0f} | {', '.join(f'{a:.3f}' for a in r['search_ms_per_gpu'])} | {r['index_h2d_reblock_s']:.3f} ({r['index_h2d_reblock_gbs']:.1f}) | "
                f"{pc} ({pg}) | {', '.join(f'{a:.2f}' for a in r['table_build_s'])} | {r['queries_h2d_pack_s']:.3f} | {r['results_d2h_s']:.4f} | `{r['res_gpu_md5'][:12]}` |\n")
    f.write("\n")
    for p in par:
        f.write(f"**Parity**: {p['what']}: {p['reads_checked']} reads over {p['shards']} shards, identical = **{p['identical']}**.  "
                f"All {len(runs)} runs wrote the same `.res.gpu` ({len(summary['distinct_res_gpu_md5'])} distinct md5), the reference's own main() included.\n\n")
    f.write("Notes.  The index H2D runs at ~11 GB/s because `loadIndex` keeps the reference's `malloc`ed (pageable) buffer; it is one 3 GB copy per run.  In these two command-line runs the peer-copy "
            "column still includes the first use of each GPU by the process (CUDA context creation, 0.3-0.4 s); inside a warm process the same copies take 9.4-10.6 ms "
            "(`r02_config5_n8.md`), and `transferCPUtoGPU` now creates the contexts first and reports them separately (`context_init_s`).  `TIME:` is what the reference prints: wall clock of `searchIndexGPU` (all shards launched, then "
            "waited for), so it contains 8 launches and 8 stream synchronisations (~0.1 ms); the CUDA-event column is the kernels alone.  GPUs of one box differ by up "
            "to 4 % at the same 1965 MHz (1.90 vs 1.98 ms for 12.5 M reads): the slowest one sets the pace -- this is also the 'N=1 -> N>=2 step' of round 1's "
            "scaling table (`bench.py` now prints `per_rank.ms_per_step`).\n")
c5 = [json.loads(x) for x in open(os.path.join(PR, "r02_config5_n8.jsonl")) if x.strip()]
pts = {}
for r in c5:
    if r["what"] == "config 5":
        pts.setdefault((r["k"], r["len"]), {})[r["kernel"]] = r
with open(os.path.join(PR, "r02_config5_n8.md"), "w") as f:
    f.write("# r02: BASELINE config 5 on 8 x B200 (and config 4, strong scaling) -- `profiles/scripts/config5_multigpu.py`\n\n"
            "Task vs Coop vs sparse-step over read lengths 12/25/50/100/250 and k in {1,2} on the 2 Gbp index; **100 M reads** sharded over 8 GPUs by the single-process host "
            "driver (`transferCPUtoGPU`: index replicated by `cudaMemcpyPeer`, sparse-step table with 14 bases per step built on every replica, `$FMGPU_MODE=sparse`), "
            "kernel family switched with `fmgpu_set_variant`; per point 4 x `fmgpu_search_index`, the best of the last 3.  Device ms = CUDA events on every shard's stream, "
            "max over the GPUs; wall ms = the whole call (8 launches + 8 syncs).  Every point verified: all reads found, Task == Coop == Sparse over the whole batch.\n\n"
            "| k | read length | Task Mq/s | Coop Mq/s | Sparse Mq/s | Sparse G LF/s | Sparse / Coop | sparse device ms | sparse wall ms | verified |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
    for (k, l), d in sorted(pts.items(), key=lambda x: (-x[0][0], x[0][1])):
        s = d["sparse"]
        ok = all(v["all_found"] and v["same_as_first_kernel"] for v in d.values())
        f.write(f"| {k} | {l} | {d['task']['mq_per_s']:.0f} | {d['coop']['mq_per_s']:.0f} | {s['mq_per_s']:.0f} | {s['g_ref_lf_steps_per_s']:.0f} | {s['mq_per_s'] / d['coop']['mq_per_s']:.1f}x | "
                f"{s['ms_device_max_over_gpus']:.3f} | {s['ms_wall_clock']:.3f} | {'yes' if ok else 'NO'} |\n")
    f.write("\nOdd lengths at k=2 (25 bp) are undefined in the reference (SURVEY App. C-5); here the last base comes from the derived 1-step rank (tail table / lead table).  "
            "12- and 25-bp reads are answered from the lead tables plus at most one sparse step, which is why they run at 300-600 G reads/s.\n\n"
            "## Config 4, strong scaling: the same 100 M x 100 bp reads on 2 / 4 / 8 GPUs (sparse-step kernel)\n\n"
            "| N | M reads/s (device, slowest GPU) | device ms | wall ms | per-GPU ms | speed-up vs N=2 | all reads found |\n|---|---:|---:|---:|---|---:|---|\n")
    strong = [r for r in c5 if r["what"].startswith("config 4")]
    base = next((r for r in strong if r["n_gpus"] == 2), None)
    for r in strong:
        f.write(f"| {r['n_gpus']} | {r['mq_per_s']:.0f} | {r['ms_device_max_over_gpus']:.3f} | {r['ms_wall_clock']:.3f} | {', '.join(f'{a:.3f}' for a in r['per_gpu_ms'])} | "
                f"{r['mq_per_s'] / base['mq_per_s'] * 2 if base else 0:.2f} (ideal {r['n_gpus']}) | {r['all_found']} |\n")
    for r in c5:
        if r["what"] == "replicas":
            f.write(f"\nReplicas, k={r['k']}: index H2D + re-block {r['index_h2d_reblock_s']:.3f} s; 7 peer copies of the block table in "
                    f"{min(r['peer_copy_s']) * 1e3:.1f}-{max(r['peer_copy_s']) * 1e3:.1f} ms each = {min(r['peer_copy_gbs']):.0f}-{max(r['peer_copy_gbs']):.0f} GB/s over NVLink; "
                    f"sparse-step table {min(r['table_build_s']):.2f}-{max(r['table_build_s']):.2f} s per GPU.\n")

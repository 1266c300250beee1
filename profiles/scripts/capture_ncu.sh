#!/bin/bash
# Per-round ncu evidence for bench.py's roofline keys (run on one B200 under gpurun AFTER bench.py has exited 0 without ncu):
#   1. launch list of the bench command itself (gpu__time_duration per launch; cold-cache, serialised: compare SHARES)
#   2. one `--set full` capture (source imported) of the timed kernel on the benchmark workload
# then `python profiles/scripts/summarize_ncu.py rNN prof_wide` here turns gpurun_out/ into profiles/rNN_* and
# `python profiles/scripts/update_ncu_summary.py rNN wide` refreshes profiles/ncu_summary.json (roofline.traffic).
# usage: capture_ncu.sh            (writes gpurun_out/launches.csv, gpurun_out/prof_wide.ncu-rep)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || { echo "bench.py failed without ncu"; tail -5 gpurun_out/bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
echo "launch list rc=$? ($(wc -l < gpurun_out/launches.csv) lines)"
# the timed kernel since the wide-step table: fm_search_wide_kernel<2, 3, 1, ...> (64-byte blocks of 96-bit entries, 46 bases per step, one read
# per lane pair) on the benchmark workload; FM_W=30 profiles the 64-bit-entry form
FM_VARIANTS=2:0:0:1 FM_REPS=3 ncu --set full --import-source on --clock-control none -k regex:fm_search_wide_kernel -f -o gpurun_out/prof_wide \
    python profiles/scripts/wide_sweep.py > gpurun_out/prof_wide.log 2>&1
echo "set full rc=$?"; ls -la gpurun_out/prof_wide.ncu-rep
# (the sparse-step kernel, timed kernel until then: -k regex:fm_search_sparse_kernel -o gpurun_out/prof_sparse python profiles/scripts/ceiling_counters.py)

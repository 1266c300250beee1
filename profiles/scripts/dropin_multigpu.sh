#!/bin/bash
# r02: the single-process host driver on PHYSICAL GPUs -- BASELINE config 4 as
# written: the 2 Gbp index FILE written by gfmi_b200, 100 M x 100 bp reads in a FASTA file, one process, index
# replicated by cudaMemcpyPeer, reads sharded over N GPUs, through
#   (a) bin/fmIndexSearchGPU_b200            our rebuild of the reference main()
#   (b) oracle/_ref/fmIndexSearchGPU_refmain the reference's OWN common/searchQueries.c, -DCUDA, linked against the library
# Parity: md5 of the two .res.gpu files must agree, and the reference CPU searcher (unmodified binary) is run on the first
# SAMPLE reads of EVERY shard and compared line by line with the same rows of .res.gpu.
# usage: dropin_multigpu.sh "<N list, e.g. 8 4 2>" [reads, default 100000000] [text bp, default 2000000000]
# Writes gpurun_out/r02_dropin_multigpu.jsonl (one line per run) and gpurun_out/r02_dropin_multigpu.log.
set -u
cd "$(dirname "$0")/../.."
ROOT=$PWD
NLIST=${1:-"8 4 2"}; NQ=${2:-100000000}; NTEXT=${3:-2000000000}; LEN=100; SAMPLE=${SAMPLE:-1000000}
W=${WORK:-/dev/shm/fm_dropin}; mkdir -p $W gpurun_out
LOG=gpurun_out/r02_dropin_multigpu.log; OUT=gpurun_out/r02_dropin_multigpu.jsonl
BIN=k-step_fm-index_b200/bin
echo "== $(date -u) N list: $NLIST, $NQ reads, $NTEXT bp, host cores $(nproc)" >> $LOG
t0=$(date +%s.%N)
$BIN/gfmi_b200 --synth $W/ref.fa $NTEXT 1 2 64 >> $LOG 2>&1 || exit 1
IDX=$W/ref.fa.$NTEXT.64fmi2steps.fmi
t1=$(date +%s.%N)
PARTS=16; PER=$(( (NQ + PARTS - 1) / PARTS ))
for p in $(seq 0 $((PARTS - 1))); do
  first=$((p * PER)); num=$PER; [ $((first + num)) -gt $NQ ] && num=$((NQ - first))
  [ $num -gt 0 ] && $BIN/fmsynth reads $W/part.$(printf %02d $p) $NTEXT 1 $num $LEN 2 $first &
done
wait
cat $W/part.?? > $W/reads.fa && rm -f $W/part.??
t2=$(date +%s.%N)
echo "index file $(stat -c %s $IDX) B in $(awk "BEGIN{print $t1 - $t0}") s; reads file $(stat -c %s $W/reads.fa) B in $(awk "BEGIN{print $t2 - $t1}") s" >> $LOG
first_run=1
for N in $NLIST; do
  DEVS=$(seq -s, 0 $((N - 1)))
  for MAIN in ours refmain; do
    [ $MAIN = refmain ] && [ $first_run != 1 ] && continue       # the reference's own main(): once, on the largest N
    EXE=$BIN/fmIndexSearchGPU_b200; [ $MAIN = refmain ] && EXE=oracle/_ref/fmIndexSearchGPU_refmain
    rm -f $W/stats.json $IDX.res.gpu
    ts=$(date +%s.%N)
    FMGPU_DEVICES=$DEVS FMGPU_STATS_FILE=$W/stats.json $EXE $IDX $W/reads.fa $LEN $NQ > $W/run.out 2>> $LOG || { echo "run failed: $MAIN N=$N" >> $LOG; tail -3 $W/run.out >> $LOG; continue; }
    te=$(date +%s.%N)
    TIME=$(grep "TIME:" $W/run.out | awk '{print $2}')
    MD5=$(md5sum $IDX.res.gpu | cut -d" " -f1)
    python - "$N" "$MAIN" "$TIME" "$MD5" "$(awk "BEGIN{print $te - $ts}")" "$W/stats.json" "$NQ" >> $OUT <<'PY'
import json, sys
n, main, t, md5, wall, stats, nq = sys.argv[1:8]
s = json.loads(open(stats).read().strip().split("\n")[-1])
print(json.dumps({"what": "config 4, single process", "n_gpus": int(n), "main": "reference common/searchQueries.c -DCUDA linked against libfmindex_b200.so" if main == "refmain" else "bin/fmIndexSearchGPU_b200",
                  "reads": int(nq), "TIME_s_per_iteration": float(t), "mqueries_per_s_from_TIME": int(nq) / float(t) / 1e6, "res_gpu_md5": md5,
                  "process_wall_s": float(wall), "mqueries_per_s_kernels_max_over_gpus": int(nq) / max(s["search_ms_per_gpu"]) / 1e3, **s}))
PY
    tail -1 $OUT | cut -c1-600 >> $LOG
    if [ $first_run = 1 ] && [ $MAIN = ours ]; then cp $IDX.res.gpu $W/res.first; fi
  done
  first_run=0
done
# parity: the reference CPU searcher on the first SAMPLE reads of every shard of the LARGEST N, against the same rows of .res.gpu
N=$(echo $NLIST | awk '{print $1}')
python - "$N" "$NQ" "$SAMPLE" "$W" "$IDX" "$LEN" >> $OUT <<'PY'
import json, os, subprocess, sys
n, nq, sample, w, idx, length = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], int(sys.argv[6])
per = ((nq + n - 1) // n + 31) & ~31
ok, checked = True, 0
res = open(os.path.join(w, "res.first"), "rb")
header = res.readline()
# offsets of result rows: index them once by scanning (text rows have variable width)
import numpy as np
data = np.fromfile(os.path.join(w, "res.first"), dtype=np.uint8)
nl = np.flatnonzero(data == 10)
for g in range(n):
    first = min(per * g, nq); cnt = min(sample, nq - first)
    if cnt <= 0: continue
    qf = os.path.join(w, f"sample{g}.fa")
    subprocess.run(f"sed -n '{2 * first + 1},{2 * (first + cnt)}p' {w}/reads.fa > {qf}", shell=True, check=True)
    link = os.path.join(w, f"idx{g}.fmi")
    if os.path.lexists(link): os.remove(link)
    os.symlink(idx, link)
    subprocess.run(["oracle/_ref/fmIndexSearchCPU_64bases_2step", link, qf, str(length), str(cnt)], check=True, stdout=subprocess.DEVNULL)
    cpu = open(link + ".res.cpu", "rb").read().split(b"\n", 1)[1]
    a, b = nl[first] + 1, nl[first + cnt] + 1                    # rows first .. first+cnt-1 of the GPU file (row r ends at newline r+1)
    same = data[a:b].tobytes() == cpu
    ok &= same; checked += cnt
print(json.dumps({"what": "parity: reference CPU searcher (oracle/_ref/fmIndexSearchCPU_64bases_2step) on the first reads of every shard vs the same rows of .res.gpu",
                  "n_gpus": n, "reads_checked": checked, "shards": n, "identical": bool(ok)}))
PY
tail -1 $OUT >> $LOG
python - $OUT <<'PY'
import json, sys
rows = [json.loads(x) for x in open(sys.argv[1]) if x.strip()]
md5s = {r["res_gpu_md5"] for r in rows if "res_gpu_md5" in r}
print("distinct .res.gpu md5 over all runs:", len(md5s), md5s)
PY
rm -rf $W

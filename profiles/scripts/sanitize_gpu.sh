#!/bin/bash
# compute-sanitizer over the CUDA side: memcheck, racecheck, synccheck and initcheck on
# __graft_entry__.smoke() (Task, Coop, fused-step and sparse-step kernels, std and AltCounters files, table builds), then
# memcheck and racecheck on a selection of the GPU parity tests (goldens incl. the quirk fixtures, the tiny-reference fuzz,
# all read lengths -- the TMA-staged kernels with their mbarrier waits and ballot-guarded rare paths).
# usage: bash profiles/scripts/sanitize_gpu.sh [out file]      (needs a B200; `make -C k-step_fm-index_b200/csrc sanitize-gpu`)
cd "$(dirname "$0")/../.."
OUT=${1:-gpurun_out/r02_sanitizers.txt}
mkdir -p "$(dirname "$OUT")"
: > "$OUT"
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool -- python -c 'import __graft_entry__ as g; g.smoke()'" >> "$OUT"
  timeout 900 "$SAN" --tool $tool --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -25 >> "$OUT"
  echo "exit code ${PIPESTATUS[0]}" >> "$OUT"
done
SEL='(test_golden or read_lengths or (test_sparse_steps_golden_all_widths and quirk_k2)) and not full_size'
FUZZ=$(for c in 0 1 2 3 6 7; do echo "tests/test_gpu_parity.py::test_fuzz_tiny_references[$c]"; done)
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool -- pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py -k \"$SEL\"" >> "$OUT"
  timeout 1200 "$SAN" --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse.py -x -q -m gpu -k "$SEL" 2>&1 | grep -v "^$" | tail -25 >> "$OUT"
  echo "exit code ${PIPESTATUS[0]}" >> "$OUT"
  echo "=== compute-sanitizer --tool $tool -- pytest test_fuzz_tiny_references[0,1,2,3,6,7]" >> "$OUT"
  timeout 900 "$SAN" --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $FUZZ -x -q -m gpu 2>&1 | grep -v "^$" | tail -25 >> "$OUT"
  echo "exit code ${PIPESTATUS[0]}" >> "$OUT"
done
grep -E "^===|ERROR SUMMARY|exit code|passed|failed" "$OUT"

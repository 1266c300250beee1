"""r02: which unit saturates at the ~46 G/s random-fetch ceiling?  Workload for ONE ncu capture.

Runs, on the benchmark's 2 Gbp index: the gather probe (fmgpu_gather_probe_ex, 16-byte and 64-byte accesses over the
sparse table's footprint) and one sparse-step search of 10 M reads of 100 bp -- the two kernels whose L2 / DRAM / fabric
counters profiles/r02_ceiling_counters.md compares.  Run it twice: once plain (prints the timings), once under

  ncu --set full --metrics $(cat profiles/scripts/ceiling_metrics.txt | tr '\n' ',') --clock-control none \
      -k regex:'fm_gather_probe_kernel|fm_search_sparse_kernel' -o gpurun_out/r02_ceiling python profiles/scripts/ceiling_counters.py
"""
import ctypes as C, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
L = pkg.lib()
n = int(float(os.environ.get("FM_N", "2e9"))); nq = int(float(os.environ.get("FM_NQ", "1e7"))); length = 100
t0 = time.time()
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free()
idx.sparsify(0, 0, 0); idx.prepare(length)
m = idx.meta
stream = torch.cuda.current_stream().cuda_stream
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), stream), "reads")
wpq = L.fmgpu_words_per_query(length)
d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
v = pkg.variant(pkg.MODE_SPARSE, int(os.environ.get("FM_QPT", "0")))      # 0 = the launcher default (3 reads per lane group, static assignment on this text)
out = {"setup_s": time.time() - t0, "sparse_bases": m.sparse_bases, "sparse_gb": m.sparse_bytes / 1e9}
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), C.byref(v), stream), "search"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
out["sparse_ms"] = min(ts)
a, s, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
pkg.check(L.fmgpu_count_fetches_sparse_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream, C.byref(a), C.byref(s), C.byref(o)), "count")
out["sparse_blocks"] = a.value; out["sb96_blocks"] = s.value; out["fetches_per_s"] = (a.value + s.value) / (min(ts) * 1e-3)
foot = int(m.sparse_bytes)
for ab in (16, 64, 128):
    out[f"probe_{ab}B_per_s"] = pkg.gather_probe(0, foot, 256, 1, ab)
print(json.dumps(out), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "r02_ceiling_timings.jsonl"), "a").write(json.dumps(out) + "\n")

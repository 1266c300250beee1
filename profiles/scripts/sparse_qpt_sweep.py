"""r02: queries per lane group and static / dynamic read assignment of the sparse-step kernel on the benchmark's uniform
2 Gbp text (10 M x 100 bp).  Appends to gpurun_out/r02_sparse_qpt_sweep.jsonl."""
import ctypes as C, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
L = pkg.lib()
n, nq, length = int(float(os.environ.get("FM_N", "2e9"))), int(float(os.environ.get("FM_NQ", "1e7"))), int(os.environ.get("FM_LEN", "100"))
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free()
idx.sparsify(0, 0, 0); idx.prepare(length)
stream = torch.cuda.current_stream().cuda_stream
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), stream), "reads")
wpq = L.fmgpu_words_per_query(length)
d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
out = open(os.path.join(ROOT, "gpurun_out", "r02_sparse_qpt_sweep.jsonl"), "a")
want = None
for dyn, rounds in ((0, 0), (1, 4), (1, 8), (1, 16)):
    for qpt in (1, 2, 3, 4):
        os.environ["FMGPU_SPARSE_DYNAMIC"] = str(dyn)
        if rounds: os.environ["FMGPU_SPARSE_ROUNDS"] = str(rounds)
        v = pkg.variant(pkg.MODE_SPARSE, qpt)
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), C.byref(v), stream), "search"); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        if want is None: want = d_res.clone()
        rec = {"dynamic": dyn, "rounds": rounds, "qpt": qpt, "ms_best": min(ts[2:]), "ms_mean": sum(ts[2:]) / len(ts[2:]), "mq_per_s": nq / min(ts[2:]) / 1e3, "equal": bool(torch.equal(d_res, want))}
        print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n")

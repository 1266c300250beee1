"""A/B of sparse-step kernel variants at full size (2 Gbp, 10 M x 100 bp): ms per launch for each queries-per-lane-group
setting in $FM_QPTS, all results compared with the first one.  Appends to gpurun_out/sparse_ab.jsonl."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
OUT = open(os.path.join(ROOT, "gpurun_out", "sparse_ab.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
n, nq = int(float(os.environ.get("FM_N", "2e9"))), int(float(os.environ.get("FM_NQ", "1e7")))
lengths = [int(x) for x in os.environ.get("FM_LENS", "100").split(",")]
qpts = [int(x) for x in os.environ.get("FM_QPTS", "4,3,5,6,8").split(",")]
L = pkg.lib()
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free()
idx.sparsify(int(os.environ.get("FM_KS", "0")), 0, 0); print(json.dumps({"ks": idx.meta.sparse_bases, "uniform_nb": idx.meta.sparse_uniform_nb, "overflow": int(idx.meta.sparse_overflow), "blocks": int(idx.meta.sparse_blocks), "gb": idx.meta.sparse_bytes / 1e9}), flush=True)
stream = torch.cuda.current_stream().cuda_stream
d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
for length in lengths:
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
    d_packed = torch.empty(nq * L.fmgpu_words_per_query(length), dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
    del d_ascii
    want = None
    for rnd in range(2):
        for qpt in qpts:
            v = pkg.variant(pkg.MODE_SPARSE, qpt)
            ts = []
            for _ in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search"); e1.record()
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            if want is None:
                want = d_res.clone()
            emit(tag=os.environ.get("FM_AB_TAG", ""), length=length, qpt=qpt, round=rnd, ms_min=min(ts[2:]), ms_mean=sum(ts[2:]) / len(ts[2:]),
                 mq_per_s=nq / min(ts[2:]) / 1e3, equals_first=bool(torch.equal(d_res, want)))

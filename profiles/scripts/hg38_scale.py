"""Beyond the reference's 2^31 limit: a 3.1 Gbp synthetic text (hg38-sized, bwtsize > 2^31) built on the GPU, searched with the
plain Coop and the sparse-step kernels, located; every check is a property (reads found, kernels agree, positions == read
starts).  Appends to gpurun_out/hg38_scale.jsonl."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
import helpers
OUT = open(os.path.join(ROOT, "gpurun_out", "r02_hg38_scale.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
n, nq, length = int(float(os.environ.get("FM_N", "3.1e9"))), int(float(os.environ.get("FM_NQ", "1e7"))), int(os.environ.get("FM_LEN", "100"))
L = pkg.lib()
t0 = time.time(); b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free(); torch.cuda.synchronize()
emit(what="build + reblock", n=n, seconds=time.time() - t0, sb96_gb=idx.meta.nbytes / 1e9, bwtsize=int(idx.meta.bwtsize))
t0 = time.time(); idx.sparsify(0, 0, 0); torch.cuda.synchronize(); m = idx.meta
idx.prepare(length)
emit(what="sparsify", seconds=time.time() - t0, sparse_gb=m.sparse_bytes / 1e9, bases=m.sparse_bases, grid_blocks_per_symbol=m.sparse_uniform_nb, overfull_buckets=int(m.sparse_overflow), tree_nodes=int(m.sparse_tree_nodes), blocks=int(m.sparse_blocks))
stream = torch.cuda.current_stream().cuda_stream
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
d_packed = torch.empty(nq * L.fmgpu_words_per_query(length), dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
res = {}
for name, v in (("coop", pkg.variant(pkg.MODE_COOP, 1, 256)), ("sparse", pkg.variant(pkg.MODE_SPARSE, 0))):
    d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[name] = d_res
    r = d_res.cpu().numpy().view(np.uint32)
    emit(what="search", kernel=name, ms=min(ts[1:]), mq_per_s=nq / min(ts[1:]) / 1e3, every_read_found=bool((r[1::2] > r[0::2]).all()),
         rows_above_2_31=int((r[0::2] >= 2 ** 31).sum()))
emit(what="kernels agree", ok=bool(torch.equal(res["coop"], res["sparse"])))
t0 = time.time(); idx.build_sa(); torch.cuda.synchronize()
emit(what="suffix array", seconds=time.time() - t0, gb=idx.meta.sa_bytes / 1e9)
d_pos = torch.empty(nq, dtype=torch.int32, device="cuda"); d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_locate_device(idx.handle, res["sparse"].data_ptr(), nq, 1, d_pos.data_ptr(), d_cnt.data_ptr(), stream), "locate"); torch.cuda.synchronize()
starts = helpers.synth_read_starts(2, nq, n, length)
cnt = d_cnt.cpu().numpy().view(np.uint32); pos = d_pos.cpu().numpy().view(np.uint32)
once = cnt == 1
emit(what="locate", found_once=float(once.mean()), positions_equal_starts=bool(np.array_equal(pos[once].astype(np.uint64), starts[once].astype(np.uint64))),
     positions_above_2_31=int((pos[once] >= 2 ** 31).sum()))
t0 = time.time(); idx.build_sa_sampled(32); torch.cuda.synchronize()
d_pos2 = torch.empty(nq, dtype=torch.int32, device="cuda")
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pkg.check(L.fmgpu_locate_device(idx.handle, res["sparse"].data_ptr(), nq, 1, d_pos2.data_ptr(), d_cnt.data_ptr(), stream), "locate"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
emit(what="sampled suffix array (rate 32)", build_seconds=time.time() - t0, gb=idx.meta.sa_bytes / 1e9, locate_ms=min(ts), mq_per_s=nq / min(ts) / 1e3,
     positions_equal_full_array=bool(torch.equal(d_pos, d_pos2)))

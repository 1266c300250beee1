"""Beyond the reference's 2^31 limit with the wide-step table: a 3.1 Gbp synthetic text (hg38-sized, bwtsize > 2^31, 32-bit row
numbers in the entries), widen_for(100) -- the 96-bit table does not fit next to its own scratch at this size, the 64-bit one
(30 bases per step) does -- searched against the plain Coop kernel.  Appends to gpurun_out/r02w_hg38_wide.jsonl."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
OUT = open(os.path.join(ROOT, "gpurun_out", "r02w_hg38_wide.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
n, nq, length = int(float(os.environ.get("FM_N", "3.1e9"))), int(float(os.environ.get("FM_NQ", "1e7"))), int(os.environ.get("FM_LEN", "100"))
L = pkg.lib()
t0 = time.time(); b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free(); torch.cuda.synchronize()
emit(what="build + reblock", n=n, seconds=time.time() - t0, sb96_gb=idx.meta.nbytes / 1e9, bwtsize=int(idx.meta.bwtsize))
emit(what="widths", best=idx.wide_bases_for(length), with_64_bit_entries=idx.wide_bases_for(length, 2))
t0 = time.time()
if os.environ.get("FM_W"): idx.widen(int(os.environ["FM_W"]))        # a given width instead of the library's choice
else: idx.widen_for(length)
idx.prepare(length); torch.cuda.synchronize(); m = idx.meta
emit(what="widen_for", seconds=time.time() - t0, bases=m.wide_bases, entry_words=m.wide_entry_words, block_entries=m.wide_block_entries, prefix_bits=m.wide_prefix_bits,
     row_bits=m.wide_row_bits, wide_gb=m.wide_bytes / 1e9, overfull_buckets=int(m.wide_overflow), tree_rows_fraction=m.wide_tree_rows / m.bwtsize, exceptional=int(m.wide_exceptional))
stream = torch.cuda.current_stream().cuda_stream
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
d_packed = torch.empty(nq * L.fmgpu_words_per_query(length), dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
res = {}
for name, v in (("coop", pkg.variant(pkg.MODE_COOP, 1, 256)), ("wide", pkg.variant(pkg.MODE_WIDE, 0))):
    d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[name] = d_res
    r = d_res.cpu().numpy().view(np.uint32)
    emit(what="search", kernel=name, ms=min(ts[1:]), mq_per_s=nq / min(ts[1:]) / 1e3, every_read_found=bool((r[1::2] > r[0::2]).all()),
         rows_above_2_31=int((r[0::2] >= 2 ** 31).sum()))
emit(what="kernels agree", ok=bool(torch.equal(res["coop"], res["wide"])))

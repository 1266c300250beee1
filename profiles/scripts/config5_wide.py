"""BASELINE config 5 on one B200 with the wide-step table: read lengths 12/25/50/75/100/101/150/250, k in {1,2}, 2 Gbp index,
10 M reads per point; the wide-step table is rebuilt with the width that serves each length in the fewest fetches
(fmgpu_wide_bases_for: 64-bit entries up to 30 bases per step, 96-bit entries up to 46), the sparse-step and plain Coop
kernels run on the same reads first (phase 1; their table is released before the wide tables are built).  Every point is
checked: all reads found, wide == sparse == Coop.  Writes gpurun_out/r02w_config5_wide.jsonl."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
L = pkg.lib()
OUT = open(os.path.join(ROOT, "gpurun_out", "r02w_config5_wide.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
n, nq = int(float(os.environ.get("FM_N", "2e9"))), int(float(os.environ.get("FM_NQ", "1e7")))
LENGTHS = (12, 25, 50, 75, 100, 101, 150, 250)
stream = torch.cuda.current_stream().cuda_stream
def packed_reads(length):
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
    wpq = L.fmgpu_words_per_query(length)
    d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
    return d_packed
def timed(idx, d_packed, d_res, length, v):
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[2:])
for k in (2, 1):
    b = pkg.IndexBuild.from_synth(n, 1, k, 64); idx = b.to_index(); b.free()
    d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    ref = {}
    idx.sparsify()
    for length in LENGTHS:                                       # phase 1: plain Coop and the sparse-step table
        d_packed = packed_reads(length); idx.prepare(length)
        for name, v in (("coop", pkg.variant(pkg.MODE_COOP, 1, 256)), ("sparse", pkg.variant(pkg.MODE_SPARSE, 0))):
            ms = timed(idx, d_packed, d_res, length, v)
            res = d_res.cpu().numpy().view(np.uint32).copy()
            if name == "coop": ref[length] = res
            emit(what="search", k=k, len=length, kernel=name, ms=ms, mq_per_s=nq / ms / 1e3, g_ref_lf_steps_per_s=nq * (length // k) / ms / 1e6,
                 all_found=bool(((res[1::2] - res[0::2]) >= 1).all()), same_as_coop=bool(np.array_equal(res, ref[length])))
            d_res.zero_()
        del d_packed
    idx.unsparsify()
    cur = 0
    for length in LENGTHS:                                       # phase 2: the wide-step table, rebuilt when the length wants another width
        wb = idx.wide_bases_for(length)
        if not wb:
            emit(what="search", k=k, len=length, kernel="wide", unavailable="no step width serves this length (the sparse-step table does)")
            continue
        build_s = None
        if wb != cur:
            if cur: idx.unwiden()
            t0 = time.time(); idx.widen(wb); build_s = time.time() - t0; cur = wb
        idx.prepare(length)
        d_packed = packed_reads(length)
        ms = timed(idx, d_packed, d_res, length, pkg.variant(pkg.MODE_WIDE, 0))
        res = d_res.cpu().numpy().view(np.uint32)
        m = idx.meta
        steps = length // m.wide_bases
        emit(what="search", k=k, len=length, kernel="wide", ms=ms, mq_per_s=nq / ms / 1e3, g_ref_lf_steps_per_s=nq * (length // k) / ms / 1e6,
             all_found=bool(((res[1::2] - res[0::2]) >= 1).all()), same_as_coop=bool(np.array_equal(res, ref[length])),
             wide_bases=m.wide_bases, entry_bits=32 * m.wide_entry_words, block_entries=m.wide_block_entries, lead_bases=length - steps * m.wide_bases, steps=steps,
             table_gb=m.wide_bytes / 1e9, build_s=build_s)
        d_res.zero_(); del d_packed
    idx.free()

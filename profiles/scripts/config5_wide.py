"""BASELINE config 5 on one B200 with the wide-step table: read lengths 12/25/50/75/100/101/150/250, k in {1,2}, 2 Gbp index,
10 M reads per point; the wide-step table is rebuilt with the width that serves each length in the fewest fetches
(fmgpu_wide_bases_for), the sparse-step and plain Coop kernels run beside it.  Every point is checked: all reads found,
wide == sparse == Coop.  Writes gpurun_out/r02w_config5_wide.jsonl."""
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
stream = torch.cuda.current_stream().cuda_stream
for k in (2, 1):
    b = pkg.IndexBuild.from_synth(n, 1, k, 64); idx = b.to_index(); b.free()
    idx.sparsify()
    cur = 0
    for length in (12, 25, 50, 75, 100, 101, 150, 250):
        d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
        pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
        wpq = L.fmgpu_words_per_query(length)
        d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
        pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize(); del d_ascii
        wb = idx.wide_bases_for(length)
        build_s = None
        if wb and wb != cur:
            if cur: idx.unwiden()
            t0 = time.time(); idx.widen(wb); build_s = time.time() - t0; cur = wb
        idx.prepare(length)
        ref = None
        for name, v in (("coop", pkg.variant(pkg.MODE_COOP, 1, 256)), ("sparse", pkg.variant(pkg.MODE_SPARSE, 0)), ("wide", pkg.variant(pkg.MODE_WIDE, 0))):
            if name == "wide" and not (wb and idx.wide_serves(length)):
                emit(what="search", k=k, len=length, kernel="wide", unavailable="no step width serves this length (the sparse-step table does)")
                continue
            ts = []
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search"); e1.record()
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            res = d_res.cpu().numpy().view(np.uint32)
            if ref is None: ref = res.copy()
            ms = min(ts[2:])
            m = idx.meta
            extra = {"wide_bases": m.wide_bases, "lead_bases": length - (length // m.wide_bases) * m.wide_bases if length >= m.wide_bases else length,
                     "steps": length // m.wide_bases, "table_gb": m.wide_bytes / 1e9, "build_s": build_s} if name == "wide" else {}
            emit(what="search", k=k, len=length, kernel=name, ms=ms, mq_per_s=nq / ms / 1e3, g_ref_lf_steps_per_s=nq * (length // k) / ms / 1e6,
                 all_found=bool(((res[1::2] - res[0::2]) >= 1).all()), same_as_coop=bool(np.array_equal(res, ref)), **extra)
            d_res.zero_()
        del d_packed, d_res
    idx.free()

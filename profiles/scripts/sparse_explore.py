"""Sparse-step kernel at full size: build time, footprint, overflow blocks, search time per (bases, lambda, qpt), fetch
counts; every result is compared with the plain Coop kernel's (whole batch) and with the reference md5 of the first
1 M reads (tests/golden/config3_2g.json).  Also random (non-matching) reads and other lengths.
Writes gpurun_out/sparse_explore.jsonl."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
import helpers
OUT = open(os.path.join(ROOT, "gpurun_out", "sparse_explore.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "config3_2g.json")))
n, nq, k = int(float(os.environ.get("FM_N", "2e9"))), int(float(os.environ.get("FM_NQ", "1e7"))), int(os.environ.get("FM_K", "2"))
CONFIGS = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("FM_SPARSE", "10:6:2,10:5:2,10:7:2,10:12:4").split(",")]
L = pkg.lib()
b = pkg.IndexBuild.from_synth(n, 1, k, 64); idx = b.to_index(); b.free()
stream = torch.cuda.current_stream().cuda_stream

def reads(length, exact=True):
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    if exact:
        pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
    else:
        d_ascii.copy_(torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (nq * length,), device="cuda")])
    wpq = L.fmgpu_words_per_query(length)
    d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
    return d_packed

def run(d_packed, length, v, out, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, out.data_ptr(), v, stream), "search"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])

sets = [(100, True), (100, False), (50, True), (25, True), (250, True), (12, True), (36, True)]
packed = {(ln, ex): reads(ln, ex) for ln, ex in sets}
want = {}
d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
for key, dp in packed.items():
    ms = run(dp, key[0], pkg.variant(pkg.MODE_COOP, 1, 256), d_res, reps=2)
    want[key] = d_res.clone()
    emit(what="plain coop", length=key[0], exact=key[1], ms=ms, mq_per_s=nq / ms / 1e3)
md5 = helpers.results_text_md5(want[(100, True)].cpu().numpy().view(np.uint32)[:2_000_000])
emit(what="plain md5 equals reference", ok=(md5 == gold["md5"]["res_cpu_std_text"]) if (n == 2_000_000_000 and k == 2) else None)

for ci, (ks, lam, lanes) in enumerate(CONFIGS):
    t0 = time.time()
    try:
        idx.sparsify(ks, lam, lanes)
    except Exception as ex:
        emit(what="sparsify failed", bases=ks, lam=lam, lanes=lanes, err=str(ex)); continue
    torch.cuda.synchronize(); t1 = time.time()
    m = idx.meta
    emit(what="sparsify", bases=ks, lam=lam, lanes=lanes, seconds=t1 - t0, sparse_gb=m.sparse_bytes / 1e9, blocks=m.sparse_blocks, overflow_blocks=m.sparse_overflow,
         start_bases=m.sparse_start_bases)
    for key, dp in packed.items():
        if ci and key != (100, True) and os.environ.get("FM_SPARSE_ALLSETS", "0") == "0":
            continue
        for qpt in ((1, 2, 3, 4) if key == (100, True) else (4,)):
            ms = run(dp, key[0], pkg.variant(pkg.MODE_SPARSE, qpt), d_res)
            same = bool(torch.equal(d_res, want[key]))
            emit(what="sparse search", bases=ks, lam=lam, lanes=lanes, length=key[0], exact=key[1], qpt=qpt, ms=ms, mq_per_s=nq / ms / 1e3, equals_plain=same)
        a, bb, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        pkg.check(L.fmgpu_count_fetches_sparse_device(idx.handle, dp.data_ptr(), nq, key[0], d_res.data_ptr(), stream, C.byref(a), C.byref(bb), C.byref(c)), "count")
        emit(what="sparse fetches", bases=ks, lam=lam, lanes=lanes, length=key[0], exact=key[1], sparse_blocks_per_read=a.value / nq, sb96_blocks_per_read=bb.value / nq,
             overflows_per_read=c.value / nq, equals_plain=bool(torch.equal(d_res, want[key])))
    idx.unsparsify()

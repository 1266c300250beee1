"""Sparse-step kernel on a repeat-rich text (genome-like): half of the text is made of diverged copies of a few
repeat families (Alu-like 300-bp units, 10 % substitutions per copy), the rest is random.  Reports overfull blocks,
fallbacks per read and speed against the plain and fused kernels; all results must equal the plain kernel's.
Writes gpurun_out/repeat_explore.jsonl."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
L = pkg.lib()
OUT = open(os.path.join(ROOT, "gpurun_out", "repeat_explore.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
n = int(float(os.environ.get("FM_N", "4e8"))); nq = int(float(os.environ.get("FM_NQ", "4e6"))); length = 100
frac = float(os.environ.get("FM_REPEAT_FRACTION", "0.5")); div = float(os.environ.get("FM_DIVERGENCE", "0.10"))
rng = np.random.default_rng(17)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
t0 = time.time()
text = ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
families = [ACGT[rng.integers(0, 4, u)] for u in (300, 300, 310, 6000, 150)]
weights = np.array([0.5, 0.15, 0.1, 0.2, 0.05])
nrep = int(n * frac)
placed = 0
while placed < nrep:
    f = families[rng.choice(len(families), p=weights)]
    pos = int(rng.integers(0, n - f.size))
    copy = f.copy()
    mut = rng.random(f.size) < div
    copy[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
    text[pos:pos + f.size] = copy
    placed += f.size
if (n + 1) % 64 == 0:
    text = text[:-1]; n -= 1
emit(what="text", n=n, repeat_fraction=frac, divergence=div, seconds=time.time() - t0)
t0 = time.time()
b = pkg.IndexBuild.from_text(text, 2, 64); idx = b.to_index(); b.free()
emit(what="index build", seconds=time.time() - t0)
starts = rng.integers(0, n - length, nq)
reads = text[(starts[:, None] + np.arange(length)[None, :])].reshape(-1)
d_ascii = torch.from_numpy(reads).cuda()
wpq = L.fmgpu_words_per_query(length)
stream = torch.cuda.current_stream().cuda_stream
d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
def run(v, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
ms = run(pkg.variant(pkg.MODE_COOP, 1, 256)); want = d_res.clone()
hits = (want[1::2] - want[0::2]).double()
emit(what="plain coop", ms=ms, mq_per_s=nq / ms / 1e3, mean_hits=float(hits.mean()), reads_with_more_than_one_hit=float((hits > 1).double().mean()))
try:
    idx.fuse(); ms = run(pkg.variant(pkg.MODE_FUSED, 2))
    emit(what="fused", ms=ms, mq_per_s=nq / ms / 1e3, equals_plain=bool(torch.equal(d_res, want))); idx.unfuse()
except pkg.FMError as ex:
    emit(what="fused unavailable", err=str(ex))
for ks, lam, lanes in [tuple(int(x) for x in c.split(":")) for c in os.environ.get("FM_SPARSE", "10:0:0,10:12:4,8:0:0,6:0:0").split(",")]:
    t0 = time.time(); idx.sparsify(ks, lam, lanes); torch.cuda.synchronize()
    m = idx.meta
    emit(what="sparsify", bases=ks, lam=m.sparse_lambda, lanes=m.sparse_lanes, seconds=time.time() - t0, sparse_gb=m.sparse_bytes / 1e9, blocks=m.sparse_blocks,
         overflow_blocks=m.sparse_overflow, overflow_fraction=m.sparse_overflow / m.sparse_blocks)
    ms = run(pkg.variant(pkg.MODE_SPARSE, 4))
    same = bool(torch.equal(d_res, want))
    a, s, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
    pkg.check(L.fmgpu_count_fetches_sparse_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream, C.byref(a), C.byref(s), C.byref(o)), "count")
    emit(what="sparse", bases=ks, lam=m.sparse_lambda, lanes=m.sparse_lanes, ms=ms, mq_per_s=nq / ms / 1e3, equals_plain=same,
         sparse_blocks_per_read=a.value / nq, sb96_blocks_per_read=s.value / nq, fallbacks_per_read=o.value / nq)
    idx.unsparsify()

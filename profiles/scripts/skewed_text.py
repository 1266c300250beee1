"""r02: the sparse-step table on NON-uniform texts.

  FM_TEXT=genome400   the r01 text (profiles/r01_repeat_text.md): 400 Mbp, half random, half 10 %-diverged copies of five
                      repeat families (same generator, same seed)
  FM_TEXT=human2g     2 Gbp with a human-like 14-mer spectrum: GC 41 % background (AT-rich 14-mers 160 x more frequent
                      than GC-rich ones), 3 % microsatellites / poly-A, and 42 % interspersed repeats in families of
                      different age: Alu-like 300 bp (10 % of the text, 8-16 % divergence), L1-like 6 kb (17 %, 4-25 %),
                      old MIR/DNA-like families (15 %, 25-35 %)
  FM_TEXT=gc2g        2 Gbp i.i.d. with GC 41 % only (skew without repeats)

For each text: GPU index build, sparse-step tables (automatic width; FM_SPARSE="ks:lambda:lanes,..." adds others), the
fused-step and plain Coop kernels beside them, block fetches per read (grid / tree), the random-access probe over the
same footprint, equality of every (L,R) with the Coop kernel's and -- FM_REF_SAMPLE reads, strided over the batch --
with the UNMODIFIED reference searcher (oracle/_ref).  Appends to gpurun_out/r02_skewed_text.jsonl."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
L = pkg.lib()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = open(os.path.join(ROOT, "gpurun_out", "r02_skewed_text.jsonl"), "a")
KIND = os.environ.get("FM_TEXT", "genome400")
def emit(**kw):
    kw["text"] = KIND
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
nq = int(float(os.environ.get("FM_NQ", "4e6"))); length = int(os.environ.get("FM_LEN", "100"))
t0 = time.time()
if KIND == "genome400":
    n = int(float(os.environ.get("FM_N", "4e8")))
    rng = np.random.default_rng(17)
    text = ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
    families = [ACGT[rng.integers(0, 4, u)] for u in (300, 300, 310, 6000, 150)]
    weights = np.array([0.5, 0.15, 0.1, 0.2, 0.05])
    nrep, placed = int(n * 0.5), 0
    while placed < nrep:
        f = families[rng.choice(len(families), p=weights)]
        pos = int(rng.integers(0, n - f.size))
        copy = f.copy()
        mut = rng.random(f.size) < 0.10
        copy[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
        text[pos:pos + f.size] = copy
        placed += f.size
else:
    n = int(float(os.environ.get("FM_N", "2e9")))
    rng = np.random.default_rng(23)
    text = np.empty(n, dtype=np.uint8)
    for a in range(0, n, 1 << 27):                               # GC 41 % background, in pieces (memory)
        b = min(n, a + (1 << 27))
        text[a:b] = ACGT[rng.choice(4, size=b - a, p=[0.295, 0.205, 0.205, 0.295]).astype(np.uint8)]
    if KIND == "human2g":
        def scatter(unit_len, total_bases, div_lo, div_hi, nfam):
            for _ in range(nfam):
                unit = ACGT[rng.choice(4, size=unit_len, p=[0.295, 0.205, 0.205, 0.295])]
                copies = max(1, int(total_bases / nfam / unit_len))
                for c0 in range(0, copies, 1 << 16):             # chunks of copies: bounded temporary
                    c = min(1 << 16, copies - c0)
                    div = rng.uniform(div_lo, div_hi, size=(c, 1))
                    m = np.tile(unit, (c, 1))
                    mut = rng.random((c, unit_len)) < div
                    m[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
                    pos = rng.integers(0, n - unit_len, c)
                    text[(pos[:, None] + np.arange(unit_len)[None, :]).reshape(-1)] = m.reshape(-1)
        scatter(300, 0.10 * n, 0.08, 0.16, 6)                    # Alu-like subfamilies
        scatter(6000, 0.17 * n, 0.04, 0.25, 8)                   # L1-like
        scatter(250, 0.15 * n, 0.25, 0.35, 40)                   # old families
        nms = int(0.03 * n / 40)                                 # microsatellites and poly-A, ~40 bp each
        pos = rng.integers(0, n - 64, nms)
        motifs = [b"A", b"CA", b"GT", b"AAT", b"GATA", b"T", b"AC", b"TTTA"]
        for mi, mo in enumerate(motifs):
            sel = pos[mi::len(motifs)]
            unit = np.frombuffer((mo * 64)[:40], dtype=np.uint8)
            text[(sel[:, None] + np.arange(40)[None, :]).reshape(-1)] = np.tile(unit, sel.size)
if (n + 1) % 64 == 0:
    text = text[:-1]; n -= 1
emit(what="text", n=n, seconds=time.time() - t0)
t0 = time.time()
b = pkg.IndexBuild.from_text(text, 2, 64); idx = b.to_index()
emit(what="index build", seconds=time.time() - t0)
rng = np.random.default_rng(5)
starts = rng.integers(0, n - length, nq)
reads = np.empty(nq * length, dtype=np.uint8)
for a in range(0, nq, 1 << 20):
    s = starts[a:a + (1 << 20)]
    reads[a * length:(a + s.size) * length] = text[(s[:, None] + np.arange(length)[None, :])].reshape(-1)
d_ascii = torch.from_numpy(reads).cuda()
wpq = L.fmgpu_words_per_query(length)
stream = torch.cuda.current_stream().cuda_stream
d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
del d_ascii
def run(v, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), C.byref(v), stream), "search"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
ms = run(pkg.variant(pkg.MODE_COOP, 1, 256)); want = d_res.clone()
hits = (want[1::2] - want[0::2]).double()
emit(what="plain coop", ms=ms, mq_per_s=nq / ms / 1e3, mean_hits=float(hits.mean()), reads_with_more_than_one_hit=float((hits > 1).double().mean()),
     reads_with_more_than_100_hits=float((hits > 100).double().mean()))
# the unmodified reference searcher on a strided sample (tag-100 image downloaded from the GPU build)
ns = int(float(os.environ.get("FM_REF_SAMPLE", "1e6")))
if ns:
    from bindings import RefSearcher
    image = b.download()
    ref = RefSearcher(2, 64, False)
    stride = max(1, nq // ns)
    sel = np.arange(0, nq, stride)[:ns]
    sample = reads.reshape(nq, length)[sel].reshape(-1)
    t0 = time.time()
    out, secs = ref.search(ref.wrap_image(image), sample, length, 1, os.cpu_count() or 1)
    got = want.cpu().numpy().view(np.uint32).reshape(nq, 2)[sel].reshape(-1)
    emit(what="reference searcher on a strided sample", reads=int(sel.size), stride=stride, seconds=secs, coop_equals_reference=bool(np.array_equal(out, got)))
    del image
b.free()
if os.environ.get("FM_FUSED", "1") != "0":
    try:
        idx.fuse(); ms = run(pkg.variant(pkg.MODE_FUSED, 2))
        emit(what="fused", ms=ms, mq_per_s=nq / ms / 1e3, fused_gb=idx.meta.fused_bytes / 1e9, equals_plain=bool(torch.equal(d_res, want))); idx.unfuse()
    except pkg.FMError as ex:
        emit(what="fused unavailable", err=str(ex))
# the wide-step table (second half of round 2): the width that serves this read length, and 30 bases per step (64-bit entries)
for wb in ([int(x) for x in os.environ["FM_WIDE"].split(",")] if os.environ.get("FM_WIDE") else [idx.wide_bases_for(length), 30]):
    t0 = time.time()
    try:
        idx.widen(wb); idx.prepare(length); torch.cuda.synchronize()
    except pkg.FMError as ex:
        emit(what="widen failed", bases=wb, err=str(ex)); continue
    m = idx.meta
    emit(what="widen", bases=m.wide_bases, entry_words=m.wide_entry_words, block_entries=m.wide_block_entries, prefix_bits=m.wide_prefix_bits, seconds=time.time() - t0,
         wide_gb=m.wide_bytes / 1e9, overfull_buckets=m.wide_overflow, tree_nodes=m.wide_tree_nodes, tree_rows_fraction=m.wide_tree_rows / m.bwtsize,
         tree_depth=m.wide_tree_depth, exceptional_buckets=m.wide_exceptional)
    a, s_, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
    pkg.check(L.fmgpu_count_fetches_wide_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream, C.byref(a), C.byref(s_), C.byref(o)), "count")
    probe = pkg.gather_probe(0, int(m.wide_bytes), 256, 2)
    for dyn, qpt in ((0, 1), (1, 1), (1, 2)):
        os.environ["FMGPU_WIDE_DYNAMIC"] = str(dyn)
        ms = run(pkg.variant(pkg.MODE_WIDE, qpt))
        emit(what="wide", bases=m.wide_bases, qpt=qpt, dynamic_assignment=dyn, ms=ms, mq_per_s=nq / ms / 1e3, equals_plain=bool(torch.equal(d_res, want)),
             grid_fetches_per_read=a.value / nq, tree_fetches_per_read=o.value / nq, sb96_blocks_per_read=s_.value / nq,
             fetches_per_s=(a.value + o.value + s_.value) / (ms * 1e-3), probe_per_s=probe, fetch_rate_over_probe=(a.value + o.value + s_.value) / (ms * 1e-3) / probe)
    del os.environ["FMGPU_WIDE_DYNAMIC"]
    idx.unwiden()
for ks, lam, lanes in [tuple(int(x) for x in c.split(":")) for c in os.environ.get("FM_SPARSE", "0:0:0,0:0:4").split(",")]:
    t0 = time.time()
    try:
        idx.sparsify(ks, lam, lanes); idx.prepare(length); torch.cuda.synchronize()
    except pkg.FMError as ex:
        emit(what="sparsify failed", bases=ks, lam=lam, lanes=lanes, err=str(ex)); continue
    m = idx.meta
    emit(what="sparsify", bases=m.sparse_bases, lam=m.sparse_lambda, lanes=m.sparse_lanes, seconds=time.time() - t0, sparse_gb=m.sparse_bytes / 1e9, blocks=m.sparse_blocks,
         grid_blocks_per_symbol=m.sparse_uniform_nb, overfull_buckets=m.sparse_overflow, tree_nodes=m.sparse_tree_nodes, tree_rows_fraction=m.sparse_tree_rows / m.bwtsize,
         tree_depth=m.sparse_tree_depth)
    a, s, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
    pkg.check(L.fmgpu_count_fetches_sparse_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream, C.byref(a), C.byref(s), C.byref(o)), "count")
    probe = pkg.gather_probe(0, int(m.sparse_bytes), 256, 2)
    rounds_list = [int(x) for x in os.environ.get("FM_ROUNDS", "0").split(",")]
    for dyn, qpt, rounds in [(0, 4, 0), (0, 2, 0)] + [(1, q, r) for r in rounds_list for q in (4, 3, 2, 1)]:
        os.environ["FMGPU_SPARSE_DYNAMIC"] = str(dyn)
        if rounds: os.environ["FMGPU_SPARSE_ROUNDS"] = str(rounds)
        ms = run(pkg.variant(pkg.MODE_SPARSE, qpt))
        emit(what="sparse", bases=m.sparse_bases, lam=m.sparse_lambda, lanes=m.sparse_lanes, qpt=qpt, dynamic_assignment=dyn, rounds=rounds or 4, ms=ms, mq_per_s=nq / ms / 1e3, equals_plain=bool(torch.equal(d_res, want)),
             grid_fetches_per_read=a.value / nq, tree_fetches_per_read=o.value / nq, sb96_blocks_per_read=s.value / nq,
             fetches_per_s=(a.value + o.value + s.value) / (ms * 1e-3), probe_per_s=probe, fetch_rate_over_probe=(a.value + o.value + s.value) / (ms * 1e-3) / probe)
    idx.unsparsify()

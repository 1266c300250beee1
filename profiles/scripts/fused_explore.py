"""Fused-step kernel at full size: build time, footprint, search time per (lanes, qpt); checks the md5 of the first
1 M reads against the reference's (tests/golden/config3_2g.json).  Writes gpurun_out/fused_explore.jsonl."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
import helpers
OUT = open(os.path.join(ROOT, "gpurun_out", "fused_explore.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "config3_2g.json")))
n, nq, length, k = int(float(os.environ.get("FM_N", "2e9"))), int(float(os.environ.get("FM_NQ", "1e7"))), 100, 2
L = pkg.lib()
b = pkg.IndexBuild.from_synth(n, 1, k, 64); idx = b.to_index(); b.free()
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
wpq = L.fmgpu_words_per_query(length)
d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize(); del d_ascii
def run(v, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search"); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
ms = run(pkg.variant(pkg.MODE_COOP, 1, 256)); emit(what="plain coop", ms=ms, mq_per_s=nq / ms / 1e3)
want = helpers.results_text_md5(d_res.cpu().numpy().view(np.uint32)[:2_000_000])
emit(what="plain md5 ok", ok=(want == gold["md5"]["res_cpu_std_text"]) if n == 2_000_000_000 else None)
for lanes in (2, 4, 1):
    t0 = time.time()
    try:
        idx.fuse(4, lanes)
    except Exception as ex:
        emit(what="fuse failed", lanes=lanes, err=str(ex)); continue
    torch.cuda.synchronize(); t1 = time.time()
    emit(what="fuse", lanes=lanes, seconds=t1 - t0, fused_gb=idx.meta.fused_bytes / 1e9)
    for qpt in (1, 2):
        ms = run(pkg.variant(pkg.MODE_FUSED, qpt))
        got = helpers.results_text_md5(d_res.cpu().numpy().view(np.uint32)[:2_000_000])
        emit(what="fused search", lanes=lanes, qpt=qpt, ms=ms, mq_per_s=nq / ms / 1e3, md5_same_as_plain=(got == want))
    idx.unfuse()

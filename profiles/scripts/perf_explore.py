"""Exploration on a B200: gather roofline probe, GPU index build at full size, kernel variant sweep.
Writes JSON lines to gpurun_out/perf_explore.jsonl.  Not a bench (bench.py is); numbers from here guide tuning."""
import hashlib
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
pkg = importlib.import_module("k-step_fm-index_b200")
import helpers  # noqa: E402

OUT = open(os.path.join(ROOT, "gpurun_out", "perf_explore.jsonl"), "a")


def emit(**kw):
    print(json.dumps(kw), flush=True)
    OUT.write(json.dumps(kw) + "\n"); OUT.flush()


def main():
    n = int(float(os.environ.get("FM_N", "2e9")))
    nq = int(float(os.environ.get("FM_NQ", "1e7")))
    length = int(os.environ.get("FM_LEN", "100"))
    k = int(os.environ.get("FM_K", "2"))
    L = pkg.lib()
    for gb in (0.25, 1, 5.4):
        r = pkg.gather_probe(0, int(gb * (1 << 30)), 512, 3)
        emit(what="gather_probe", table_gb=gb, gloads_per_s=r / 1e9, sector_gbs=r * 32 / 1e9)
    t0 = time.time()
    b = pkg.IndexBuild.from_synth(n, 1, k, 64)
    torch.cuda.synchronize()
    t1 = time.time()
    emit(what="build", n=n, k=k, seconds=t1 - t0, image_bytes=b.image_words * 4)
    idx = b.to_index()
    t2 = time.time()
    emit(what="reblock", seconds=t2 - t1, sb96_bytes=int(idx.meta.nbytes))
    if os.environ.get("FM_MD5", "1") == "1":
        img = b.download()
        t3 = time.time()
        emit(what="image_md5", md5=hashlib.md5(img.tobytes()).hexdigest(), seconds_d2h=t3 - t2, dpos=[int(v) for v in img[6:6 + k]],
             dbase=[int(v) for v in img[6 + k:6 + 2 * k]])
    b.free()
    # reads generated on the device
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "synth reads")
    wpq = L.fmgpu_words_per_query(length)
    d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
    d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack")
    torch.cuda.synchronize()
    best = None
    for mode in (pkg.MODE_TASK, pkg.MODE_COOP):
        for qpt in (1, 2, 4):
            for tpb in (128, 256, 512):
                v = pkg.variant(mode, qpt, tpb)
                times = []
                for it in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), v, stream), "search")
                    e1.record(); torch.cuda.synchronize()
                    times.append(e0.elapsed_time(e1))
                ms = min(times[1:])
                emit(what="search", mode=mode, qpt=qpt, tpb=tpb, ms=ms, mq_per_s=nq / ms / 1e3, glf_steps_per_s=nq * (length // k) / ms / 1e6)
                if best is None or ms < best[0]:
                    best = (ms, mode, qpt, tpb)
    res = d_res.cpu().numpy().view(np.uint32)
    emit(what="check", all_hit=bool(((res[1::2] - res[0::2]) >= 1).all()), singletons=float(((res[1::2] - res[0::2]) == 1).mean()),
         res_md5_first_1m=helpers.results_text_md5(res[:2_000_000]), best=best)


if __name__ == "__main__":
    main()

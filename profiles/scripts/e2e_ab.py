"""A/B of end-to-end feed options in ONE process on one box (alternating, so box-to-box variation cancels)."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200"); L = pkg.lib()
n, nq, length = 2_000_000_000, 10_000_000, 100
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free(); idx.sparsify()
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
h_ascii = torch.empty(nq * length, dtype=torch.uint8, pin_memory=True); h_ascii.copy_(d_ascii); torch.cuda.synchronize(); del d_ascii
h_res = torch.empty(2 * nq, dtype=torch.int32, pin_memory=True)
handles = (C.c_void_p * 1)(idx.handle)
def run(feed, reps=6):
    v = pkg.variant(pkg.MODE_SPARSE, 4, feed=feed)
    for _ in range(2):
        pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, length, h_res.data_ptr(), C.byref(v)), "e2e")
    t0 = time.perf_counter()
    for _ in range(reps):
        pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, length, h_res.data_ptr(), C.byref(v)), "e2e")
    return (time.perf_counter() - t0) / reps * 1e3
for rnd in range(3):
    for pf in (0, 4096, 2048, 8192):
        L.fm_hostpack_set_prefetch(pf)
        for feed in (2, 3):
            ms = run(feed)
            print(json.dumps({"round": rnd, "feed": feed, "prefetch_bytes": pf, "ms": ms, "mq_per_s": nq / ms / 1e3}), flush=True)

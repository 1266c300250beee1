"""Host stream packer: interleaved sub-streams per thread x prefetch distance, alone and inside the end-to-end search
(feed 2 = host packing only, 3 = hybrid).  One process; knobs set through fm_hostpack_set_streams / _set_prefetch.
Appends to gpurun_out/hostpack_streams.jsonl."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200"); L = pkg.lib()
OUT = open(os.path.join(ROOT, "gpurun_out", "hostpack_streams.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
n, nq, length = int(float(os.environ.get("FM_N", "2e9"))), 10_000_000, 100
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free(); idx.sparsify()
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
h_ascii = torch.empty(nq * length, dtype=torch.uint8, pin_memory=True); h_ascii.copy_(d_ascii); torch.cuda.synchronize()
h_res = torch.empty(2 * nq, dtype=torch.int32, pin_memory=True)
want = None
handles = (C.c_void_p * 1)(idx.handle)
out = torch.empty(nq * length // 4 + 64, dtype=torch.uint8, pin_memory=True)
configs = [(1, 8192), (1, 0), (2, 2048), (4, 0), (4, 1024), (4, 2048), (4, 4096), (8, 1024), (8, 2048), (1, 8192), (4, 2048)]
for streams, pf in configs:
    L.fm_hostpack_set_streams(streams); L.fm_hostpack_set_prefetch(pf)
    for _ in range(2): L.fm_hostpack_stream(h_ascii.data_ptr(), nq * length, out.data_ptr(), 0)
    t = time.perf_counter()
    for _ in range(5): L.fm_hostpack_stream(h_ascii.data_ptr(), nq * length, out.data_ptr(), 0)
    dt = (time.perf_counter() - t) / 5
    rec = {"streams": streams, "prefetch": pf, "packer_alone_mreads_per_s": nq / dt / 1e6, "packer_alone_ascii_gb_per_s": nq * length / dt / 1e9}
    for feed in (2, 3):
        v = pkg.variant(pkg.MODE_SPARSE, 4); v.reserved = feed
        for _ in range(3):
            pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, length, h_res.data_ptr(), C.byref(v)), "e2e")
        t = time.perf_counter()
        for _ in range(8):
            pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, length, h_res.data_ptr(), C.byref(v)), "e2e")
        dt = (time.perf_counter() - t) / 8
        r = h_res.numpy().view(np.uint32).copy()
        if want is None: want = r
        rec[f"e2e_feed{feed}_mreads_per_s"] = nq / dt / 1e6
        rec[f"e2e_feed{feed}_same_results"] = bool(np.array_equal(r, want))
    emit(**rec)

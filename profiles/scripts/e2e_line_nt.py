"""A/B of the host stream packer's whole-line non-temporal stores ($FM_HOSTPACK_LINE) inside the end-to-end search: the
process is started once per setting (the knob is read once); prints M reads/s for feed modes 2 (host pack) and 3 (hybrid)."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200"); L = pkg.lib()
n, nq, length = int(float(os.environ.get("FM_N", "2e9"))), 10_000_000, 100
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free(); idx.sparsify()
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
h_ascii = torch.empty(nq * length, dtype=torch.uint8, pin_memory=True); h_ascii.copy_(d_ascii); torch.cuda.synchronize()
h_res = torch.empty(2 * nq, dtype=torch.int32, pin_memory=True)
handles = (C.c_void_p * 1)(idx.handle)
# packer alone
out = torch.empty(nq * length // 4 + 64, dtype=torch.uint8, pin_memory=True)
for _ in range(2): L.fm_hostpack_stream(h_ascii.data_ptr(), nq * length, out.data_ptr(), 0)
t = time.perf_counter()
for _ in range(5): L.fm_hostpack_stream(h_ascii.data_ptr(), nq * length, out.data_ptr(), 0)
dt = (time.perf_counter() - t) / 5
print(json.dumps({"line_nt": os.environ.get("FM_HOSTPACK_LINE", "1"), "what": "stream packer alone", "mreads_per_s": nq / dt / 1e6, "ascii_gb_per_s": nq * length / dt / 1e9}), flush=True)
for feed in (2, 3, 2, 3):
    v = pkg.variant(pkg.MODE_SPARSE, 4); v.reserved = feed
    for _ in range(3):
        pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, length, h_res.data_ptr(), C.byref(v)), "e2e")
    t = time.perf_counter()
    for _ in range(10):
        pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, length, h_res.data_ptr(), C.byref(v)), "e2e")
    dt = (time.perf_counter() - t) / 10
    print(json.dumps({"line_nt": os.environ.get("FM_HOSTPACK_LINE", "1"), "feed": feed, "mreads_per_s": nq / dt / 1e6}), flush=True)

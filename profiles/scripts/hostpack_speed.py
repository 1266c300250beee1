"""Host packer throughput on the GPU box's CPU (per-read reversed packer and pure stream packer)."""
import ctypes as C, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("k-step_fm-index_b200"); L = pkg.lib()
L.fm_hostpack_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
nq, ln = 10_000_000, 100
reads = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(0).integers(0, 4, nq * ln, dtype=np.uint8)].copy()
out = np.zeros(nq * 7, dtype=np.uint32)
for th in (1, 4, 8, 16):
    for name, fn in (("per_read", lambda: L.fm_hostpack_reads(reads.ctypes.data, nq, ln, out.ctypes.data, th)),
                     ("stream", lambda: L.fm_hostpack_stream(reads.ctypes.data, nq * ln, out.ctypes.data, th))):
        fn(); t = time.time(); fn(); fn(); dt = (time.time() - t) / 2
        print(json.dumps({"packer": name, "threads": th, "ms": dt * 1e3, "ascii_gbs": nq * ln / dt / 1e9, "mreads_per_s": nq / dt / 1e6}), flush=True)

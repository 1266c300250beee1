#!/bin/bash
# stream packer throughput per software-prefetch distance (one process per setting: the knob is read once)
for d in 0 256 512 1024 2048 4096; do
  FM_HOSTPACK_PREFETCH=$d python - <<PY
import ctypes as C, importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("k-step_fm-index_b200"); L = pkg.lib()
L.fm_hostpack_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
nq, ln = 10_000_000, 100
reads = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(0).integers(0, 4, nq * ln, dtype=np.uint8)].copy()
out = np.zeros(nq * 7, dtype=np.uint32)
for th in (8, 16):
    fn = lambda: L.fm_hostpack_stream(reads.ctypes.data, nq * ln, out.ctypes.data, th)
    fn(); best = 1e9
    for _ in range(5):
        t = time.time(); fn(); best = min(best, time.time() - t)
    print(json.dumps({"prefetch_bytes": int(os.environ["FM_HOSTPACK_PREFETCH"]), "threads": th, "ms": best * 1e3, "ascii_gbs": nq * ln / best / 1e9}), flush=True)
PY
done

"""Is the random-gather ceiling address translation or DRAM?  All 32 lanes of a warp-level load fall in one
random window (e.g. one 2 MB page) of a 5.4 GB table; compare with fully random lanes."""
import ctypes as C, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("k-step_fm-index_b200")
out = open(os.path.join(ROOT, "gpurun_out", "probe_locality.jsonl"), "a")
table = int(5.4 * (1 << 30))
r = pkg.gather_probe(0, table, 256, 2)
print(json.dumps({"window": "none (fully random lanes)", "gloads_per_s": r / 1e9}), flush=True)
for win in (4096, 65536, 1 << 20, 2 << 20, 32 << 20, 512 << 20):
    v = C.c_double()
    pkg.check(pkg.lib().fmgpu_gather_probe_local(0, table, win, 256, 2, C.byref(v)), "probe_local")
    rec = {"window_bytes": win, "gloads_per_s": v.value / 1e9, "sector_gbs": v.value * 32 / 1e9}
    print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n"); out.flush()

"""Random-access roofline probe over table footprints and access widths (gpurun_out/probe_sweep.jsonl)."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("k-step_fm-index_b200")
out = open(os.path.join(ROOT, "gpurun_out", "probe_sweep.jsonl"), "a")
for gb in (0.0625, 0.25, 1, 5.4, 21, 43, 86, 128):
    for width in (16, 32, 64, 128):
        r = pkg.gather_probe(0, int(gb * (1 << 30)), 256, 2, access_bytes=width)
        rec = {"table_gb": gb, "access_bytes": width, "gaccess_per_s": r / 1e9, "useful_gbs": r * width / 1e9}
        print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n"); out.flush()

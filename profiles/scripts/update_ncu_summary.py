"""Refreshes one entry of profiles/ncu_summary.json (bench.py's roofline.traffic) from profiles/<tag>_prof_<key>.json: the
whole-batch launch (largest grid) of the timed kernel.   usage: python profiles/scripts/update_ncu_summary.py r02 [wide|sparse]"""
import json, os, sys
PR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
key = sys.argv[2] if len(sys.argv) > 2 else "wide"
recs = json.load(open(os.path.join(PR, f"{tag}_prof_{key}.json")))
def num(s): return float(s.split()[0].replace(",", ""))
def nbytes(s):
    v, u = s.split()[0].replace(",", ""), s.split()[1]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
big = max(recs, key=lambda r: num(r["launch__grid_size"]) * (1 if "Lb0" in r["kernel"] or ", 0>" in r["kernel"] or "false" in r["kernel"] else 0.001))
what = {"sparse": "14 bases per 64-byte block, uniform grid of 2 blocks per symbol + search trees, one state machine per read, 2-base lead table, 34.4 GB table",
        "wide": "46 bases per 64-byte block of 5 packed 96-bit entries, block = top 30 bits of the wide symbol, 8-base lead table, 70.4 GB table"}[key]
summ = json.load(open(os.path.join(PR, "ncu_summary.json")))
summ[key] = {"source": f"profiles/{tag}_prof_{key}.json (ncu --set full via profiles/scripts/capture_ncu.sh, {big['kernel'][:90]}, grid {big['launch__grid_size']}: {what}, 10 M x 100 bp reads)",
             "dram_bytes_per_launch_10m_reads": nbytes(big["dram__bytes_read.sum"]) + nbytes(big["dram__bytes_write.sum"]),
             "gpu_time_ms": num(big["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "s": 1e3}[big["gpu__time_duration.sum"].split()[1]]}
json.dump(summ, open(os.path.join(PR, "ncu_summary.json"), "w"), indent=1)
print(json.dumps(summ[key], indent=1))

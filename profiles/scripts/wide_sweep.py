"""r02: wide-step kernel variants on the benchmark's uniform 2 Gbp text (10 M x 100 bp): chained state machine vs burst
(all grid blocks of a read in flight at once), reads per lane group, blocks per chunk.  Appends to gpurun_out/r02_wide_sweep.jsonl."""
import ctypes as C, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
L = pkg.lib()
n, nq, length = int(float(os.environ.get("FM_N", "2e9"))), int(float(os.environ.get("FM_NQ", "1e7"))), int(os.environ.get("FM_LEN", "100"))
b = pkg.IndexBuild.from_synth(n, 1, 2, 64); idx = b.to_index(); b.free()
stream = torch.cuda.current_stream().cuda_stream
d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), stream), "reads")
wpq = L.fmgpu_words_per_query(length)
d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda"); d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack"); torch.cuda.synchronize()
out = open(os.path.join(ROOT, "gpurun_out", "r02_wide_sweep.jsonl"), "a")
want = None
variants = [tuple(int(x) for x in v.split(":")) for v in os.environ["FM_VARIANTS"].split(",")] if os.environ.get("FM_VARIANTS") else \
    [(l_, b_, pf_, q_) for l_ in (2, 4) for b_, pf_ in ((0, 0), (1, 3)) for q_ in (1, 2, 3, 4)]      # lanes:burst:pf:qpt
reps = int(os.environ.get("FM_REPS", "10"))
cur_lanes = 0
for var in variants:
    lanes, burst, pf, qpt = var[:4]
    dyn = var[4] if len(var) > 4 else -1                      # lanes:burst:pf:qpt[:dynamic]  (dynamic -1 = the library's choice)
    if dyn >= 0: os.environ["FMGPU_WIDE_DYNAMIC"] = str(dyn)
    else: os.environ.pop("FMGPU_WIDE_DYNAMIC", None)
    if lanes != cur_lanes:
        if cur_lanes: idx.unwiden()
        idx.widen(int(os.environ.get("FM_W", "0")) or idx.wide_bases_for(length), int(os.environ.get("FM_PB", "0")), lanes); idx.prepare(length)
        cur_lanes = lanes
        m = idx.meta
        print(json.dumps({"wide_bases": m.wide_bases, "lanes": m.wide_lanes, "prefix_bits": m.wide_prefix_bits, "table_gb": m.wide_bytes / 1e9, "overflow": m.wide_overflow,
                          "tree_rows": m.wide_tree_rows, "exceptional": m.wide_exceptional}), flush=True)
    if True:
        os.environ["FMGPU_WIDE_BURST"] = str(burst)
        if pf: os.environ["FMGPU_WIDE_PF"] = str(pf)
        v = pkg.variant(pkg.MODE_WIDE, qpt)
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pkg.check(L.fmgpu_search_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), C.byref(v), stream), "search"); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        if want is None: want = d_res.clone()
        rec = {"len": length, "bases": idx.meta.wide_bases, "lanes": lanes, "dynamic": dyn, "rounds": os.environ.get("FMGPU_WIDE_ROUNDS"), "burst": burst, "pf": pf, "qpt": qpt, "ms_best": min(ts[-8:]), "ms_mean": sum(ts[-8:]) / len(ts[-8:]), "mq_per_s": nq / min(ts[-8:]) / 1e3, "equal": bool(torch.equal(d_res, want))}
        print(json.dumps(rec), flush=True); out.write(json.dumps(rec) + "\n")
        d_res.zero_()

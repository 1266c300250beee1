"""r02: BASELINE config 5 on N physical GPUs through the single-process host driver (csrc/fm_host.c): Task vs Coop vs
Sparse over read lengths 12/25/50/100/250 and k in {1,2} on the 2 Gbp index, FM_NQ reads (default 100 M) sharded over
FMGPU devices 0..N-1 by transferCPUtoGPU; per point 3 x fmgpu_search_index (= searchIndexGPU with an error code), the
device-timed ms of every GPU (CUDA events on each shard's stream) and the wall clock of the call.  Every point is
checked: all reads found, and Task == Coop == Sparse over the whole batch (md5).  Also config 4 (strong scaling) when
FM_STRONG="2,4,8": the same 100 M reads on fewer GPUs.  Appends to gpurun_out/r02_config5_multigpu.jsonl."""
import ctypes as C, hashlib, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("k-step_fm-index_b200")
L = pkg.lib()
OUT = open(os.path.join(ROOT, "gpurun_out", "r02_config5_multigpu.jsonl"), "a")
def emit(**kw):
    print(json.dumps(kw), flush=True); OUT.write(json.dumps(kw) + "\n"); OUT.flush()
n, nq = int(float(os.environ.get("FM_N", "2e9"))), int(float(os.environ.get("FM_NQ", "1e8")))
ndev = int(os.environ.get("FM_GPUS", str(L.fmgpu_device_count())))
lengths = [int(x) for x in os.environ.get("FM_LENGTHS", "12,25,50,100,250").split(",")]
strong = [int(x) for x in os.environ.get("FM_STRONG", "").split(",") if x]
maxlen = max(lengths)
h_ascii = torch.empty(nq * maxlen, dtype=torch.uint8, pin_memory=True)
st = pkg.fmgpu_transfer_stats_t()

def make_reads(length):
    per = ((nq + ndev - 1) // ndev + 31) & ~31
    for g in range(ndev):
        a, b = min(per * g, nq), min(per * (g + 1), nq)
        if b <= a: continue
        with torch.cuda.device(g):
            d = torch.empty((b - a) * length, dtype=torch.uint8, device=f"cuda:{g}")
            pkg.check(L.fmgpu_synth_reads_device(g, n, 1, b - a, length, 2, a, d.data_ptr(), None), "reads")
            torch.cuda.synchronize(g)
            h_ascii[a * length:b * length].copy_(d)
            del d

def run_point(fmi, qry, res, devs, what, k, length, kernels):
    nd = len(devs)
    arr = (C.c_int32 * nd)(*devs)
    pkg.check(L.fmgpu_set_devices(arr, nd), "set_devices")
    t0 = time.time()
    pkg.check(L.transferCPUtoGPU(C.byref(fmi), C.byref(qry), res), "transferCPUtoGPU")
    transfer_s = time.time() - t0
    md5s = {}
    for name, v in kernels:
        L.fmgpu_set_variant(C.byref(v))
        walls, dev_ms = [], []
        for _ in range(4):
            t0 = time.perf_counter()
            pkg.check(L.fmgpu_search_index(C.byref(fmi), C.byref(qry), res), "search")
            walls.append((time.perf_counter() - t0) * 1e3)
            L.fmgpu_get_transfer_stats(C.byref(st))
            dev_ms.append(max(st.search_ms[g] for g in range(nd)))
        pkg.check(L.transferGPUtoCPU(res), "transferGPUtoCPU")
        lr = pkg.resultsArray(res, copy=False)
        v64 = lr.view(np.uint64)
        md5s[name] = f"{int(v64.sum(dtype=np.uint64)):016x}{int(np.bitwise_xor.reduce(v64)):016x}"   # sum and xor of the (L,R) pairs: cheap equality check
        ms = min(dev_ms[1:])
        emit(what=what, k=k, len=length, n_gpus=nd, kernel=name, reads=nq, ms_device_max_over_gpus=ms, ms_wall_clock=min(walls[1:]),
             mq_per_s=nq / ms / 1e3, mq_per_s_wall=nq / min(walls[1:]) / 1e3, g_ref_lf_steps_per_s=nq * (length // k) / ms / 1e6,
             per_gpu_ms=[st.search_ms[g] for g in range(nd)], all_found=bool(((lr[1::2] - lr[0::2]) >= 1).all()),
             same_as_first_kernel=md5s[name] == list(md5s.values())[0], results_checksum=md5s[name], transfer_s=transfer_s)
    L.fmgpu_set_variant(None)

for k in (2, 1):
    t0 = time.time()
    b = pkg.IndexBuild.from_synth(n, 1, k, 64, device=0); image = b.download(); b.free()
    fmi = pkg.index_from_image(image)
    res = pkg.initResults(nq)
    emit(what="index", k=k, image_gb=image.nbytes / 1e9, build_s=time.time() - t0)
    os.environ["FMGPU_MODE"] = "sparse"                      # the sparse-step table on every replica (built once per k, kept across lengths)
    kernels = (("task", pkg.variant(pkg.MODE_TASK, 2, 512)), ("coop", pkg.variant(pkg.MODE_COOP, 1, 256)), ("sparse", pkg.variant(pkg.MODE_SPARSE, 0)))
    for length in lengths:
        make_reads(length)
        qry = pkg.qrys_t(nq, length, h_ascii.data_ptr(), None)
        run_point(fmi, qry, res, list(range(ndev)), "config 5", k, length, kernels)
        qp = C.c_void_p(C.addressof(qry)); L.freeQueriesGPU(C.byref(qp))
    L.fmgpu_get_transfer_stats(C.byref(st))
    emit(what="replicas", k=k, n_gpus=ndev, index_h2d_reblock_s=st.index_h2d_reblock_s, peer_copy_s=[st.peer_copy_s[g] for g in range(1, ndev)],
         peer_copy_gbs=[st.table_bytes / st.peer_copy_s[g] / 1e9 for g in range(1, ndev) if st.peer_copy_s[g] > 0], table_build_s=[st.table_build_s[g] for g in range(ndev)])
    L.freeIndexGPU(C.byref(C.c_void_p(C.addressof(fmi))))
    if k == 2 and strong:
        # config 4, strong scaling: the same reads (length 100) over fewer GPUs; a fresh replica set per device count
        make_reads(100)
        for nd in strong:
            if nd > ndev: continue
            qry = pkg.qrys_t(nq, 100, h_ascii.data_ptr(), None)
            run_point(fmi, qry, res, list(range(nd)), "config 4 (strong scaling)", 2, 100, (("sparse", pkg.variant(pkg.MODE_SPARSE, 0)),))
            qp = C.c_void_p(C.addressof(qry)); L.freeQueriesGPU(C.byref(qp))
            L.freeIndexGPU(C.byref(C.c_void_p(C.addressof(fmi))))
    L.freeResultsGPU(C.byref(res)); L.freeResults(C.byref(res))
    del image

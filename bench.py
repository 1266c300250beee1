#!/usr/bin/env python
"""bench.py -- batched 2-step FM-index backward search on B200 (the one hot path of this repo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): synthetic uniform-random
2 000 000 000-bp reference (fm_synth.h, seed 1), k=2, d=64 index (3.0 GB tag-100 image == what the
reference's gfmiBaseLine writes, re-blocked to the 5.33 GB SB96 device layout), 10 000 000 exact 100-bp
reads PER GPU (seed 2; rank r takes reads [r*10M, (r+1)*10M)) -> weak scaling; N=8 is 80 M reads,
BASELINE configs[3]'s 100 M-read shape.  A "step" = one pass of the search over the rank's 10 M reads.

  value      Mqueries/s, whole job, kernels only, reads packed and resident in HBM (the reference's own
             timed region, common/searchQueries.c:78-98), CUDA events on the launching stream.  Timed
             kernel: the sparse-step kernel (10 bases per 64-byte block fetch, table built on the GPU from
             the 2-step index); the fused-step kernel (4 bases per fetch) and the plain 2-step Coop kernel
             are timed beside it as fused_4base_kernel / plain_2step_kernel; $FM_BENCH_MODE=fused|coop|task
             makes one of those the timed kernel instead.
  e2e        same metric through the C-ABI call fmgpu_search_host with HOST buffers: pinned ASCII reads
             in, (L,R) in pinned host memory out, everything in between (H2D, 2-bit packing on the GPU
             and/or the host, search, D2H) inside the timed region, chunk-pipelined.
  roofline   algorithmic bytes = (exact count of distinct 32-byte sectors the 2-step search must touch,
             counted by an instrumented kernel run) x 32 B, over the timed kernel's mean duration,
             against the measured HBM copy bandwidth of MEASURED_PEAKS.json; the measured random-access
             ceiling (gather probe over the same footprint) is reported next to it.
  cpu_baseline / --impl reference
             the reference's own searchIndexCPU (oracle/_ref/libref_search_k2_d64_std.so, compiled
             from /root/reference) on all host cores, on a bounded sample of the same reads.

Inputs are larger than L2 (26 GB sparse table / 68 GB fused table / 5.33 GB index, 250 MB packed reads vs 126 MB
L2), so no flush between steps.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TEXT = int(float(os.environ.get("FM_BENCH_N", "2e9")))
NQ_PER_GPU = int(float(os.environ.get("FM_BENCH_NQ", "1e7")))
READ_LEN = int(os.environ.get("FM_BENCH_LEN", "100"))
K_STEPS = int(os.environ.get("FM_BENCH_K", "2"))
CHUNK = 64
SEED_REF, SEED_READS = 1, 2
CPU_SAMPLE = int(float(os.environ.get("FM_BENCH_CPU_SAMPLE", "1e6")))
MODE = os.environ.get("FM_BENCH_MODE", "sparse")          # sparse | fused | coop | task
INDEX_TAG = int(os.environ.get("FM_BENCH_TAG", "100"))    # on-disk layout the device index is derived from


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled WHILE a timed region runs.  The `value` region lasts tens of
    milliseconds, so the sampler polls NVML from a thread every millisecond (nvidia_ml_py); if NVML cannot be loaded it
    falls back to `nvidia-smi -lms 200`, started early enough to be running when the region begins."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.rows = gpu_index, None, []
        self.nvml, self.handle, self.thread, self.run, self.samples, self.reason_bits = None, None, None, False, [], 0
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while self.run:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nvml:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self.run = False
            self.thread.join(timeout=1.0)
            n, bits = self.nvml, self.reason_bits
            names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
            reasons = sorted(name for name, const in names if bits & int(getattr(n, const, 0)))
            return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.samples), "how": "NVML polled every ms during the timed region"}
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "how": "nvidia-smi -lms 200"}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, HBM copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(which):
    """dram bytes per launch of the timed search kernel from the committed ncu capture (profiles/), if there is one."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        return d[which]["dram_bytes_per_launch_10m_reads"] * NQ_PER_GPU / 1e7
    except Exception:
        return None


CPU_KIND = "reference"


def reference_search_rate(pkg, image, sample_ascii, steps, warmup, threads=0):
    """The reference's own searchIndexCPU (oracle/_ref, kind "reference") on `sample_ascii`; when the compiled reference is
    missing, the C port of it (oracle/liboracle.so, kind "port").  Returns (Mq/s, seconds/step, cores, (L,R))."""
    global CPU_KIND
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from bindings import Oracle, RefSearcher
    if threads == 0:
        # all host cores, whatever OMP_NUM_THREADS says (torchrun forces it to 1 on its workers)
        threads = int(os.environ.get("FM_BENCH_CPU_THREADS", str(os.cpu_count() or 1)))
    cores = threads
    nq = sample_ascii.size // READ_LEN
    try:
        ref = RefSearcher(K_STEPS, CHUNK, False)
        idx = ref.wrap_image(image)

        def one_pass():
            return ref.search(idx, sample_ascii, READ_LEN, 1, threads)
    except OSError:
        CPU_KIND = "port"
        os.environ["OMP_NUM_THREADS"] = str(threads)
        orc = Oracle()
        idx = orc.wrap(image)

        def one_pass():
            t0 = time.perf_counter()
            res = orc.search(idx, sample_ascii, READ_LEN)
            return res, time.perf_counter() - t0
    out = None
    for _ in range(warmup):
        out, _s = one_pass()
    times = []
    for _ in range(steps):
        out, secs = one_pass()
        times.append(secs)
    sec = sum(times) / len(times)
    return nq / sec / 1e6, sec, cores, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: whatever libraries print there (NCCL's version banner under
    # torchrun, for one) is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0                                           # rank 0 alone runs the CPU arm

    # host packer threads: the ranks of one box share its cores
    # (torchrun forces OMP_NUM_THREADS=1 on its workers; $FM_BENCH_HOST_THREADS overrides our split)
    if args.impl == "ours" and ("OMP_NUM_THREADS" not in os.environ or "TORCHELASTIC_RUN_ID" in os.environ or "FM_BENCH_HOST_THREADS" in os.environ):
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        os.environ["OMP_NUM_THREADS"] = os.environ.get("FM_BENCH_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, local_world))))

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("k-step_fm-index_b200")
    L = pkg.lib()
    if L.fmgpu_device_count() < 1:
        raise SystemExit("bench.py: no sm_100 GPU visible; this product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    distributed = world > 1 and args.impl == "ours"
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.current_stream().cuda_stream
    workload = (f"synthetic {N_TEXT}-bp uniform ACGT reference (seed {SEED_REF}), k={K_STEPS} d={CHUNK} index, "
                f"{NQ_PER_GPU} exact {READ_LEN}-bp reads per GPU (seed {SEED_READS})")

    # ------------------------------------------------------------------ setup (untimed)
    t_setup = time.time()
    image = None
    setup = {}
    if rank == 0:
        t0 = time.time()
        build = pkg.IndexBuild.from_synth(N_TEXT, SEED_REF, K_STEPS, CHUNK, device=dev)
        setup["index_build_s"] = round(time.time() - t0, 3)
        t0 = time.time()
        if INDEX_TAG != 100:
            # search a transformed layout (101 interleaved, 200/201 AltCounters): same files tfmiBMP_* / tfmiAC_* write
            tbuild = build.transform(INDEX_TAG)
            index = tbuild.to_index()
            tbuild.free()
        else:
            index = build.to_index()
        setup["reblock_s"] = round(time.time() - t0, 3)
        setup["index_tag"] = INDEX_TAG
        if world == 1 or args.impl == "reference":
            image = build.download()                       # host copy of the tag-100 file image for the CPU arm
        build.free()
        meta = index.meta
    if distributed:
        # one build + NCCL broadcast of the 5.33 GB block table over NVLink to every other GPU
        sharding = importlib.import_module("k-step_fm-index_b200.sharding")
        meta = sharding.broadcast_meta(meta if rank == 0 else None, 0, dist, pkg.fmgpu_index_meta_t)
        if rank != 0:
            index = pkg.DeviceIndex.alloc_like(meta, device=dev)
        table = torch.as_tensor(index, device=f"cuda:{dev}")
        t0 = time.time()
        dist.broadcast(table, src=0)
        torch.cuda.synchronize()
        setup["index_broadcast_s"] = round(time.time() - t0, 3)

    nq = NQ_PER_GPU if args.impl == "ours" else CPU_SAMPLE
    first = rank * NQ_PER_GPU
    d_ascii = torch.empty(nq * READ_LEN, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(dev, N_TEXT, SEED_REF, nq, READ_LEN, SEED_READS, first, d_ascii.data_ptr(), stream), "synth reads")
    h_ascii = torch.empty(nq * READ_LEN, dtype=torch.uint8, pin_memory=True)
    h_ascii.copy_(d_ascii)
    torch.cuda.synchronize()

    if args.impl == "reference":
        # -------------------------------------------------------------- reference arm: CPU, rank 0 only
        sample = h_ascii.numpy()
        mq, sec, cores, _ = reference_search_rate(pkg, image, sample, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "Mqueries/s", "value": mq, "unit": "Mqueries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "lf_steps_per_s": mq * 1e6 * (READ_LEN // K_STEPS),
                "config": {"workload": workload, "timing": "reference searchIndexCPU under its own omp parallel region, wall clock per pass"},
                "cpu_baseline": {"value": mq, "unit": "Mqueries/s", "cores": cores, "kind": CPU_KIND,
                                 "sample": f"first {nq} reads of the workload per step, {cores} OpenMP threads, index image built on the GPU (byte-identical to gfmiBaseLine's)"},
                "e2e": {"value": mq, "unit": "Mqueries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    wpq = L.fmgpu_words_per_query(READ_LEN)
    d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
    d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(dev, d_ascii.data_ptr(), nq, READ_LEN, d_packed.data_ptr(), stream), "pack")
    torch.cuda.synchronize()
    del d_ascii
    plain_var = pkg.variant(pkg.MODE_TASK if MODE == "task" else pkg.MODE_COOP, int(os.environ.get("FM_BENCH_QPT", "1")),
                            int(os.environ.get("FM_BENCH_TPB", "256")))
    var, fused, sparse = plain_var, False, False
    if MODE == "sparse":
        # sparse-step table built on this replica from its own 2-step block table (every rank builds its own)
        try:
            t0 = time.time()
            index.sparsify(int(os.environ.get("FM_BENCH_SPARSE_BASES", "0")), int(os.environ.get("FM_BENCH_SPARSE_LAMBDA", "0")),
                           int(os.environ.get("FM_BENCH_SPARSE_LANES", "0")))
            torch.cuda.synchronize()
            setup["sparsify_s"] = round(time.time() - t0, 3)
            var, sparse = pkg.variant(pkg.MODE_SPARSE, int(os.environ.get("FM_BENCH_QPT", "4"))), True
        except pkg.FMError as ex:
            setup["sparse_unavailable"] = str(ex)
    if MODE == "fused":
        # fused-step table composed on this replica from its own 2-step block table (every rank builds its own)
        try:
            t0 = time.time()
            index.fuse()
            torch.cuda.synchronize()
            setup["fuse_s"] = round(time.time() - t0, 3)
            var, fused = pkg.variant(pkg.MODE_FUSED, int(os.environ.get("FM_BENCH_QPT", "2"))), True
        except pkg.FMError as ex:
            setup["fuse_unavailable"] = str(ex)
    meta = index.meta

    def search_step(v=None):
        pkg.check(L.fmgpu_search_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), C.byref(v or var), stream), "search")

    # algorithmic bytes: exact count of the 32-byte sectors the 2-step search of THIS rank's reads must touch
    nblk, nsec = C.c_uint64(), C.c_uint64()
    pkg.check(L.fmgpu_count_fetches_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), stream,
                                           C.byref(nblk), C.byref(nsec)), "count fetches")
    lf_steps = nq * (READ_LEN // K_STEPS)
    algo_bytes = nsec.value * 32
    nfb, nlb = C.c_uint64(), C.c_uint64()
    novf = C.c_uint64()
    if fused:
        pkg.check(L.fmgpu_count_fetches_fused_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), stream,
                                                     C.byref(nfb), C.byref(nlb)), "count fused fetches")
    if sparse:
        pkg.check(L.fmgpu_count_fetches_sparse_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), stream,
                                                      C.byref(nfb), C.byref(nlb), C.byref(novf)), "count sparse fetches")

    # measured random-access ceiling over the footprint the timed kernel walks (rank 0, once)
    footprint = int(meta.sparse_bytes) if sparse else int(meta.fused_bytes) if fused else int(meta.nbytes)
    probe = pkg.gather_probe(dev, footprint, 256, 2) if rank == 0 else 0.0

    # the plain 2-step kernel on the same reads, for reference next to the fused one (rank-local, not the headline)
    plain, fused_extra = None, None
    if sparse and os.environ.get("FM_BENCH_ALSO_FUSED", "1") != "0":
        # the fused-step kernel (previous headline kernel) on the same reads; its 68 GB table is released again
        try:
            index.fuse()
            fv = pkg.variant(pkg.MODE_FUSED, 2)
            for _ in range(3):
                search_step(fv)
            fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fe0.record()
            for _ in range(5):
                search_step(fv)
            fe1.record(); torch.cuda.synchronize()
            fused_extra = {"kernel": "fused: 4 bases/step, 64-byte blocks", "table_gb": index.meta.fused_bytes / 1e9,
                           "ms_per_step": fe0.elapsed_time(fe1) / 5, "mqueries_per_s_per_gpu": nq / (fe0.elapsed_time(fe1) / 5) / 1e3}
            index.unfuse()
        except pkg.FMError as ex:
            fused_extra = {"unavailable": str(ex)}
    if fused or sparse:
        for _ in range(3):
            search_step(plain_var)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(5):
            search_step(plain_var)
        pe1.record(); torch.cuda.synchronize()
        plain = {"kernel": "coop" if plain_var.mode == pkg.MODE_COOP else "task", "ms_per_step": pe0.elapsed_time(pe1) / 5,
                 "mqueries_per_s_per_gpu": nq / (pe0.elapsed_time(pe1) / 5) / 1e3}
    setup["setup_s"] = round(time.time() - t_setup, 2)

    # ------------------------------------------------------------------ timed: device-resident
    for _ in range(args.warmup):
        search_step()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]      # per-step marks (recorded, never waited on inside the region)
    e0.record()
    for i in range(args.steps):
        search_step()
        marks[i].record()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    per_step = [(marks[i - 1] if i else e0).elapsed_time(marks[i]) for i in range(args.steps)]
    ms_best = max_over_ranks(min(per_step))                 # best single step (the reference reports best and mean of its iterations)
    clocks = sampler.stop() if rank == 0 else None
    res_dev = d_res.cpu().numpy().view(np.uint32).copy()

    # ------------------------------------------------------------------ timed: end to end through the C ABI, host buffers
    h_res = torch.empty(2 * nq, dtype=torch.int32, pin_memory=True)
    handles = (C.c_void_p * 1)(index.handle)

    def e2e_step():
        pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, READ_LEN, h_res.data_ptr(), C.byref(var)), "search_host")

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()                                         # synchronous: returns when the (L,R) are in host memory
    torch.cuda.synchronize()
    e2e_ms_step = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    barrier()
    same = bool(np.array_equal(h_res.numpy().view(np.uint32), res_dev))

    # extra (not the headline): the same end-to-end call fed with reads that already are 2-bit packed on the host
    h_packed = torch.empty(nq * wpq, dtype=torch.int32, pin_memory=True)
    h_packed.copy_(d_packed)
    torch.cuda.synchronize()
    h_res2 = torch.empty(2 * nq, dtype=torch.int32, pin_memory=True)

    def e2e_packed_step():
        pkg.check(L.fmgpu_search_host_packed(handles, 1, h_packed.data_ptr(), nq, READ_LEN, h_res2.data_ptr(), C.byref(var)), "search_host_packed")

    for _ in range(args.warmup):
        e2e_packed_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_packed_step()
    torch.cuda.synchronize()
    e2e_packed_ms_step = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    barrier()
    same_packed = bool(np.array_equal(h_res2.numpy().view(np.uint32), res_dev))
    hits_ok = bool(((res_dev[1::2] - res_dev[0::2]) >= 1).all())      # every exact read occurs in the text

    # extra (not the headline): locate -- (L,R) -> text positions through the suffix array derived from the index table
    locate = None
    if os.environ.get("FM_BENCH_LOCATE", "1") != "0":
        try:
            t0 = time.time()
            index.build_sa()
            torch.cuda.synchronize()
            sa_s = time.time() - t0
            d_pos = torch.empty(nq, dtype=torch.int32, device="cuda")
            d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")

            def locate_step():
                pkg.check(L.fmgpu_locate_device(index.handle, d_res.data_ptr(), nq, 1, d_pos.data_ptr(), d_cnt.data_ptr(), stream), "locate")

            for _ in range(3):
                locate_step()
            le0, le1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            le0.record()
            for _ in range(5):
                locate_step()
            le1.record(); torch.cuda.synchronize()
            # reads found exactly once must be located where they were cut from (fm_synth.h: fm_synth_read_start)
            j = np.arange(first, first + nq, dtype=np.uint64)
            with np.errstate(over="ignore"):
                x = ((np.uint64(SEED_READS) ^ np.uint64(0xA5A5A5A55A5A5A5A)) + j + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
                x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
                x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
                x = x ^ (x >> np.uint64(31))
            starts = x % np.uint64(N_TEXT - READ_LEN + 1)
            cnt = d_cnt.cpu().numpy().view(np.uint32)
            pos = d_pos.cpu().numpy().view(np.uint32)
            once = cnt == 1
            locate = {"suffix_array_build_s": round(sa_s, 3), "suffix_array_gb": index.meta.sa_bytes / 1e9,
                      "ms_per_step": le0.elapsed_time(le1) / 5, "mqueries_per_s_per_gpu": nq / (le0.elapsed_time(le1) / 5) / 1e3,
                      "reads_found_once": float(once.mean()),
                      "positions_equal_read_starts": bool(np.array_equal(pos[once].astype(np.uint64), starts[once])),
                      "note": "extra, not the headline: SA derived on the GPU from the index table by list ranking over its LF mapping "
                              "(no SA in the reference's files), one gather per read"}
            index.drop_sa()
        except pkg.FMError as ex:
            locate = {"unavailable": str(ex)}

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N=1)
    cpu = None
    parity = None
    if rank == 0 and world == 1 and image is not None:
        ns = min(CPU_SAMPLE, nq)
        mq_cpu, sec_cpu, cores, out = reference_search_rate(pkg, image, h_ascii.numpy()[: ns * READ_LEN], 2, 1)
        parity = bool(np.array_equal(out, res_dev[: 2 * ns]))         # GPU (L,R) == reference CPU (L,R) on the sample
        cpu = {"value": mq_cpu, "unit": "Mqueries/s", "cores": cores, "kind": CPU_KIND,
               "sample": f"first {ns} of the {nq} reads, 2 timed passes after 1 warm-up, {cores} OpenMP threads, reference searchIndexCPU from oracle/_ref",
               "gpu_matches_reference_on_sample": parity}

    if rank == 0:
        peak, peak_src = measured_peak()
        mq = world * nq / ms_step / 1e3
        achieved = algo_bytes / (ms_step * 1e-3) / 1e9
        line = {
            "metric": "Mqueries/s", "value": mq, "unit": "Mqueries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "ms_per_step_best": ms_best, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "lf_steps_per_s": world * lf_steps / (ms_step * 1e-3),
            "config": {"workload": workload,
                       "kernel": (f"sparse: {meta.sparse_bases} bases/step, {32 * meta.sparse_lanes}-byte blocks of occurrence rows (lambda {meta.sparse_lambda}), {meta.sparse_lanes} x 256-bit loads, "
                                  + (f"{meta.sparse_start_bases}-base start table + lead tables, " if meta.sparse_start_bases else "lead tables (no start table at this width), ")
                                  + (f"uniform grid ({meta.sparse_uniform_nb} blocks per symbol, no directory lookup: the text's symbol counts are even), " if meta.sparse_uniform_nb
                                     else "per-symbol block counts + L2-resident directory, ")
                                  + f"qpt={var.queries_per_thread}" if sparse else
                                  f"fused: {meta.fused_bases} bases/step, {32 * meta.fused_lanes}-byte blocks, {meta.fused_lanes} x 256-bit loads, qpt={var.queries_per_thread}"
                                  if fused else f"{MODE} qpt={var.queries_per_thread} tpb={var.threads_per_block}"),
                       "device_layout": (f"sparse-step table {meta.sparse_bytes / 1e9:.1f} GB ({meta.sparse_blocks} blocks, {meta.sparse_overflow} overfull -> SB96 steps) "
                                         f"built on the GPU from the 2-step SB96 table ({meta.nbytes / 1e9:.2f} GB)" if sparse else
                                         f"fused-step table {meta.fused_bytes / 1e9:.1f} GB composed on the GPU from the 2-step SB96 table ({meta.nbytes / 1e9:.2f} GB)"
                                         if fused else "SB96 (16-byte per-symbol blocks: u32 rank + 96 indicator bits)"),
                       "l2": f"inputs larger than L2 ({footprint / 1e9:.1f} GB table, 250 MB packed reads vs 126 MB L2), no flush",
                       "parallelism": f"index replicated, reads sharded x{world}, no collective in the search", "setup": setup},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic("sparse" if sparse else "fused" if fused else "plain"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                         "algorithmic_model": "SURVEY 8(d): 32 B x exact count of sectors the 2-step search must touch (LF steps x |{sector(L),sector(R)}|)",
                         "note": ("the algorithmic bytes are those of the reference's 2-step algorithm; the sparse-step layout needs fewer block fetches than that "
                                  "algorithm has LF steps, so frac may exceed 1 -- table_block_bytes_per_launch and traffic are what this kernel really moves, "
                                  "block_fetches_per_s_over_ceiling is its distance from the measured random-access ceiling") if sparse else None,
                         "sectors_per_lf_step": nsec.value / lf_steps, "blocks_per_lf_step": nblk.value / lf_steps,
                         "table_blocks_per_launch": nfb.value if (fused or sparse) else None,
                         "table_block_bytes_per_launch": nfb.value * 32 * (meta.sparse_lanes if sparse else meta.fused_lanes) if (fused or sparse) else None,
                         "sb96_blocks_per_launch": nlb.value if (fused or sparse) else nblk.value,
                         "overflow_fallbacks_per_launch": novf.value if sparse else None,
                         "block_fetches_per_s": ((nfb.value + nlb.value) if (fused or sparse) else nblk.value) / (ms_step * 1e-3),
                         "dram_fill_bytes_per_fetch": 64 if (fused or (sparse and meta.sparse_lanes == 2)) else 128,
                         "random_access_ceiling": {"accesses_per_s": probe,
                                                   "how": "independent uniform random 16-byte loads over a table of the same footprint; the ceiling is a miss RATE "
                                                          "(requests/s), the same for 64- and 128-byte fills (profiles/r01_prefetch_variants.md)",
                                                   "block_fetches_per_s_over_ceiling": (((nfb.value + nlb.value) if (fused or sparse) else nblk.value) / (ms_step * 1e-3)) / probe if probe else None},
                         "frac_of_nominal_8tbs": achieved / 8000.0},
            "plain_2step_kernel": plain,
            "fused_4base_kernel": fused_extra,
            "cpu_baseline": cpu,
            "e2e": {"value": world * nq / e2e_ms_step / 1e3, "unit": "Mqueries/s", "ms_per_step": e2e_ms_step,
                    "h2d_bytes_per_step": nq * READ_LEN, "d2h_bytes_per_step": nq * 8, "matches_device_resident_result": same,
                    "feed": os.environ.get("FMGPU_FEED", "auto (self-tuning during warm-up: hybrid of ASCII-over-PCIe and AVX-512 host packing, or ASCII only)"),
                    "host_pack_threads_per_rank": int(L.fm_hostpack_threads()),
                    "note": "h2d_bytes_per_step counts the ASCII reads handed to the call; host-packed chunks cross PCIe as 2-bit (25 B/read)"},
            "e2e_packed_input": {"value": world * nq / e2e_packed_ms_step / 1e3, "unit": "Mqueries/s", "ms_per_step": e2e_packed_ms_step,
                                 "h2d_bytes_per_step": nq * wpq * 4, "d2h_bytes_per_step": nq * 8, "matches_device_resident_result": same_packed,
                                 "note": "extra, not the headline: fmgpu_search_host_packed, host reads already in the 2-bit binary format (28 B per 100-bp read)"},
            "locate": locate,
            "gpu_launches": args.steps,
            "clocks": clocks,
            "checks": {"every_read_found": hits_ok, "e2e_equals_resident": same, "gpu_equals_reference_cpu_on_sample": parity},
        }
        emit(line)
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

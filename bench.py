#!/usr/bin/env python
"""bench.py -- batched 2-step FM-index backward search on B200 (the one hot path of this repo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): synthetic uniform-random
2 000 000 000-bp reference (fm_synth.h, seed 1), k=2, d=64 index (3.0 GB tag-100 image == what the
reference's gfmiBaseLine writes, md5-checked against the reference's own 27-minute build, re-blocked to the 5.33 GB
SB96 device layout), 10 000 000 exact 100-bp reads PER GPU (seed 2; rank r takes reads [r*10M, (r+1)*10M)) -> weak
scaling; N=8 is 80 M reads, BASELINE configs[3]'s 100 M-read shape.  A "step" = one pass of the search over the
rank's 10 M reads -- in BOTH arms: `--impl reference` runs the reference's own searchIndexCPU over the same 10 M reads.

  value      Mqueries/s, whole job, kernels only, reads packed and resident in HBM (the reference's own
             timed region, common/searchQueries.c:78-98), CUDA events on the launching stream.  Timed
             kernel: the wide-step kernel (46 bases per 64-byte block fetch: 100 bp = 8 + 2 x 46; table built on the GPU from
             the 2-step index); the sparse-step kernel (14 bases per 64-byte fetch), the fused-step kernel
             (4 bases per fetch) and the plain 2-step Coop kernel are timed beside it as sparse_14base_kernel /
             fused_4base_kernel / plain_2step_kernel; $FM_BENCH_MODE=sparse|fused|coop|task makes one of
             those the timed kernel instead.
  e2e        same metric through the C-ABI call fmgpu_search_host with HOST buffers: pinned ASCII reads
             in, (L,R) in pinned host memory out, everything in between (H2D, 2-bit packing on the GPU
             and/or the host, search, D2H) inside the timed region, chunk-pipelined.  e2e.host_ceiling =
             measured host DRAM read bandwidth / 100 bytes per read: the bound of ANY feed of ASCII reads.
  roofline   achieved = bytes the TIMED kernel must move (its block fetches x block size + packed reads in +
             results out, counted by an instrumented run of the same kernel) / its mean launch time; frac =
             achieved / measured HBM copy bandwidth (MEASURED_PEAKS.json); traffic = dram bytes of the same kernel
             from the committed ncu capture.  vs_reference_algorithm_bytes keeps SURVEY 8(d)'s yardstick (sectors
             the reference's 2-step algorithm must touch) and request_rate_frac the distance from the measured
             random-access ceiling (gather probe over the same footprint) -- the limit that actually binds.
  cpu_baseline / --impl reference
             the reference's own searchIndexCPU (oracle/_ref/libref_search_k2_d64_std.so, compiled
             from /root/reference) on all host cores.

$FM_BENCH_SINGLE_PROCESS=1: ONE process drives all N GPUs through the reference's own call sequence
(transferCPUtoGPU: one H2D + cudaMemcpyPeer replicas, N x 10 M reads sharded; searchIndexGPU; transferGPUtoCPU) --
the drop-in host driver of csrc/fm_host.c instead of one rank per GPU.  Under torchrun only rank 0 works.

Inputs are larger than L2 (74 GB wide table / 34 GB sparse table / 68 GB fused table / 5.33 GB index, 280 MB packed
reads vs 126 MB L2), so no flush between steps.
"""
import argparse
import ctypes as C
import hashlib
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
CPU_SAMPLE = int(float(os.environ.get("FM_BENCH_CPU_SAMPLE", "1e6")))   # reads of the strided parity sample / cpu_baseline at N=1
MODE = os.environ.get("FM_BENCH_MODE", "wide")            # wide | sparse | fused | coop | task
INDEX_TAG = int(os.environ.get("FM_BENCH_TAG", "100"))    # on-disk layout the device index is derived from
SINGLE_PROCESS = os.environ.get("FM_BENCH_SINGLE_PROCESS", "0") == "1"


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
        return d[which]["dram_bytes_per_launch_10m_reads"] * NQ_PER_GPU / 1e7, d[which]["source"]
    except Exception:
        return None, None


def shared_config(n_gpus):
    """The `config` dict: identical in both arms (the driver compares them), so nothing implementation-specific here."""
    return {"workload": (f"synthetic {N_TEXT}-bp uniform ACGT reference (seed {SEED_REF}), k={K_STEPS} d={CHUNK} index, "
                         f"{NQ_PER_GPU} exact {READ_LEN}-bp reads per GPU (seed {SEED_READS})"),
            "reads_per_step": NQ_PER_GPU * (n_gpus if n_gpus > 0 else 1), "read_length": READ_LEN, "k": K_STEPS, "d": CHUNK,
            "text_bp": N_TEXT, "n_gpus": n_gpus,
            "l2": "inputs larger than L2 (multi-GB tables, 280 MB packed reads vs 126 MB L2), no flush",
            "parallelism": f"index replicated, reads sharded x{n_gpus}, no collective in the search"}


def golden_index_md5():
    """md5 of the tag-100 file the UNMODIFIED reference builder wrote for this text (27 minutes; tests/golden/config3_2g.json)."""
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "config3_2g.json")))
        if (g["text"]["n"], g["text"]["seed"], g["k"], g["d"]) == (N_TEXT, SEED_REF, K_STEPS, CHUNK):
            return g["md5"]["tag100_fmi"]
    except Exception:
        pass
    return None


def image_md5(image):
    h = hashlib.md5()
    b = memoryview(image).cast("B")
    for a in range(0, len(b), 1 << 26):
        h.update(b[a:a + (1 << 26)])
    return h.hexdigest()


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


def skewed_text_extra(pkg, L, torch, dev, stream):
    """Extra key `skewed_text`: the same kernel family on a NON-uniform text -- 400 Mbp, half of it 10 %-diverged copies of
    five repeat families (the shape of profiles/r01_repeat_text.md, generated on the GPU here) -- where wide symbols
    occur thousands of times and their buckets become search trees.  Sparse-step vs fused-step vs plain Coop, all (L,R)
    equal to the plain kernel's, which is checked against the reference searcher on a strided sample."""
    n, nq, length = int(float(os.environ.get("FM_BENCH_SKEWED_N", "4e8"))), 4_000_000, READ_LEN
    g = torch.Generator(device="cuda"); g.manual_seed(17)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    text = acgt[torch.randint(0, 4, (n,), device="cuda", generator=g)]
    fam_len = (300, 300, 310, 6000, 150)
    fam_share = (0.5, 0.15, 0.1, 0.2, 0.05)
    for flen, share in zip(fam_len, fam_share):
        unit = acgt[torch.randint(0, 4, (flen,), device="cuda", generator=g)]
        copies = int(n * 0.5 * share / flen)
        for c0 in range(0, copies, 1 << 15):
            c = min(1 << 15, copies - c0)
            m = unit.repeat(c, 1)
            mut = torch.rand((c, flen), device="cuda", generator=g) < 0.10
            m[mut] = acgt[torch.randint(0, 4, (int(mut.sum().item()),), device="cuda", generator=g)]
            pos = torch.randint(0, n - flen, (c,), device="cuda", generator=g)
            text[(pos[:, None] + torch.arange(flen, device="cuda")[None, :]).reshape(-1)] = m.reshape(-1)
    if (n + 1) % CHUNK == 0:
        text = text[:-1]; n -= 1
    starts = torch.randint(0, n - length, (nq,), device="cuda", generator=g)
    d_ascii = text[(starts[:, None] + torch.arange(length, device="cuda")[None, :])].reshape(-1).contiguous()
    h_text = text.cpu().numpy()
    del text
    t0 = time.time()
    build = pkg.IndexBuild.from_text(h_text, K_STEPS, CHUNK, device=dev)
    index = build.to_index()
    build_s = time.time() - t0
    wpq = L.fmgpu_words_per_query(length)
    d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
    d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(dev, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack")
    torch.cuda.synchronize()

    def timed(v, reps=5):
        best = 1e30
        for i in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pkg.check(L.fmgpu_search_device(index.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), C.byref(v), stream), "search")
            e1.record(); torch.cuda.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1))
        return best
    out = {"text": f"{n} bp, half random, half 10 %-diverged copies of 5 repeat families (150-6000 bp units), {nq} exact {length}-bp reads",
           "index_build_s": round(build_s, 3)}
    ms = timed(pkg.variant(pkg.MODE_COOP, 1, 256))
    want = d_res.clone()
    out["plain_coop"] = {"ms": ms, "mqueries_per_s": nq / ms / 1e3}
    # the plain kernel against the reference searcher, strided sample
    ns = min(CPU_SAMPLE, nq)
    sel = np.arange(0, nq, max(1, nq // ns))[:ns]
    image = build.download()
    build.free()
    sample = d_ascii.cpu().numpy().reshape(nq, length)[sel].reshape(-1)
    _mq, _s, _c, ref_out = reference_search_rate(pkg, image, sample, 1, 0)
    out["plain_equals_reference_on_strided_sample"] = bool(np.array_equal(ref_out, want.cpu().numpy().view(np.uint32).reshape(nq, 2)[sel].reshape(-1)))
    del image
    try:
        index.fuse()
        ms = timed(pkg.variant(pkg.MODE_FUSED, 2))
        out["fused_4base"] = {"ms": ms, "mqueries_per_s": nq / ms / 1e3, "table_gb": index.meta.fused_bytes / 1e9, "equals_plain": bool(torch.equal(d_res, want))}
        index.unfuse()
    except pkg.FMError as ex:
        out["fused_4base"] = {"unavailable": str(ex)}
    a, s, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
    try:
        wb = index.wide_bases_for(length)
        index.widen(wb); index.prepare(length)
        m = index.meta
        pkg.check(L.fmgpu_count_fetches_wide_device(index.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream, C.byref(a), C.byref(s), C.byref(o)), "count")
        probe = pkg.gather_probe(dev, int(m.wide_bytes), 256, 2)
        by_assignment = {}
        for dyn in ("0", "1"):
            os.environ["FMGPU_WIDE_DYNAMIC"] = dyn
            by_assignment["dynamic" if dyn == "1" else "static"] = min(timed(pkg.variant(pkg.MODE_WIDE, q)) for q in (1, 2))
        del os.environ["FMGPU_WIDE_DYNAMIC"]
        ms = min(timed(pkg.variant(pkg.MODE_WIDE, q)) for q in (1, 2))          # the library's own choice of read assignment
        fetches = a.value + s.value + o.value
        out["wide"] = {"ms": ms, "ms_by_read_assignment": by_assignment, "entry_words": m.wide_entry_words, "mqueries_per_s": nq / ms / 1e3, "equals_plain": bool(torch.equal(d_res, want)), "bases_per_step": m.wide_bases,
                       "table_gb": m.wide_bytes / 1e9, "rows_in_search_trees": m.wide_tree_rows / m.bwtsize, "tree_depth": m.wide_tree_depth,
                       "exceptional_buckets": m.wide_exceptional,
                       "grid_fetches_per_read": a.value / nq, "tree_fetches_per_read": o.value / nq, "sb96_fetches_per_read": s.value / nq,
                       "fetches_per_s": fetches / (ms * 1e-3), "probe_accesses_per_s": probe, "request_rate_frac": fetches / (ms * 1e-3) / probe}
        index.unwiden()
    except pkg.FMError as ex:
        out["wide"] = {"unavailable": str(ex)}
    index.sparsify(); index.prepare(length)
    m = index.meta
    pkg.check(L.fmgpu_count_fetches_sparse_device(index.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream, C.byref(a), C.byref(s), C.byref(o)), "count")
    probe = pkg.gather_probe(dev, int(m.sparse_bytes), 256, 2)
    ms = min(timed(pkg.variant(pkg.MODE_SPARSE, q)) for q in (1, 2, 3))   # reads handed out dynamically (more than 1 % of the rows in trees)
    fetches = a.value + s.value + o.value
    out["sparse"] = {"ms": ms, "mqueries_per_s": nq / ms / 1e3, "equals_plain": bool(torch.equal(d_res, want)), "bases_per_step": m.sparse_bases,
                     "table_gb": m.sparse_bytes / 1e9, "rows_in_search_trees": m.sparse_tree_rows / m.bwtsize, "tree_depth": m.sparse_tree_depth,
                     "grid_fetches_per_read": a.value / nq, "tree_fetches_per_read": o.value / nq, "sb96_fetches_per_read": s.value / nq,
                     "fetches_per_s": fetches / (ms * 1e-3), "probe_accesses_per_s": probe, "request_rate_frac": fetches / (ms * 1e-3) / probe}
    index.free()
    return out


def run_reference_arm(args, pkg, L, torch, dev, stream, emit):
    """--impl reference: the UNMODIFIED searchIndexCPU on this box's host cores, the SAME 10 M reads per step as the GPU arm's
    rank 0, on the tag-100 image built on the GPU and proven (md5) to be the file the reference's own builder writes."""
    t0 = time.time()
    build = pkg.IndexBuild.from_synth(N_TEXT, SEED_REF, K_STEPS, CHUNK, device=dev)
    image = build.download()
    build.free()
    md5, want_md5 = image_md5(image), golden_index_md5()
    nq = NQ_PER_GPU
    d_ascii = torch.empty(nq * READ_LEN, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(dev, N_TEXT, SEED_REF, nq, READ_LEN, SEED_READS, 0, d_ascii.data_ptr(), stream), "synth reads")
    sample = d_ascii.cpu().numpy()
    del d_ascii
    torch.cuda.empty_cache()
    setup_s = time.time() - t0
    mq, sec, cores, out = reference_search_rate(pkg, image, sample, args.steps, args.warmup)
    hits_ok = bool(((out[1::2] - out[0::2]) >= 1).all())
    emit({"impl": "reference", "metric": "Mqueries/s", "value": mq, "unit": "Mqueries/s", "n_gpus": args.gpus,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
          "lf_steps_per_s": mq * 1e6 * (READ_LEN // K_STEPS),
          "config": shared_config(args.gpus),
          "reference_arm": {"timing": "reference searchIndexCPU under its own omp parallel region, wall clock per pass over the step's reads",
                            "reads_per_step": nq, "setup_s": round(setup_s, 1),
                            "index_md5": md5, "index_md5_reference_build": want_md5,
                            "index_is_the_reference_builders_file": (md5 == want_md5) if want_md5 else None,
                            "every_read_found": hits_ok,
                            "note": "the index image is built by this repo's GPU builder only to avoid the reference's 27-minute CPU build; its md5 equals the "
                                    "md5 of the file gfmiBaseLine_64bases_2step wrote for the same text (tests/golden/config3_2g.json); the timed code is the "
                                    "unmodified reference searcher (oracle/_ref/libref_search_k2_d64_std.so). At N > 1 the CPU arm still searches one rank's "
                                    "10 M reads per step: its Mqueries/s does not depend on the batch size"},
          "cpu_baseline": {"value": mq, "unit": "Mqueries/s", "cores": cores, "kind": CPU_KIND,
                           "sample": f"{nq} reads per step (the whole step of one GPU rank), {cores} OpenMP threads"},
          "e2e": {"value": mq, "unit": "Mqueries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})
    return 0


def run_single_process(args, pkg, L, torch, emit):
    """$FM_BENCH_SINGLE_PROCESS=1: the reference's own call sequence, one process, all N GPUs (csrc/fm_host.c)."""
    ndev = args.gpus
    if L.fmgpu_device_count() < ndev:
        raise SystemExit(f"bench.py: {ndev} GPUs asked, {L.fmgpu_device_count()} visible")
    stream0 = None
    t_setup = time.time()
    build = pkg.IndexBuild.from_synth(N_TEXT, SEED_REF, K_STEPS, CHUNK, device=0)
    image = build.download()
    build.free()
    md5, want_md5 = image_md5(image), golden_index_md5()
    nq_total = NQ_PER_GPU * ndev
    h_ascii = torch.empty(nq_total * READ_LEN, dtype=torch.uint8, pin_memory=True)
    for g in range(ndev):
        with torch.cuda.device(g):
            d = torch.empty(NQ_PER_GPU * READ_LEN, dtype=torch.uint8, device=f"cuda:{g}")
            pkg.check(L.fmgpu_synth_reads_device(g, N_TEXT, SEED_REF, NQ_PER_GPU, READ_LEN, SEED_READS, g * NQ_PER_GPU, d.data_ptr(), stream0), "synth reads")
            torch.cuda.synchronize(g)
            h_ascii[g * NQ_PER_GPU * READ_LEN:(g + 1) * NQ_PER_GPU * READ_LEN].copy_(d)
            del d
    fmi = pkg.index_from_image(image)
    qry = pkg.qrys_t(nq_total, READ_LEN, h_ascii.data_ptr(), None)
    res = pkg.initResults(nq_total)
    devs = (C.c_int32 * ndev)(*range(ndev))
    pkg.check(L.fmgpu_set_devices(devs, ndev), "fmgpu_set_devices")
    L.fmgpu_set_variant(None)
    t0 = time.time()
    pkg.check(L.transferCPUtoGPU(C.byref(fmi), C.byref(qry), res), "transferCPUtoGPU")
    transfer_s = time.time() - t0
    setup_s = time.time() - t_setup
    for _ in range(args.warmup):
        pkg.check(L.fmgpu_search_index(C.byref(fmi), C.byref(qry), res), "search")
    samplers = [ClockSampler(g) for g in range(ndev)]
    for s in samplers:
        s.start()
    per_gpu = [[] for _ in range(ndev)]
    st = pkg.fmgpu_transfer_stats_t()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pkg.check(L.fmgpu_search_index(C.byref(fmi), C.byref(qry), res), "search")   # launches every shard, waits for all
        L.fmgpu_get_transfer_stats(C.byref(st))
        for g in range(ndev):
            per_gpu[g].append(st.search_ms[g])
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = [s.stop() for s in samplers]
    ms_dev = max(sum(v) / len(v) for v in per_gpu)            # device-timed (CUDA events per shard), max over GPUs
    t0 = time.time()
    pkg.check(L.transferGPUtoCPU(res), "transferGPUtoCPU")
    d2h_s = time.time() - t0
    L.fmgpu_get_transfer_stats(C.byref(st))
    lr = pkg.resultsArray(res, copy=False)
    hits_ok = bool(((lr[1::2] - lr[0::2]) >= 1).all())
    # parity: the reference searcher on a strided sample over ALL shards
    ns = min(CPU_SAMPLE, nq_total)
    sel = np.arange(0, nq_total, max(1, nq_total // ns))[:ns]
    sample = h_ascii.numpy().reshape(nq_total, READ_LEN)[sel].reshape(-1)
    mq_cpu, sec_cpu, cores, out = reference_search_rate(pkg, image, sample, 2, 1)
    parity = bool(np.array_equal(out, lr.reshape(nq_total, 2)[sel].reshape(-1)))
    res_md5 = hashlib.md5(memoryview(lr).cast("B")).hexdigest()
    L.freeIndexGPU(C.byref(C.c_void_p(C.addressof(fmi))))
    qp = C.c_void_p(C.addressof(qry)); L.freeQueriesGPU(C.byref(qp))
    L.freeResultsGPU(C.byref(res))
    # e2e: the host-buffer call over replicas made the same way (one upload + peer copies)
    first = pkg.DeviceIndex.from_image(image, device=0)
    reps = [first] + [first.replicate(g) for g in range(1, ndev)]
    wb = reps[0].wide_bases_for(READ_LEN)
    for r in reps:
        r.widen(wb) if wb else r.sparsify()
    var = pkg.variant(pkg.MODE_WIDE, 0) if wb else pkg.variant(pkg.MODE_SPARSE, 3)
    h_res = torch.empty(2 * nq_total, dtype=torch.int32, pin_memory=True)
    handles = (C.c_void_p * ndev)(*[r.handle for r in reps])

    def e2e_step():
        pkg.check(L.fmgpu_search_host(handles, ndev, h_ascii.data_ptr(), nq_total, READ_LEN, h_res.data_ptr(), C.byref(var)), "search_host")
    for _ in range(max(args.warmup, 6)):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    same = bool(np.array_equal(h_res.numpy().view(np.uint32), lr))
    host_bw = L.fm_host_read_bandwidth(h_ascii.data_ptr(), min(h_ascii.numel(), 1 << 31), 0, 3)
    meta = reps[0].meta
    peak, peak_src = measured_peak()
    emit({"metric": "Mqueries/s", "value": nq_total / ms_dev / 1e3, "unit": "Mqueries/s", "n_gpus": ndev, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": ms_dev, "ms_per_step_wall_clock_incl_launch_and_sync": wall_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "u32", "data": "synthetic", "lf_steps_per_s": nq_total * (READ_LEN // K_STEPS) / (ms_dev * 1e-3),
          "config": shared_config(ndev),
          "single_process": {"driver": "transferCPUtoGPU -> searchIndexGPU x steps -> transferGPUtoCPU (csrc/fm_host.c), FMGPU devices 0.." + str(ndev - 1),
                             "per_gpu_ms_per_step": [sum(v) / len(v) for v in per_gpu], "per_gpu_sm_mhz": [c["sm_mhz"] for c in clocks],
                             "index_h2d_reblock_s": st.index_h2d_reblock_s, "index_file_gb": st.index_file_bytes / 1e9,
                             "peer_copy_s": [st.peer_copy_s[g] for g in range(1, ndev)],
                             "peer_copy_gbs": [st.table_bytes / st.peer_copy_s[g] / 1e9 if st.peer_copy_s[g] > 0 else None for g in range(1, ndev)],
                             "table_build_s_per_gpu": [st.table_build_s[g] for g in range(ndev)], "queries_h2d_pack_s": st.queries_h2d_pack_s,
                             "results_d2h_s": d2h_s, "transferCPUtoGPU_s": transfer_s, "setup_s": round(setup_s, 1),
                             "index_md5": md5, "index_is_the_reference_builders_file": (md5 == want_md5) if want_md5 else None, "results_md5": res_md5},
          "kernel_config": {"kernel": (f"wide: {meta.wide_bases} bases/step, {32 * meta.wide_lanes}-byte blocks of {32 * meta.wide_entry_words}-bit entries, 2^{meta.wide_prefix_bits} buckets" if meta.wide_bases else
                                       f"sparse: {meta.sparse_bases} bases/step, grid of {meta.sparse_uniform_nb} blocks per symbol"),
                            "table_gb": (meta.wide_bytes if meta.wide_bases else meta.sparse_bytes) / 1e9},
          "cpu_baseline": {"value": mq_cpu, "unit": "Mqueries/s", "cores": cores, "kind": CPU_KIND,
                           "sample": f"{ns} reads strided over all {ndev} shards, 2 timed passes after 1 warm-up", "gpu_matches_reference_on_sample": parity},
          "e2e": {"value": nq_total / e2e_ms / 1e3, "unit": "Mqueries/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": nq_total * READ_LEN,
                  "d2h_bytes_per_step": nq_total * 8, "matches_device_resident_result": same,
                  "host_ceiling": {"host_read_gbs": host_bw, "bytes_per_read": READ_LEN, "mqueries_per_s": host_bw * 1e3 / READ_LEN,
                                   "e2e_over_ceiling": (nq_total / e2e_ms / 1e3) / (host_bw * 1e3 / READ_LEN) if host_bw > 0 else None}},
          "roofline": {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                       "note": "kernel roofline is reported by the default (one rank per GPU) mode; this mode measures the host driver"},
          "gpu_launches": args.steps * ndev,
          "clocks": clocks[0],
          "checks": {"every_read_found": hits_ok, "e2e_equals_resident": same, "gpu_equals_reference_cpu_on_strided_sample": parity}})
    for r in reps:
        r.free()
    return 0


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
    if (args.impl == "reference" or SINGLE_PROCESS) and rank != 0:
        return 0                                           # rank 0 alone runs the CPU arm / the single-process driver

    # host packer threads: the ranks of one box share its cores
    # (torchrun forces OMP_NUM_THREADS=1 on its workers; $FM_BENCH_HOST_THREADS overrides our split)
    if args.impl == "ours" and ("OMP_NUM_THREADS" not in os.environ or "TORCHELASTIC_RUN_ID" in os.environ or "FM_BENCH_HOST_THREADS" in os.environ):
        local_world = 1 if SINGLE_PROCESS else int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        os.environ["OMP_NUM_THREADS"] = os.environ.get("FM_BENCH_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, local_world))))

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("k-step_fm-index_b200")
    L = pkg.lib()
    if L.fmgpu_device_count() < 1:
        raise SystemExit("bench.py: no sm_100 GPU visible; this product has no CPU path")
    if args.impl == "ours" and SINGLE_PROCESS:
        return run_single_process(args, pkg, L, torch, emit)
    torch.cuda.set_device(local_rank)
    dev = local_rank
    stream = torch.cuda.current_stream().cuda_stream
    if args.impl == "reference":
        return run_reference_arm(args, pkg, L, torch, dev, stream, emit)
    distributed = world > 1
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

    def all_ranks(x):
        """x of every rank, in rank order (rank 0 reports them: max_over_ranks alone hides which GPU is the slow one)"""
        if not distributed:
            return [x]
        t = torch.tensor([float(x) if x is not None else -1.0], dtype=torch.float64, device="cuda")
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(p.item()) for p in parts]

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
        if world == 1:
            image = build.download()                       # host copy of the tag-100 file image for the CPU baseline
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

    nq = NQ_PER_GPU
    first = rank * NQ_PER_GPU
    d_ascii = torch.empty(nq * READ_LEN, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(dev, N_TEXT, SEED_REF, nq, READ_LEN, SEED_READS, first, d_ascii.data_ptr(), stream), "synth reads")
    h_ascii = torch.empty(nq * READ_LEN, dtype=torch.uint8, pin_memory=True)
    h_ascii.copy_(d_ascii)
    torch.cuda.synchronize()

    wpq = L.fmgpu_words_per_query(READ_LEN)
    d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
    d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(dev, d_ascii.data_ptr(), nq, READ_LEN, d_packed.data_ptr(), stream), "pack")
    torch.cuda.synchronize()
    del d_ascii
    plain_var = pkg.variant(pkg.MODE_TASK if MODE == "task" else pkg.MODE_COOP, int(os.environ.get("FM_BENCH_QPT", "1")),
                            int(os.environ.get("FM_BENCH_TPB", "256")))
    var, fused, sparse, wide = plain_var, False, False, False
    if MODE == "wide":
        # wide-step table built on this replica from its own 2-step block table (every rank builds its own); the step
        # width is the one that serves this read length with the fewest fetches (30 bases for 100-bp reads)
        try:
            t0 = time.time()
            wb = int(os.environ.get("FM_BENCH_WIDE_BASES", "0"))
            if wb:
                index.widen(wb, int(os.environ.get("FM_BENCH_WIDE_PREFIX_BITS", "0")), int(os.environ.get("FM_BENCH_WIDE_LANES", "0")))
            else:
                index.widen_for(READ_LEN)                       # the width with the fewest steps (the 64-bit-entry one when memory is short)
            index.prepare(READ_LEN)
            torch.cuda.synchronize()
            setup["widen_s"] = round(time.time() - t0, 3)
            var, wide = pkg.variant(pkg.MODE_WIDE, int(os.environ.get("FM_BENCH_QPT", "0"))), True
        except pkg.FMError as ex:
            setup["wide_unavailable"] = str(ex)
    if MODE == "sparse" or (MODE == "wide" and not wide):
        # sparse-step table built on this replica from its own 2-step block table (every rank builds its own)
        try:
            t0 = time.time()
            index.sparsify(int(os.environ.get("FM_BENCH_SPARSE_BASES", "0")), int(os.environ.get("FM_BENCH_SPARSE_LAMBDA", "0")),
                           int(os.environ.get("FM_BENCH_SPARSE_LANES", "0")))
            index.prepare(READ_LEN)
            torch.cuda.synchronize()
            setup["sparsify_s"] = round(time.time() - t0, 3)
            var, sparse = pkg.variant(pkg.MODE_SPARSE, int(os.environ.get("FM_BENCH_QPT", "3"))), True
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

    # the reference algorithm's yardstick: exact count of the 32-byte sectors the 2-step search of THIS rank's reads must touch
    nblk, nsec = C.c_uint64(), C.c_uint64()
    pkg.check(L.fmgpu_count_fetches_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), stream,
                                           C.byref(nblk), C.byref(nsec)), "count fetches")
    lf_steps = nq * (READ_LEN // K_STEPS)
    ref_algo_bytes = nsec.value * 32
    # the TIMED kernel's own block fetches (instrumented instantiation of the same kernel)
    nfb, nlb, ntree = C.c_uint64(), C.c_uint64(), C.c_uint64()
    if fused:
        pkg.check(L.fmgpu_count_fetches_fused_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), stream,
                                                     C.byref(nfb), C.byref(nlb)), "count fused fetches")
    if sparse:
        pkg.check(L.fmgpu_count_fetches_sparse_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), stream,
                                                      C.byref(nfb), C.byref(nlb), C.byref(ntree)), "count sparse fetches")
    if wide:
        pkg.check(L.fmgpu_count_fetches_wide_device(index.handle, d_packed.data_ptr(), nq, READ_LEN, d_res.data_ptr(), stream,
                                                    C.byref(nfb), C.byref(nlb), C.byref(ntree)), "count wide fetches")
    tabled = fused or sparse or wide
    block_bytes = 32 * meta.wide_lanes if wide else 32 * (meta.sparse_lanes if sparse else meta.fused_lanes) if tabled else 16
    table_fetches = (nfb.value + ntree.value) if tabled else nblk.value
    sb96_fetches = nlb.value if tabled else 0
    all_fetches = table_fetches + sb96_fetches
    # bytes the timed kernel must move per launch: its block fetches (whole blocks; an SB96 block costs its 32-byte sector), the packed reads in, (L,R) out
    kernel_bytes = table_fetches * (block_bytes if tabled else 32) + sb96_fetches * 32 + nq * wpq * 4 + nq * 8

    # measured random-access ceiling over the footprint the timed kernel walks (rank 0, once)
    footprint = int(meta.wide_bytes) if wide else int(meta.sparse_bytes) if sparse else int(meta.fused_bytes) if fused else int(meta.nbytes)
    probe = pkg.gather_probe(dev, footprint, 256, 2) if rank == 0 else 0.0

    # the plain 2-step kernel on the same reads, for reference next to the timed one (rank-local, not the headline); its (L,R)
    # over the WHOLE batch are the full-size parity check of the timed kernel (the plain kernel itself is pinned on the reference below)
    plain, fused_extra, plain_res, sparse_extra = None, None, None, None
    if wide and os.environ.get("FM_BENCH_ALSO_SPARSE", "1") != "0":
        # the sparse-step kernel (round-2 headline until the wide-step table) on the same reads; its table is released again
        try:
            index.sparsify(); index.prepare(READ_LEN)
            sv = pkg.variant(pkg.MODE_SPARSE, 3)
            for _ in range(3):
                search_step(sv)
            se0, se1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            se0.record()
            for _ in range(5):
                search_step(sv)
            se1.record(); torch.cuda.synchronize()
            sm_ = index.meta
            sparse_extra = {"kernel": f"sparse: {sm_.sparse_bases} bases/step, {32 * sm_.sparse_lanes}-byte blocks", "table_gb": sm_.sparse_bytes / 1e9,
                            "ms_per_step": se0.elapsed_time(se1) / 5, "mqueries_per_s_per_gpu": nq / (se0.elapsed_time(se1) / 5) / 1e3}
            index.unsparsify()
        except pkg.FMError as ex:
            sparse_extra = {"unavailable": str(ex)}
    if (sparse or wide) and os.environ.get("FM_BENCH_ALSO_FUSED", "1") != "0":
        # the fused-step kernel (round-1 headline kernel) on the same reads; its 68 GB table is released again
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
    if tabled:
        for _ in range(3):
            search_step(plain_var)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(5):
            search_step(plain_var)
        pe1.record(); torch.cuda.synchronize()
        plain = {"kernel": "coop" if plain_var.mode == pkg.MODE_COOP else "task", "ms_per_step": pe0.elapsed_time(pe1) / 5,
                 "mqueries_per_s_per_gpu": nq / (pe0.elapsed_time(pe1) / 5) / 1e3}
        plain_res = d_res.clone()
    setup["setup_s"] = round(time.time() - t_setup, 2)

    # ------------------------------------------------------------------ timed: device-resident
    for _ in range(args.warmup):
        search_step()
    barrier()
    sampler = ClockSampler(dev)
    sampler.start()                                          # every rank samples its own GPU
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]      # per-step marks (recorded, never waited on inside the region)
    e0.record()
    for i in range(args.steps):
        search_step()
        marks[i].record()
    e1.record()
    barrier()
    ms_mine = e0.elapsed_time(e1) / args.steps
    ms_step = max_over_ranks(ms_mine)
    per_step = [(marks[i - 1] if i else e0).elapsed_time(marks[i]) for i in range(args.steps)]
    ms_best = max_over_ranks(min(per_step))                 # best single step (the reference reports best and mean of its iterations)
    clocks = sampler.stop()
    per_rank_ms = all_ranks(ms_mine)
    per_rank_mhz = all_ranks(clocks["sm_mhz"])
    res_dev = d_res.cpu().numpy().view(np.uint32).copy()
    whole_batch_equal = bool(torch.equal(d_res, plain_res)) if plain_res is not None else None
    whole_batch_md5 = hashlib.md5(memoryview(res_dev).cast("B")).hexdigest()
    del plain_res

    # ------------------------------------------------------------------ timed: end to end through the C ABI, host buffers
    h_res = torch.empty(2 * nq, dtype=torch.int32, pin_memory=True)
    handles = (C.c_void_p * 1)(index.handle)

    def e2e_step():
        pkg.check(L.fmgpu_search_host(handles, 1, h_ascii.data_ptr(), nq, READ_LEN, h_res.data_ptr(), C.byref(var)), "search_host")

    for _ in range(max(args.warmup, 6)):                   # (the auto feed measures both of its modes twice in its first large calls)
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()                                         # synchronous: returns when the (L,R) are in host memory
    torch.cuda.synchronize()
    e2e_ms_step = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    barrier()
    same = bool(np.array_equal(h_res.numpy().view(np.uint32), res_dev))
    # the host-side ceiling of any ASCII feed: host DRAM read bandwidth over the same pinned buffer (all ranks at once, like the feed)
    barrier()
    host_bw_mine = L.fm_host_read_bandwidth(h_ascii.data_ptr(), h_ascii.numel(), 0, 3)
    host_bw = sum(all_ranks(host_bw_mine))                   # what the whole box delivered to all ranks at the same time
    barrier()
    # ... and what the DMA engines pull out of that memory: the same buffer copied H2D by every rank at once, nothing else running
    d_land = torch.empty(min(h_ascii.numel(), 1 << 29), dtype=torch.uint8, device="cuda")
    d_land.copy_(h_ascii[: d_land.numel()], non_blocking=True)
    barrier()
    he0, he1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    he0.record()
    for a in range(0, h_ascii.numel() - d_land.numel() + 1, d_land.numel()):
        d_land.copy_(h_ascii[a:a + d_land.numel()], non_blocking=True)
    he1.record(); torch.cuda.synchronize()
    copied = (h_ascii.numel() // d_land.numel()) * d_land.numel()
    h2d_bw = sum(all_ranks(copied / (he0.elapsed_time(he1) * 1e-3) / 1e9))
    del d_land
    barrier()

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

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N=1) + parity on a STRIDED sample
    cpu = None
    parity = None
    index_md5_ok = None
    if rank == 0 and world == 1 and image is not None:
        ns = min(CPU_SAMPLE, nq)
        sel = np.arange(0, nq, max(1, nq // ns))[:ns]
        sample = h_ascii.numpy().reshape(nq, READ_LEN)[sel].reshape(-1)
        mq_cpu, sec_cpu, cores, out = reference_search_rate(pkg, image, sample, 2, 1)
        parity = bool(np.array_equal(out, res_dev.reshape(nq, 2)[sel].reshape(-1)))   # GPU (L,R) == reference CPU (L,R) on the sample
        want_md5 = golden_index_md5()
        index_md5_ok = (image_md5(image) == want_md5) if (want_md5 and os.environ.get("FM_BENCH_INDEX_MD5", "1") != "0") else None
        cpu = {"value": mq_cpu, "unit": "Mqueries/s", "cores": cores, "kind": CPU_KIND,
               "sample": f"{ns} of the {nq} reads, every {max(1, nq // ns)}th read of the batch, 2 timed passes after 1 warm-up, {cores} OpenMP threads, reference searchIndexCPU from oracle/_ref",
               "gpu_matches_reference_on_sample": parity}
        del image

    # ------------------------------------------------------------------ extra: the same kernels on a non-uniform text
    skewed = None
    if rank == 0 and world == 1 and (sparse or wide) and os.environ.get("FM_BENCH_SKEWED", "1") != "0":
        index.unsparsify(); index.unwiden()                  # make room; the timed table is not needed any more
        torch.cuda.empty_cache()
        try:
            skewed = skewed_text_extra(pkg, L, torch, dev, stream)
        except Exception as ex:                              # extra key only: never lose the headline line over it
            skewed = {"failed": repr(ex)}

    if rank == 0:
        peak, peak_src = measured_peak()
        mq = world * nq / ms_step / 1e3
        achieved = kernel_bytes / (ms_step * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic("wide" if wide else "sparse" if sparse else "fused" if fused else "plain")
        e2e_value = world * nq / e2e_ms_step / 1e3
        ceiling_mq = host_bw * 1e3 / READ_LEN if host_bw > 0 else None
        kernel_desc = (f"wide: {meta.wide_bases} bases/step, {32 * meta.wide_lanes}-byte blocks of {meta.wide_block_entries} "
                       f"{32 * meta.wide_entry_words}-bit entries (rest of the symbol + row), block = top {meta.wide_prefix_bits} bits of the "
                       f"wide symbol (computed from the read, shared by both interval ends), {READ_LEN - (READ_LEN // meta.wide_bases) * meta.wide_bases}-base lead table, "
                       f"{meta.wide_overflow} overfull buckets as search trees ({meta.wide_tree_nodes} blocks, depth {meta.wide_tree_depth}), {meta.wide_exceptional} exceptional buckets on plain steps, "
                       f"{meta.wide_lanes} x 256-bit loads, one state machine per read, qpt={var.queries_per_thread or 1}" if wide else
                       f"sparse: {meta.sparse_bases} bases/step, {32 * meta.sparse_lanes}-byte blocks of occurrence rows (lambda {meta.sparse_lambda}), {meta.sparse_lanes} x 256-bit loads, "
                       + (f"{meta.sparse_start_bases}-base start table + lead tables, " if meta.sparse_start_bases else "lead tables (no start table at this width), ")
                       + f"uniform grid of {meta.sparse_uniform_nb} blocks per symbol (no directory lookup), {meta.sparse_overflow} overfull buckets as search trees "
                       + f"({meta.sparse_tree_nodes} blocks, depth {meta.sparse_tree_depth}), one state machine per read, qpt={var.queries_per_thread}" if sparse else
                       f"fused: {meta.fused_bases} bases/step, {32 * meta.fused_lanes}-byte blocks, {meta.fused_lanes} x 256-bit loads, qpt={var.queries_per_thread}"
                       if fused else f"{MODE} qpt={var.queries_per_thread} tpb={var.threads_per_block}")
        line = {
            "metric": "Mqueries/s", "value": mq, "unit": "Mqueries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "ms_per_step_best": ms_best, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "lf_steps_per_s": world * lf_steps / (ms_step * 1e-3),
            "config": shared_config(world),
            "kernel_config": {"kernel": kernel_desc,
                              "device_layout": (f"wide-step table {meta.wide_bytes / 1e9:.1f} GB ({meta.wide_blocks} blocks) "
                                                f"built on the GPU from the 2-step SB96 table ({meta.nbytes / 1e9:.2f} GB)" if wide else
                                                f"sparse-step table {meta.sparse_bytes / 1e9:.1f} GB ({meta.sparse_blocks} blocks) "
                                                f"built on the GPU from the 2-step SB96 table ({meta.nbytes / 1e9:.2f} GB)" if sparse else
                                                f"fused-step table {meta.fused_bytes / 1e9:.1f} GB composed on the GPU from the 2-step SB96 table ({meta.nbytes / 1e9:.2f} GB)"
                                                if fused else "SB96 (16-byte per-symbol blocks: u32 rank + 96 indicator bits)"),
                              "footprint_gb": footprint / 1e9, "derived_tables_gb": meta.derived_bytes / 1e9, "setup": setup},
            "per_rank": {"ms_per_step": per_rank_ms, "sm_mhz": per_rank_mhz,
                         "note": "device-timed ms per step and median SM clock of every rank during the timed region; value uses the max"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": kernel_bytes,
                         "algorithmic_model": (f"bytes the timed kernel must move: {table_fetches} table-block fetches x {block_bytes} B + {sb96_fetches} SB96 fetches x 32 B "
                                               f"(counted by the instrumented instantiation of the same kernel) + {nq * wpq * 4} B packed reads in + {nq * 8} B (L,R) out"),
                         "traffic_over_algorithmic": (traffic / kernel_bytes) if traffic else None,
                         "request_rate_frac": (all_fetches / (ms_step * 1e-3)) / probe if probe else None,
                         "request_rate": {"block_fetches_per_s": all_fetches / (ms_step * 1e-3), "probe_accesses_per_s": probe,
                                          "how": "independent uniform random 16-byte loads over a table of the same footprint; the binding limit is a request RATE "
                                                 "(L2 miss path, profiles/r02_ceiling_counters.md), the same for 64- and 128-byte fills -- which is why frac of the byte "
                                                 "roofline stays near one half although the kernel wastes nothing (traffic_over_algorithmic ~ 1)"},
                         "vs_reference_algorithm_bytes": {"bytes_per_launch": ref_algo_bytes, "achieved": ref_algo_bytes / (ms_step * 1e-3) / 1e9,
                                                          "frac": ref_algo_bytes / (ms_step * 1e-3) / 1e9 / peak,
                                                          "sectors_per_lf_step": nsec.value / lf_steps, "blocks_per_lf_step": nblk.value / lf_steps,
                                                          "model": "SURVEY 8(d) yardstick: 32 B x exact count of sectors the reference's 2-step search must touch; "
                                                                   "exceeds 1 because this kernel consumes up to 30 bases per fetch where that algorithm consumes 2"},
                         "table_blocks_per_launch": table_fetches, "sb96_blocks_per_launch": sb96_fetches,
                         "tree_blocks_per_launch": ntree.value if (sparse or wide) else None,
                         "frac_of_nominal_8tbs": achieved / 8000.0},
            "plain_2step_kernel": plain,
            "sparse_14base_kernel": sparse_extra,
            "fused_4base_kernel": fused_extra,
            "skewed_text": skewed,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mqueries/s", "ms_per_step": e2e_ms_step,
                    "h2d_bytes_per_step": nq * READ_LEN, "d2h_bytes_per_step": nq * 8, "matches_device_resident_result": same,
                    "feed": os.environ.get("FMGPU_FEED", "auto (self-tuning during warm-up: hybrid of ASCII-over-PCIe and AVX-512 host packing, or ASCII only)"),
                    "host_pack_threads_per_rank": int(L.fm_hostpack_threads()),
                    "host_ceiling": {"host_read_gbs": host_bw, "bytes_per_read": READ_LEN, "mqueries_per_s": ceiling_mq,
                                     "e2e_over_ceiling": (e2e_value / ceiling_mq) if ceiling_mq else None,
                                     "pinned_h2d_gbs_all_ranks": h2d_bw,
                                     "how": "fm_host_read_bandwidth over the pinned ASCII buffer, all ranks at once, all host threads: every read costs its 100 bytes of host DRAM "
                                            "reads whoever fetches them (DMA engine or packer thread), so this is the bound of any feed of ASCII reads on this host"},
                    "note": "h2d_bytes_per_step counts the ASCII reads handed to the call; host-packed chunks cross PCIe as 2-bit (25 B/read)"},
            "e2e_packed_input": {"value": world * nq / e2e_packed_ms_step / 1e3, "unit": "Mqueries/s", "ms_per_step": e2e_packed_ms_step,
                                 "h2d_bytes_per_step": nq * wpq * 4, "d2h_bytes_per_step": nq * 8, "matches_device_resident_result": same_packed,
                                 "note": "extra, not the headline: fmgpu_search_host_packed, host reads already in the 2-bit binary format (28 B per 100-bp read)"},
            "locate": locate,
            "gpu_launches": args.steps,
            "clocks": clocks,
            "checks": {"every_read_found": hits_ok, "e2e_equals_resident": same,
                       "timed_kernel_equals_plain_kernel_on_all_reads": whole_batch_equal, "results_md5_rank0": whole_batch_md5,
                       "gpu_equals_reference_cpu_on_strided_sample": parity, "index_is_the_reference_builders_file": index_md5_ok},
        }
        emit(line)
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

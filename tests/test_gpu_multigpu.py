"""SURVEY.md 8 row a-9 on PHYSICAL GPUs: the single-process host driver of the reference flow
(common/searchQueries.c:64-125: transferCPUtoGPU -> 5 x searchIndexGPU -> transferGPUtoCPU -> saveResults) with the index
replicated over every visible device by peer copies and the batch sharded over them.

Every test uses ALL visible sm_100 devices (FMGPU_DEVICES=0..n-1); on a one-GPU box the same code runs with one
replica and the assertions about peer copies are skipped.  The checker is the UNMODIFIED reference: its own main()
linked against the product library (oracle/_ref/fmIndexSearchGPU_refmain) produces "<index>.res.gpu", its own CPU
searcher binary produces "<index>.res.cpu" from the same files, and the two text files must be identical.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

REFMAIN = os.path.join(helpers.REF_DIR, "fmIndexSearchGPU_refmain")


@pytest.fixture(scope="module")
def pkg():
    helpers.ensure_built()
    p = helpers.pkg()
    if p.lib().fmgpu_device_count() < 1:
        pytest.fail("no sm_100 GPU visible: GPU tests cannot run (there is no CPU fallback)")
    return p


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    """3 Mbp text, reference-built 2-step index in all four layouts, 200 000 reads (exact, plus random ones that mostly miss)."""
    if not helpers.has_ref_tools():
        pytest.skip("oracle/_ref tools are not built")
    d = str(tmp_path_factory.mktemp("multigpu"))
    n, length, nq = 3_000_017, 100, 200_000
    text = helpers.synth_text(n, seed=77)
    paths = helpers.build_reference_indexes(d, text, 2, 64)
    rng = np.random.default_rng(5)
    reads = np.concatenate([helpers.synth_reads(text, 3, nq - 5000, length),
                            np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 5000 * length)]])
    qfa = os.path.join(d, "reads.fa")
    helpers.write_fasta_reads(qfa, reads, length)
    return {"dir": d, "paths": paths, "qfa": qfa, "length": length, "nq": nq, "reads": reads}


def _file_md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()


def _devices_env(pkg):
    ndev = pkg.lib().fmgpu_device_count()
    return ndev, ",".join(str(i) for i in range(ndev))


@pytest.mark.parametrize("mode", ["auto", "wide", "sparse", "fused", "task"])
def test_reference_main_linked_against_the_library_on_all_gpus(pkg, dataset, mode, tmp_path):
    """common/searchQueries.c, unmodified, -DCUDA, linked against libfmindex_b200.so: its .res.gpu equals the .res.cpu
    the reference's own CPU searcher writes for the same index and reads (std and AltCounters layouts)."""
    if not os.path.exists(REFMAIN):
        pytest.skip("oracle/_ref/fmIndexSearchGPU_refmain is not built")
    ndev, devs = _devices_env(pkg)
    stats_file = str(tmp_path / "stats.jsonl")
    # (layout searched on the GPUs, layout the reference CPU binary accepts, that binary): tags 200 and 201 hold the same index
    for tag, cpu_tag, cpu_bin in ((100, 100, "fmIndexSearchCPU_64bases_2step"), (201, 200, "fmIndexSearchCPU_64bases_2step-ac")):
        fn, cpu_fn = dataset["paths"][tag], dataset["paths"][cpu_tag]
        for f in (fn + ".res.gpu", cpu_fn + ".res.cpu"):
            if os.path.exists(f):
                os.remove(f)
        env = dict(os.environ, FMGPU_DEVICES=devs, FMGPU_MODE=mode, FMGPU_STATS_FILE=stats_file)
        p = subprocess.run([REFMAIN, fn, dataset["qfa"], str(dataset["length"]), str(dataset["nq"])], env=env, capture_output=True, text=True)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        assert "TIME:" in p.stdout
        helpers.run([os.path.join(helpers.REF_DIR, cpu_bin), cpu_fn, dataset["qfa"], str(dataset["length"]), str(dataset["nq"])])
        assert _file_md5(fn + ".res.gpu") == _file_md5(cpu_fn + ".res.cpu"), f"tag {tag} mode {mode} on {ndev} GPU(s)"
    import json
    lines = [json.loads(x) for x in open(stats_file)]
    assert len(lines) == 2 and all(s["ndev"] == ndev and s["searches"] == 5 for s in lines)
    assert all(len(s["search_ms_per_gpu"]) == ndev and min(s["search_ms_per_gpu"]) > 0 for s in lines)
    if ndev > 1:
        assert all(len(s["peer_copy_s"]) == ndev - 1 and min(s["peer_copy_s"]) > 0 for s in lines)


def test_driver_binary_and_library_flow_agree_on_all_gpus(pkg, dataset):
    """bin/fmIndexSearchGPU_b200 (our rebuild of that main) and the in-process flow give the same (L,R); the transfer
    statistics name every device, every peer copy and every GPU's kernel time."""
    ndev, devs = _devices_env(pkg)
    fn = dataset["paths"][100]
    exe = os.path.join(helpers.ROOT, helpers.PKG_NAME, "bin", "fmIndexSearchGPU_b200")
    env = dict(os.environ, FMGPU_DEVICES=devs, FMGPU_MODE="sparse")
    p = subprocess.run([exe, fn, dataset["qfa"], str(dataset["length"]), str(dataset["nq"])], env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    md5_bin = _file_md5(fn + ".res.gpu")
    os.environ["FMGPU_MODE"] = "sparse"
    try:
        got = pkg.search_files(fn, dataset["qfa"], dataset["length"], dataset["nq"], devices=list(range(ndev)))
    finally:
        del os.environ["FMGPU_MODE"]
        pkg.lib().fmgpu_set_devices(None, 0)
    assert helpers.results_text_md5(got) == md5_bin
    st = pkg.fmgpu_transfer_stats_t()
    assert pkg.lib().fmgpu_get_transfer_stats(C.byref(st)) == 0
    assert st.ndev == ndev and st.searches == 1 and st.index_h2d_reblock_s > 0 and st.queries_h2d_pack_s > 0 and st.results_d2h_s > 0
    assert all(st.search_ms[g] > 0 for g in range(ndev)) and all(st.table_build_s[g] > 0 for g in range(ndev))
    assert all(st.peer_copy_s[g] > 0 for g in range(1, ndev))
    oracle = helpers.Oracle()
    oh = oracle.wrap(np.fromfile(fn, dtype=np.uint32))
    assert np.array_equal(got, oracle.search(oh, dataset["reads"], dataset["length"]))
    oracle.free(oh)


def test_replicas_are_bit_identical_and_every_gpu_answers_alone(pkg, dataset):
    """fmgpu_index_replicate (cudaMemcpyPeer): every replica's block table equals the first one's, and each device alone
    returns the reference result for the whole batch."""
    import torch
    ndev, _ = _devices_env(pkg)
    image = np.fromfile(dataset["paths"][100], dtype=np.uint32)
    first = pkg.DeviceIndex.from_image(image, device=0)
    want = None
    reps = [first] + [first.replicate(g) for g in range(1, ndev)]
    ref_table = torch.as_tensor(first, device="cuda:0").cpu()
    for g, rep in enumerate(reps):
        assert rep.device == g
        assert torch.equal(torch.as_tensor(rep, device=f"cuda:{g}").cpu(), ref_table), f"replica on GPU {g} differs"
        rep.sparsify(8, 0, 0)
        b = pkg.DeviceBatch(g, dataset["nq"], dataset["length"], 2)
        b.upload_ascii(dataset["reads"])
        b.search(rep, pkg.variant(pkg.MODE_SPARSE, 4))
        got = b.download()
        if want is None:
            oracle = helpers.Oracle()
            oh = oracle.wrap(image)
            want = oracle.search(oh, dataset["reads"], dataset["length"])
            oracle.free(oh)
        assert np.array_equal(got, want), f"GPU {g}"
        rep.widen(rep.wide_bases_for(dataset["length"]))
        b.search(rep, pkg.variant(pkg.MODE_WIDE))
        assert np.array_equal(b.download(), want), f"GPU {g}, wide-step table"
        b.free()
    # end to end over all replicas at once: chunks go round-robin over the GPUs
    got = pkg.search_host(reps, dataset["reads"], dataset["length"], pkg.variant(pkg.MODE_SPARSE, 4))
    assert np.array_equal(got, want)
    got = pkg.search_host(reps, dataset["reads"], dataset["length"], pkg.variant(pkg.MODE_WIDE))
    assert np.array_equal(got, want)
    for rep in reps:
        rep.free()
    pkg.lib().fmgpu_release_pipeline()


def test_retransfer_with_mixed_query_and_result_handles(pkg, dataset):
    """transferCPUtoGPU called again with the same queries and ANOTHER results handle, then with new queries and the same
    results handle: no handle keeps a pointer to a released shard set, nothing leaks, every download is right
    (shard-set lifetime tied to one pair)."""
    L = pkg.lib()
    ndev, _ = _devices_env(pkg)
    devs = (C.c_int32 * ndev)(*range(ndev))
    assert L.fmgpu_set_devices(devs, ndev) == 0
    L.fmgpu_set_variant(None)
    length, nq = dataset["length"], 20_000
    idx = pkg.loadIndex(dataset["paths"][100])
    q1 = pkg.queriesFromArray(dataset["reads"][: nq * length], length)
    q2 = pkg.queriesFromArray(dataset["reads"][nq * length: 2 * nq * length], length)
    r1, r2 = pkg.initResults(nq), pkg.initResults(nq)
    oracle = helpers.Oracle()
    oh = oracle.wrap(np.fromfile(dataset["paths"][100], dtype=np.uint32))
    want1 = oracle.search(oh, dataset["reads"][: nq * length], length)
    want2 = oracle.search(oh, dataset["reads"][nq * length: 2 * nq * length], length)
    try:
        def go(q, r):
            pkg.check(L.transferCPUtoGPU(idx, C.byref(q), r), "transferCPUtoGPU")
            pkg.check(L.fmgpu_search_index(idx, C.byref(q), r), "fmgpu_search_index")
            pkg.check(L.transferGPUtoCPU(r), "transferGPUtoCPU")
            return pkg.resultsArray(r)
        assert np.array_equal(go(q1, r1), want1)
        assert np.array_equal(go(q1, r2), want1)                       # same queries, other results: r1 must be detached
        assert C.cast(r1, C.POINTER(pkg.res_t)).contents.d_results is None
        assert L.transferGPUtoCPU(r1) == pkg.FM_E_BAD_ARGUMENT          # ... and say so instead of touching freed memory
        assert np.array_equal(go(q2, r2), want2)                       # new queries, same results: q1 must be detached
        assert q1.d_queries is None
        assert L.fmgpu_search_index(idx, C.byref(q1), r2) == pkg.FM_E_BAD_ARGUMENT
        assert np.array_equal(go(q1, r1), want1)
        assert np.array_equal(pkg.resultsArray(r2), want2)             # r2's host results are untouched by the other pair
    finally:
        for q in (q1, q2):
            qp = C.c_void_p(C.addressof(q))
            L.freeQueriesGPU(C.byref(qp))
        for r in (r1, r2):
            L.freeResultsGPU(C.byref(r))
        L.freeIndexGPU(C.byref(idx)); L.freeIndex(C.byref(idx))
        for r in (r1, r2):
            L.freeResults(C.byref(r))
        L.fmgpu_set_devices(None, 0)
        oracle.free(oh)

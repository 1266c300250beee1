"""GPU parity tests of the sparse-step layout (csrc/fm_sparse.cuh): up to 14 bases per block fetch, a uniform grid of
blocks of occurrence rows, overfull buckets as search trees of blocks, one state machine per read.  Everything
goes through the C ABI (ctypes) and is compared bit for bit with committed reference outputs, the oracle, the
reference searcher itself, or the plain kernels on the same index at sizes the CPU cannot reach.   pytest -m gpu"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
GOLDEN = sorted(os.path.join(helpers.ROOT, "tests", "golden", f) for f in os.listdir(os.path.join(helpers.ROOT, "tests", "golden")) if f.endswith(".npz"))
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="module")
def pkg(built):
    p = helpers.pkg()
    assert p.lib().fmgpu_device_count() >= 1, "no sm_100 GPU: the product has no CPU fallback"
    return p


def widths(k):
    return [2, 3, 5, 6, 10, 12] if k == 1 else [4, 6, 10, 12]          # (14 bases: test_sparse_steps_read_lengths and the full-size test)


@pytest.mark.parametrize("dynamic", ["0", "1"], ids=["static_assignment", "dynamic_assignment"])
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p))
def test_sparse_steps_golden_all_widths(pkg, monkeypatch, path, dynamic):
    """Committed outputs of the unmodified reference searchers, the AltCounters quirk fixtures included (phantom
    occurrences); every width, both block sizes, every qpt, lambda from 1 (big grid, no trees) to the slot count (many
    overfull buckets -> search trees)."""
    monkeypatch.setenv("FMGPU_SPARSE_DYNAMIC", dynamic)         # reads handed to lane groups statically / from a queue (plans of sparse steps only)
    g = np.load(path)
    reads, length, k = g["reads"], int(g["length"]), int(g["k"])
    nq = reads.size // length
    quirk = "quirk" in os.path.basename(path)
    b = pkg.DeviceBatch(0, nq, length, k)
    b.upload_ascii(reads)
    for tag, key in ((100, "expected_std"), (101, "expected_std"), (200, "expected_ac"), (201, "expected_ac")):
        for ks in ([w for w in widths(k) if w <= 10] if quirk else widths(k)):
            for lanes, lams in ((2, (1, 5, 15)), (4, (12, 31))):
                for lam in lams:
                    idx = pkg.DeviceIndex.from_image(g[f"image_{tag}"]).sparsify(ks, lam, lanes)
                    m = idx.meta
                    assert (m.sparse_bases, m.sparse_lambda, m.sparse_lanes) == (ks, lam, lanes)
                    assert m.sparse_uniform_nb >= 1 and m.sparse_blocks == m.sparse_uniform_nb * 4 ** ks + m.sparse_tree_nodes
                    assert (m.sparse_overflow > 0) == (m.sparse_tree_nodes > 0) == (m.sparse_tree_depth > 0)
                    assert m.sparse_bytes == m.sparse_blocks * 32 * lanes + (8 * 4 ** m.sparse_start_bases if m.sparse_start_bases else 0)
                    assert m.derived_bytes >= m.sparse_bytes
                    for qpt in (1, 2, 3, 4):
                        b.search(idx, pkg.variant(pkg.MODE_SPARSE, qpt))
                        assert np.array_equal(b.download(), g[key]), f"tag {tag} ks {ks} lanes {lanes} lambda {lam} qpt {qpt}"
                    idx.free()
    b.free()


@pytest.mark.parametrize("k,length", [(1, 1), (1, 3), (1, 5), (1, 10), (1, 17), (1, 33), (1, 100), (2, 2), (2, 6), (2, 10), (2, 20), (2, 30), (2, 34),
                                      (2, 100), (2, 126), (2, 128), (2, 250), (1, 251), (2, 25), (2, 101)])
def test_sparse_steps_read_lengths(pkg, k, length):
    """Lengths that are not a multiple of the sparse width run their leading steps on the SB96 table, odd lengths on a
    2-step index end with the derived 1-step rank; bit fields of the packed read straddle 32-bit words."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", f"small_k{k}_d64.npz"))
    text = helpers.synth_text(int(g["n"]), seed=7 + k)
    reads = np.concatenate([helpers.synth_reads(text, 21, 700, length), ACGT[np.random.default_rng(length).integers(0, 4, 68 * length)]])
    idx = pkg.DeviceIndex.from_image(g["image_101"])
    b = pkg.DeviceBatch(0, reads.size // length, length, k)
    b.upload_ascii(reads)
    if length % k == 0:
        o = helpers.Oracle()
        h = o.wrap(g["image_101"])
        want = o.search(h, reads, length)
        o.free(h)
    else:                                                        # defined by the plain kernels (tested against the 1-step reference elsewhere)
        b.search(idx, pkg.variant(pkg.MODE_COOP))
        want = b.download()
    for ks in ([3, 10, 14] if k == 1 else [4, 10, 14]):
        for lanes in ((2, 4) if ks < 14 else (2,)):
            idx.sparsify(ks, 0, lanes)
            b.search(idx, pkg.variant(pkg.MODE_SPARSE))
            assert np.array_equal(b.download(), want), f"k={k} len={length} ks={ks} lanes={lanes}"
            idx.unsparsify()
    b.free(); idx.free()


def test_sparse_unavailable_and_errors(pkg):
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "quirk_k2_n124.npz"))
    idx = pkg.DeviceIndex.from_image(g["image_200"])           # (AltCounters padding quirk: served since round 2)
    b = pkg.DeviceBatch(0, 4, 8, 2)
    b.upload_ascii(g["reads"][:32])
    with pytest.raises(pkg.FMError) as ei:
        b.search(idx, pkg.variant(pkg.MODE_SPARSE))            # no table: loud failure, no silent fallback
    assert ei.value.code == pkg.FM_E_BAD_ARGUMENT
    b.free(); idx.free()
    idx = pkg.DeviceIndex.from_image(g["image_100"])           # same text, standard file: fine
    for bad in ((3, 0, 0), (2, 0, 0), (16, 0, 0), (4, 16, 2), (4, 32, 4), (4, 0, 3)):
        with pytest.raises(pkg.FMError) as ei:
            idx.sparsify(*bad)
        assert ei.value.code == pkg.FM_E_BAD_ARGUMENT, bad
    idx.sparsify(4, 0, 2)
    assert idx.meta.sparse_bases == 4
    b = pkg.DeviceBatch(0, g["reads"].size // 8, 8, 2)
    b.upload_ascii(g["reads"])
    b.search(idx, pkg.variant(pkg.MODE_SPARSE))
    assert np.array_equal(b.download(), g["expected_std"])
    idx.unsparsify()
    assert idx.meta.sparse_bases == 0 and idx.meta.sparse_bytes == 0
    b.free(); idx.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["polyA", "ACGT_period4", "two_letter", "repeat_x40", "random_plus_repeat"])
@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("dynamic", ["0", "1"], ids=["static_assignment", "dynamic_assignment"])
def test_sparse_steps_repetitive_texts_search_trees(pkg, tmp_path, monkeypatch, name, k, dynamic):
    """Repeats put hundreds of occurrences of one wide symbol into consecutive BWT rows: those buckets are overfull and
    become search trees of blocks (several levels deep for poly-A), which the kernel walks one fetch per iteration.
    Index files and expected (L,R) from the unmodified reference tools."""
    monkeypatch.setenv("FMGPU_SPARSE_DYNAMIC", dynamic)
    rng = np.random.default_rng(11)
    n = 30011
    unit = ACGT[rng.integers(0, 4, 700)]
    text = {"polyA": np.full(n, ord("A"), dtype=np.uint8),
            "ACGT_period4": np.tile(ACGT, n // 4 + 1)[:n],
            "two_letter": np.frombuffer(b"AC", dtype=np.uint8)[rng.integers(0, 2, n)],
            "repeat_x40": np.tile(unit, n // 700 + 1)[:n],
            "random_plus_repeat": np.concatenate([ACGT[rng.integers(0, 4, n // 2)], np.tile(unit, n // 1400 + 1)])[:n]}[name].copy()
    d = 64
    if (n + 1) % d == 0:
        text = text[:-1]
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, d)
    length = 20
    starts = rng.integers(0, text.size - length + 1, 2000)
    reads = np.concatenate([text[s:s + length] for s in starts] + [ACGT[rng.integers(0, 4, 500 * length)]])
    ref = helpers.RefSearcher(k, d, False)
    want, _ = ref.search(ref.load(paths[100]), reads, length)
    image = np.fromfile(paths[101], dtype=np.uint32)
    idx = pkg.DeviceIndex.from_image(image)
    b = pkg.DeviceBatch(0, reads.size // length, length, k)
    b.upload_ascii(reads)
    saw_overflow, deepest = False, 0
    for ks in ([5, 10] if k == 1 else [4, 10]):
        for lanes in (2, 4):
            idx.sparsify(ks, 0, lanes)
            m = idx.meta
            saw_overflow |= m.sparse_overflow > 0
            deepest = max(deepest, m.sparse_tree_depth)
            assert m.sparse_tree_rows <= m.bwtsize and (m.sparse_tree_nodes > 0) == (m.sparse_overflow > 0)
            for qpt in (1, 4):
                b.search(idx, pkg.variant(pkg.MODE_SPARSE, qpt))
                assert np.array_equal(b.download(), want), f"{name} k={k} ks={ks} lanes={lanes} qpt={qpt}"
            idx.unsparsify()
    assert saw_overflow or name == "two_letter", "these texts are meant to overflow buckets"
    if name == "polyA":
        assert deepest >= 2                                      # 30 000 occurrences of one symbol: 15 * 15 * 15 < 30 000
    b.free(); idx.free()


@pytest.mark.parametrize("k", [1, 2])
def test_sparse_start_table_and_fetch_counter(pkg, k):
    """The sparse kernel's start table ((L,R) of every 10-mer, computed by the kernel itself) must not change any
    result: exact, mutated and random reads of several lengths against the plain Coop kernel on the same index; the
    instrumented kernel counts one block fetch per sparse step once L and R share a bucket."""
    import torch
    n = 20_000_003
    os.environ["FMGPU_START_TABLE"] = "1"
    try:
        b = pkg.IndexBuild.from_synth(n, 3, k, 64)
        idx = b.to_index().sparsify(10, 0, 0)
        b.free()
    finally:
        del os.environ["FMGPU_START_TABLE"]
    m = idx.meta
    assert (m.sparse_bases, m.sparse_lanes, m.sparse_lambda, m.sparse_start_bases) == (10, 2, 5, 10)
    L = pkg.lib()
    rng = np.random.default_rng(5)
    stream = torch.cuda.current_stream().cuda_stream
    # 36..39, 76, 77, 126: 6 or more leftover bases start from their lead table; 6, 8: shorter than one sparse step
    # and odd widths / 11 bases (lengths = 1 mod 10) come out of a lead table without any tail fetch
    # lengths = 2 .. 5 mod 10 with two sparse steps or more build the wide tables (12 .. 13 bases here) on first use
    for length in (10, 20, 100, 12, 14, 16, 50, 101, 99, 36, 37, 38, 39, 76, 77, 126, 6, 8, 11, 21, 31, 151, 75, 32, 33, 22, 23, 15):
        nq = 200_000
        d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
        pkg.check(L.fmgpu_synth_reads_device(0, n, 3, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
        torch.cuda.synchronize()
        reads = d_ascii.cpu().numpy().copy()
        mut = rng.integers(0, nq * length, nq // 3)
        reads[mut] = ACGT[rng.integers(0, 4, mut.size)]
        batch = pkg.DeviceBatch(0, nq, length, k)
        batch.upload_ascii(reads)
        batch.search(idx, pkg.variant(pkg.MODE_COOP))
        want = batch.download()
        for qpt in (1, 2, 3, 4):
            batch.search(idx, pkg.variant(pkg.MODE_SPARSE, qpt))
            assert np.array_equal(batch.download(), want), f"k={k} len={length} qpt={qpt}"
        batch.free()
        if length == 100:
            d_ascii.copy_(torch.from_numpy(reads))
            wpq = L.fmgpu_words_per_query(length)
            d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
            d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
            pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack")
            a, s, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
            pkg.check(L.fmgpu_count_fetches_sparse_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream,
                                                          C.byref(a), C.byref(s), C.byref(o)), "count")
            assert np.array_equal(d_res.cpu().numpy().view(np.uint32), want)
            assert 9 * nq <= a.value <= 9.05 * nq                # 10-base start table + 9 sparse steps, L and R in one bucket
            assert s.value == 0 and o.value <= 0.05 * nq         # no SB96 steps at this length; overfull buckets (tree fetches) are rare on a random text
    idx.free()


def test_sparse_automatic_width_follows_the_text_length(pkg):
    """Default width: the widest multiple of k (up to 14 bases) that leaves lambda rows per wide symbol on average, so the
    grid costs ~12.8 bytes per base whatever the text length -- and whatever the text: there is no "uneven counts" case
    any more.  Same (L,R) as the plain kernel."""
    import torch
    for n, k, expect in ((4_000_000, 2, 8), (4_000_000, 1, 9), (90_000_000, 2, 12), (30_000, 2, 6)):
        build = pkg.IndexBuild.from_synth(n, 1, k, 64)
        idx = build.to_index()
        build.free()
        d_ascii = torch.empty(60_000 * 50, dtype=torch.uint8, device="cuda")
        pkg.check(pkg.lib().fmgpu_synth_reads_device(0, n, 1, 60_000, 50, 2, 0, d_ascii.data_ptr(), None), "reads")
        torch.cuda.synchronize()
        reads = np.concatenate([d_ascii.cpu().numpy(), ACGT[np.random.default_rng(3).integers(0, 4, 40_000 * 50)]])
        b = pkg.DeviceBatch(0, reads.size // 50, 50, k)
        b.upload_ascii(reads)
        b.search(idx, pkg.variant(pkg.MODE_COOP))
        want = b.download()
        assert (want[1:120_000:2] > want[0:120_000:2]).all()     # the exact reads are found
        idx.sparsify()
        m = idx.meta
        assert m.sparse_bases == expect, (n, k, m.sparse_bases)
        assert m.sparse_bytes < 26 * n + (1 << 20), "grid + trees: 12.8 .. 25.6 bytes per base on a random text (blocks per symbol round up)"
        assert m.sparse_tree_rows < n // 100
        for qpt in (1, 4):
            b.search(idx, pkg.variant(pkg.MODE_SPARSE, qpt))
            assert np.array_equal(b.download(), want), (n, k, qpt)
        b.free(); idx.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
def test_sparse_quirk_fuzz_against_the_altcounters_reference(pkg, tmp_path):
    """AltCounters files whose padding-entry quirk is ACTIVE (a '$' row in the last chunk whose counter
    lives in the padding entry) on the sparse-step path: the composed rank function then has phantom occurrences, which
    the table stores as repeated list entries.  Tiny references make the quirk frequent; the checker is the reference's
    own AltCounters searcher, which has the quirk by construction."""
    rng = np.random.default_rng(77)
    active = 0
    for case in range(48):
        k = 1 + case % 2
        d = (32, 64)[(case // 2) % 2]
        n = int(rng.integers(d + 2, 5 * d))
        if (n + 1) % d == 0:
            n += 1
        alphabet = (b"ACGT", b"AC", b"AAAAAACGT")[(case // 4) % 3]
        text = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)]
        paths = helpers.build_reference_indexes(str(tmp_path / f"c{case}"), text, k, d)
        ref = helpers.RefSearcher(k, d, True)
        for length in (2 * k, 4 * k, 6 * k, 12, 18):
            starts = rng.integers(0, n - length + 1, 200)
            reads = np.concatenate([text[s:s + length] for s in starts] + [text[:length], text[-length:],
                                   np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), 150 * length)]])
            want, _ = ref.search(ref.load(paths[200]), reads, length)
            batch = pkg.DeviceBatch(0, reads.size // length, length, k)
            batch.upload_ascii(reads)
            for tag in (200, 201):
                idx = pkg.DeviceIndex.from_image(np.fromfile(paths[tag], dtype=np.uint32))
                if idx.meta.quirk_mask == 0:
                    idx.free()
                    continue
                active += 1
                for kf, fl in (((2, 1), (3, 2), (4, 2), (4, 4)) if k == 1 else ((4, 1), (4, 2), (4, 4))):
                    idx.fuse(kf, fl)                                  # the fused-step table carries the quirk as a phantom list
                    for qpt in (1, 2):
                        batch.search(idx, pkg.variant(pkg.MODE_FUSED, qpt))
                        assert np.array_equal(batch.download(), want), f"case {case}: k={k} d={d} n={n} len={length} tag={tag} fused kf={kf} lanes={fl}"
                    idx.unfuse()
                for ks in ((2, 3, 4, 6) if k == 1 else (4, 6, 8)):
                    for lanes, lam in ((2, 0), (2, 1), (4, 0)):
                        idx.sparsify(ks, lam, lanes)
                        for qpt in (1, 4):
                            batch.search(idx, pkg.variant(pkg.MODE_SPARSE, qpt))
                            assert np.array_equal(batch.download(), want), f"case {case}: k={k} d={d} n={n} len={length} tag={tag} ks={ks} lanes={lanes} lam={lam}"
                        idx.unsparsify()
                idx.free()
            batch.free()
    assert active >= 20, f"only {active} (case, length, tag) combinations had an active quirk: the fuzz does not exercise it"


def test_sparse_config3_full_size_against_reference_checksums(pkg):
    """BASELINE config 3 at FULL size (2 Gbp, k=2, d=64): (L,R) of the first 1 M reads from the sparse-step kernel,
    default table (10 bases per step, 64-byte blocks, start table), must have the md5 recorded from the UNMODIFIED
    reference searcher (tests/golden/config3_2g.json); tags 100 and 201."""
    import torch
    gold = json.load(open(os.path.join(helpers.ROOT, "tests", "golden", "config3_2g.json")))
    n, k, d = gold["text"]["n"], gold["k"], gold["d"]
    b = pkg.IndexBuild.from_synth(n, gold["text"]["seed"], k, d)
    nq, length = gold["reads"]["num"], gold["reads"]["len"]
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(pkg.lib().fmgpu_synth_reads_device(0, n, gold["text"]["seed"], nq, length, gold["reads"]["seed"], 0, d_ascii.data_ptr(), None), "reads")
    torch.cuda.synchronize()
    reads = d_ascii.cpu().numpy()
    del d_ascii
    batch = pkg.DeviceBatch(0, nq, length, k)
    batch.upload_ascii(reads)
    for tag, key in ((100, "res_cpu_std_text"), (201, "res_cpu_ac_text")):
        t = b if tag == 100 else b.transform(tag)
        idx = t.to_index().sparsify()
        m = idx.meta                                             # automatic choice on this text: 14 bases per step, uniform grid
        assert (m.sparse_bases, m.sparse_lanes, m.sparse_start_bases) == (14, 2, 0) and m.sparse_uniform_nb == 2 and m.sparse_bytes < 45e9
        for qpt in (2, 4):
            batch.search(idx, pkg.variant(pkg.MODE_SPARSE, qpt))
            assert helpers.results_text_md5(batch.download()) == gold["md5"][key], f"tag {tag} qpt {qpt}"
        idx.unsparsify()
        idx.sparsify(12, 0, 0)
        m = idx.meta
        assert (m.sparse_bases, m.sparse_start_bases) == (12, 12) and m.sparse_uniform_nb > 0
        batch.search(idx, pkg.variant(pkg.MODE_SPARSE, 4))
        assert helpers.results_text_md5(batch.download()) == gold["md5"][key], f"tag {tag} 12 bases"
        idx.unsparsify()
        idx.sparsify(10, 0, 0)                                   # and the 10-base table
        m = idx.meta
        assert (m.sparse_bases, m.sparse_lanes, m.sparse_start_bases) == (10, 2, 10) and m.sparse_bytes < 30e9
        batch.search(idx, pkg.variant(pkg.MODE_SPARSE, 4))
        assert helpers.results_text_md5(batch.download()) == gold["md5"][key], f"tag {tag} 10 bases"
        idx.free()
        if tag != 100:
            t.free()
    batch.free(); b.free()


def test_sparse_end_to_end_host_calls(pkg):
    """fmgpu_search_host / fmgpu_search_host_packed with the sparse-step variant: host buffers in, host (L,R) out."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    reads, length = g["reads"], int(g["length"])
    idx = pkg.DeviceIndex.from_image(g["image_100"]).sparsify()
    got = pkg.search_host([idx], reads, length, var=pkg.variant(pkg.MODE_SPARSE, 4))
    assert np.array_equal(got, g["expected_std"])
    idx.free()


def test_sparse_dropin_flow_env_modes(pkg, tmp_path):
    """The reference-shaped file flow with $FMGPU_MODE=sparse (table on every replica), 1 and 2 logical shards."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    length = int(g["length"])
    nq = 2001
    reads = g["reads"][: nq * length]
    qfa = str(tmp_path / "q.fa")
    helpers.write_fasta_reads(qfa, reads, length)
    ndev = pkg.lib().fmgpu_device_count()
    for tag, key in ((101, "expected_std"), (200, "expected_ac")):
        fn = str(tmp_path / f"i{tag}.fmi")
        g[f"image_{tag}"].tofile(fn)
        for shards in (1, 2):
            os.environ["FMGPU_MODE"] = "sparse"
            try:
                got = pkg.search_files(fn, qfa, length, nq, devices=[i % ndev for i in range(shards)], var=None)
            finally:
                del os.environ["FMGPU_MODE"]
            assert np.array_equal(got, g[key][: 2 * nq]), f"tag {tag} shards {shards}"


def test_dropin_auto_mode_picks_table_by_text(pkg, tmp_path, capfd):
    """transferCPUtoGPU in auto mode on indexes larger than L2: the wide-step table (a step width serves 60-bp reads: 2 x 30)
    for a random text AND for a repeat-rich one, the sparse-step table when $FMGPU_MODE asks for it; same (L,R) as the plain kernel."""
    rng = np.random.default_rng(23)
    n, length, nq = 48_000_001, 60, 20_000
    for kind in ("random", "repeats"):
        text = ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
        if kind == "repeats":
            unit = ACGT[rng.integers(0, 4, 300)]
            for pos in rng.integers(0, n - 300, n // 600):       # half of the text: 10 %-diverged copies of one family
                copy = unit.copy()
                mut = rng.random(300) < 0.10
                copy[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
                text[pos:pos + 300] = copy
        b = pkg.IndexBuild.from_text(text, 2, 64)
        fn = str(tmp_path / f"{kind}.fmi")
        b.download().tofile(fn)
        b.free()
        starts = rng.integers(0, n - length, nq)
        reads = text[(starts[:, None] + np.arange(length)[None, :])].reshape(-1)
        qfa = str(tmp_path / f"{kind}.fa")
        helpers.write_fasta_reads(qfa, reads, length)
        want = pkg.search_files(fn, qfa, length, nq, devices=[0], var=pkg.variant(pkg.MODE_COOP))
        assert ((want[1::2] - want[0::2]) >= 1).all()
        capfd.readouterr()
        os.environ["FMGPU_VERBOSE"] = "1"
        try:
            got = pkg.search_files(fn, qfa, length, nq, devices=[0], var=None)
        finally:
            del os.environ["FMGPU_VERBOSE"]
        err = capfd.readouterr().err
        assert np.array_equal(got, want), kind
        assert "search table: wide-step" in err, err             # on both: repeats become search trees, not a reason to change tables
        os.environ["FMGPU_VERBOSE"] = "1"; os.environ["FMGPU_MODE"] = "sparse"
        try:
            got = pkg.search_files(fn, qfa, length, nq, devices=[0], var=None)
        finally:
            del os.environ["FMGPU_VERBOSE"], os.environ["FMGPU_MODE"]
        err = capfd.readouterr().err
        assert np.array_equal(got, want), kind
        assert "search table: sparse-step" in err, err

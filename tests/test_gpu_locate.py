"""GPU tests of locate (csrc/fm_locate.cuh): the suffix array derived from the index table by list ranking over its LF
mapping, and SA[L..R) gathers.  The reference stops at (L,R) (src/fmIndexCPUBaseline.c:288-290), so the checker here is
brute force on the text: sorted suffixes for the suffix array, a scan for the occurrences of a read.   pytest -m gpu"""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="module")
def pkg(built):
    p = helpers.pkg()
    assert p.lib().fmgpu_device_count() >= 1, "no sm_100 GPU: the product has no CPU fallback"
    return p


def brute_force_sa(text):
    """Suffix array of text + '$' with '$' (end of text) smallest: row 0 is the empty suffix n."""
    t = text.tobytes()
    return np.array(sorted(range(len(t) + 1), key=lambda i: t[i:]), dtype=np.uint32)


def occurrences(text_bytes, read_bytes):
    out, i = [], text_bytes.find(read_bytes)
    while i >= 0:
        out.append(i)
        i = text_bytes.find(read_bytes, i + 1)
    return out


@pytest.mark.parametrize("k", [1, 2])
def test_suffix_array_from_index_files_equals_sorted_suffixes(pkg, k):
    """Committed index files written by the reference tools (all four layouts): the SA derived from the device table
    must be the suffix array of the text they index; reads cut to 6 bases occur several times each."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", f"small_k{k}_d64.npz"))
    n = int(g["n"])
    text = helpers.synth_text(n, seed=7 + k)
    want = brute_force_sa(text)
    tb = text.tobytes()
    length = int(g["length"])
    short = np.ascontiguousarray(g["reads"].reshape(-1, length)[:600, :6]).reshape(-1)
    for tag in (100, 101, 200, 201):
        idx = pkg.DeviceIndex.from_image(g[f"image_{tag}"])
        assert idx.meta.sa_bytes == 0
        idx.build_sa()
        assert idx.meta.sa_bytes == 4 * (n + 1)
        assert np.array_equal(idx.download_sa(), want), f"tag {tag}"
        b = pkg.DeviceBatch(0, 600, 6, k)
        b.upload_ascii(short)
        b.search(idx)
        lr = b.download().reshape(-1, 2)
        pos, nhits = b.locate(idx, 64)
        assert np.array_equal(nhits, np.maximum(lr[:, 1].astype(np.int64) - lr[:, 0], 0).astype(np.uint32))
        for q in range(600):
            occ = occurrences(tb, short[6 * q: 6 * q + 6].tobytes())
            assert nhits[q] == len(occ)
            got = pos[q, : min(len(occ), 64)]
            if len(occ) <= 64:
                assert sorted(got.tolist()) == occ, f"tag {tag} read {q}"
                assert (pos[q, len(occ):] == 0xFFFFFFFF).all()
            else:
                assert set(got.tolist()) <= set(occ)
        b.free()
        idx.drop_sa()
        assert idx.meta.sa_bytes == 0
        idx.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["polyA", "ACGT_period4", "two_letter", "random_plus_repeat"])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_suffix_array_repetitive_texts_and_wide_k(pkg, tmp_path, name, k):
    """Texts whose LF chains run through long BWT runs, index files from the unmodified reference builder; k = 3 files
    are located through their 2-step projection."""
    rng = np.random.default_rng(3)
    n = 3001
    unit = ACGT[rng.integers(0, 4, 150)]
    text = {"polyA": np.full(n, ord("A"), dtype=np.uint8),
            "ACGT_period4": np.tile(ACGT, n // 4 + 1)[:n],
            "two_letter": np.frombuffer(b"AC", dtype=np.uint8)[rng.integers(0, 2, n)],
            "random_plus_repeat": np.concatenate([ACGT[rng.integers(0, 4, n // 2)], np.tile(unit, n // 300 + 1)])[:n]}[name].copy()
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, 64)
    want = brute_force_sa(text)
    for tag in (100, 201):
        idx = pkg.DeviceIndex.from_image(np.fromfile(paths[tag], dtype=np.uint32))
        # (an active AltCounters padding quirk does not matter: the table holds the text's own quirk-free ranks)
        assert np.array_equal(idx.build_sa().download_sa(), want), f"{name} k={k} tag {tag}"
        idx.free()


def test_locate_errors(pkg):
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "quirk_k2_n124.npz"))
    idx = pkg.DeviceIndex.from_image(g["image_100"])
    with pytest.raises(pkg.FMError) as ei:
        idx.build_sa_sampled(5000)                               # rate out of range
    assert ei.value.code == pkg.FM_E_BAD_ARGUMENT
    b = pkg.DeviceBatch(0, g["reads"].size // 8, 8, 2)
    b.upload_ascii(g["reads"])
    b.search(idx)
    with pytest.raises(pkg.FMError) as ei:
        b.locate(idx, 4)                                         # no suffix array yet: loud failure
    assert ei.value.code == pkg.FM_E_BAD_ARGUMENT
    idx.build_sa()
    pos, nhits = b.locate(idx, 4)
    lr = b.download().reshape(-1, 2)
    assert np.array_equal(nhits, np.maximum(lr[:, 1].astype(np.int64) - lr[:, 0], 0).astype(np.uint32))
    b.free(); idx.free()


def test_locate_config3_full_size(pkg):
    """2 Gbp index built on the GPU, 1 M exact 100-bp reads: every read found exactly once must be located at the
    position it was cut from (fm_synth.h read starts); the suffix array is a permutation of 0..n."""
    import torch
    n, nq, length = 2_000_000_000, 1_000_000, 100
    build = pkg.IndexBuild.from_synth(n, 1, 2, 64)
    idx = build.to_index()
    build.free()
    idx.sparsify(0, 0, 0)
    idx.build_sa()
    L = pkg.lib()
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
    torch.cuda.synchronize()
    b = pkg.DeviceBatch(0, nq, length, 2)
    b.upload_ascii(d_ascii.cpu().numpy())
    b.search(idx, pkg.variant(pkg.MODE_SPARSE, 4))
    pos, nhits = b.locate(idx, 2)
    starts = helpers.synth_read_starts(2, nq, n, length)
    assert (nhits >= 1).all()
    once = nhits == 1
    assert once.mean() > 0.99                                    # 100-mers of a random 2 Gbp text are unique
    assert np.array_equal(pos[once, 0].astype(np.uint64), starts[once].astype(np.uint64))
    twice = nhits == 2
    if twice.any():                                              # a repeated 100-mer: the start is one of the two
        assert ((pos[twice, 0] == starts[twice]) | (pos[twice, 1] == starts[twice])).all()
    # permutation check on the device: every value 0..n exactly once
    cai = {"shape": (n + 1,), "typestr": "<i4", "data": (int(L.fmgpu_index_sa(idx.handle)), False), "version": 2}
    holder = type("SA", (), {"__cuda_array_interface__": cai})()
    sa = torch.as_tensor(holder, device="cuda")
    seen = torch.zeros(n + 1, dtype=torch.uint8, device="cuda")
    seen[sa.to(torch.int64)] = 1
    assert int(seen.sum(dtype=torch.int64)) == n + 1
    b.free(); idx.free()


def test_text_beyond_2_31_rows(pkg):
    """hg38-sized text (3.1 Gbp: BWT rows above 2^31, where the reference's builder stops): GPU-built index, plain Coop
    and sparse-step kernels agree, every exact read is found, and locate returns the position each read was cut from."""
    import torch
    n, nq, length = 3_100_000_000, 2_000_000, 100
    build = pkg.IndexBuild.from_synth(n, 1, 2, 64)
    idx = build.to_index()
    build.free()
    assert idx.meta.bwtsize == n + 1
    idx.sparsify(0, 0, 0)
    L = pkg.lib()
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(L.fmgpu_synth_reads_device(0, n, 1, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
    torch.cuda.synchronize()
    b = pkg.DeviceBatch(0, nq, length, 2)
    b.upload_ascii(d_ascii.cpu().numpy())
    b.search(idx, pkg.variant(pkg.MODE_COOP))
    want = b.download()
    assert (want[1::2] > want[0::2]).all()
    assert (want[0::2] >= 2 ** 31).any()                         # rows in the upper half of the 32-bit range are exercised
    b.search(idx, pkg.variant(pkg.MODE_SPARSE, 4))
    assert np.array_equal(b.download(), want)
    idx.build_sa()
    pos, nhits = b.locate(idx, 1)
    starts = helpers.synth_read_starts(2, nq, n, length)
    once = nhits == 1
    assert once.mean() > 0.99
    assert np.array_equal(pos[once, 0].astype(np.uint64), starts[once].astype(np.uint64))
    assert (pos[once, 0] >= 2 ** 31).any()
    b.free(); idx.free()


@pytest.mark.parametrize("k", [1, 2])
def test_sampled_suffix_array_locates_like_the_full_one(pkg, k):
    """fmgpu_index_build_sa_sampled: only every rate-th text position keeps its SA value, locate walks the 1-step LF mapping
    on the 64-byte walk table.  Positions must equal the full array's for every rate, every layout (an AltCounters file
    with an active padding quirk included: the table holds the text's own quirk-free ranks), wide and empty intervals."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", f"small_k{k}_d64.npz"))
    n, length = int(g["n"]), int(g["length"])
    rng = np.random.default_rng(3 + k)
    for tag in (100, 201):
        idx = pkg.DeviceIndex.from_image(g[f"image_{tag}"])
        for qlen, nq, max_hits in ((2 * k, 300, 2048), (6, 600, 64), (length, 600, 4)):
            reads = np.ascontiguousarray(g["reads"].reshape(-1, length)[:nq, :qlen]).reshape(-1).copy()
            reads[rng.integers(0, reads.size, nq // 5)] = ACGT[rng.integers(0, 4, nq // 5)]        # some reads miss
            b = pkg.DeviceBatch(0, nq, qlen, k)
            b.upload_ascii(reads)
            b.search(idx)
            idx.build_sa()
            assert idx.meta.sa_rate == 1
            want_pos, want_n = b.locate(idx, max_hits)
            for rate in (2, 7, 32, 129):
                idx.build_sa_sampled(rate)
                m = idx.meta
                assert m.sa_rate == rate and m.sa_bytes < 4 * (n + 1) + 4096 and (rate < 32 or m.sa_bytes < 0.4 * 4 * (n + 1) + 4096)
                pos, nhits = b.locate(idx, max_hits)
                assert np.array_equal(nhits, want_n) and np.array_equal(pos, want_pos), f"k={k} tag={tag} len={qlen} rate={rate}"
            idx.drop_sa()
            assert idx.meta.sa_bytes == 0 and idx.meta.sa_rate == 0
            b.free()
        idx.free()
    # quirk fixture: locate works on the AltCounters file too, and gives the positions of the standard file's SA
    gq = np.load(os.path.join(helpers.ROOT, "tests", "golden", "quirk_k2_n124.npz" if k == 2 else "quirk_k1_n124.npz"))
    text = helpers.synth_text(int(gq["n"]), seed=100 + int(gq["n"]))
    want = brute_force_sa(text)
    for tag in (100, 200):
        idx = pkg.DeviceIndex.from_image(gq[f"image_{tag}"])
        assert np.array_equal(idx.build_sa().download_sa(), want), f"quirk fixture tag {tag}"
        idx.free()


def test_sampled_suffix_array_at_scale(pkg):
    """20 Mbp: sampled (rate 32) and full arrays locate every exact read where it was cut from; the sampled form is
    more than 5 x smaller."""
    import torch
    n, nq, length = 20_000_003, 200_000, 40
    b = pkg.IndexBuild.from_synth(n, 3, 2, 64)
    idx = b.to_index()
    b.free()
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(pkg.lib().fmgpu_synth_reads_device(0, n, 3, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
    torch.cuda.synchronize()
    batch = pkg.DeviceBatch(0, nq, length, 2)
    batch.upload_ascii(d_ascii.cpu().numpy())
    batch.search(idx)
    starts = helpers.synth_read_starts(2, nq, n, length)
    idx.build_sa()
    full_bytes = idx.meta.sa_bytes
    pos_full, nh = batch.locate(idx, 1)
    idx.build_sa_sampled(32)
    assert idx.meta.sa_bytes * 5 < full_bytes
    pos, nh2 = batch.locate(idx, 1)
    once = nh == 1
    assert once.mean() > 0.99 and np.array_equal(nh, nh2) and np.array_equal(pos, pos_full)
    assert np.array_equal(pos[once, 0].astype(np.int64), starts[once])
    batch.free(); idx.free()

"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI of
libfmindex_b200.so; the checker is the C oracle, the committed reference outputs, and the reference CPU
searcher itself run in-process from oracle/_ref.  Integer work: the bar is bit-exact."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(helpers.ROOT, "tests", "golden", "*.npz")))
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="module")
def pkg(built):
    p = helpers.pkg()
    assert p.lib().fmgpu_device_count() >= 1, "no sm_100 GPU: the product has no CPU fallback"
    return p


def gpu_search(pkg, image, reads, length, var=None, device=0):
    idx = pkg.DeviceIndex.from_image(image, device=device)
    k = int(image[1])
    b = pkg.DeviceBatch(device, reads.size // length, length, k)
    b.upload_ascii(reads)
    b.search(idx, var)
    out = b.download()
    b.free(); idx.free()
    return out


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_all_tags_all_variants(pkg, path):
    """Committed outputs of the unmodified reference searchers, incl. the AltCounters quirk fixtures."""
    g = np.load(path)
    reads, length, k = g["reads"], int(g["length"]), int(g["k"])
    nq = reads.size // length
    for tag, key in ((100, "expected_std"), (101, "expected_std"), (200, "expected_ac"), (201, "expected_ac")):
        idx = pkg.DeviceIndex.from_image(g[f"image_{tag}"])
        b = pkg.DeviceBatch(0, nq, length, k)
        b.upload_ascii(reads)
        for mode in (pkg.MODE_TASK, pkg.MODE_COOP):
            for qpt in (1, 2, 4):
                for tpb in (128, 256, 512):
                    b.search(idx, pkg.variant(mode, qpt, tpb))
                    assert np.array_equal(b.download(), g[key]), f"tag {tag} mode {mode} qpt {qpt} tpb {tpb}"
        b.free(); idx.free()


def test_quirk_metadata(pkg):
    """AltCounters padding-entry quirk is detected only for AC files whose '$' row is in the last chunk."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "quirk_k2_n124.npz"))
    i_std = pkg.DeviceIndex.from_image(g["image_100"]); i_ac = pkg.DeviceIndex.from_image(g["image_201"])
    assert i_std.meta.quirk_mask == 0 and i_std.meta.quirk_start == 0xFFFFFFFF
    assert i_ac.meta.quirk_mask != 0 and i_ac.meta.quirk_start == 64
    g2 = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    i2 = pkg.DeviceIndex.from_image(g2["image_200"])
    assert i2.meta.quirk_mask == 0
    m = i2.meta
    assert m.nsymbols == 16 and m.nblocks % 8 == 0 and m.nblocks >= m.bwtsize // 96 + 1
    assert m.nbytes == 16 * m.nsymbols * m.nblocks
    for i in (i_std, i_ac, i2):
        i.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,d", [(1, 32), (1, 64), (1, 128), (2, 32), (2, 64), (2, 128)])
def test_fresh_data_vs_reference_searcher(pkg, tmp_path, k, d):
    """Fresh seeded text, reference index builder + transformers, reference CPU searcher in-process."""
    n = 300_007 + 64 * d
    text = helpers.synth_text(n, seed=40 + k + d)
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, d)
    length = 48
    rng = np.random.default_rng(d)
    reads = np.concatenate([helpers.synth_reads(text, 6, 20_000, length), text[:length], text[-length:],
                            ACGT[rng.integers(0, 4, 2_000 * length)]])
    for ac, tags in ((False, (100, 101)), (True, (200, 201))):
        ref = helpers.RefSearcher(k, d, ac)
        want, _ = ref.search(ref.load(paths[tags[0]]), reads, length)
        for tag in tags:
            image = np.fromfile(paths[tag], dtype=np.uint32)
            for mode in (pkg.MODE_TASK, pkg.MODE_COOP):
                got = gpu_search(pkg, image, reads, length, pkg.variant(mode))
                assert np.array_equal(got, want), f"k={k} d={d} tag={tag} mode={mode}"


@pytest.mark.parametrize("k,length", [(1, 1), (1, 15), (1, 16), (1, 17), (1, 33), (1, 100), (2, 2), (2, 16), (2, 30),
                                      (2, 32), (2, 34), (2, 100), (2, 128), (2, 250), (1, 250), (2, 1000)])
def test_read_lengths(pkg, k, length):
    """Word boundaries of the 2-bit packing (16 bases/word), even word counts (bank padding), long reads."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", f"small_k{k}_d64.npz"))
    text = helpers.synth_text(int(g["n"]), seed=7 + k)
    reads = np.concatenate([helpers.synth_reads(text, 21, 700, length), ACGT[np.random.default_rng(length).integers(0, 4, 68 * length)]])
    o = helpers.Oracle()
    for tag in (100, 201):
        h = o.wrap(g[f"image_{tag}"])
        want = o.search(h, reads, length)
        o.free(h)
        for mode in (pkg.MODE_TASK, pkg.MODE_COOP):
            for qpt in (1, 4):
                got = gpu_search(pkg, g[f"image_{tag}"], reads, length, pkg.variant(mode, qpt, 256))
                assert np.array_equal(got, want), f"k={k} len={length} tag={tag} mode={mode} qpt={qpt}"


@pytest.mark.parametrize("nq", [0, 1, 2, 31, 32, 33, 255, 257, 1025])
def test_ragged_batches(pkg, nq):
    """Empty and ragged batches (the reference requires num % 32 == 0, common/common.c:143-153)."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    length = int(g["length"])
    reads = g["reads"][: nq * length]
    want = g["expected_std"][: 2 * nq]
    for mode in (pkg.MODE_TASK, pkg.MODE_COOP):
        got = gpu_search(pkg, g["image_101"], reads, length, pkg.variant(mode, 2, 128))
        assert np.array_equal(got, want)


def test_dropin_file_flow_and_logical_shards(pkg, tmp_path):
    """loadIndex / loadQueries / initResults / transferCPUtoGPU / searchIndexGPU / transferGPUtoCPU /
    saveResults through files, with the batch sharded over 1, 2 and 3 replicas (all on GPU 0 when the box
    has one GPU): the result must not depend on the shard count."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    length = int(g["length"])
    nq = 2001                                                   # ragged on purpose
    reads = g["reads"][: nq * length]
    qfa = str(tmp_path / "q.fa")
    helpers.write_fasta_reads(qfa, reads, length)
    ndev = pkg.lib().fmgpu_device_count()
    for tag, key in ((100, "expected_std"), (200, "expected_ac")):
        fn = str(tmp_path / f"i{tag}.fmi")
        g[f"image_{tag}"].tofile(fn)
        for shards in (1, 2, 3):
            devices = [i % ndev for i in range(shards)]
            got = pkg.search_files(fn, qfa, length, nq, devices=devices, var=pkg.variant(pkg.MODE_COOP if shards == 2 else pkg.MODE_TASK))
            assert np.array_equal(got, g[key][: 2 * nq]), f"tag {tag} shards {shards}"
            # driver defaults: $FMGPU_MODE auto (small index -> plain Coop) and fused (fused-step table on every replica)
            for env_mode in ("auto", "fused", "task"):
                os.environ["FMGPU_MODE"] = env_mode
                try:
                    got = pkg.search_files(fn, qfa, length, nq, devices=devices, var=None)
                finally:
                    del os.environ["FMGPU_MODE"]
                assert np.array_equal(got, g[key][: 2 * nq]), f"tag {tag} shards {shards} (FMGPU_MODE={env_mode})"
    # the reference-shaped driver binary writes the reference's text format
    exe = os.path.join(helpers.ROOT, helpers.PKG_NAME, "bin", "fmIndexSearchGPU_b200")
    out = helpers.run([exe, fn, qfa, str(length), str(nq)])
    assert "TIME:" in out
    txt = open(fn + ".res.gpu").read().split("\n")
    assert txt[0] == str(nq)
    flat = np.array([int(v) for line in txt[1:nq + 1] for v in line.split()], dtype=np.uint32)
    assert np.array_equal(flat, g["expected_ac"][: 2 * nq])


def test_end_to_end_host_pipeline(pkg):
    """fmgpu_search_host: host ASCII in, host (L,R) out, chunks overlapped on several streams."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    length = int(g["length"])
    reps = np.tile(g["reads"], 40)                              # ~85k reads
    want = np.tile(g["expected_std"], 40)
    idx = pkg.DeviceIndex.from_image(g["image_100"])
    got = pkg.search_host([idx], reps, length)
    assert np.array_equal(got, want)
    big = np.tile(g["reads"], 1200)                             # ~2.5 M reads: several 512K-read chunks per lane
    wantbig = np.tile(g["expected_std"], 1200)
    for feed in (pkg.FEED_ASCII, pkg.FEED_HOSTPACK, pkg.FEED_HYBRID):
        for mode in (pkg.MODE_TASK, pkg.MODE_COOP):
            got = pkg.search_host([idx], big, length, pkg.variant(mode, feed=feed))
            assert np.array_equal(got, wantbig), f"feed {feed} mode {mode}"
    for _ in range(4):                                           # auto feed: probes hybrid and ASCII, then keeps the faster
        assert np.array_equal(pkg.search_host([idx], big, length, pkg.variant(pkg.MODE_COOP)), wantbig)
    rep2 = idx.replicate(0)                                      # second replica (same GPU) = second shard lane
    got2 = pkg.search_host([idx, rep2], reps[: 1000 * length], length, pkg.variant(pkg.MODE_COOP))
    assert np.array_equal(got2, want[:2000])
    rep2.free(); idx.free()
    assert pkg.lib().fmgpu_release_pipeline() == 0                 # staging buffers are re-created on demand
    idx = pkg.DeviceIndex.from_image(g["image_100"])
    assert np.array_equal(pkg.search_host([idx], reps[: 5000 * length], length), want[:10000])
    idx.free()
    assert pkg.lib().fmgpu_release_pipeline() == 0


def test_pack_kernel_bit_layout(pkg):
    """Reversed 2-bit packing: field s of the packed read is the k-step symbol of LF step s."""
    import torch
    length, nq = 37, 50
    rng = np.random.default_rng(1)
    reads = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)[rng.integers(0, 9, nq * length)]
    d_ascii = torch.from_numpy(reads.copy()).cuda()
    wpq = pkg.lib().fmgpu_words_per_query(length)
    assert wpq == 3
    d_packed = torch.zeros(nq * wpq, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    pkg.check(pkg.lib().fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), s), "pack")
    torch.cuda.synchronize()
    got = d_packed.cpu().numpy().view(np.uint32).reshape(nq, wpq)
    code = lambda c: (((c >> 2) & 1) << 1) | (((c >> 2) & 1) ^ ((c >> 1) & 1))
    r = reads.reshape(nq, length).astype(np.uint32)
    want = np.zeros((nq, wpq), dtype=np.uint32)
    for t in range(length):
        want[:, t // 16] |= code(r[:, length - 1 - t]) << np.uint32(2 * (t % 16))
    assert np.array_equal(got, want)


def test_fetch_counter_matches_oracle(pkg):
    """The in-kernel count of necessary block / sector fetches (algorithmic bytes of the roofline) equals the
    oracle's instrumented walk."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    length, reads = int(g["length"]), g["reads"]
    o = helpers.Oracle()
    h = o.wrap(g["image_100"])
    idx = pkg.DeviceIndex.from_image(g["image_100"])
    b = pkg.DeviceBatch(0, reads.size // length, length, 2)
    b.upload_ascii(reads)
    nb, ns = b.count_fetches(idx)
    assert nb == o.count_sectors(h, reads, length, 96, 1)
    assert ns == o.count_sectors(h, reads, length, 96, 2)
    assert np.array_equal(b.download(), g["expected_std"])     # the counting variant is still a correct search
    b.free(); idx.free(); o.free(h)


def test_error_paths(pkg):
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    idx = pkg.DeviceIndex.from_image(g["image_100"])
    gq = np.load(os.path.join(helpers.ROOT, "tests", "golden", "quirk_k2_n124.npz"))
    qidx = pkg.DeviceIndex.from_image(gq["image_200"])          # odd length at k=2 on an index with the AC quirk: refused
    b = pkg.DeviceBatch(0, 10, 7, 2)
    b.upload_ascii(gq["reads"][: 70])
    with pytest.raises(pkg.FMError) as ei:
        b.search(qidx)
    assert ei.value.code == pkg.FM_E_QUERY_SHAPE
    b.free(); qidx.free()
    b = pkg.DeviceBatch(0, 10, 32, 2)
    b.upload_ascii(g["reads"][: 320])
    for bad_variant in (pkg.variant(pkg.MODE_TASK, 3, 256), pkg.variant(pkg.MODE_TASK, 2, 96), pkg.variant(7, 1, 128)):
        with pytest.raises(pkg.FMError) as ei:
            b.search(idx, bad_variant)
        assert ei.value.code == pkg.FM_E_BAD_ARGUMENT
    b.free()
    bad = g["image_100"].copy(); bad[1] = 3; bad[3] = 64         # k = 3 header over a k = 2 body: size mismatch
    with pytest.raises(pkg.FMError) as ei:
        pkg.DeviceIndex.from_image(bad)
    assert ei.value.code == pkg.FM_E_READING_FMI
    bad = g["image_100"].copy(); bad[1] = 5; bad[3] = 1024       # k = 5: no such index anywhere in the reference
    with pytest.raises(pkg.FMError) as ei:
        pkg.DeviceIndex.from_image(bad)
    assert ei.value.code == pkg.FM_E_UNSUPPORTED_INDEX
    bad = g["image_100"].copy(); bad[4] += 1                     # entry count inconsistent with bwtsize/d
    with pytest.raises(pkg.FMError):
        pkg.DeviceIndex.from_image(bad)
    idx.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
def test_bwtsize_multiple_of_d(pkg, tmp_path):
    """bwtsize % d == 0 makes the reference read past its last entry.  The device
    layout carries a block for X = bwtsize; result = the oracle's defined extension, and no fault."""
    n = 64 * 40 - 1
    text = helpers.synth_text(n, seed=77)
    paths = helpers.build_reference_indexes(str(tmp_path), text, 2, 64)
    reads = helpers.synth_reads(text, 3, 512, 12)
    o = helpers.Oracle()
    h = o.load(paths[100])
    want = o.search(h, reads, 12)
    o.free(h)
    got = gpu_search(pkg, np.fromfile(paths[100], dtype=np.uint32), reads, 12)
    assert np.array_equal(got, want)
    assert ((got[1::2] - got[0::2]) >= 1).all()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
def test_config2_full_size(pkg, tmp_path):
    """BASELINE config 2: 4 Mbp index, 2-step, AltCounters layout, 1M 100-bp reads on one B200, bit-exact vs
    the reference CPU searchers (std and AC), Task and Coop kernels; checksum of the first 100k pinned."""
    n, nq, length = 4_000_000, 1_000_000, 100
    text = helpers.synth_text(n, seed=1)
    reads = helpers.synth_reads(text, 2, nq, length)
    paths = helpers.build_reference_indexes(str(tmp_path), text, 2, 64)
    pinned = open(os.path.join(helpers.ROOT, "tests", "golden", "config1_100k.md5")).read().split()[0]
    for ac, tags in ((False, (100, 101)), (True, (200, 201))):
        ref = helpers.RefSearcher(2, 64, ac)
        want, _ = ref.search(ref.load(paths[tags[0]]), reads, length)
        assert helpers.results_text_md5(want[:200_000]) == pinned
        for tag in tags:
            image = np.fromfile(paths[tag], dtype=np.uint32)
            for mode in (pkg.MODE_TASK, pkg.MODE_COOP):
                got = gpu_search(pkg, image, reads, length, pkg.variant(mode))
                assert np.array_equal(got, want), f"tag {tag} mode {mode}"
    # 1-step index of the same text gives the same intervals (survey's G1 invariant)
    p1 = helpers.build_reference_indexes(str(tmp_path / "k1"), text, 1, 64)
    got1 = gpu_search(pkg, np.fromfile(p1[101], dtype=np.uint32), reads, length)
    assert np.array_equal(got1, want)


# --------------------------------------------------------------------------- #
# GPU index builder: byte-identical to genFMindex's .fmi
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_builder_reproduces_reference_index_files(pkg, path):
    g = np.load(path)
    n, k, d = int(g["n"]), int(g["k"]), int(g["d"])
    seed = (100 + n) if os.path.basename(path).startswith("quirk") else 7 + k
    want = g["image_100"]
    b = pkg.IndexBuild.from_synth(n, seed, k, d)
    assert np.array_equal(b.download(), want), "synthetic-text build != reference gfmiBaseLine output"
    b.free()
    b = pkg.IndexBuild.from_text(helpers.synth_text(n, seed), k, d)
    assert np.array_equal(b.download(), want), "ASCII-text build != reference gfmiBaseLine output"
    idx = b.to_index()                                          # straight into the searchable layout
    length = int(g["length"])
    batch = pkg.DeviceBatch(0, g["reads"].size // length, length, k)
    batch.upload_ascii(g["reads"])
    batch.search(idx)
    assert np.array_equal(batch.download(), g["expected_std"])
    batch.free(); idx.free(); b.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,d,n", [(1, 64, 4_000_000), (2, 64, 4_000_000), (2, 128, 1_000_003), (1, 32, 777_777)])
def test_gpu_builder_vs_reference_builder_config1_size(pkg, tmp_path, k, d, n):
    text = helpers.synth_text(n, seed=1)
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, d)
    want = np.fromfile(paths[100], dtype=np.uint32)
    b = pkg.IndexBuild.from_synth(n, 1, k, d)
    got = b.download()
    b.free()
    assert got.size == want.size and np.array_equal(got[:6 + 2 * k], want[:6 + 2 * k]), "header / '$' rows differ"
    assert np.array_equal(got, want)


def test_gpu_builder_repetitive_text(pkg):
    """Equal-key runs are ordered by full suffix comparison; hopeless texts fail loudly, not slowly."""
    unit = helpers.synth_text(500, seed=5)
    text = np.concatenate([unit, unit, unit[:123], helpers.synth_text(300, seed=6), unit[:77]])   # long repeats
    b = pkg.IndexBuild.from_text(text, 2, 64)
    img = b.download()
    idx = b.to_index()
    # brute-force suffix order on the host
    t = text.tobytes()
    sa = sorted(range(len(t) + 1), key=lambda i: t[i:])
    reads = np.concatenate([text[i:i + 12] for i in range(0, text.size - 12, 7)])
    batch = pkg.DeviceBatch(0, reads.size // 12, 12, 2)
    batch.upload_ascii(reads)
    batch.search(idx)
    got = batch.download().reshape(-1, 2)
    for qi in range(0, got.shape[0], 5):
        pat = reads[qi * 12:(qi + 1) * 12].tobytes()
        rows = [r for r, i in enumerate(sa) if t[i:i + 12] == pat]
        assert (got[qi, 0], got[qi, 1]) == (rows[0], rows[-1] + 1)
    assert img[0] == 100 and img[2] == text.size + 1
    batch.free(); idx.free(); b.free()
    pass


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_transformers_reproduce_reference_files(pkg, path):
    """tfmiBMP / tfmiAC on the GPU: tag 101, 200, 201 images byte-identical to the reference tools' files
    (including the AltCounters padding entry), and searchable with the matching semantics."""
    g = np.load(path)
    n, k, d = int(g["n"]), int(g["k"]), int(g["d"])
    seed = (100 + n) if os.path.basename(path).startswith("quirk") else 7 + k
    b = pkg.IndexBuild.from_synth(n, seed, k, d)
    length = int(g["length"])
    for tag, key in ((101, "expected_std"), (200, "expected_ac"), (201, "expected_ac")):
        t = b.transform(tag)
        assert np.array_equal(t.download(), g[f"image_{tag}"]), f"tag {tag} image differs from the reference tool's file"
        idx = t.to_index()
        batch = pkg.DeviceBatch(0, g["reads"].size // length, length, k)
        batch.upload_ascii(g["reads"])
        batch.search(idx, pkg.variant(pkg.MODE_COOP))
        assert np.array_equal(batch.download(), g[key])
        batch.free(); idx.free(); t.free()
    b.free()


def test_config3_full_size_against_reference_checksums(pkg):
    """BASELINE config 3 at FULL size (2 Gbp, k=2, d=64): the GPU-built index image, its three transformed
    layouts and the (L,R) of the first 1 M reads must have the md5s recorded from the UNMODIFIED reference
    tools (27-minute gfmiBaseLine build + tfmiBMP/tfmiAC + fmIndexSearchCPU[-ac]; tests/golden/config3_2g.json)."""
    import hashlib
    import json
    gold = json.load(open(os.path.join(helpers.ROOT, "tests", "golden", "config3_2g.json")))
    n, k, d = gold["text"]["n"], gold["k"], gold["d"]
    b = pkg.IndexBuild.from_synth(n, gold["text"]["seed"], k, d)
    img = b.download()
    assert [int(v) for v in img[:10]] == gold["header_words"]
    assert hashlib.md5(img.data).hexdigest() == gold["md5"]["tag100_fmi"]
    del img
    nq, length = gold["reads"]["num"], gold["reads"]["len"]
    import torch
    d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
    pkg.check(pkg.lib().fmgpu_synth_reads_device(0, n, gold["text"]["seed"], nq, length, gold["reads"]["seed"], 0, d_ascii.data_ptr(), None), "reads")
    torch.cuda.synchronize()
    reads = d_ascii.cpu().numpy()
    del d_ascii
    batch = pkg.DeviceBatch(0, nq, length, k)
    batch.upload_ascii(reads)
    for tag, key in ((100, "res_cpu_std_text"), (101, "res_cpu_std_text"), (200, "res_cpu_ac_text"), (201, "res_cpu_ac_text")):
        t = b if tag == 100 else b.transform(tag)
        if tag != 100:
            timg = t.download()
            name = {101: "tag101_interleaving", 200: "tag200_ac", 201: "tag201_interleaving_ac"}[tag]
            assert hashlib.md5(timg.data).hexdigest() == gold["md5"][name], f"tag {tag} image md5"
            del timg
        idx = t.to_index()
        modes = [pkg.MODE_TASK, pkg.MODE_COOP]
        if tag in (100, 201):
            idx.fuse()                                           # 68 GB fused table + start table (auto at this size)
            assert idx.meta.fused_bases == 4 and idx.meta.start_bases == 12
            modes.append(pkg.MODE_FUSED)
        for mode in modes:
            batch.search(idx, pkg.variant(mode))
            assert helpers.results_text_md5(batch.download()) == gold["md5"][key], f"tag {tag} mode {mode}"
        idx.free()
        if tag != 100:
            t.free()
    batch.free(); b.free()


# --------------------------------------------------------------------------- #
# fused-step layout (fm_fused.cuh): up to 4 bases per rank, composed from the index's own LF mapping
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p))
def test_fused_steps_all_widths(pkg, path):
    g = np.load(path)
    reads, length, k = g["reads"], int(g["length"]), int(g["k"])
    nq = reads.size // length
    for tag, key in ((100, "expected_std"), (201, "expected_ac")):
        for kf in ([2, 3, 4] if k == 1 else [4]):
            for lanes in (1, 2, 4):
                idx = pkg.DeviceIndex.from_image(g[f"image_{tag}"]).fuse(kf, lanes)
                m = idx.meta
                assert (m.fused_bases, m.fused_lanes) == (kf, lanes) and m.fused_bytes == 4 ** kf * (m.bwtsize // (32 * (8 * lanes - 1)) + 1) * 32 * lanes
                b = pkg.DeviceBatch(0, nq, length, k)
                b.upload_ascii(reads)
                for qpt in (1, 2):
                    b.search(idx, pkg.variant(pkg.MODE_FUSED, qpt))
                    assert np.array_equal(b.download(), g[key]), f"tag {tag} kf {kf} lanes {lanes} qpt {qpt}"
                b.free(); idx.free()


@pytest.mark.parametrize("k,length", [(1, 1), (1, 3), (1, 5), (1, 17), (1, 33), (1, 100), (2, 2), (2, 6), (2, 30), (2, 34), (2, 100),
                                      (2, 126), (2, 128), (2, 250), (1, 251)])
def test_fused_steps_read_lengths(pkg, k, length):
    """Lengths that are not a multiple of the fused width run their leading steps on the SB96 table; bit fields
    of the packed read straddle 32-bit words."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", f"small_k{k}_d64.npz"))
    text = helpers.synth_text(int(g["n"]), seed=7 + k)
    reads = np.concatenate([helpers.synth_reads(text, 21, 700, length), ACGT[np.random.default_rng(length).integers(0, 4, 68 * length)]])
    o = helpers.Oracle()
    h = o.wrap(g["image_101"])
    want = o.search(h, reads, length)
    o.free(h)
    for kf in ([3, 4] if k == 1 else [4]):
        idx = pkg.DeviceIndex.from_image(g["image_101"]).fuse(kf, 2)
        b = pkg.DeviceBatch(0, reads.size // length, length, k)
        b.upload_ascii(reads)
        b.search(idx, pkg.variant(pkg.MODE_FUSED))
        assert np.array_equal(b.download(), want), f"k={k} len={length} kf={kf}"
        b.free(); idx.free()


def test_fused_unavailable_cases(pkg):
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "quirk_k2_n124.npz"))
    idx = pkg.DeviceIndex.from_image(g["image_200"])           # (AltCounters padding quirk: fusable since round 2, see the goldens test)
    b = pkg.DeviceBatch(0, 4, 8, 2)
    b.upload_ascii(g["reads"][:32])
    with pytest.raises(pkg.FMError) as ei:
        b.search(idx, pkg.variant(pkg.MODE_FUSED))             # not fused: loud failure, no silent fallback
    assert ei.value.code == pkg.FM_E_BAD_ARGUMENT
    b.free(); idx.free()
    idx = pkg.DeviceIndex.from_image(g["image_100"])           # same text, standard file: fusable
    idx.fuse(4, 2)
    b = pkg.DeviceBatch(0, g["reads"].size // 8, 8, 2)
    b.upload_ascii(g["reads"])
    b.search(idx, pkg.variant(pkg.MODE_FUSED))
    assert np.array_equal(b.download(), g["expected_std"])
    b.free(); idx.free()


@pytest.mark.parametrize("length", [1, 2, 15, 16, 17, 31, 33, 100, 101, 250])
def test_unstream_kernel_equals_pack_kernel(pkg, length):
    """host stream packer + fm_unstream_kernel == fm_pack_kernel (device ASCII packer) == host per-read packer."""
    import torch
    L = pkg.lib()
    nq = 1237
    rng = np.random.default_rng(length)
    reads = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)[rng.integers(0, 9, nq * length)].copy()
    wpq = L.fmgpu_words_per_query(length)
    want = np.zeros(nq * wpq, dtype=np.uint32)
    L.fm_hostpack_reads_scalar(reads.ctypes.data, nq, length, want.ctypes.data)
    stream_bytes = np.zeros(((nq * length + 3) // 4 + 19) & ~15, dtype=np.uint8)
    L.fm_hostpack_stream(reads.ctypes.data, nq * length, stream_bytes.ctypes.data, 0)
    d_stream = torch.from_numpy(stream_bytes).cuda()
    d_packed = torch.zeros(nq * wpq, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    pkg.check(L.fmgpu_unstream_device(0, d_stream.data_ptr(), nq, length, d_packed.data_ptr(), s), "unstream")
    d_ascii = torch.from_numpy(reads).cuda()
    d_packed2 = torch.zeros(nq * wpq, dtype=torch.int32, device="cuda")
    pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed2.data_ptr(), s), "pack")
    torch.cuda.synchronize()
    assert np.array_equal(d_packed.cpu().numpy().view(np.uint32), want)
    assert np.array_equal(d_packed2.cpu().numpy().view(np.uint32), want)


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("tail_table", [True, False])
@pytest.mark.parametrize("length", [1, 3, 15, 17, 25, 33, 99, 101])
def test_odd_read_length_on_2step_index_equals_1step_index(pkg, tmp_path, monkeypatch, length, tail_table):
    """len % 2 == 1 on a 2-step index is undefined in the reference (it reads query[-1]).  Here the
    last base is consumed by a 1-step rank derived from the 2-step table; the result must be what the reference
    searcher returns on the 1-step index of the SAME text (k=1 and k=2 agree wherever both are defined).
    Both forms of that rank are checked: the tail table (one block fetch, built by the first odd-length search on the
    replica) and the four-fetch derivation it is made from ($FMGPU_TAIL_TABLE=0, or no memory for the table)."""
    monkeypatch.setenv("FMGPU_TAIL_TABLE", "1" if tail_table else "0")
    n = 40_009
    text = helpers.synth_text(n, seed=31)
    p1 = helpers.build_reference_indexes(str(tmp_path / "k1"), text, 1, 64)
    p2 = helpers.build_reference_indexes(str(tmp_path / "k2"), text, 2, 64)
    rng = np.random.default_rng(length)
    reads = np.concatenate([helpers.synth_reads(text, 9, 3000, length), text[:length], text[-length:], text[1:length + 1],
                            ACGT[rng.integers(0, 4, 1500 * length)]])
    ref = helpers.RefSearcher(1, 64, False)
    want, _ = ref.search(ref.load(p1[100]), reads, length)
    for tag in (100, 101, 200, 201):
        idx = pkg.DeviceIndex.from_image(np.fromfile(p2[tag], dtype=np.uint32))
        assert idx.meta.tail_valid == 1
        idx.fuse(4, 2)
        idx.sparsify(4, 0, 0)
        assert idx.meta.tail_bytes == 0                          # nothing is built before an odd length asks for it
        b = pkg.DeviceBatch(0, reads.size // length, length, 2)
        b.upload_ascii(reads)
        for v in (pkg.variant(pkg.MODE_TASK, 1), pkg.variant(pkg.MODE_TASK, 4), pkg.variant(pkg.MODE_COOP, 2), pkg.variant(pkg.MODE_FUSED, 1),
                  pkg.variant(pkg.MODE_FUSED, 2), pkg.variant(pkg.MODE_SPARSE, 1), pkg.variant(pkg.MODE_SPARSE, 4)):
            b.search(idx, v)
            assert np.array_equal(b.download(), want), f"len {length} tag {tag} mode {v.mode} qpt {v.queries_per_thread}"
        assert idx.meta.tail_bytes == (idx.meta.nbytes // 4 if tail_table else 0)
        b.free(); idx.free()


def test_end_to_end_packed_input(pkg):
    """fmgpu_search_host_packed: host reads already in the 2-bit binary format (fm_hostpack_reads words)."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    length = int(g["length"])
    reads = np.tile(g["reads"], 300)
    want = np.tile(g["expected_std"], 300)
    nq = reads.size // length
    L = pkg.lib()
    wpq = L.fmgpu_words_per_query(length)
    packed = np.zeros(nq * wpq, dtype=np.uint32)
    L.fm_hostpack_reads(reads.ctypes.data, nq, length, packed.ctypes.data, 0)
    idx = pkg.DeviceIndex.from_image(g["image_100"]).fuse()
    out = np.zeros(2 * nq, dtype=np.uint32)
    for mode in (pkg.MODE_TASK, pkg.MODE_COOP, pkg.MODE_FUSED):
        handles = (C.c_void_p * 1)(idx.handle)
        v = pkg.variant(mode)
        pkg.check(L.fmgpu_search_host_packed(handles, 1, packed.ctypes.data, nq, length, out.ctypes.data, C.byref(v)), "packed e2e")
        assert np.array_equal(out, want), f"mode {mode}"
    idx.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,d", [(1, 64), (2, 64), (2, 128)])
def test_index_build_cli_writes_the_reference_files(pkg, tmp_path, k, d):
    """bin/gfmi_b200 (the reference's generateIndex main rebuilt on the GPU builder): same file names, same bytes as
    gfmiBaseLine_* + tfmiBMP_* + tfmiAC_*, and the reference CPU searcher accepts the files."""
    n = 123_457
    text = helpers.synth_text(n, seed=17)
    refdir, mydir = tmp_path / "ref", tmp_path / "mine"
    paths = helpers.build_reference_indexes(str(refdir), text, k, d)
    os.makedirs(mydir)
    fa = str(mydir / "ref.fa")
    helpers.write_fasta_ref(fa, text)
    exe = os.path.join(helpers.ROOT, helpers.PKG_NAME, "bin", "gfmi_b200")
    out = helpers.run([exe, fa, str(n), str(k), str(d), "--all"])
    assert "BUILD TIME" in out
    for tag, suf in helpers.TAG_SUFFIX.items():
        mine = f"{fa}.{n}.{d}fmi{k}steps.fmi{suf}"
        assert open(mine, "rb").read() == open(paths[tag], "rb").read(), f"tag {tag} file differs"
    reads = helpers.synth_reads(text, 4, 2000, 20 if k == 1 else 24)
    ref = helpers.RefSearcher(k, d, False)
    got, _ = ref.search(ref.load(f"{fa}.{n}.{d}fmi{k}steps.fmi"), reads, reads.size // 2000)
    assert ((got[1::2] - got[0::2]) >= 1).all()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["polyA", "ACGT_period4", "two_letter", "long_unit_x5", "repeat_plus_random", "AAAC_tail"])
@pytest.mark.parametrize("k", [1, 2])
def test_gpu_builder_general_texts_prefix_doubling(pkg, tmp_path, name, k):
    """Highly repetitive texts take the prefix-doubling path of the GPU builder; the image must still be
    byte-identical to the reference builder's file (divsufsort handles any text)."""
    rng = np.random.default_rng(3)
    unit = helpers.synth_text(700, seed=5)
    text = {
        "polyA": np.full(20_001, ord("A"), dtype=np.uint8),
        "ACGT_period4": np.tile(np.frombuffer(b"ACGT", dtype=np.uint8), 6_000)[:23_999],
        "two_letter": np.frombuffer(b"AC", dtype=np.uint8)[rng.integers(0, 2, 30_011)],
        "long_unit_x5": np.concatenate([unit] * 5 + [unit[:333]]),
        "repeat_plus_random": np.concatenate([helpers.synth_text(5_000, 1), np.full(3_000, ord("T"), dtype=np.uint8), helpers.synth_text(5_000, 1),
                                              np.tile(np.frombuffer(b"GA", dtype=np.uint8), 2_000), helpers.synth_text(777, 2)]),
        "AAAC_tail": np.concatenate([helpers.synth_text(9_000, 4), np.full(200, ord("A"), dtype=np.uint8)]),
    }[name]
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, 64)
    want = np.fromfile(paths[100], dtype=np.uint32)
    b = pkg.IndexBuild.from_text(text, k, 64)
    got = b.download()
    b.free()
    assert np.array_equal(got[:6 + 2 * k], want[:6 + 2 * k]), "header / '$' rows differ"
    assert np.array_equal(got, want)


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", range(36))
def test_fuzz_tiny_references(pkg, tmp_path, case):
    """The survey's fuzz, on the GPU: tiny references (62...5000 bp; ACGT, AC and A-rich alphabets), d in {32,64,128},
    k in {1,2}, read lengths 2...20 (random, prefixes and suffixes of the text).  Index files come from the reference
    tools; every tag, every kernel family, against the reference searcher of that flavour -- '$' rows in every
    position, AltCounters quirk cases included -- and the GPU builder / transformers must reproduce the files."""
    rng = np.random.default_rng(1000 + case)
    k = 1 + case % 2
    d = (32, 64, 128)[(case // 2) % 3]
    n = int(rng.integers(62, 5000))
    if (n + 1) % d == 0:
        n += 1                                                   # reference reads past its last entry there (App. C-2)
    alphabet = (b"ACGT", b"AC", b"AAAAAACGT")[(case // 6) % 3]
    text = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)]
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, d)
    length = int(rng.integers(1, 11)) * 2
    starts = rng.integers(0, n - length + 1, 300)
    reads = np.concatenate([text[s:s + length] for s in starts] + [text[:length], text[-length:], text[1:length + 1],
                           np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), 200 * length)]])
    b = pkg.IndexBuild.from_text(text, k, d)
    for ac, tags in ((False, (100, 101)), (True, (200, 201))):
        ref = helpers.RefSearcher(k, d, ac)
        want, _ = ref.search(ref.load(paths[tags[0]]), reads, length)
        for tag in tags:
            image = np.fromfile(paths[tag], dtype=np.uint32)
            mine = b if tag == 100 else b.transform(tag)
            assert np.array_equal(mine.download(), image), f"case {case}: GPU-built tag {tag} differs from the reference file"
            if tag != 100:
                mine.free()
            idx = pkg.DeviceIndex.from_image(image)
            modes = [pkg.MODE_TASK, pkg.MODE_COOP, pkg.MODE_SPARSE]
            idx.sparsify(2 * k * (1 + case % 3), 0, 2 + 2 * (case % 2))  # sparse-step table (quirk files included: phantom occurrences)
            idx.fuse(4, 2)                                           # fused-step table (quirk files: phantom occurrences in the kernel parameters)
            modes.append(pkg.MODE_FUSED)
            batch = pkg.DeviceBatch(0, reads.size // length, length, k)
            batch.upload_ascii(reads)
            for mode in modes:
                batch.search(idx, pkg.variant(mode))
                assert np.array_equal(batch.download(), want), f"case {case}: k={k} d={d} n={n} tag={tag} mode={mode}"
            batch.free(); idx.free()
    b.free()


@pytest.mark.parametrize("k", [1, 2])
def test_fused_start_table(pkg, k):
    """The fused kernel's start table ((L,R) of all 4^12 12-mers, computed by the kernel itself) must not change
    any result: exact, mutated and random reads of several lengths (table used when len % fused width == 0 and
    len >= 12), against the plain Coop kernel on the same index."""
    n = 20_000_003                                               # bwtsize > 4^12, so the table is meaningful
    os.environ["FMGPU_START_TABLE"] = "1"
    try:
        b = pkg.IndexBuild.from_synth(n, 3, k, 64)
        idx = b.to_index().fuse(4, 2)
        b.free()
    finally:
        del os.environ["FMGPU_START_TABLE"]
    assert idx.meta.start_bases == 12
    text_seed = 3
    import torch
    L = pkg.lib()
    rng = np.random.default_rng(5)
    for length in (12, 16, 24, 52, 100, 14, 10, 101 if k == 2 else 99):
        nq = 200_000
        d_ascii = torch.empty(nq * length, dtype=torch.uint8, device="cuda")
        pkg.check(L.fmgpu_synth_reads_device(0, n, text_seed, nq, length, 2, 0, d_ascii.data_ptr(), None), "reads")
        torch.cuda.synchronize()
        reads = d_ascii.cpu().numpy().copy()
        mut = rng.integers(0, nq * length, nq // 3)              # a third of the reads get one random base: many empty intervals
        reads[mut] = ACGT[rng.integers(0, 4, mut.size)]
        batch = pkg.DeviceBatch(0, nq, length, k)
        batch.upload_ascii(reads)
        batch.search(idx, pkg.variant(pkg.MODE_COOP))
        want = batch.download()
        for qpt in (1, 2):
            batch.search(idx, pkg.variant(pkg.MODE_FUSED, qpt))
            assert np.array_equal(batch.download(), want), f"k={k} len={length} qpt={qpt}"
        batch.free()
    idx.free()

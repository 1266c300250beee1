"""GPU parity tests of the wide-step layout (csrc/fm_wide.cuh): up to 30 bases per 128-byte block fetch, the block
selected by a prefix of the wide symbol (computed from the read, the same for both interval ends), 64-bit entries
carrying the rest of the symbol with the row, overfull buckets as search trees, buckets a short suffix sorts into
answered by plain steps.  Everything goes through the C ABI (ctypes) and is compared bit for bit with committed
reference outputs, the oracle, the reference searcher itself, or the plain kernels on the same index at sizes the CPU
cannot reach.   pytest -m gpu"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
GOLDEN = sorted(os.path.join(helpers.ROOT, "tests", "golden", f) for f in os.listdir(os.path.join(helpers.ROOT, "tests", "golden")) if f.endswith(".npz"))
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
FM_E_NOT_IMPLEMENTED = 19


@pytest.fixture(scope="module")
def pkg(built):
    p = helpers.pkg()
    assert p.lib().fmgpu_device_count() >= 1, "no sm_100 GPU: the product has no CPU fallback"
    return p


def serving_widths(k, length, lead_max):
    """step widths W (multiples of k, 2k..46) with length = b + S*W, S >= 1, b <= lead_max"""
    out = []
    for w in range(2 * k, 47, k):
        if any((length - s * w) >= 0 and (length - s * w) <= lead_max for s in range(1, length // w + 1)):
            out.append(w)
    return out


@pytest.mark.parametrize("dynamic", ["0", "1"], ids=["static_assignment", "dynamic_assignment"])
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p))
def test_wide_steps_golden_all_widths(pkg, monkeypatch, path, dynamic):
    """Committed outputs of the unmodified reference searchers; every serving width, prefix sizes from one bit (everything
    in two buckets: deep search trees) over the automatic choice to the whole symbol (one symbol per bucket), every qpt,
    lead tables of several widths, and every 3rd bucket forced onto the exceptional path.  AltCounters files with an
    active padding quirk are refused (the sparse-step table serves them)."""
    monkeypatch.setenv("FMGPU_WIDE_DYNAMIC", dynamic)           # reads handed to the lane groups statically / from a queue
    g = np.load(path)
    reads, length, k = g["reads"], int(g["length"]), int(g["k"])
    nq = reads.size // length
    b = pkg.DeviceBatch(0, nq, length, k)
    b.upload_ascii(reads)
    served = 0
    for tag, key in ((100, "expected_std"), (101, "expected_std"), (200, "expected_ac"), (201, "expected_ac")):
        probe = pkg.DeviceIndex.from_image(g[f"image_{tag}"])
        quirk = probe.meta.quirk_mask != 0
        probe.free()
        if quirk:
            idx = pkg.DeviceIndex.from_image(g[f"image_{tag}"])
            with pytest.raises(pkg.FMError) as ei:
                idx.widen()
            assert ei.value.code == FM_E_NOT_IMPLEMENTED
            assert idx.wide_bases_for(100) == 0
            idx.free()
            continue
        for lead_max in (0, 3, 15):
            monkeypatch.setenv("FMGPU_WIDE_LEAD_MAX", str(lead_max))
            ws = serving_widths(k, length, lead_max)
            for w in sorted(set(ws[:: max(1, len(ws) // 3)] + ws[-1:])):
                for pbits, force in ((0, 0), (1, 0), (5, 0), (2 * w, 0), (0, 3)):
                    if 2 * w - min(pbits or 16, 16, 2 * w) + (int(g["n"]) + 1).bit_length() > 96:
                        continue                                 # entry = rest of the symbol + row number: 64 bits, else 96, at most
                    if force:
                        monkeypatch.setenv("FMGPU_WIDE_FORCE_EXC", str(force))
                    else:
                        monkeypatch.delenv("FMGPU_WIDE_FORCE_EXC", raising=False)
                    lanes = (2, 4)[(served + tag) % 2] if pbits != 5 else (4, 2)[(served + tag) % 2]
                    idx = pkg.DeviceIndex.from_image(g[f"image_{tag}"]).widen(w, min(pbits, 16), lanes)
                    m = idx.meta
                    assert m.wide_bases == w and m.wide_lanes == lanes and m.wide_blocks == (1 << m.wide_prefix_bits) + m.wide_tree_nodes
                    assert (m.wide_overflow > 0) == (m.wide_tree_nodes > 0) == (m.wide_tree_depth > 0)
                    assert m.wide_bytes == m.wide_blocks * 32 * lanes and m.derived_bytes >= m.wide_bytes
                    bits = 2 * w - m.wide_prefix_bits + m.wide_row_bits
                    assert m.wide_entry_words == (2 if bits <= 64 and w <= 31 else 3)
                    if force:
                        assert m.wide_exceptional >= (1 << m.wide_prefix_bits) // 3
                    idx.prepare(length)
                    assert idx.wide_serves(length)
                    for qpt in (1, 2, 3, 4):
                        b.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
                        assert np.array_equal(b.download(), g[key]), f"tag {tag} W {w} prefix {pbits} lanes {lanes} lead_max {lead_max} force {force} qpt {qpt}"
                    served += 1
                    idx.free()
    assert served or "quirk" in os.path.basename(path)
    b.free()


@pytest.mark.parametrize("k,length", [(1, 8), (1, 17), (1, 33), (1, 100), (2, 16), (2, 20), (2, 30), (2, 34), (2, 60), (2, 100), (2, 126),
                                      (2, 128), (2, 250), (1, 251), (2, 25), (2, 101), (2, 45), (2, 61)])
def test_wide_steps_read_lengths(pkg, monkeypatch, k, length):
    """length = lead bases + whole wide steps, bit fields of the packed read straddle 32-bit words (60-bit keys span three);
    odd lengths on a 2-step index take an odd lead table (derived 1-step rank).  The width the library proposes for the
    length, the widest, and a narrow one; a table whose width does not serve the length says so instead of answering."""
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
    monkeypatch.setenv("FMGPU_WIDE_DYNAMIC", str(length % 2))   # (the default follows the share of rows in trees)
    proposed = idx.wide_bases_for(length)
    assert (proposed != 0) == (length >= 16)
    tried = 0
    for w in sorted({proposed, 30, 8 * k, 14, 22, 46, 40, 36} - {0}):
        idx.widen(w, 0, (2, 4)[(w // 2) % 2])
        idx.prepare(length)
        if idx.wide_serves(length):
            for qpt in (1, 2, 4):
                b.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
                assert np.array_equal(b.download(), want), f"k={k} len={length} W={w} qpt={qpt}"
            tried += 1
        else:
            with pytest.raises(pkg.FMError) as ei:
                b.search(idx, pkg.variant(pkg.MODE_WIDE))
            assert ei.value.code == FM_E_NOT_IMPLEMENTED
        idx.unwiden()
        assert idx.meta.wide_bases == 0 and idx.meta.wide_bytes == 0
    assert tried >= 1
    b.free(); idx.free()


def test_wide_unavailable_and_errors(pkg):
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    idx = pkg.DeviceIndex.from_image(g["image_100"])
    length = int(g["length"])
    b = pkg.DeviceBatch(0, g["reads"].size // length, length, 2)
    b.upload_ascii(g["reads"])
    with pytest.raises(pkg.FMError) as ei:
        b.search(idx, pkg.variant(pkg.MODE_WIDE))              # no table: loud failure, no silent fallback
    assert ei.value.code == pkg.FM_E_BAD_ARGUMENT
    for bad in ((3, 0, 0), (2, 0, 0), (48, 0, 0), (16, 31, 0), (16, 0, 3), (16, 0, 8)):
        with pytest.raises(pkg.FMError) as ei:
            idx.widen(*bad)
        assert ei.value.code == pkg.FM_E_BAD_ARGUMENT, bad
    idx.widen(16)
    assert idx.meta.wide_bases == 16
    b.search(idx, pkg.variant(pkg.MODE_WIDE), prepare=False)   # 32 = 2 x 16 needs no lead table ...
    assert np.array_equal(b.download(), g["expected_std"])
    idx.unwiden()
    idx.widen(22)                                               # ... 32 = 10 + 22 does, and a launch never builds one
    with pytest.raises(pkg.FMError) as ei:
        b.search(idx, pkg.variant(pkg.MODE_WIDE), prepare=False)
    assert ei.value.code == FM_E_NOT_IMPLEMENTED
    b.search(idx, pkg.variant(pkg.MODE_WIDE))
    assert np.array_equal(b.download(), g["expected_std"])
    b.free(); idx.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["polyA", "ACGT_period4", "two_letter", "repeat_x40", "random_plus_repeat"])
@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("dynamic", ["0", "1"], ids=["static_assignment", "dynamic_assignment"])
def test_wide_steps_repetitive_texts_search_trees(pkg, tmp_path, monkeypatch, name, k, dynamic):
    """Repeats put thousands of rows into one bucket: those become search trees of 128-byte blocks (three levels for
    poly-A), walked one fetch per iteration, the two interval ends apart once they leave the root.  Index files and
    expected (L,R) from the unmodified reference tools."""
    monkeypatch.setenv("FMGPU_WIDE_DYNAMIC", dynamic)
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
    ref = helpers.RefSearcher(k, d, False)
    image = np.fromfile(paths[101], dtype=np.uint32)
    idx = pkg.DeviceIndex.from_image(image)
    saw_overflow, deepest = False, 0
    monkeypatch.setenv("FMGPU_WIDE_PACK", str(k % 2))          # (k = 2: the unpacked form of the 96-bit entries)
    for length, widths in ((20, (10, 20)), (40, (30, 10)), (66, (30,)), (98, (46,))):
        starts = rng.integers(0, text.size - length + 1, 2000)
        reads = np.concatenate([text[s:s + length] for s in starts] + [ACGT[rng.integers(0, 4, 500 * length)]])
        want, _ = ref.search(ref.load(paths[100]), reads, length)
        b = pkg.DeviceBatch(0, reads.size // length, length, k)
        b.upload_ascii(reads)
        for w in widths:
            for pbits in (0, 4, max(4, 2 * w + 15 - 64)) if w <= 30 else (0, 12):   # (few prefix bits: deep trees, and 96-bit entries when 64 do not hold the rest of the symbol + a 15-bit row)
                idx.widen(w, pbits, (2, 4)[(w + pbits + length) % 2])
                m = idx.meta
                saw_overflow |= m.wide_overflow > 0
                deepest = max(deepest, m.wide_tree_depth)
                assert m.wide_tree_rows <= m.bwtsize and (m.wide_tree_nodes > 0) == (m.wide_overflow > 0)
                for qpt in (1, 4):
                    b.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
                    assert np.array_equal(b.download(), want), f"{name} k={k} len={length} W={w} prefix={pbits} qpt={qpt}"
                idx.unwiden()
        b.free()
    assert saw_overflow, "these texts are meant to overflow buckets"
    if name == "polyA":
        assert deepest >= 3                                      # 30 000 rows in one bucket: 15 * 16 * 16 < 30 000 (7 * 8 * 8 * 8 for 64-byte blocks)
    idx.free()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
def test_wide_fuzz_tiny_references_against_the_reference_searcher(pkg, tmp_path):
    """Tiny references: the suffixes shorter than a wide step are a visible share of the text, they sort INTO buckets
    (which the builder must detect and mark exceptional), reads run over the text's ends, absent symbols are frequent.
    std and AltCounters files (quirk-free ones), k = 1 and 2, widths up to 30; the checker is the reference searcher."""
    rng = np.random.default_rng(99)
    exceptional = 0
    cases = 0
    for case in range(36):
        k = 1 + case % 2
        d = (32, 64)[(case // 2) % 2]
        n = int(rng.integers(d + 2, 6 * d))
        if (n + 1) % d == 0:
            n += 1
        alphabet = (b"ACGT", b"AC", b"AAAAAACGT")[(case // 4) % 3]
        text = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)]
        paths = helpers.build_reference_indexes(str(tmp_path / f"c{case}"), text, k, d)
        for tag, ac in ((100, False), (201, True)):
            idx = pkg.DeviceIndex.from_image(np.fromfile(paths[tag], dtype=np.uint32))
            if idx.meta.quirk_mask:
                idx.free()
                continue
            ref = helpers.RefSearcher(k, d, ac)
            for length, w, pbits in ((8, 8, 0), (12, 6, 0), (12, 12, 3), (24, 12, 0), (30, 30, 0), (34, 30, 0), (40, 20, 6), (60, 30, 6),
                                     (60, 30, 2), (46, 46, 0), (50, 46, 0), (92, 46, 6), (80, 40, 0)):
                if length > n:
                    continue
                starts = rng.integers(0, n - length + 1, 200)
                reads = np.concatenate([text[s:s + length] for s in starts] + [text[:length], text[-length:],
                                       np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), 150 * length)]])
                want, _ = ref.search(ref.load(paths[200 if ac else 100]), reads, length)
                batch = pkg.DeviceBatch(0, reads.size // length, length, k)
                batch.upload_ascii(reads)
                os.environ["FMGPU_WIDE_LEAD_MAX"] = "5"
                os.environ["FMGPU_WIDE_DYNAMIC"] = str((case + length) % 2)
                os.environ["FMGPU_WIDE_PACK"] = str((case // 3 + w) % 2)      # 96-bit entries on 64-byte blocks: five packed / four per block
                try:
                    idx.widen(w, pbits, (2, 4)[(case // 2 + length) % 2])
                    exceptional += idx.meta.wide_exceptional
                    for qpt in (1, 3):
                        batch.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
                        assert np.array_equal(batch.download(), want), f"case {case}: k={k} d={d} n={n} tag={tag} len={length} W={w} prefix={pbits} qpt={qpt}"
                finally:
                    del os.environ["FMGPU_WIDE_LEAD_MAX"], os.environ["FMGPU_WIDE_DYNAMIC"], os.environ["FMGPU_WIDE_PACK"]
                idx.unwiden()
                batch.free()
                cases += 1
            idx.free()
    print(f"wide fuzz: {cases} (index, length, width) cases, {exceptional} exceptional buckets")
    assert cases >= 300 and exceptional >= 1, (cases, exceptional)


@pytest.mark.parametrize("k", [1, 2])
def test_wide_lead_table_and_fetch_counter(pkg, k):
    """20 Mbp synthetic text, 30 bases per step: exact, mutated and random reads of several lengths against the plain Coop
    kernel on the same index; the instrumented kernel counts ONE block fetch per wide step and read (both interval ends
    share the block), plus the rare tree node."""
    import torch
    n = 20_000_003
    bld = pkg.IndexBuild.from_synth(n, 3, k, 64)
    idx = bld.to_index().widen()
    bld.free()
    m = idx.meta
    assert (m.wide_bases, m.wide_lanes, m.wide_prefix_bits, m.wide_row_bits) == (30, 2, 24, 25)   # (roomy grid: 1.2 rows per 7-entry bucket)
    assert m.wide_exceptional <= 64 and m.wide_tree_rows < n // 20
    L = pkg.lib()
    rng = np.random.default_rng(5)
    stream = torch.cuda.current_stream().cuda_stream
    for length in (100, 30, 31, 40, 60, 70, 72, 250, 43, 101):
        nq = 200_000
        if not (length % k == 0 or k == 2):
            continue
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
        assert idx.wide_serves(length)
        for qpt in (1, 2, 3, 4):
            batch.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
            assert np.array_equal(batch.download(), want), f"k={k} len={length} qpt={qpt}"
        batch.free()
        if length == 100:
            d_ascii.copy_(torch.from_numpy(reads))
            wpq = L.fmgpu_words_per_query(length)
            d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
            d_res = torch.zeros(2 * nq, dtype=torch.int32, device="cuda")
            pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), stream), "pack")
            a, s, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
            pkg.check(L.fmgpu_count_fetches_wide_device(idx.handle, d_packed.data_ptr(), nq, length, d_res.data_ptr(), stream,
                                                        C.byref(a), C.byref(s), C.byref(o)), "count")
            assert np.array_equal(d_res.cpu().numpy().view(np.uint32), want)
            assert a.value == 3 * nq                              # 10-base lead table + 3 wide steps, one grid block each
            assert s.value <= 0.001 * nq * 60 and o.value <= 0.2 * nq   # exceptional buckets: a handful; overfull buckets: the Poisson tail
    # 96-bit entries: 46 bases per step, four entries per 64-byte block
    assert idx.wide_bases_for(100) == 46
    idx.unwiden()
    idx.widen(46)
    m = idx.meta
    assert (m.wide_bases, m.wide_lanes, m.wide_entry_words, m.wide_block_entries, m.wide_prefix_bits) == (46, 2, 3, 5, 24)
    for length in (100, 92, 46, 50, 101, 146):
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
        assert idx.wide_serves(length)
        for qpt in (1, 2, 3, 4):
            batch.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
            assert np.array_equal(batch.download(), want), f"k={k} len={length} qpt={qpt} (46 bases per step)"
        if length == 100:
            d_packed = torch.from_numpy(np.zeros(1, dtype=np.int32))   # (the shard's own packed reads are used below)
            a, s, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
            pkg.check(L.fmgpu_count_fetches_wide_device(idx.handle, L.fmgpu_batch_packed(batch.handle), nq, length, L.fmgpu_batch_results(batch.handle),
                                                        L.fmgpu_batch_stream(batch.handle), C.byref(a), C.byref(s), C.byref(o)), "count")
            assert np.array_equal(batch.download(), want)
            assert a.value == 2 * nq and o.value <= 0.3 * nq      # 8-base lead table + 2 wide steps
        batch.free()
    idx.free()


def test_wide_config3_full_size_against_reference_checksums(pkg):
    """BASELINE config 3 at FULL size (2 Gbp, k=2, d=64): (L,R) of the first 1 M reads from the wide-step kernel,
    default table (30 bases per step, 2^28 buckets, 10-base lead table), must have the md5 recorded from the UNMODIFIED
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
        idx = t.to_index()
        assert idx.wide_bases_for(length) == 46                  # 100 = 8 + 2 x 46: 96-bit entries
        idx.widen(46)
        m = idx.meta
        assert (m.wide_bases, m.wide_lanes, m.wide_entry_words, m.wide_block_entries, m.wide_prefix_bits, m.wide_row_bits) == (46, 2, 3, 5, 30, 31)
        assert m.wide_bytes < 76e9 and m.wide_exceptional <= 96 and m.wide_tree_rows < n // 10
        for qpt in (1, 2):
            batch.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
            assert helpers.results_text_md5(batch.download()) == gold["md5"][key], f"tag {tag} qpt {qpt} (46 bases per step)"
        idx.unwiden()
        idx.widen()                                             # widest step with 64-bit entries: 100 = 10 + 3 x 30
        m = idx.meta
        assert (m.wide_bases, m.wide_lanes, m.wide_entry_words, m.wide_prefix_bits, m.wide_row_bits) == (30, 2, 2, 30, 31) and m.wide_bytes < 72e9   # roomy grid
        assert m.wide_exceptional <= 64 and m.wide_tree_rows < n // 100
        for qpt in (1, 2, 3):
            batch.search(idx, pkg.variant(pkg.MODE_WIDE, qpt))
            assert helpers.results_text_md5(batch.download()) == gold["md5"][key], f"tag {tag} qpt {qpt}"
        idx.unwiden()
        os.environ["FMGPU_WIDE_ROOMY"] = "0"
        try:
            idx.widen(0, 0, 4)                                  # 128-byte blocks, compact grid: 2^28 buckets of 15 entries
        finally:
            del os.environ["FMGPU_WIDE_ROOMY"]
        m = idx.meta
        assert (m.wide_bases, m.wide_lanes, m.wide_prefix_bits) == (30, 4, 28) and m.wide_bytes < 36e9
        batch.search(idx, pkg.variant(pkg.MODE_WIDE, 1))
        assert helpers.results_text_md5(batch.download()) == gold["md5"][key], f"tag {tag} 128-byte blocks"
        idx.free()
        if tag != 100:
            t.free()
    batch.free(); b.free()


def test_wide_end_to_end_host_calls(pkg):
    """fmgpu_search_host with the wide-step variant: host buffers in, host (L,R) out."""
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    reads, length = g["reads"], int(g["length"])
    idx = pkg.DeviceIndex.from_image(g["image_100"]).widen(22)
    assert idx.meta.wide_bases == 22
    got = pkg.search_host([idx], reads, length, var=pkg.variant(pkg.MODE_WIDE, 2))
    assert np.array_equal(got, g["expected_std"])
    idx.free()


def test_wide_dropin_flow_env_modes(pkg, tmp_path):
    """The reference-shaped file flow with $FMGPU_MODE=wide (table on every replica, width chosen for the read length),
    1 and 2 logical shards; a quirk AltCounters file falls back to the sparse-step table."""
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
            os.environ["FMGPU_MODE"] = "wide"
            os.environ["FMGPU_VERBOSE"] = "1"
            try:
                got = pkg.search_files(fn, qfa, length, nq, devices=[i % ndev for i in range(shards)], var=None)
            finally:
                del os.environ["FMGPU_MODE"], os.environ["FMGPU_VERBOSE"]
            assert np.array_equal(got, g[key][: 2 * nq]), f"tag {tag} shards {shards}"
    q = np.load(os.path.join(helpers.ROOT, "tests", "golden", "quirk_k2_n124.npz"))
    fn = str(tmp_path / "quirk.fmi")
    q["image_200"].tofile(fn)
    qfa2 = str(tmp_path / "q2.fa")
    helpers.write_fasta_reads(qfa2, q["reads"], 8)
    os.environ["FMGPU_MODE"] = "wide"
    try:
        got = pkg.search_files(fn, qfa2, 8, q["reads"].size // 8, devices=[0], var=None)
    finally:
        del os.environ["FMGPU_MODE"]
    assert np.array_equal(got, q["expected_ac"])

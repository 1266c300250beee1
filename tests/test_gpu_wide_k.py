"""GPU parity tests for index files with k = 3 and k = 4 (built and searched on the CPU only by the reference,
makefile:226-230): the library searches them through the 2-step index that their first two BWT layers define.
Expected (L,R) come from the unmodified reference searchers compiled for that k (oracle/_ref).   pytest -m gpu"""
import os

import numpy as np
import pytest

import helpers

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")]
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="module")
def pkg(built):
    p = helpers.pkg()
    assert p.lib().fmgpu_device_count() >= 1, "no sm_100 GPU: the product has no CPU fallback"
    return p


def make_reads(text, rng, length, nexact=1500, nrandom=400):
    starts = rng.integers(0, text.size - length + 1, nexact)
    reads = [text[s:s + length].copy() for s in starts]
    for r in reads[::3]:                                           # a third of them with one substituted base
        r[rng.integers(0, length)] = ACGT[rng.integers(0, 4)]
    return np.concatenate(reads + [text[:length], text[-length:], text[1:length + 1], ACGT[rng.integers(0, 4, nrandom * length)]])


@pytest.mark.parametrize("k", [3, 4])
@pytest.mark.parametrize("d", [32, 64, 128])
def test_wide_k_files_all_tags_all_kernels(pkg, tmp_path, k, d):
    rng = np.random.default_rng(100 * k + d)
    n = 60013 + 7 * d
    text = ACGT[rng.integers(0, 4, n)]
    paths = helpers.build_reference_indexes(str(tmp_path / "wide"), text, k, d)
    paths2 = helpers.build_reference_indexes(str(tmp_path / "two"), text, 2, d)
    import torch
    twostep = pkg.DeviceIndex.from_image(np.fromfile(paths2[100], dtype=np.uint32))
    table2 = torch.as_tensor(twostep, device="cuda").clone()
    twostep.free()
    for length in (12, 24, 36):
        reads = make_reads(text, rng, length)
        nq = reads.size // length
        for ac, tags in ((False, (100, 101)), (True, (200, 201))):
            ref = helpers.RefSearcher(k, d, ac)
            want, _ = ref.search(ref.load(paths[tags[0]]), reads, length)
            for tag in tags:
                image = np.fromfile(paths[tag], dtype=np.uint32)
                assert int(image[1]) == k
                try:
                    idx = pkg.DeviceIndex.from_image(image)
                except pkg.FMError as ex:                            # active AltCounters padding quirk: refused, never wrong
                    assert ac and ex.code == 52
                    continue
                m = idx.meta
                assert (m.steps, m.source_steps, m.source_tag, m.nsymbols) == (2, k, tag, 16)
                # the derived table IS the table of the reference's own 2-step file of the same text
                assert torch.equal(torch.as_tensor(idx, device="cuda"), table2), f"k={k} d={d} tag={tag}: projected table differs"
                idx.fuse(4, 2)
                idx.sparsify(6 if length % 6 == 0 else 4, 0, 0)
                b = pkg.DeviceBatch(0, nq, length, 2)
                b.upload_ascii(reads)
                for v in (pkg.variant(pkg.MODE_TASK, 2), pkg.variant(pkg.MODE_COOP, 1), pkg.variant(pkg.MODE_FUSED, 2), pkg.variant(pkg.MODE_SPARSE, 4)):
                    b.search(idx, v)
                    assert np.array_equal(b.download(), want), f"k={k} d={d} tag={tag} len={length} mode={v.mode}"
                b.free(); idx.free()


@pytest.mark.parametrize("k", [3, 4])
def test_wide_k_dropin_file_flow(pkg, tmp_path, k):
    """loadIndex ... saveResults on a k = 3 / 4 file; a read length that is not a multiple of k is refused like any
    length the reference leaves undefined."""
    rng = np.random.default_rng(k)
    n, d, length, nq = 30029, 64, 24, 999
    text = ACGT[rng.integers(0, 4, n)]
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, d)
    reads = make_reads(text, rng, length, nexact=nq - 3, nrandom=0)
    qfa = str(tmp_path / "q.fa")
    helpers.write_fasta_reads(qfa, reads, length)
    for ac, tag in ((False, 101), (True, 200)):
        ref = helpers.RefSearcher(k, d, ac)
        want, _ = ref.search(ref.load(paths[200 if ac else 100]), reads, length)
        try:
            got = pkg.search_files(paths[tag], qfa, length, nq, devices=[0], var=None)
        except pkg.FMError as ex:
            assert ac and ex.code == 52
            continue
        assert np.array_equal(got, want), f"k={k} tag={tag}"
    bad = 22 if k == 3 else 26
    helpers.write_fasta_reads(qfa, reads[: 10 * bad], bad)
    with pytest.raises(pkg.FMError) as ei:
        pkg.search_files(paths[101], qfa, bad, 10, devices=[0], var=None)
    assert ei.value.code == pkg.FM_E_QUERY_SHAPE


@pytest.mark.parametrize("case", range(24))
def test_wide_k_fuzz_tiny_references(pkg, tmp_path, case):
    """Tiny references put the '$' rows everywhere, including the last chunk where the AltCounters padding quirk lives:
    std files always match the std searcher; AC files either match the AC searcher or are refused (code 52) --
    and they are only ever refused when the quirk could bite (a '$' row in the last chunk)."""
    rng = np.random.default_rng(5000 + case)
    k = 3 + case % 2
    d = (32, 64, 128)[(case // 2) % 3]
    n = int(rng.integers(4 * k + 40, 3000))
    if (n + 1) % d == 0:
        n += 1
    alphabet = (b"ACGT", b"AC", b"AAAAAACGT")[(case // 6) % 3]
    text = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)]
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, d)
    length = k * int(rng.integers(1, 6))
    starts = rng.integers(0, n - length + 1, 200)
    reads = np.concatenate([text[s:s + length] for s in starts] + [text[:length], text[-length:],
                           np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), 100 * length)]])
    nq = reads.size // length
    for ac, tags in ((False, (100, 101)), (True, (200, 201))):
        ref = helpers.RefSearcher(k, d, ac)
        want, _ = ref.search(ref.load(paths[tags[0]]), reads, length)
        for tag in tags:
            image = np.fromfile(paths[tag], dtype=np.uint32)
            try:
                idx = pkg.DeviceIndex.from_image(image)
            except pkg.FMError as ex:
                assert ac and ex.code == 52, f"case {case} tag {tag}: {ex}"
                dpos = image[6:6 + k]
                assert any(int(p) // d == (n + 1 - 1) // d for p in dpos), "refused without a '$' row in the last chunk"
                continue
            b = pkg.DeviceBatch(0, nq, length, 2)
            b.upload_ascii(reads)
            for v in (pkg.variant(pkg.MODE_TASK, 1), pkg.variant(pkg.MODE_COOP, 2)):
                b.search(idx, v)
                assert np.array_equal(b.download(), want), f"case {case}: k={k} d={d} n={n} tag={tag} len={length} mode={v.mode}"
            b.free(); idx.free()


@pytest.mark.parametrize("k", [3, 4])
@pytest.mark.parametrize("d", [32, 64, 128])
def test_gpu_builder_writes_the_reference_files_for_k3_k4(pkg, tmp_path, k, d):
    """fmgpu_build_from_text with k = 3, 4 (src/genFMindex.c:184-260,327-455 for any K_STEPS): the image and its three
    transformed layouts are byte-identical to what gfmiBaseLine_<d>bases_<k>step / tfmiBMP / tfmiAC write, on a random
    text and on a repetitive one (prefix-doubling path); gfmi_b200 writes the same file under the same name."""
    rng = np.random.default_rng(7 * k + d)
    unit = ACGT[rng.integers(0, 4, 500)]
    texts = {"random": ACGT[rng.integers(0, 4, 30011 + d)],
             "repeats": np.concatenate([ACGT[rng.integers(0, 4, 4000)], np.tile(unit, 9), np.full(700, ord("A"), dtype=np.uint8), ACGT[rng.integers(0, 4, 777)]])}
    for name, text in texts.items():
        if (text.size + 1) % d == 0:
            text = text[:-1]
        paths = helpers.build_reference_indexes(str(tmp_path / name), text, k, d)
        b = pkg.IndexBuild.from_text(text, k, d)
        got = b.download()
        want = np.fromfile(paths[100], dtype=np.uint32)
        assert np.array_equal(got[:6 + 2 * k], want[:6 + 2 * k]), f"{name}: header / '$' rows differ"
        assert np.array_equal(got, want), f"{name}: k={k} d={d} tag-100 image differs from the reference builder's file"
        for tag in (101, 200, 201):
            t = b.transform(tag)
            assert np.array_equal(t.download(), np.fromfile(paths[tag], dtype=np.uint32)), f"{name}: k={k} d={d} tag {tag}"
            t.free()
        # searched like any k >= 3 file: through the 2-step index its first two layers define
        length = 4 * k
        reads = make_reads(text, rng, length, 600, 200)
        ref = helpers.RefSearcher(k, d, False)
        want_lr, _ = ref.search(ref.load(paths[100]), reads, length)
        idx = b.to_index()
        batch = pkg.DeviceBatch(0, reads.size // length, length, 2)
        batch.upload_ascii(reads)
        batch.search(idx, pkg.variant(pkg.MODE_COOP))
        assert np.array_equal(batch.download(), want_lr), f"{name}: k={k} d={d}"
        batch.free(); idx.free(); b.free()
    # the CLI: same file name, same bytes as the reference tool
    text = texts["random"] if (texts["random"].size + 1) % d else texts["random"][:-1]
    wd = tmp_path / "cli"
    wd.mkdir()
    helpers.write_fasta_ref(str(wd / "ref.fa"), text)
    exe = os.path.join(helpers.ROOT, helpers.PKG_NAME, "bin", "gfmi_b200")
    helpers.run([exe, "ref.fa", str(text.size), str(k), str(d), "--all"], cwd=str(wd))
    ref_paths = helpers.build_reference_indexes(str(tmp_path / "random"), text, k, d)
    for tag, suf in helpers.TAG_SUFFIX.items():
        mine = wd / f"ref.fa.{text.size}.{d}fmi{k}steps.fmi{suf}"
        assert mine.read_bytes() == open(ref_paths[tag], "rb").read(), f"gfmi_b200 k={k} d={d} tag {tag}"

"""CPU tests: the C oracle against (a) the committed outputs of the unmodified reference searchers and
(b) the reference searchers run in-process here (oracle/_ref) on fresh seeded inputs."""
import glob
import os

import numpy as np
import pytest

import helpers

GOLDEN = sorted(glob.glob(os.path.join(helpers.ROOT, "tests", "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_committed_reference_outputs(built, path):
    g = np.load(path)
    o = helpers.Oracle()
    length = int(g["length"])
    for tag, key in ((100, "expected_std"), (101, "expected_std"), (200, "expected_ac"), (201, "expected_ac")):
        h = o.wrap(g[f"image_{tag}"])
        got = o.search(h, g["reads"], length)
        o.free(h)
        assert np.array_equal(got, g[key]), f"oracle != reference output for tag {tag}"


def test_golden_set_contains_altcounters_quirk(built):
    """At least one fixture must exercise SURVEY.md App. C-3 (AC result != std result)."""
    assert any(not np.array_equal(np.load(p)["expected_std"], np.load(p)["expected_ac"]) for p in GOLDEN)


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("d", [32, 64, 128])
def test_oracle_matches_reference_in_process(built, tmp_path, k, d):
    n = 50021 + 13 * d
    text = helpers.synth_text(n, seed=3 * k + d)
    paths = helpers.build_reference_indexes(str(tmp_path), text, k, d)
    length = 24
    reads = np.concatenate([helpers.synth_reads(text, 5, 3000, length), text[:length], text[-length:],
                            np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(k + d).integers(0, 4, 500 * length)]])
    o = helpers.Oracle()
    for ac, tags in ((False, (100, 101)), (True, (200, 201))):
        ref = helpers.RefSearcher(k, d, ac)
        want, _ = ref.search(ref.load(paths[tags[0]]), reads, length)
        for tag in tags:
            h = o.load(paths[tag])
            got = o.search(h, reads, length)
            o.free(h)
            assert np.array_equal(got, want), f"k={k} d={d} tag={tag}"
    # every exact read occurs at least once
    r = want.reshape(-1, 2)[:3000]
    assert (r[:, 1] > r[:, 0]).all()


@pytest.mark.skipif(not helpers.has_ref_tools(), reason="oracle/_ref not built")
def test_config1_full_size_checksum(built, tmp_path):
    """BASELINE config 1/2 index (4 Mbp, d=64) with 100k of its reads: k=1/k=2 x std/AC agree (the survey's
    G1 invariant) and equal the reference searcher; md5 of the result dump is pinned."""
    n = 4_000_000
    text = helpers.synth_text(n, seed=1)
    reads = helpers.synth_reads(text, 2, 100_000, 100)
    o = helpers.Oracle()
    outs = []
    for k in (1, 2):
        paths = helpers.build_reference_indexes(str(tmp_path / f"k{k}"), text, k, 64)
        for ac, tag in ((False, 100), (True, 200)):
            ref = helpers.RefSearcher(k, 64, ac)
            want, _ = ref.search(ref.load(paths[tag]), reads, 100)
            h = o.load(paths[tag]); got = o.search(h, reads, 100); o.free(h)
            assert np.array_equal(got, want)
            outs.append(want)
    for w in outs[1:]:
        assert np.array_equal(w, outs[0])
    assert ((outs[0][1::2] - outs[0][0::2]) >= 1).all()
    pinned = open(os.path.join(helpers.ROOT, "tests", "golden", "config1_100k.md5")).read().split()[0]
    assert helpers.results_text_md5(outs[0]) == pinned


def test_synth_generator_matches_c_tool(built, tmp_path):
    """numpy restatement of fm_synth.h == the C tool's FASTA output."""
    fa, rd = str(tmp_path / "r.fa"), str(tmp_path / "q.fa")
    helpers.run([helpers.FMSYNTH, "ref", fa, "1000", "9"])
    helpers.run([helpers.FMSYNTH, "reads", rd, "1000", "9", "64", "20", "4"])
    text = helpers.synth_text(1000, 9)
    lines = open(fa).read().split("\n")
    assert lines[0] == "> 1000" and "".join(lines[1:]) == text.tobytes().decode()
    o = helpers.Oracle()
    got = o.load_queries(rd, 20, 64)
    assert np.array_equal(got, helpers.synth_reads(text, 4, 64, 20))
    assert open(rd).readline().startswith(">rid1 ")

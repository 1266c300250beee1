"""CPU check of the wide-step table's DESIGN (csrc/fm_wide.cuh), without a GPU: a numpy restatement of what the
builder does -- wide symbols by composing the index's own LF mapping, entries sorted by (symbol, row), buckets by a prefix of
the symbol, one base value per bucket, the three verification rules that mark a bucket exceptional -- on top of the ORACLE's
k-step LF function (oracle/fm_oracle.c fmo_lf, pinned to the reference searchers by tests/test_oracle.py), and then, EXHAUSTIVELY
for every wide symbol and every row boundary X of small reference indexes:

    bucket lookup(sigma, X)  ==  `hops` consecutive reference LF steps of X for the symbols of sigma.

This is the claim the CUDA kernels rest on (the GPU suite checks it through reads; here it is checked for all (sigma, X),
present and absent symbols alike).  It also shows that each verification rule is needed: dropping any of them breaks a fixture.
Test infrastructure only: nothing here is product code."""
import os

import numpy as np
import pytest

import helpers

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden")


def lf_table(oracle, h, nsym, bwtsize):
    """LF[s, X] of the reference searcher for every k-step symbol and every row boundary 0 .. bwtsize."""
    t = np.empty((nsym, bwtsize + 1), dtype=np.int64)
    for s in range(nsym):
        for x in range(bwtsize + 1):
            t[s, x] = oracle.lf(h, s, x)
    return t


class WideModel:
    """The wide-step table of csrc/fm_wide.cuh in numpy (64-bit keys are plenty for the widths used here)."""

    def __init__(self, LF, kbits, hops, prefix_bits, rules=("entries", "buckets", "last")):
        nsym, m = LF.shape
        self.LF, self.kbits, self.hops, self.bwtsize = LF, kbits, hops, m - 1
        self.wbits = kbits * hops
        self.pb = min(prefix_bits, self.wbits)
        self.sub_bits = self.wbits - self.pb
        n = self.bwtsize
        # k-step symbol of every row: the one whose rank steps by one across the row ('$' rows: none)
        step = LF[:, 1:] - LF[:, :-1]
        assert ((step == 0) | (step == 1)).all() and (step.sum(axis=0) <= 1).all()
        sym = np.where(step.sum(axis=0) == 1, step.argmax(axis=0), -1)
        # compose: wide symbol F(i) (hop consumed last in the top bits) and y(i) = the row the chain ends in
        row = np.arange(n)
        F = np.zeros(n, dtype=np.int64)
        ok = np.ones(n, dtype=bool)
        for h in range(hops):
            s = np.where(ok, sym[np.where(ok, row, 0)], -1)
            ok &= s >= 0
            F |= np.where(ok, s, 0).astype(np.int64) << (kbits * h)
            row = np.where(ok, LF[np.where(ok, s, 0), np.where(ok, row, 0)], 0)
        order = np.lexsort((np.arange(n)[ok], F[ok]))              # (F, i) ascending
        self.keys, self.rows, self.y = F[ok][order], np.arange(n)[ok][order], row[ok][order]
        nroots = 1 << self.pb
        bucket = self.keys >> self.sub_bits
        self.bstart = np.searchsorted(bucket, np.arange(nroots + 1))
        # G of every bucket's smallest symbol: the composed rank at X = 0
        self.g0 = np.array([self.compose(b << self.sub_bits, np.zeros(1, dtype=np.int64))[0] for b in range(nroots)])
        cnt = np.diff(self.bstart)
        exc = np.zeros(nroots, dtype=bool)
        if "entries" in rules:                                     # every entry sits where the composed walk says
            expect = self.g0[bucket] + (np.arange(self.keys.size) - self.bstart[bucket])
            np.logical_or.at(exc, bucket[self.y != expect], True)
        if "buckets" in rules:                                     # the next bucket's G continues the count
            exc[:-1] |= self.g0[:-1] + cnt[:-1] != self.g0[1:]
        if "last" in rules:                                        # behind the last bucket the count has reached every row
            exc[-1] |= self.g0[-1] + cnt[-1] != self.bwtsize
        self.exc = exc
        rb = int(self.bwtsize).bit_length()
        self.rb = rb
        self.entries = ((self.keys & ((1 << self.sub_bits) - 1)) << rb) | self.rows

    def compose(self, sigma, X):
        """`hops` consecutive reference LF steps of the row boundaries X for the symbols of sigma (hop 0 = lowest bits)."""
        x = X
        for h in range(self.hops):
            x = self.LF[(sigma >> (self.kbits * h)) & ((1 << self.kbits) - 1), x]
        return x

    def lookup(self, sigma, X):
        """What a kernel computes from the bucket's block (a search tree only changes where the entries are stored)."""
        b = sigma >> self.sub_bits
        if self.exc[b]:
            return self.compose(sigma, X)                          # plain steps
        e = self.entries[self.bstart[b]:self.bstart[b + 1]]
        key = ((sigma & ((1 << self.sub_bits) - 1)) << self.rb) | X
        return self.g0[b] + np.searchsorted(e, key, side="left")


def fixtures():
    out = []
    for name, tags in (("quirk_k1_n124.npz", (100,)), ("quirk_k2_n100.npz", (100, 101)), ("quirk_k1_n250.npz", (100,)), ("quirk_k2_n124.npz", (100,))):
        for tag in tags:
            out.append((name, tag))
    return out


@pytest.mark.parametrize("name,tag", fixtures(), ids=lambda v: str(v))
def test_wide_lookup_equals_composed_reference_lf_for_every_symbol_and_row(name, tag):
    g = np.load(os.path.join(GOLDEN, name))
    k, n = int(g["k"]), int(g["n"])
    o = helpers.Oracle()
    h = o.wrap(g[f"image_{tag}"])
    LF = lf_table(o, h, 4 ** k, n + 1)
    o.free(h)
    X = np.arange(n + 2, dtype=np.int64)
    exceptional = 0
    for bases, pbits in ((4, 3), (6, 5), (6, 12), (8, 7)):
        hops = bases // k
        model = WideModel(LF, 2 * k, hops, pbits)
        exceptional += int(model.exc.sum())
        # the short suffixes are what the exceptional buckets are about: at most `bases` rows carry no wide symbol
        assert n + 1 - model.keys.size <= bases
        for sigma in range(1 << model.wbits):
            got, want = model.lookup(sigma, X), model.compose(sigma, X)
            assert np.array_equal(got, want), f"{name} tag {tag} bases {bases} prefix {pbits} sigma {sigma:#x}"
    assert exceptional > 0, "tiny references are meant to have buckets a short suffix sorts into"


def test_each_verification_rule_is_needed(tmp_path):
    """Without the bucket rule or the last-bucket rule some index answers wrongly: the builder's checks are not decoration.
    (The last-bucket rule was missing at first; the GPU fuzz on tiny references found it.  Its witness here: a text whose
    longest run of T sits at its very end, so that a suffix shorter than the step sorts behind every full context.)"""
    broken = {"entries": False, "buckets": False, "last": False}
    images = [(int(np.load(os.path.join(GOLDEN, name))["k"]), np.load(os.path.join(GOLDEN, name))[f"image_{tag}"]) for name, tag in fixtures()]
    if helpers.has_ref_tools():
        rng = np.random.default_rng(4)
        body = np.frombuffer(b"ACG", dtype=np.uint8)[rng.integers(0, 3, 150)]
        text = np.concatenate([body, np.frombuffer(b"TGTTACTTTT", dtype=np.uint8)])
        for k in (1, 2):
            paths = helpers.build_reference_indexes(str(tmp_path / f"k{k}"), text, k, 32)
            images.append((k, np.fromfile(paths[100], dtype=np.uint32)))
    for k, image in images:
        n = int(image[2]) - 1
        o = helpers.Oracle()
        h = o.wrap(image)
        LF = lf_table(o, h, 4 ** k, n + 1)
        o.free(h)
        X = np.arange(n + 2, dtype=np.int64)
        for bases, pbits in ((4, 2), (4, 3), (6, 3), (6, 5)):
            for dropped in broken:
                rules = tuple(r for r in broken if r != dropped)
                model = WideModel(LF, 2 * k, bases // k, pbits, rules)
                for sigma in range(1 << model.wbits):
                    if not np.array_equal(model.lookup(sigma, X), model.compose(sigma, X)):
                        broken[dropped] = True
                        break
    assert broken["buckets"] and (broken["last"] or not helpers.has_ref_tools()), broken
    # (a short suffix inside a bucket shifts every entry behind it AND the next bucket's G, so the bucket rule usually catches
    #  what the entry rule catches; the entry rule is kept because it is the direct statement of what a leaf assumes)


def test_model_lf_is_the_reference_search():
    """The LF table the model is built on reproduces the committed reference outputs when it is used as a searcher."""
    g = np.load(os.path.join(GOLDEN, "quirk_k2_n124.npz"))
    k, n, length = int(g["k"]), int(g["n"]), int(g["length"])
    o = helpers.Oracle()
    h = o.wrap(g["image_100"])
    LF = lf_table(o, h, 4 ** k, n + 1)
    o.free(h)
    code = np.zeros(256, dtype=np.int64)
    for c, v in zip(b"ACGT", range(4)):
        code[c] = v
        code[c + 32] = v
    reads = g["reads"].reshape(-1, length)
    want = g["expected_std"].reshape(-1, 2)
    for q in range(reads.shape[0]):
        L, R = 0, n + 1
        for s in range(length // k):
            sym = 0
            for j in range(k):                                     # symbol of LF step s: base len-1-(k*s+j) at bits 2j
                sym |= int(code[reads[q, length - 1 - (k * s + j)]]) << (2 * j)
            L, R = int(LF[sym, L]), int(LF[sym, R])
        assert (L, R) == (int(want[q, 0]), int(want[q, 1])), q

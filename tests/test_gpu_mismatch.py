"""GPU tests of the one-mismatch search (csrc/fm_mismatch.cu).  The reference matches exactly only,
so the checker is brute force: every window of the text is compared with the read (Hamming distance 0 / 1).   pytest -m gpu"""
import ctypes as C
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


def brute_force(text, reads, length):
    """per read: (occurrences at distance 0, occurrences at distance 1, distinct distance-1 variants that occur)"""
    win = np.lib.stride_tricks.sliding_window_view(text, length)
    out = []
    for r in reads.reshape(-1, length):
        d = (win != r[None, :]).sum(axis=1)
        one = np.flatnonzero(d == 1)
        variants = {bytes(win[i]) for i in one}
        out.append((int((d == 0).sum()), int(one.size), len(variants)))
    return np.array(out, dtype=np.int64)


@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("mode", ["coop", "sparse", "fused"])
def test_one_mismatch_search_against_brute_force(pkg, k, mode):
    import torch
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", f"small_k{k}_d64.npz"))
    n = int(g["n"])
    text = helpers.synth_text(n, seed=7 + k)
    L = pkg.lib()
    rng = np.random.default_rng(9)
    idx = pkg.DeviceIndex.from_image(g["image_100"])
    if mode == "sparse":
        idx.sparsify(4, 0, 0)
    if mode == "fused":
        idx.fuse(4, 2)
    v = pkg.variant({"coop": pkg.MODE_COOP, "sparse": pkg.MODE_SPARSE, "fused": pkg.MODE_FUSED}[mode])
    for length in (8, 12, 20):
        nq = 150
        starts = rng.integers(0, n - length, nq)
        reads = np.concatenate([text[s:s + length] for s in starts]).copy()
        mut = rng.integers(0, nq, nq // 2) * length + rng.integers(0, length, nq // 2)     # half of the reads get one substitution
        reads[mut] = ACGT[rng.integers(0, 4, mut.size)]
        want = brute_force(text, reads, length)
        idx.prepare(length)
        d_ascii = torch.from_numpy(reads).cuda()
        wpq = L.fmgpu_words_per_query(length)
        d_packed = torch.empty(nq * wpq, dtype=torch.int32, device="cuda")
        pkg.check(L.fmgpu_pack_queries_device(0, d_ascii.data_ptr(), nq, length, d_packed.data_ptr(), None), "pack")
        d_out = torch.zeros(nq * 4, dtype=torch.int32, device="cuda")
        d_var = torch.zeros(nq * 3 * length * 2, dtype=torch.int32, device="cuda")
        for keep_variants in (False, True):
            pkg.check(L.fmgpu_search_device_mm1(idx.handle, d_packed.data_ptr(), nq, length, d_out.data_ptr(),
                                                d_var.data_ptr() if keep_variants else None, C.byref(v), None), "mm1")
            out = d_out.cpu().numpy().view(np.uint32).reshape(nq, 4).astype(np.int64)
            assert np.array_equal(np.maximum(out[:, 1] - out[:, 0], 0), want[:, 0]), f"exact hits, k={k} {mode} len={length}"
            assert np.array_equal(out[:, 3], want[:, 1]), f"distance-1 occurrences, k={k} {mode} len={length}"
            assert np.array_equal(out[:, 2], want[:, 2]), f"distance-1 variants, k={k} {mode} len={length}"
        var = d_var.cpu().numpy().view(np.uint32).reshape(nq, 3 * length, 2).astype(np.int64)
        assert np.array_equal(np.maximum(var[:, :, 1] - var[:, :, 0], 0).sum(axis=1), want[:, 1])
        # variant j = 3 t + s changes base len-1-t: check one read's variant intervals against exact searches of the variant strings
        q = 3
        b = pkg.DeviceBatch(0, 3 * length, length, k)
        vs = []
        code = {65: 0, 67: 1, 71: 2, 84: 3}
        for t in range(length):
            for s in range(3):
                r = reads[q * length:(q + 1) * length].copy()
                r[length - 1 - t] = ACGT[(code[int(r[length - 1 - t])] + 1 + s) & 3]
                vs.append(r)
        b.upload_ascii(np.concatenate(vs))
        b.search(idx, pkg.variant(pkg.MODE_COOP))
        assert np.array_equal(b.download().reshape(-1, 2), var[q].astype(np.uint32))
        b.free()
    idx.free()

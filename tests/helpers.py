"""Shared test plumbing: loads the product library, the C oracle and the
reference tools (oracle/_ref), and builds small seeded data sets with the
UNMODIFIED reference index builders.

Nothing here reads /root/reference at run time: oracle/_ref holds binaries that
were compiled from it in the build container and travel with the repo.
"""
import ctypes as C
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
PKG_NAME = "k-step_fm-index_b200"
FMSYNTH = os.path.join(ROOT, PKG_NAME, "bin", "fmsynth")
TAG_SUFFIX = {100: "", 101: ".interleaving", 200: ".ac", 201: ".interleaving.ac"}

_built = False


def ensure_built():
    global _built
    if _built:
        return
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    _built = True


def pkg():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module(PKG_NAME)


def has_ref_tools():
    return os.path.exists(os.path.join(REF_DIR, "gfmiBaseLine_64bases_2step"))


# --------------------------------------------------------------------------- #
# C oracle (oracle/liboracle.so)
# --------------------------------------------------------------------------- #
class Oracle:
    def __init__(self):
        self._keep = {}
        self.lib = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        L = self.lib
        L.fmo_load_index.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.fmo_wrap_image.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.fmo_free_index.argtypes = [C.c_void_p]
        L.fmo_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.fmo_lf.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.fmo_lf.restype = C.c_uint32
        L.fmo_count_sectors.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.fmo_count_sectors.restype = C.c_uint64
        L.fmo_load_queries.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.c_void_p]
        L.fmo_write_results.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32]

    def load(self, fn):
        h = C.c_void_p()
        rc = self.lib.fmo_load_index(os.fsencode(fn), C.byref(h))
        assert rc == 0, f"fmo_load_index({fn}) -> {rc}"
        return h

    def wrap(self, image):
        image = np.ascontiguousarray(image, dtype=np.uint32)
        h = C.c_void_p()
        rc = self.lib.fmo_wrap_image(image.ctypes.data, image.size, C.byref(h))
        assert rc == 0, f"fmo_wrap_image -> {rc}"
        self._keep[h.value] = image            # the oracle index points into this array
        return h

    def free(self, h):
        self.lib.fmo_free_index(h)
        self._keep.pop(h.value, None)

    def search(self, h, ascii_bases, length):
        a = np.ascontiguousarray(ascii_bases, dtype=np.uint8).reshape(-1)
        nq = a.size // length
        out = np.zeros(2 * nq, dtype=np.uint32)
        self.lib.fmo_search(h, a.ctypes.data, nq, length, out.ctypes.data)
        return out

    def lf(self, h, sigma, x):
        return self.lib.fmo_lf(h, sigma, x)

    def count_sectors(self, h, ascii_bases, length, block_rows, blocks_per_sector):
        a = np.ascontiguousarray(ascii_bases, dtype=np.uint8).reshape(-1)
        return self.lib.fmo_count_sectors(h, a.ctypes.data, a.size // length, length, block_rows, blocks_per_sector)

    def load_queries(self, fn, length, num):
        out = np.empty(num * length, dtype=np.uint8)
        rc = self.lib.fmo_load_queries(os.fsencode(fn), length, num, out.ctypes.data)
        assert rc == 0, f"fmo_load_queries -> {rc}"
        return out


# --------------------------------------------------------------------------- #
# the reference's own CPU searcher, in-process (oracle/_ref/libref_search_*.so)
# --------------------------------------------------------------------------- #
class RefSearcher:
    def __init__(self, k, d, ac):
        path = os.path.join(REF_DIR, f"libref_search_k{k}_d{d}_{'ac' if ac else 'std'}.so")
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L = self.lib
        L.loadIndex.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ref_wrap_index_image.argtypes = [C.c_void_p]
        L.ref_wrap_index_image.restype = C.c_void_p
        L.ref_wrap_queries.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.ref_wrap_queries.restype = C.c_void_p
        L.ref_wrap_results.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_wrap_results.restype = C.c_void_p
        L.ref_search_parallel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.ref_search_parallel.restype = C.c_double
        L.ref_max_threads.restype = C.c_int32
        assert L.ref_cfg_steps() == k and L.ref_cfg_chunk() == d and L.ref_cfg_ac() == int(ac)

    def load(self, fn):
        h = C.c_void_p()
        # the reference loader prints the header to stdout
        rc = self.lib.loadIndex(os.fsencode(fn), C.byref(h))
        assert rc == 0, f"reference loadIndex({fn}) -> {rc}"
        return h

    def wrap_image(self, image):
        image = np.ascontiguousarray(image, dtype=np.uint32)
        self._keep = getattr(self, "_keep", []) + [image]
        return C.c_void_p(self.lib.ref_wrap_index_image(image.ctypes.data))

    def search(self, index, ascii_bases, length, iters=1, threads=0):
        a = np.ascontiguousarray(ascii_bases, dtype=np.uint8).reshape(-1)
        nq = a.size // length
        out = np.zeros(2 * nq, dtype=np.uint32)
        q = C.c_void_p(self.lib.ref_wrap_queries(a.ctypes.data, nq, length))
        r = C.c_void_p(self.lib.ref_wrap_results(out.ctypes.data, nq))
        secs = self.lib.ref_search_parallel(index, q, r, iters, threads)
        return out, secs


# --------------------------------------------------------------------------- #
# data sets
# --------------------------------------------------------------------------- #
def run(cmd, cwd=None):
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, f"{cmd} failed:\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}"
    return p.stdout


def synth_text(n, seed):
    """The text of fm_synth.h as a numpy uint8 array of ASCII bases (vectorised restatement)."""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[(x >> np.uint64(62)).astype(np.int64)]


def synth_read_starts(seed, num, n, length, first=0):
    j = np.arange(first, first + num, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = ((np.uint64(seed) ^ np.uint64(0xA5A5A5A55A5A5A5A)) + j + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x % np.uint64(n - length + 1)).astype(np.int64)


def synth_reads(text, seed, num, length, first=0):
    starts = synth_read_starts(seed, num, text.size, length, first)
    return text[starts[:, None] + np.arange(length)[None, :]].reshape(-1)


def write_fasta_ref(path, text):
    with open(path, "wb") as f:
        f.write(b"> %d" % text.size)
        for i in range(0, text.size, 70):
            f.write(b"\n" + text[i:i + 70].tobytes())
        f.write(b"\n")


def write_fasta_reads(path, reads, length):
    r = np.ascontiguousarray(reads, dtype=np.uint8).reshape(-1, length)
    with open(path, "wb") as f:
        for i in range(r.shape[0]):
            f.write(b">rid%d\n" % (i + 1) + r[i].tobytes() + b"\n")


def build_reference_indexes(workdir, text, k, d, name="ref.fa"):
    """Runs the reference builder + both transformers on `text`; returns {tag: path}."""
    os.makedirs(workdir, exist_ok=True)
    fa = os.path.join(workdir, name)
    write_fasta_ref(fa, text)
    n = text.size
    run([os.path.join(REF_DIR, f"gfmiBaseLine_{d}bases_{k}step"), name, str(n)], cwd=workdir)
    base = f"{fa}.{n}.{d}fmi{k}steps.fmi"
    run([os.path.join(REF_DIR, f"tfmiBMP_{d}bases_{k}step"), os.path.basename(base)], cwd=workdir)
    run([os.path.join(REF_DIR, f"tfmiAC_{d}bases_{k}step"), os.path.basename(base)], cwd=workdir)
    return {tag: base + suf for tag, suf in TAG_SUFFIX.items()}


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def results_text_md5(res):
    """md5 of the reference's text dump (common/common.c:201-220) of an (L,R) array."""
    r = np.asarray(res, dtype=np.uint32).reshape(-1, 2)
    h = hashlib.md5()
    h.update(b"%d\n" % r.shape[0])
    for i in range(0, r.shape[0], 1 << 16):
        h.update("".join(f"{a} {b}\n" for a, b in r[i:i + (1 << 16)].tolist()).encode())
    return h.hexdigest()

"""oracle/bindings.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings of the two checkers: our C restatement (oracle/liboracle.so) and the reference's own CPU
searcher compiled from /root/reference into oracle/_ref/libref_search_*.so.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


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

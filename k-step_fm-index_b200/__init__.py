"""k-step_fm-index_b200 -- Python host-side mirror of the reference's search interface.

Thin ctypes binding over ``lib/libfmindex_b200.so`` (C host layer + hand-written
CUDA for sm_100a).  The function names are the reference's own
(``common/interface.h:27-41``, ``common/common.h:83-94``): ``loadIndex``,
``loadQueries``, ``initResults``, ``transferCPUtoGPU``, ``searchIndexGPU``,
``transferGPUtoCPU``, ``saveResults``, ``free*``; the ``fmgpu_*`` functions are
the thin C ABI to the CUDA side (``include/fmindex_b200.h`` PART 2).

There is no CPU search path here: if the shared library is missing the import
fails loudly, and on a box without an sm_100 GPU every device entry point
returns ``FM_E_CUDA``.

The directory name contains hyphens, so import it with
``importlib.import_module("k-step_fm-index_b200")`` (repo root on ``sys.path``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FMGPU_LIB") or os.path.join(_HERE, "lib", "libfmindex_b200.so")   # $FMGPU_LIB: e.g. the -DFM_DEBUG_BOUNDS build
CSRC = os.path.join(_HERE, "csrc")
BIN = os.path.join(_HERE, "bin")

FM_SUCCESS = 0
FM_E_READING_FMI = 5
FM_E_CUDA = 50
FM_E_BAD_ARGUMENT = 51
FM_E_UNSUPPORTED_INDEX = 52
FM_E_QUERY_SHAPE = 53
MODE_TASK, MODE_COOP, MODE_FUSED, MODE_SPARSE, MODE_WIDE = 0, 1, 2, 3, 4


def build_native(verbose=False):
    """Compiles the C host layer and the CUDA kernels for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise RuntimeError("building libfmindex_b200.so failed")
    return LIB_PATH


class qrys_t(C.Structure):          # common/common.h:64-69
    _fields_ = [("num", C.c_uint32), ("size", C.c_uint32), ("h_queries", C.c_void_p), ("d_queries", C.c_void_p)]


class res_t(C.Structure):           # common/common.h:77-81
    _fields_ = [("num", C.c_uint32), ("h_results", C.c_void_p), ("d_results", C.c_void_p)]


class fmi_t(C.Structure):           # src/fmIndexCPUBaseline.c:54-69 + extensions
    _fields_ = [("steps", C.c_uint32), ("bwtsize", C.c_uint32), ("ncounters", C.c_uint32), ("nentries", C.c_uint32),
                ("chunk", C.c_uint32), ("nbitmaps", C.c_uint32),
                ("h_dollarPositionBWT", C.POINTER(C.c_uint32)), ("h_dollarBaseBWT", C.POINTER(C.c_uint32)),
                ("h_modposdollarBWT", C.POINTER(C.c_uint32)), ("h_index", C.c_void_p),
                ("d_dollarPositionBWT", C.c_void_p), ("d_dollarBaseBWT", C.c_void_p), ("d_modposdollarBWT", C.c_void_p),
                ("d_index", C.c_void_p), ("tag", C.c_uint32), ("entry_words", C.c_uint32)]


class fmgpu_variant_t(C.Structure):
    _fields_ = [("mode", C.c_int32), ("queries_per_thread", C.c_int32), ("threads_per_block", C.c_int32), ("feed", C.c_int32)]


class fmgpu_index_meta_t(C.Structure):
    _fields_ = [("steps", C.c_uint32), ("bwtsize", C.c_uint32), ("nsymbols", C.c_uint32), ("nblocks", C.c_uint32),
                ("source_tag", C.c_uint32), ("quirk_start", C.c_uint32), ("quirk_mask", C.c_uint32), ("source_steps", C.c_uint32),
                ("nbytes", C.c_uint64), ("fused_bases", C.c_uint32), ("fused_lanes", C.c_uint32), ("fused_bytes", C.c_uint64),
                ("tail_valid", C.c_uint32), ("tail_row", C.c_uint32), ("tail_base", C.c_uint32), ("tail_const", C.c_uint32 * 4),
                ("start_bases", C.c_uint32),
                ("sparse_bases", C.c_uint32), ("sparse_lambda", C.c_uint32), ("sparse_bytes", C.c_uint64),
                ("sparse_blocks", C.c_uint64), ("sparse_overflow", C.c_uint64), ("sparse_start_bases", C.c_uint32),
                ("sparse_lanes", C.c_uint32), ("tail_bytes", C.c_uint64),
                ("sparse_uniform_nb", C.c_uint32), ("sa_rate", C.c_uint32), ("sa_bytes", C.c_uint64),
                ("sparse_tree_nodes", C.c_uint64), ("sparse_tree_rows", C.c_uint64), ("sparse_tree_depth", C.c_uint32),
                ("reserved1", C.c_uint32), ("derived_bytes", C.c_uint64), ("budget_bytes", C.c_uint64),
                ("wide_bases", C.c_uint32), ("wide_prefix_bits", C.c_uint32), ("wide_row_bits", C.c_uint32), ("wide_tree_depth", C.c_uint32),
                ("wide_bytes", C.c_uint64), ("wide_blocks", C.c_uint64), ("wide_overflow", C.c_uint64),
                ("wide_tree_nodes", C.c_uint64), ("wide_tree_rows", C.c_uint64), ("wide_exceptional", C.c_uint64),
                ("wide_lanes", C.c_uint32), ("wide_entry_words", C.c_uint32),
                ("wide_block_entries", C.c_uint32), ("reserved3", C.c_uint32)]


class fmgpu_transfer_stats_t(C.Structure):
    _fields_ = [("ndev", C.c_int32), ("searches", C.c_int32), ("index_h2d_reblock_s", C.c_double), ("peer_copy_s", C.c_double * 16),
                ("table_build_s", C.c_double * 16), ("queries_h2d_pack_s", C.c_double), ("results_d2h_s", C.c_double),
                ("search_ms", C.c_float * 16), ("index_file_bytes", C.c_uint64), ("table_bytes", C.c_uint64),
                ("query_bytes", C.c_uint64), ("result_bytes", C.c_uint64), ("context_init_s", C.c_double * 16), ("replicate_s", C.c_double * 16)]


class fmgpu_mm1_t(C.Structure):
    _fields_ = [("L", C.c_uint32), ("R", C.c_uint32), ("variants_found", C.c_uint32), ("occurrences_1mm", C.c_uint32)]


class fmgpu_pipeline_stats_t(C.Structure):
    _fields_ = [("calls", C.c_uint64), ("last_feed", C.c_int32), ("pad", C.c_int32), ("last_seconds", C.c_double),
                ("last_reads_host_packed", C.c_uint64), ("last_reads_ascii_over_link", C.c_uint64),
                ("host_pack_seconds_per_read", C.c_double)]


_VP, _VPP = C.c_void_p, C.POINTER(C.c_void_p)
_U32P = C.POINTER(C.c_uint32)

# name -> (restype, argtypes): every symbol include/fmindex_b200.h declares
PROTOTYPES = {
    # PART 1 -- reference interface
    "loadIndex": (C.c_int32, [C.c_char_p, _VPP]),
    "freeIndex": (C.c_int32, [_VPP]),
    "loadQueries": (C.c_int32, [C.c_char_p, C.c_uint32, C.c_uint32, _VPP]),
    "initResults": (C.c_int32, [C.c_uint32, _VPP]),
    "writeResults": (C.c_int32, [C.c_char_p, _VP, C.c_uint32]),
    "loadResults": (C.c_int32, [C.c_char_p, _VPP]),
    "saveResults": (C.c_int32, [C.c_char_p, _VP, _VP]),
    "freeQueries": (C.c_int32, [_VPP]),
    "freeResults": (C.c_int32, [_VPP]),
    "errorCommon": (C.c_char_p, [C.c_int32]),
    "sampleTime": (C.c_double, []),
    "transferCPUtoGPU": (C.c_int32, [_VP, _VP, _VP]),
    "searchIndexGPU": (None, [_VP, _VP, _VP]),
    "transferGPUtoCPU": (C.c_int32, [_VP]),
    "freeIndexGPU": (C.c_int32, [_VPP]),
    "freeQueriesGPU": (C.c_int32, [_VPP]),
    "freeResultsGPU": (C.c_int32, [_VPP]),
    # PART 2 -- thin C ABI to CUDA
    "fmgpu_device_count": (C.c_int32, []),
    "fmgpu_device_warmup": (C.c_int32, [C.c_int32]),
    "fmgpu_set_devices": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32]),
    "fmgpu_set_variant": (C.c_int32, [C.POINTER(fmgpu_variant_t)]),
    "fmgpu_last_error": (C.c_char_p, []),
    "fmgpu_index_create": (C.c_int32, [C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       _U32P, _U32P, _VP, _VPP]),
    "fmgpu_index_create_from_device": (C.c_int32, [C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                   C.c_uint32, _U32P, _U32P, _VP, _VPP]),
    "fmgpu_index_replicate": (C.c_int32, [_VP, C.c_int32, _VPP]),
    "fmgpu_last_peer_copy_seconds": (C.c_double, []),
    "fmgpu_index_alloc_like": (C.c_int32, [C.c_int32, C.POINTER(fmgpu_index_meta_t), _VPP]),
    "fmgpu_index_fuse": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, C.c_uint64]),
    "fmgpu_index_unfuse": (C.c_int32, [_VP]),
    "fmgpu_index_sparsify": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, C.c_uint32]),
    "fmgpu_index_unsparsify": (C.c_int32, [_VP]),
    "fmgpu_index_widen": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, C.c_uint32]),
    "fmgpu_index_unwiden": (C.c_int32, [_VP]),
    "fmgpu_wide_bases_for": (C.c_uint32, [_VP, C.c_uint32]),
    "fmgpu_wide_bases_for_words": (C.c_uint32, [_VP, C.c_uint32, C.c_uint32]),
    "fmgpu_index_wide_serves": (C.c_int32, [_VP, C.c_uint32]),
    "fmgpu_count_fetches_wide_device": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, _VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                    C.POINTER(C.c_uint64)]),
    "fmgpu_index_build_sa": (C.c_int32, [_VP]),
    "fmgpu_index_build_sa_sampled": (C.c_int32, [_VP, C.c_uint32]),
    "fmgpu_index_drop_sa": (C.c_int32, [_VP]),
    "fmgpu_index_sa": (_VP, [_VP]),
    "fmgpu_index_download_sa": (C.c_int32, [_VP, _VP]),
    "fmgpu_locate_device": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, _VP, _VP]),
    "fmgpu_batch_locate": (C.c_int32, [_VP, _VP, C.c_uint32, _VP, _VP]),
    "fmgpu_index_get_meta": (C.c_int32, [_VP, C.POINTER(fmgpu_index_meta_t)]),
    "fmgpu_index_blocks": (_VP, [_VP]),
    "fmgpu_index_device": (C.c_int32, [_VP]),
    "fmgpu_index_free": (C.c_int32, [_VPP]),
    "fmgpu_batch_create": (C.c_int32, [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, _VPP]),
    "fmgpu_batch_upload_ascii": (C.c_int32, [_VP, _VP]),
    "fmgpu_batch_search": (C.c_int32, [_VP, _VP, C.POINTER(fmgpu_variant_t)]),
    "fmgpu_batch_sync": (C.c_int32, [_VP]),
    "fmgpu_batch_download": (C.c_int32, [_VP, _VP]),
    "fmgpu_batch_search_timed": (C.c_int32, [_VP, _VP, C.POINTER(fmgpu_variant_t), C.c_int32, C.POINTER(C.c_float)]),
    "fmgpu_batch_count_fetches": (C.c_int32, [_VP, _VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fmgpu_batch_packed": (_VP, [_VP]),
    "fmgpu_batch_results": (_VP, [_VP]),
    "fmgpu_batch_stream": (_VP, [_VP]),
    "fmgpu_batch_free": (C.c_int32, [_VPP]),
    "fmgpu_words_per_query": (C.c_uint32, [C.c_uint32]),
    "fmgpu_pack_queries_device": (C.c_int32, [C.c_int32, _VP, C.c_uint64, C.c_uint32, _VP, _VP]),
    "fmgpu_search_device": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, C.POINTER(fmgpu_variant_t), _VP]),
    "fmgpu_search_host": (C.c_int32, [_VPP, C.c_int32, _VP, C.c_uint64, C.c_uint32, _VP, C.POINTER(fmgpu_variant_t)]),
    "fmgpu_host_alloc": (_VP, [C.c_size_t]),
    "fmgpu_host_free": (None, [_VP]),
    "fmgpu_host_register": (C.c_int32, [_VP, C.c_size_t]),
    "fmgpu_host_unregister": (C.c_int32, [_VP]),
    "fmgpu_gather_probe": (C.c_int32, [C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]),
    "fmgpu_gather_probe_ex": (C.c_int32, [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]),
    "fm_hostpack_reads": (None, [_VP, C.c_uint64, C.c_uint32, _VP, C.c_int]),
    "fm_hostpack_reads_scalar": (None, [_VP, C.c_uint64, C.c_uint32, _VP]),
    "fmgpu_search_host_packed": (C.c_int32, [_VPP, C.c_int32, _VP, C.c_uint64, C.c_uint32, _VP, C.POINTER(fmgpu_variant_t)]),
    "fmgpu_release_pipeline": (C.c_int32, []),
    "fmgpu_pipeline_create": (C.c_int32, [_VPP]),
    "fmgpu_pipeline_search_host": (C.c_int32, [_VP, _VPP, C.c_int32, _VP, C.c_uint64, C.c_uint32, _VP, C.POINTER(fmgpu_variant_t)]),
    "fmgpu_pipeline_search_host_packed": (C.c_int32, [_VP, _VPP, C.c_int32, _VP, C.c_uint64, C.c_uint32, _VP, C.POINTER(fmgpu_variant_t)]),
    "fmgpu_pipeline_get_stats": (C.c_int32, [_VP, C.POINTER(fmgpu_pipeline_stats_t)]),
    "fmgpu_pipeline_free": (C.c_int32, [_VPP]),
    "fmgpu_index_prepare": (C.c_int32, [_VP, C.c_uint32]),
    "fmgpu_search_device_mm1": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, _VP, C.POINTER(fmgpu_variant_t), _VP]),
    "fmgpu_set_table_budget": (C.c_int32, [C.c_uint64]),
    "fmgpu_get_transfer_stats": (C.c_int32, [C.POINTER(fmgpu_transfer_stats_t)]),
    "fmgpu_search_index": (C.c_int32, [_VP, _VP, _VP]),
    "fmgpu_batch_search_timed_async": (C.c_int32, [_VP, _VP, C.POINTER(fmgpu_variant_t)]),
    "fmgpu_batch_last_ms": (C.c_int32, [_VP, C.POINTER(C.c_float)]),
    "fm_hostpack_stream": (None, [_VP, C.c_uint64, _VP, C.c_int]),
    "fmgpu_unstream_device": (C.c_int32, [C.c_int32, _VP, C.c_uint64, C.c_uint32, _VP, _VP]),
    "fm_hostpack_set_prefetch": (None, [C.c_int]),
    "fm_hostpack_set_streams": (None, [C.c_int]),
    "fm_hostpack_has_simd": (C.c_int, []),
    "fm_host_read_bandwidth": (C.c_double, [_VP, C.c_uint64, C.c_int, C.c_int]),
    "fm_hostpack_threads": (C.c_int, []),
    "fmgpu_gather_probe_local": (C.c_int32, [C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]),
    "fmgpu_count_fetches_device": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, _VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fmgpu_count_fetches_fused_device": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, _VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fmgpu_count_fetches_sparse_device": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, _VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                      C.POINTER(C.c_uint64)]),
    "fmgpu_build_from_text": (C.c_int32, [C.c_int32, _VP, C.c_uint64, C.c_uint32, C.c_uint32, _VPP]),
    "fmgpu_build_from_synth": (C.c_int32, [C.c_int32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _VPP]),
    "fmgpu_build_image_words": (C.c_uint64, [_VP]),
    "fmgpu_build_image_device": (_VP, [_VP]),
    "fmgpu_build_download": (C.c_int32, [_VP, _VP]),
    "fmgpu_build_to_index": (C.c_int32, [_VP, _VPP]),
    "fmgpu_build_transform": (C.c_int32, [_VP, C.c_uint32, _VPP]),
    "fmgpu_build_save": (C.c_int32, [_VP, C.c_char_p]),
    "fmgpu_build_free": (C.c_int32, [_VPP]),
    "fmgpu_build_last_error": (C.c_char_p, []),
    "fmgpu_synth_reads_device": (C.c_int32, [C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, _VP, _VP]),
}

_lib = None


def lib():
    """The loaded C-ABI library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


class FMError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().errorCommon(code)
        detail = lib().fmgpu_last_error() if code >= 19 else None       # the fmgpu_* layer says why (per thread)
        super().__init__(f"{where}: error {code}: {msg.decode() if msg else '?'}" + (f" ({detail.decode()})" if detail and detail != b"no error" else ""))


def check(code, where):
    if code != FM_SUCCESS:
        raise FMError(code, where)


FEED_AUTO, FEED_ASCII, FEED_HOSTPACK, FEED_HYBRID = 0, 1, 2, 3


def variant(mode=MODE_TASK, queries_per_thread=0, threads_per_block=0, feed=FEED_AUTO):
    """feed only matters to search_host: how host ASCII reads reach the GPU (see fmindex_b200.h)."""
    return fmgpu_variant_t(mode, queries_per_thread, threads_per_block, feed)


# --------------------------------------------------------------------------- #
# reference-shaped convenience wrappers (same call sequence as
# common/searchQueries.c:64-125); handles are c_void_p like the reference's void*
# --------------------------------------------------------------------------- #
def loadIndex(fn):
    h = C.c_void_p()
    check(lib().loadIndex(os.fsencode(fn), C.byref(h)), f"loadIndex({fn})")
    return h


def loadQueries(fn, size, num):
    h = C.c_void_p()
    check(lib().loadQueries(os.fsencode(fn), size, num, C.byref(h)), f"loadQueries({fn})")
    return h


def queriesFromArray(ascii_bases, size):
    """qrys_t around a numpy uint8 array of num*size ASCII bases (no file involved)."""
    a = np.ascontiguousarray(ascii_bases, dtype=np.uint8).reshape(-1)
    assert a.size % size == 0
    q = qrys_t(a.size // size, size, a.ctypes.data, None)
    q._keep = a
    return q


def index_from_image(image_u32):
    """fmi_t (what loadIndex returns) around the words of an index FILE held in a numpy uint32 array: no file involved.
    The array and the '$' tables are kept alive by the returned structure; pass C.byref(fmi) where a void* index goes."""
    im = np.ascontiguousarray(image_u32, dtype=np.uint32)
    tag, k, bwtsize, ncnt, nent, d = (int(v) for v in im[:6])
    f = fmi_t()
    f.tag, f.steps, f.bwtsize, f.ncounters, f.nentries, f.chunk = tag, k, bwtsize, ncnt, nent, d
    f.nbitmaps = 2 * (d // 32)
    f.entry_words = f.nbitmaps * k + ncnt
    if im.size != 6 + 2 * k + nent * f.entry_words:
        raise FMError(FM_E_READING_FMI, "index_from_image (image size does not match its header)")
    f._dpos = (C.c_uint32 * k)(*[int(v) for v in im[6:6 + k]])
    f._dbase = (C.c_uint32 * k)(*[int(v) for v in im[6 + k:6 + 2 * k]])
    f._dmod = (C.c_uint32 * k)(*[int(v) // d for v in im[6:6 + k]])
    f.h_dollarPositionBWT = C.cast(f._dpos, C.POINTER(C.c_uint32))
    f.h_dollarBaseBWT = C.cast(f._dbase, C.POINTER(C.c_uint32))
    f.h_modposdollarBWT = C.cast(f._dmod, C.POINTER(C.c_uint32))
    f._keep = im
    f.h_index = im[6 + 2 * k:].ctypes.data
    return f


def initResults(num):
    h = C.c_void_p()
    check(lib().initResults(num, C.byref(h)), "initResults")
    return h


def resultsArray(results, copy=True):
    r = C.cast(results, C.POINTER(res_t)).contents
    a = np.ctypeslib.as_array(C.cast(r.h_results, _U32P), shape=(2 * r.num,))
    return a.copy() if copy else a


def index_fields(index):
    return C.cast(index, C.POINTER(fmi_t)).contents


def search_files(index_fn, queries_fn, size, num, devices=None, var=None):
    """The reference main()'s GPU flow; returns the (L,R) array of 2*num uint32."""
    L = lib()
    if devices is not None:
        arr = (C.c_int32 * len(devices))(*devices)
        check(L.fmgpu_set_devices(arr, len(devices)), "fmgpu_set_devices")
    L.fmgpu_set_variant(C.byref(var) if var is not None else None)
    idx = loadIndex(index_fn)
    qry = loadQueries(queries_fn, size, num)
    res = initResults(num)
    try:
        check(L.transferCPUtoGPU(idx, qry, res), "transferCPUtoGPU")
        L.searchIndexGPU(idx, qry, res)
        check(L.transferGPUtoCPU(res), "transferGPUtoCPU")
        return resultsArray(res)
    finally:
        L.freeIndexGPU(C.byref(idx)); L.freeQueriesGPU(C.byref(qry)); L.freeResultsGPU(C.byref(res))
        L.freeIndex(C.byref(idx)); L.freeQueries(C.byref(qry)); L.freeResults(C.byref(res))


class DeviceIndex:
    """Device-resident SB96 index on one GPU (fmgpu_index_t)."""

    def __init__(self, handle):
        self.handle = handle

    @classmethod
    def from_file(cls, fn, device=0):
        idx = loadIndex(fn)
        try:
            f = index_fields(idx)
            h = C.c_void_p()
            check(lib().fmgpu_index_create(device, f.tag, f.steps, f.chunk, f.bwtsize, f.ncounters, f.nentries,
                                           f.h_dollarPositionBWT, f.h_dollarBaseBWT, f.h_index, C.byref(h)),
                  "fmgpu_index_create")
        finally:
            lib().freeIndex(C.byref(idx))
        return cls(h)

    @classmethod
    def from_image(cls, image_u32, device=0):
        """image = the words of an index FILE (header + entries) as a numpy uint32 array."""
        im = np.ascontiguousarray(image_u32, dtype=np.uint32)
        if im.ndim != 1 or im.size < 6:
            raise FMError(FM_E_READING_FMI, "DeviceIndex.from_image (image shorter than a header)")
        tag, k, bwtsize, ncnt, nent, d = (int(v) for v in im[:6])
        if not 1 <= k <= 4 or d == 0 or d % 32:
            raise FMError(FM_E_UNSUPPORTED_INDEX, "DeviceIndex.from_image")
        # fmgpu_index_create takes a bare pointer: the words it will read must be there (loadIndex checks the file the same way)
        if im.size != 6 + 2 * k + nent * (2 * (d // 32) * k + ncnt):
            raise FMError(FM_E_READING_FMI, "DeviceIndex.from_image (image size does not match its header)")
        dpos = (C.c_uint32 * k)(*[int(v) for v in im[6:6 + k]])
        dbase = (C.c_uint32 * k)(*[int(v) for v in im[6 + k:6 + 2 * k]])
        h = C.c_void_p()
        check(lib().fmgpu_index_create(device, tag, k, d, bwtsize, ncnt, nent, dpos, dbase,
                                       im[6 + 2 * k:].ctypes.data, C.byref(h)), "fmgpu_index_create")
        return cls(h)

    @classmethod
    def alloc_like(cls, meta, device=0):
        h = C.c_void_p()
        check(lib().fmgpu_index_alloc_like(device, C.byref(meta), C.byref(h)), "fmgpu_index_alloc_like")
        return cls(h)

    def fuse(self, fused_bases=0, lanes=0, budget_bytes=0):
        """Builds the fused-step table (up to 4 bases per rank, 256-bit loads) for MODE_FUSED searches."""
        check(lib().fmgpu_index_fuse(self.handle, fused_bases, lanes, budget_bytes), "fmgpu_index_fuse")
        return self

    def unfuse(self):
        check(lib().fmgpu_index_unfuse(self.handle), "fmgpu_index_unfuse")

    def sparsify(self, sparse_bases=0, lam=0, lanes=0):
        """Builds the sparse-step table (up to 14 bases per 64/128-byte block fetch) for MODE_SPARSE searches."""
        check(lib().fmgpu_index_sparsify(self.handle, sparse_bases, lam, lanes), "fmgpu_index_sparsify")
        return self

    def unsparsify(self):
        check(lib().fmgpu_index_unsparsify(self.handle), "fmgpu_index_unsparsify")

    def widen(self, wide_bases=0, prefix_bits=0, lanes=0):
        """Builds the wide-step table (up to 30 bases per 64/128-byte block fetch, both interval ends in one block) for MODE_WIDE searches."""
        check(lib().fmgpu_index_widen(self.handle, wide_bases, prefix_bits, lanes), "fmgpu_index_widen")
        return self

    def unwiden(self):
        check(lib().fmgpu_index_unwiden(self.handle), "fmgpu_index_unwiden")

    def wide_bases_for(self, length, max_entry_words=3):
        """Step width a wide-step table should have to serve reads of `length` bases with the fewest fetches (0 = none);
        max_entry_words=2 keeps 64-bit entries (steps up to 30 bases; less memory to build)."""
        return int(lib().fmgpu_wide_bases_for_words(self.handle, length, max_entry_words))

    def widen_for(self, length):
        """The wide-step table for reads of `length` bases: the best width, else the 64-bit-entry width when memory is short."""
        wb = self.wide_bases_for(length)
        if not wb:
            raise FMError(19, "fmgpu_wide_bases_for (no step width serves this read length)")
        try:
            return self.widen(wb)
        except FMError as ex:
            wb2 = self.wide_bases_for(length, 2)
            if ex.code != 19 or not wb2 or wb2 == wb:
                raise
            return self.widen(wb2)

    def wide_serves(self, length):
        return bool(lib().fmgpu_index_wide_serves(self.handle, length))

    def build_sa(self):
        """Derives the suffix array of the indexed text from this replica's own table (needed by locate)."""
        check(lib().fmgpu_index_build_sa(self.handle), "fmgpu_index_build_sa")
        return self

    def build_sa_sampled(self, rate=32):
        """Keeps every `rate`-th suffix array value + an LF-walk table instead of the whole array (same locate results)."""
        check(lib().fmgpu_index_build_sa_sampled(self.handle, rate), "fmgpu_index_build_sa_sampled")
        return self

    def drop_sa(self):
        check(lib().fmgpu_index_drop_sa(self.handle), "fmgpu_index_drop_sa")

    def download_sa(self):
        out = np.empty(int(self.meta.bwtsize), dtype=np.uint32)
        check(lib().fmgpu_index_download_sa(self.handle, out.ctypes.data), "fmgpu_index_download_sa")
        return out

    def prepare(self, length):
        """Builds whatever tables a search of `length`-base reads uses (tail table, lead tables); searches never build."""
        check(lib().fmgpu_index_prepare(self.handle, length), "fmgpu_index_prepare")
        return self

    def replicate(self, device):
        h = C.c_void_p()
        check(lib().fmgpu_index_replicate(self.handle, device, C.byref(h)), "fmgpu_index_replicate")
        return DeviceIndex(h)

    @property
    def meta(self):
        m = fmgpu_index_meta_t()
        check(lib().fmgpu_index_get_meta(self.handle, C.byref(m)), "fmgpu_index_get_meta")
        return m

    @property
    def device(self):
        return lib().fmgpu_index_device(self.handle)

    @property
    def blocks_ptr(self):
        return lib().fmgpu_index_blocks(self.handle)

    @property
    def __cuda_array_interface__(self):
        """The block table as a flat uint8 device array (lets torch wrap it for an NCCL broadcast)."""
        return {"shape": (int(self.meta.nbytes),), "typestr": "|u1", "data": (int(self.blocks_ptr), False), "version": 2}

    def free(self):
        if self.handle:
            lib().fmgpu_index_free(C.byref(self.handle))
            self.handle = C.c_void_p()


class DeviceBatch:
    """Device-resident query shard + its results (fmgpu_batch_t)."""

    def __init__(self, device, nq, length, steps):
        self.nq, self.len = nq, length
        self.handle = C.c_void_p()
        check(lib().fmgpu_batch_create(device, nq, length, steps, C.byref(self.handle)), "fmgpu_batch_create")

    def upload_ascii(self, ascii_bases):
        a = np.ascontiguousarray(ascii_bases, dtype=np.uint8).reshape(-1)
        assert a.size == self.nq * self.len
        check(lib().fmgpu_batch_upload_ascii(self.handle, a.ctypes.data), "fmgpu_batch_upload_ascii")

    def search(self, index, var=None, sync=True, prepare=True):
        if prepare:
            index.prepare(self.len)
        check(lib().fmgpu_batch_search(index.handle, self.handle, C.byref(var) if var is not None else None), "fmgpu_batch_search")
        if sync:
            check(lib().fmgpu_batch_sync(self.handle), "fmgpu_batch_sync")

    def search_timed(self, index, iters, var=None):
        index.prepare(self.len)
        ms = C.c_float()
        check(lib().fmgpu_batch_search_timed(index.handle, self.handle, C.byref(var) if var is not None else None, iters, C.byref(ms)),
              "fmgpu_batch_search_timed")
        return ms.value

    def count_fetches(self, index):
        nb, ns = C.c_uint64(), C.c_uint64()
        check(lib().fmgpu_batch_count_fetches(index.handle, self.handle, C.byref(nb), C.byref(ns)), "fmgpu_batch_count_fetches")
        return nb.value, ns.value

    def download(self):
        out = np.empty(2 * self.nq, dtype=np.uint32)
        check(lib().fmgpu_batch_download(self.handle, out.ctypes.data), "fmgpu_batch_download")
        return out

    def locate(self, index, max_hits):
        """Text positions of the occurrences of every read after a search: (positions [nq, max_hits] in suffix-array
        order, 0xFFFFFFFF beyond the hits; nhits [nq] = R - L)."""
        pos = np.empty((self.nq, max_hits), dtype=np.uint32)
        nhits = np.empty(self.nq, dtype=np.uint32)
        check(lib().fmgpu_batch_locate(index.handle, self.handle, max_hits, pos.ctypes.data, nhits.ctypes.data), "fmgpu_batch_locate")
        return pos, nhits

    def free(self):
        if self.handle:
            lib().fmgpu_batch_free(C.byref(self.handle))
            self.handle = C.c_void_p()


class IndexBuild:
    """Tag-100 index image built on the GPU (fmgpu_build_t)."""

    def __init__(self, handle):
        self.handle = handle

    @classmethod
    def from_text(cls, ascii_text, k, d, device=0):
        t = np.ascontiguousarray(ascii_text, dtype=np.uint8).reshape(-1)
        h = C.c_void_p()
        rc = lib().fmgpu_build_from_text(device, t.ctypes.data, t.size, k, d, C.byref(h))
        if rc:
            raise FMError(rc, "fmgpu_build_from_text: " + lib().fmgpu_build_last_error().decode())
        return cls(h)

    @classmethod
    def from_synth(cls, n, seed, k, d, device=0):
        h = C.c_void_p()
        rc = lib().fmgpu_build_from_synth(device, n, seed, k, d, C.byref(h))
        if rc:
            raise FMError(rc, "fmgpu_build_from_synth: " + lib().fmgpu_build_last_error().decode())
        return cls(h)

    @property
    def image_words(self):
        return lib().fmgpu_build_image_words(self.handle)

    def download(self, out=None):
        if out is None:
            out = np.empty(self.image_words, dtype=np.uint32)
        check(lib().fmgpu_build_download(self.handle, out.ctypes.data), "fmgpu_build_download")
        return out

    def to_index(self):
        h = C.c_void_p()
        check(lib().fmgpu_build_to_index(self.handle, C.byref(h)), "fmgpu_build_to_index")
        return DeviceIndex(h)

    def transform(self, tag):
        """Image of the reference's tfmiBMP (101) / tfmiAC (200, 201) output for this tag-100 build."""
        h = C.c_void_p()
        check(lib().fmgpu_build_transform(self.handle, tag, C.byref(h)), "fmgpu_build_transform")
        return IndexBuild(h)

    def free(self):
        if self.handle:
            lib().fmgpu_build_free(C.byref(self.handle))
            self.handle = C.c_void_p()


def search_host(replicas, ascii_bases, length, var=None, out=None):
    """End-to-end: host ASCII reads in, host (L,R) out, pipelined over the replicas' GPUs."""
    a = np.ascontiguousarray(ascii_bases, dtype=np.uint8).reshape(-1)
    nq = a.size // length
    if out is None:
        out = np.empty(2 * nq, dtype=np.uint32)
    arr = (C.c_void_p * len(replicas))(*[r.handle for r in replicas])
    check(lib().fmgpu_search_host(arr, len(replicas), a.ctypes.data, nq, length, out.ctypes.data,
                                  C.byref(var) if var is not None else None), "fmgpu_search_host")
    return out


def gather_probe(device, table_bytes, loads_per_thread=256, iters=3, access_bytes=16):
    """Random-access roofline: independent uniformly random aligned accesses per second."""
    v = C.c_double()
    check(lib().fmgpu_gather_probe_ex(device, table_bytes, access_bytes, loads_per_thread, iters, C.byref(v)), "fmgpu_gather_probe_ex")
    return v.value

"""CPU tests of the host layer: the C-ABI library loads, exports every declared symbol, and the file
readers/writers keep the reference's formats and error behaviour (no compute calls: no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers


def test_library_exports_every_declared_symbol(built):
    pkg = helpers.pkg()
    L = pkg.lib()
    header = open(os.path.join(helpers.ROOT, "include", "fmindex_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", header))
    declared -= {"defined"}
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/fmindex_b200.h but not exported"
        assert name in pkg.PROTOTYPES, f"{name} has no ctypes prototype"
    assert not hasattr(L, "searchIndexCPU"), "the product must not carry a CPU search path"


def test_no_oracle_in_product(built):
    """The product library must not link or reference anything under oracle/."""
    pkg = helpers.pkg()
    out = helpers.run(["nm", "-D", pkg.LIB_PATH])
    assert "fmo_" not in out and "ref_search_parallel" not in out
    for root, _, files in os.walk(os.path.join(helpers.ROOT, helpers.PKG_NAME)):
        for f in files:
            if f.endswith((".c", ".cu", ".cuh", ".h", ".py")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "fm_oracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f


@pytest.mark.parametrize("fixture", ["small_k1_d64", "small_k2_d64", "small_k2_d128"])
def test_load_index_all_tags(built, tmp_path, fixture):
    pkg = helpers.pkg()
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", fixture + ".npz"))
    for tag in (100, 101, 200, 201):
        im = g[f"image_{tag}"]
        fn = str(tmp_path / f"x{tag}.fmi")
        im.tofile(fn)
        h = pkg.loadIndex(fn)
        f = pkg.index_fields(h)
        k = int(g["k"])
        assert (f.tag, f.steps, f.bwtsize, f.chunk) == (tag, k, int(g["n"]) + 1, int(g["d"]))
        assert f.ncounters == (4 ** k) // (2 if tag >= 200 else 1)
        assert f.nentries == -(-f.bwtsize // f.chunk) + (1 if tag >= 200 else 0)
        assert [f.h_dollarPositionBWT[i] for i in range(k)] == [int(v) for v in im[6:6 + k]]
        assert [f.h_modposdollarBWT[i] for i in range(k)] == [int(v) // f.chunk for v in im[6:6 + k]]
        raw = np.ctypeslib.as_array(C.cast(f.h_index, C.POINTER(C.c_uint32)), shape=(f.nentries * f.entry_words,))
        assert np.array_equal(raw, im[6 + 2 * k:])
        pkg.lib().freeIndex(C.byref(h))
        assert pkg.index_fields(h).h_index is None


def test_load_index_errors(built, tmp_path):
    pkg = helpers.pkg()
    L = pkg.lib()
    h = C.c_void_p()
    assert L.loadIndex(b"/nonexistent/file.fmi", C.byref(h)) == 1          # E_OPENING_INDEX_FILE
    bad = tmp_path / "bad.fmi"
    np.array([777, 1, 100, 4, 2, 64, 0, 0], dtype=np.uint32).tofile(bad)
    assert L.loadIndex(os.fsencode(str(bad)), C.byref(h)) == 100             # wrong type -> "use gfmiBaseLine"
    assert b"gfmiBaseLine" in L.errorCommon(100)
    trunc = tmp_path / "trunc.fmi"
    np.array([100, 1, 1000, 4, 16, 64, 5, 0, 1, 2, 3], dtype=np.uint32).tofile(trunc)
    assert L.loadIndex(os.fsencode(str(trunc)), C.byref(h)) == 5            # E_READING_FMI
    short = tmp_path / "short.fmi"
    np.array([100, 1], dtype=np.uint32).tofile(short)
    assert L.loadIndex(os.fsencode(str(short)), C.byref(h)) == 5
    assert L.errorCommon(0) == b"No error" and L.errorCommon(19) == b"Not implemented"
    assert b"tfmiAC" in L.errorCommon(201) and L.errorCommon(12345) == b"Unknown error"


def test_load_queries_and_results_roundtrip(built, tmp_path):
    pkg = helpers.pkg()
    L = pkg.lib()
    text = helpers.synth_text(5000, 3)
    reads = helpers.synth_reads(text, 8, 37, 20)                # 37: not a multiple of 32 (ragged batch)
    fa = str(tmp_path / "q.fa")
    helpers.write_fasta_reads(fa, reads, 20)
    q = pkg.loadQueries(fa, 20, 37)
    qs = C.cast(q, C.POINTER(pkg.qrys_t)).contents
    assert (qs.num, qs.size) == (37, 20)
    got = np.ctypeslib.as_array(C.cast(qs.h_queries, C.POINTER(C.c_uint8)), shape=(37 * 20,))
    assert np.array_equal(got, reads)
    L.freeQueries(C.byref(q))
    # wrong length / too few reads are errors, not a silent mis-stride
    h = C.c_void_p()
    assert L.loadQueries(os.fsencode(fa), 21, 37, C.byref(h)) == 12
    assert L.loadQueries(os.fsencode(fa), 20, 38, C.byref(h)) == 12
    assert L.loadQueries(b"/nonexistent.fa", 20, 1, C.byref(h)) == 14

    r = pkg.initResults(5)
    arr = pkg.resultsArray(r, copy=False)
    assert arr.shape == (10,) and (arr == 0).all()
    arr[:] = np.array([1, 2, 30, 40, 0, 0, 4294967295, 7, 5, 5], dtype=np.uint32)
    out = str(tmp_path / "idx")
    assert L.saveResults(os.fsencode(out), r, None) == 0
    assert open(out + ".res.gpu").read() == "5\n1 2\n30 40\n0 0\n4294967295 7\n5 5\n"
    back = C.c_void_p()
    assert L.loadResults(os.fsencode(out + ".res.gpu"), C.byref(back)) == 0
    assert np.array_equal(pkg.resultsArray(back), arr)
    L.freeResults(C.byref(r)); L.freeResults(C.byref(back))
    e = C.c_void_p()
    assert L.initResults(0, C.byref(e)) == 0                    # empty batch
    assert pkg.resultsArray(e).size == 0


def test_device_entry_points_fail_loudly_without_gpu(built):
    """No CPU fallback: on a box without an sm_100 device the GPU entry points return FM_E_CUDA."""
    pkg = helpers.pkg()
    L = pkg.lib()
    if L.fmgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    with pytest.raises(pkg.FMError) as ei:
        pkg.DeviceIndex.from_image(g["image_100"], device=0)
    assert ei.value.code == pkg.FM_E_CUDA
    h = C.c_void_p()
    assert L.fmgpu_batch_create(0, 32, 32, 2, C.byref(h)) == pkg.FM_E_CUDA
    v = C.c_double()
    assert L.fmgpu_gather_probe(0, 1 << 20, 16, 1, C.byref(v)) == pkg.FM_E_CUDA


@pytest.mark.parametrize("length", [1, 3, 4, 15, 16, 17, 36, 63, 64, 65, 100, 127, 128, 129, 250, 1000])
def test_host_packer_simd_equals_scalar_equals_definition(built, length):
    """fm_hostpack_reads (AVX-512 path when available) == scalar path == the packing definition:
    packed position t = base len-1-t, 16 codes per word, code from ASCII bits 2 and 1."""
    pkg = helpers.pkg()
    L = pkg.lib()
    nq = 257
    rng = np.random.default_rng(length)
    reads = np.frombuffer(b"ACGTacgtNn$", dtype=np.uint8)[rng.integers(0, 11, nq * length)].copy()
    wpq = L.fmgpu_words_per_query(length)
    a = np.full(nq * wpq, 0xDEADBEEF, dtype=np.uint32)
    b = np.full(nq * wpq, 0xDEADBEEF, dtype=np.uint32)
    L.fm_hostpack_reads(reads.ctypes.data, nq, length, a.ctypes.data, 0)
    L.fm_hostpack_reads_scalar(reads.ctypes.data, nq, length, b.ctypes.data)
    r = reads.reshape(nq, length).astype(np.uint32)
    code = (((r >> 2) & 1) << 1) | (((r >> 2) & 1) ^ ((r >> 1) & 1))
    want = np.zeros((nq, wpq), dtype=np.uint32)
    for t in range(length):
        want[:, t // 16] |= code[:, length - 1 - t] << np.uint32(2 * (t % 16))
    assert np.array_equal(b.reshape(nq, wpq), want)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("out_offset", [0, 16, 5])
@pytest.mark.parametrize("nbases", [1, 3, 4, 63, 64, 65, 255, 256, 257, 4095, 4096, 4097, 100_003])
def test_host_stream_packer(built, nbases, out_offset):
    """fm_hostpack_stream: base g at bits 2(g%4) of byte g/4, threads split at 4096-base slices; with
    $FM_HOSTPACK_LINE=1 (read once per process, so set for the whole test module) a 64-byte aligned output takes the
    whole-line non-temporal path (256 bases per store), other alignments the 16-byte path."""
    pkg = helpers.pkg()
    L = pkg.lib()
    rng = np.random.default_rng(nbases)
    a = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)[rng.integers(0, 9, nbases)].copy()
    buf = np.zeros((nbases + 3) // 4 + 8 + 128, dtype=np.uint8)
    shift = (-buf.ctypes.data) % 64 + out_offset
    out = buf[shift:]
    assert (out.ctypes.data - out_offset) % 64 == 0
    L.fm_hostpack_stream(a.ctypes.data, nbases, out.ctypes.data, 0)
    r = np.zeros(((nbases + 3) // 4) * 4, dtype=np.uint32)
    r[:nbases] = a
    code = (((r >> 2) & 1) << 1) | (((r >> 2) & 1) ^ ((r >> 1) & 1))
    code[nbases:] = 0
    want = (code[0::4] | code[1::4] << 2 | code[2::4] << 4 | code[3::4] << 6).astype(np.uint8)
    assert np.array_equal(out[: want.size], want)


def test_fast_ingest_matches_reference_reader_semantics(built, tmp_path):
    """mmap/OpenMP query parser: headers skipped, CRLF tolerated, missing final newline, extra reads ignored,
    short file / wrong length rejected; large enough to be cut into per-thread slices."""
    pkg = helpers.pkg()
    L = pkg.lib()
    text = helpers.synth_text(200_000, 3)
    nq, length = 60_000, 37
    reads = helpers.synth_reads(text, 8, nq, length)
    r = reads.reshape(nq, length)
    fa = tmp_path / "big.fa"
    with open(fa, "wb") as f:
        for i in range(nq):
            f.write(b">rid%d %d-%d\n" % (i + 1, i, i + length))
            f.write(r[i].tobytes() + (b"\r\n" if i % 7 == 0 else b"\n"))
        f.seek(f.tell() - 1); f.truncate()                      # no newline at the end of the file
    assert os.path.getsize(fa) > (1 << 20)
    for want_n in (nq, nq - 12345, 1):
        q = pkg.loadQueries(str(fa), length, want_n)
        qs = C.cast(q, C.POINTER(pkg.qrys_t)).contents
        got = np.ctypeslib.as_array(C.cast(qs.h_queries, C.POINTER(C.c_uint8)), shape=(want_n * length,))
        assert np.array_equal(got, reads[: want_n * length])
        L.freeQueries(C.byref(q))
    h = C.c_void_p()
    assert L.loadQueries(os.fsencode(str(fa)), length, nq + 1, C.byref(h)) == 12      # too few reads
    assert L.loadQueries(os.fsencode(str(fa)), length + 1, nq, C.byref(h)) == 12      # wrong length
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    assert L.loadQueries(os.fsencode(str(empty)), 10, 1, C.byref(h)) == 12


def test_fast_result_writer_is_byte_identical(built, tmp_path):
    pkg = helpers.pkg()
    L = pkg.lib()
    rng = np.random.default_rng(0)
    n = 300_001
    res = rng.integers(0, 2 ** 32, 2 * n, dtype=np.uint64).astype(np.uint32)
    res[:6] = [0, 0, 4294967295, 4294967295, 10, 9]
    out = str(tmp_path / "r.txt")
    assert L.writeResults(os.fsencode(out), res.ctypes.data, n) == 0
    want = "%d\n" % n + "".join(f"{a} {b}\n" for a, b in res.reshape(-1, 2).tolist())
    assert open(out).read() == want
    assert helpers.results_text_md5(res) == __import__("hashlib").md5(want.encode()).hexdigest()


def test_header_is_plain_c(built, tmp_path):
    """include/fmindex_b200.h must compile as C99 with no C++/CUDA/torch types, and link against the library."""
    src = tmp_path / "use.c"
    src.write_text('#include "fmindex_b200.h"\n'
                   'int main(void) { fmgpu_variant_t v = {0, 0, 0, 0}; fmgpu_index_meta_t m; (void) v; (void) m;\n'
                   '  return fmgpu_device_count() < 0 || errorCommon(0) == 0 || sizeof(fmi_t) < 14 * 4; }\n')
    pkg = helpers.pkg()
    exe = tmp_path / "use"
    helpers.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(helpers.ROOT, "include"), str(src),
                 "-L", os.path.dirname(pkg.LIB_PATH), "-lfmindex_b200", "-Wl,-rpath," + os.path.dirname(pkg.LIB_PATH), "-o", str(exe)])
    helpers.run([str(exe)])


def test_ctypes_mirrors_match_the_header(built, tmp_path):
    """The ctypes structures in the package must have the size and field offsets of include/fmindex_b200.h."""
    import ctypes as C
    pkg = helpers.pkg()
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fmindex_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(fmgpu_index_meta_t), offsetof(fmgpu_index_meta_t, nbytes),\n'
                   '  offsetof(fmgpu_index_meta_t, tail_const), offsetof(fmgpu_index_meta_t, sparse_bytes), offsetof(fmgpu_index_meta_t, tail_bytes),\n'
                   '  sizeof(fmgpu_variant_t), sizeof(fmi_t), sizeof(qrys_t), offsetof(fmgpu_index_meta_t, derived_bytes),\n'
                   '  sizeof(fmgpu_transfer_stats_t), offsetof(fmgpu_transfer_stats_t, search_ms), offsetof(fmgpu_transfer_stats_t, result_bytes),\n'
                   '  sizeof(fmgpu_pipeline_stats_t)); return 0; }\n')
    exe = tmp_path / "sizes"
    helpers.run(["gcc", "-std=c99", "-I", os.path.join(helpers.ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    m = pkg.fmgpu_index_meta_t
    assert got == [C.sizeof(m), m.nbytes.offset, m.tail_const.offset, m.sparse_bytes.offset, m.tail_bytes.offset,
                   C.sizeof(pkg.fmgpu_variant_t), C.sizeof(pkg.fmi_t), C.sizeof(pkg.qrys_t), m.derived_bytes.offset,
                   C.sizeof(pkg.fmgpu_transfer_stats_t), pkg.fmgpu_transfer_stats_t.search_ms.offset,
                   pkg.fmgpu_transfer_stats_t.result_bytes.offset, C.sizeof(pkg.fmgpu_pipeline_stats_t)]


def test_bench_and_entry_scripts_compile_without_warnings():
    """bench.py and __graft_entry__.py must byte-compile with warnings as errors (a missing '+' between string pieces
    is only a SyntaxWarning until the line runs on the GPU box), and bench.py --help must work without a GPU."""
    import warnings
    for name in ("bench.py", "__graft_entry__.py"):
        path = os.path.join(helpers.ROOT, name)
        with open(path) as f:
            src = f.read()
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            compile(src, path, "exec")
    out = subprocess.run([os.sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "--impl" in out.stdout


def test_host_layer_under_asan_and_ubsan(built, tmp_path):
    """`make sanitize`: fm_host.c, fm_ingest.c and fm_hostpack.c compiled with AddressSanitizer + UndefinedBehaviorSanitizer
    and driven through loaders, writers, packers, error paths and handle lifetimes (tests/c/host_sanitize.c): zero findings,
    leak check included."""
    if not os.path.exists("/usr/bin/gcc"):
        pytest.skip("system gcc with the sanitizer runtimes is not installed")
    pkg = helpers.pkg()
    helpers.run(["make", "-C", pkg.CSRC, "sanitize"])
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz"))
    for tag in (100, 201):
        fn = str(tmp_path / f"i{tag}.fmi")
        g[f"image_{tag}"].tofile(fn)
        p = subprocess.run([os.path.join(pkg.CSRC, "build", "host_sanitize"), fn, "2", str(tmp_path)], capture_output=True, text=True,
                           env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1"))
        assert p.returncode == 0 and "host_sanitize OK" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]
        assert "AddressSanitizer" not in p.stderr and "runtime error" not in p.stderr and "LeakSanitizer" not in p.stderr, p.stderr[-3000:]

/*
 * fm_hostpack.c -- host-side ASCII -> 2-bit packing of reads (plain C, OpenMP, AVX-512 fast paths).
 *
 * Code A=0 C=1 G=2 T=3 from ASCII bits 2 and 1 (the reference's bit trick, src/fmIndexCPUBaseline.c:213-226;
 * case-insensitive, other bytes alias).  Two packers:
 *   fm_hostpack_stream  the whole batch as one 2-bit sequence, 64 bases per AVX-512 iteration and no per-read
 *                       work; used by fmgpu_search_host so that 25 instead of 100 bytes per 100-bp read cross
 *                       PCIe (the GPU's fm_unstream_kernel cuts, reverses and word-aligns the reads);
 *   fm_hostpack_reads   per-read reversed words, byte-identical to the device kernel fm_pack_kernel (packed
 *                       position t = base len-1-t, 16 bases per word); for callers that upload packed batches.
 * Both replace the host-side warp interleave of the reference query loader (common/common.c:175-194).  This is
 * data-format conversion only -- no search arithmetic here.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <immintrin.h>
#include <omp.h>

static void pack_read_scalar(const unsigned char *rd, uint32_t len, uint32_t wpq, uint32_t *out)
{
  uint32_t w, i;
  for (w = 0; w < wpq; w++) {
    uint32_t v = 0;
    for (i = 0; i < 16; i++) {
      const uint32_t t = 16 * w + i;
      if (t < len) {
        const uint32_t c = rd[len - 1 - t];
        const uint32_t hi = (c >> 2) & 1u, mid = (c >> 1) & 1u;
        v |= ((hi << 1) | (hi ^ mid)) << (2 * i);
      }
    }
    out[w] = v;
  }
}

/* 64 bases per iteration: masked load of the block's bytes, byte reversal with vpermb, 2-bit codes, then
 * four codes per byte with pmaddubsw / pmaddwd, narrowed with vpmovdb */
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi")))
static void pack_read_avx512(const unsigned char *rd, uint32_t len, uint32_t wpq, uint32_t *out)
{
  const __m512i iota = _mm512_set_epi8(63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 46, 45, 44, 43, 42, 41, 40,
                                       39, 38, 37, 36, 35, 34, 33, 32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16,
                                       15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  unsigned char *o = (unsigned char *) out;
  const uint32_t out_bytes = wpq * 4;
  uint32_t done = 0, ob = 0;
  while (done < len) {
    const uint32_t m = (len - done < 64) ? len - done : 64;          /* reversed positions [done, done+m) */
    const __mmask64 km = (m == 64) ? ~(__mmask64) 0 : (((__mmask64) 1 << m) - 1);
    const __m512i x = _mm512_maskz_loadu_epi8(km, rd + (len - done - m));    /* bases len-done-m .. len-done-1, ascending */
    const __m512i idx = _mm512_sub_epi8(_mm512_set1_epi8((char)(m - 1)), iota);
    const __m512i r = _mm512_maskz_permutexvar_epi8(km, idx, x);      /* byte j = base len-1-(done+j) */
    const __m512i u = _mm512_and_si512(_mm512_srli_epi16(r, 1), _mm512_set1_epi8(3));            /* bit2<<1 | bit1 */
    const __m512i c = _mm512_xor_si512(u, _mm512_and_si512(_mm512_srli_epi16(u, 1), _mm512_set1_epi8(1)));
    const __m512i cz = _mm512_maskz_mov_epi8(km, c);                  /* positions past the read pack as 0 */
    const __m512i p16 = _mm512_maddubs_epi16(cz, _mm512_set1_epi16(0x0401));                     /* c0 + 4 c1 */
    const __m512i p32 = _mm512_madd_epi16(p16, _mm512_set1_epi32(0x00100001));                   /* + 16 c2 + 64 c3 */
    const __m128i b = _mm512_cvtepi32_epi8(p32);                      /* 16 bytes = 64 bases */
    const uint32_t nb = (m + 3) / 4;
    _mm_mask_storeu_epi8(o + ob, (__mmask16)((1u << nb) - 1u), b);
    ob += nb; done += m;
  }
  if (ob < out_bytes) memset(o + ob, 0, out_bytes - ob);
}

int fm_hostpack_has_simd(void)
{
  return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
         __builtin_cpu_supports("avx512vbmi");
}

int fm_hostpack_threads(void) { return omp_get_max_threads(); }

/* packed must hold nq * ceil(len/16) words; nthreads <= 0 means all OpenMP threads */
void fm_hostpack_reads(const char *ascii, uint64_t nq, uint32_t len, uint32_t *packed, int nthreads)
{
  const uint32_t wpq = (len + 15u) / 16u;
  const int simd = fm_hostpack_has_simd();
  int64_t q;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  #pragma omp parallel for schedule(static) num_threads(nthreads)
  for (q = 0; q < (int64_t) nq; q++) {
    const unsigned char *rd = (const unsigned char *) ascii + (uint64_t) q * len;
    uint32_t *out = packed + (uint64_t) q * wpq;
    if (simd) pack_read_avx512(rd, len, wpq, out);
    else      pack_read_scalar(rd, len, wpq, out);
  }
}

/* ------------------------------------------------------------------------ *
 * Stream packing: the whole batch as ONE sequence of bases, base g (= q*len + i) at bits [2(g%4), 2(g%4)+2) of
 * byte g/4 -- no per-read work on the host at all (64 ASCII bytes -> 16 packed bytes per iteration); the
 * per-read reversal / word alignment is done by fm_unstream_kernel on the GPU.
 * ------------------------------------------------------------------------ */
static void stream_scalar(const unsigned char *in, uint64_t nbases, unsigned char *out)
{
  uint64_t g;
  for (g = 0; g + 4 <= nbases; g += 4) {
    uint32_t v = 0, j;
    for (j = 0; j < 4; j++) {
      const uint32_t c = in[g + j], hi = (c >> 2) & 1u, mid = (c >> 1) & 1u;
      v |= ((hi << 1) | (hi ^ mid)) << (2 * j);
    }
    out[g / 4] = (unsigned char) v;
  }
  if (g < nbases) {
    uint32_t v = 0, j;
    for (j = 0; g + j < nbases; j++) {
      const uint32_t c = in[g + j], hi = (c >> 2) & 1u, mid = (c >> 1) & 1u;
      v |= ((hi << 1) | (hi ^ mid)) << (2 * j);
    }
    out[g / 4] = (unsigned char) v;
  }
}

/* software prefetch distance of the stream packer in bytes: $FM_HOSTPACK_PREFETCH or fm_hostpack_set_prefetch(),
 * default 8192; 0 = off.  Worth ~5 % on the hybrid end-to-end feed of the benchmark host (1 041 -> 1 107 M reads/s,
 * means of three alternating rounds, profiles/r01_hostpack_prefetch.md); the packer alone is within noise. */
static int g_pf_dist = -1;
static int fm_hostpack_prefetch_distance(void)
{
  if (g_pf_dist < 0) { const char *e = getenv("FM_HOSTPACK_PREFETCH"); g_pf_dist = e && *e ? atoi(e) : 8192; }
  return g_pf_dist;
}
void fm_hostpack_set_prefetch(int bytes) { g_pf_dist = bytes < 0 ? 0 : bytes; }

__attribute__((target("avx512f,avx512bw,avx512vl")))
static void stream_avx512(const unsigned char *in, uint64_t nbases, unsigned char *out)
{
  const __m512i three = _mm512_set1_epi8(3), one = _mm512_set1_epi8(1);
  const __m512i w16 = _mm512_set1_epi16(0x0401), w32 = _mm512_set1_epi32(0x00100001);
  /* non-temporal 16-byte stores measured SLOWER than plain stores on the benchmark host (hybrid feed 977 vs
   * 1100 M reads/s, profiles/r01_e2e_ab.jsonl): partial-line write combining.  Off unless $FM_HOSTPACK_NT=1. */
  static int nt_allowed = -1;
  if (nt_allowed < 0) { const char *e = getenv("FM_HOSTPACK_NT"); nt_allowed = (e && e[0] == '1'); }
  const int nt = nt_allowed && (((uintptr_t) out) & 15u) == 0;
  const int pf_dist = fm_hostpack_prefetch_distance();
  uint64_t g = 0;
  /* whole output lines: 256 bases -> one 64-byte non-temporal store (no read-for-ownership, no partial-line write
   * combining).  Measured SLOWER too on the benchmark host (hybrid feed 960-1030 vs 1070-1160 M reads/s, host packing
   * alone equal, profiles/r01_e2e_line_nt.jsonl): with plain stores the 50 MB of staging buffers stay in the 60 MB L3 and
   * the DMA engine reads them from there, non-temporal stores send them through DRAM.  Off unless $FM_HOSTPACK_LINE=1. */
  static int line_nt = -1;
  if (line_nt < 0) { const char *e = getenv("FM_HOSTPACK_LINE"); line_nt = (e && e[0] == '1'); }
  if (line_nt && (((uintptr_t) out) & 63u) == 0) {
    for (; g + 256 <= nbases; g += 256) {
      if (pf_dist) {
        _mm_prefetch((const char *)(in + g + pf_dist), _MM_HINT_NTA);       _mm_prefetch((const char *)(in + g + pf_dist + 64), _MM_HINT_NTA);
        _mm_prefetch((const char *)(in + g + pf_dist + 128), _MM_HINT_NTA); _mm_prefetch((const char *)(in + g + pf_dist + 192), _MM_HINT_NTA);
      }
      __m128i q[4];
      for (int j = 0; j < 4; j++) {
        const __m512i x = _mm512_loadu_si512((const void *)(in + g + 64 * j));
        const __m512i u = _mm512_and_si512(_mm512_srli_epi16(x, 1), three);
        const __m512i c = _mm512_xor_si512(u, _mm512_and_si512(_mm512_srli_epi16(u, 1), one));
        q[j] = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(c, w16), w32));
      }
      __m512i line = _mm512_castsi128_si512(q[0]);
      line = _mm512_inserti32x4(line, q[1], 1); line = _mm512_inserti32x4(line, q[2], 2); line = _mm512_inserti32x4(line, q[3], 3);
      _mm512_stream_si512((void *)(out + g / 4), line);
    }
    _mm_sfence();
  }
  for (; g + 64 <= nbases; g += 64) {
    if (pf_dist) _mm_prefetch((const char *)(in + g + pf_dist), _MM_HINT_NTA);
    const __m512i x = _mm512_loadu_si512((const void *)(in + g));
    const __m512i u = _mm512_and_si512(_mm512_srli_epi16(x, 1), three);
    const __m512i c = _mm512_xor_si512(u, _mm512_and_si512(_mm512_srli_epi16(u, 1), one));
    const __m512i p32 = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w16), w32);
    /* non-temporal when the destination is 16-byte aligned: the packed bytes are only read back by the DMA
     * engine, so skip the read-for-ownership of the output lines */
    if (nt) _mm_stream_si128((__m128i *)(out + g / 4), _mm512_cvtepi32_epi8(p32));
    else    _mm_storeu_si128((__m128i *)(out + g / 4), _mm512_cvtepi32_epi8(p32));
  }
  if (nt) _mm_sfence();
  if (g < nbases) stream_scalar(in + g, nbases - g, out + g / 4);
}

/* Several interleaved sub-streams per thread: one sequential stream per core leaves the core's line-fill buffers and the
 * L2 streamer under-used (a core reads ~9 GB/s from one stream, ~12 from four), so every thread cuts its range into
 * `streams` pieces and converts 64 bases of each in turn, with a software prefetch into L1 `pf_dist` bytes ahead of every
 * piece.  $FM_HOSTPACK_STREAMS / fm_hostpack_set_streams(): default 4, 1 = one stream per thread (the 4096-base slice loop). */
static int g_streams = -1;
static int fm_hostpack_streams(void)
{
  if (g_streams < 0) { const char *e = getenv("FM_HOSTPACK_STREAMS"); g_streams = e && *e ? atoi(e) : 4; if (g_streams < 1) g_streams = 1; if (g_streams > 8) g_streams = 8; }
  return g_streams;
}
void fm_hostpack_set_streams(int streams) { g_streams = streams < 1 ? 1 : (streams > 8 ? 8 : streams); }

__attribute__((target("avx512f,avx512bw,avx512vl")))
static void stream_avx512_multi(const unsigned char *in, uint64_t nbases, unsigned char *out, int K, int pf_dist)
{
  const __m512i three = _mm512_set1_epi8(3), one = _mm512_set1_epi8(1);
  const __m512i w16 = _mm512_set1_epi16(0x0401), w32 = _mm512_set1_epi32(0x00100001);
  const uint64_t seg = (nbases / (uint64_t) K) & ~(uint64_t) 255;    /* bases per piece: whole 64-byte output lines */
  uint64_t g;
  int k;
  for (g = 0; g < seg; g += 64)
    for (k = 0; k < K; k++) {
      const unsigned char *p = in + (uint64_t) k * seg + g;
      if (pf_dist) _mm_prefetch((const char *)(p + pf_dist), _MM_HINT_T0);
      const __m512i x = _mm512_loadu_si512((const void *) p);
      const __m512i u = _mm512_and_si512(_mm512_srli_epi16(x, 1), three);
      const __m512i c = _mm512_xor_si512(u, _mm512_and_si512(_mm512_srli_epi16(u, 1), one));
      _mm_storeu_si128((__m128i *)(out + ((uint64_t) k * seg + g) / 4), _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(c, w16), w32)));
    }
  if ((uint64_t) K * seg < nbases) stream_avx512(in + (uint64_t) K * seg, nbases - (uint64_t) K * seg, out + (uint64_t) K * seg / 4);
}

/* out must hold (nbases + 3) / 4 bytes (+ up to 3 bytes of slack are NOT written); threads split at 4096-base
 * boundaries so every thread writes whole bytes */
void fm_hostpack_stream(const char *ascii, uint64_t nbases, unsigned char *out, int nthreads)
{
  const int simd = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
  const uint64_t nblk = (nbases + 4095) / 4096;                     /* 4096-base slices */
  const int streams = fm_hostpack_streams();
  int64_t b;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (simd && streams > 1 && nblk >= (uint64_t) nthreads * 8) {
    /* one contiguous range of slices per thread, several sub-streams inside it */
    const uint64_t per = (nblk + (uint64_t) nthreads - 1) / (uint64_t) nthreads;
    int pf = fm_hostpack_prefetch_distance();
    if (pf > 4096) pf = 4096;                                       /* the single-stream default (8192, non-temporal hint) is too far for L1 */
    #pragma omp parallel for schedule(static, 1) num_threads(nthreads)
    for (b = 0; b < (int64_t) nthreads; b++) {
      const uint64_t b0 = (uint64_t) b * per, b1 = (b0 + per < nblk) ? b0 + per : nblk;
      if (b0 < b1) {
        const uint64_t g0 = b0 * 4096, g1 = (b1 * 4096 < nbases) ? b1 * 4096 : nbases;
        stream_avx512_multi((const unsigned char *) ascii + g0, g1 - g0, out + g0 / 4, streams, pf);
      }
    }
    return;
  }
  #pragma omp parallel for schedule(static) num_threads(nthreads)
  for (b = 0; b < (int64_t) nblk; b++) {
    const uint64_t g0 = (uint64_t) b * 4096, n = (nbases - g0 < 4096) ? nbases - g0 : 4096;
    if (simd) stream_avx512((const unsigned char *) ascii + g0, n, out + g0 / 4);
    else      stream_scalar((const unsigned char *) ascii + g0, n, out + g0 / 4);
  }
}

/* scalar-only entry point so tests can check the SIMD path against it */
void fm_hostpack_reads_scalar(const char *ascii, uint64_t nq, uint32_t len, uint32_t *packed)
{
  const uint32_t wpq = (len + 15u) / 16u;
  uint64_t q;
  for (q = 0; q < nq; q++) pack_read_scalar((const unsigned char *) ascii + q * len, len, wpq, packed + q * wpq);
}

/* Host DRAM read bandwidth over a caller's buffer (GB/s, best of `iters` passes, all threads): the ceiling of ANY feed
 * of ASCII reads -- every read costs its `len` bytes of host DRAM reads whoever fetches them, the DMA engine or a
 * packer thread.  bench.py prints e2e next to (this / bytes per read).  The buffer should be far larger than L3. */
double fm_host_read_bandwidth(const void *buf, uint64_t bytes, int nthreads, int iters)
{
  const uint64_t nw = bytes / 8;
  const uint64_t *p = (const uint64_t *) buf;
  double best = 0.0;
  int it;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (iters < 1) iters = 1;
  for (it = 0; it < iters; it++) {
    uint64_t sink = 0;
    int64_t t;
    const double t0 = omp_get_wtime();
    #pragma omp parallel for schedule(static, 1) num_threads(nthreads) reduction(^ : sink)
    for (t = 0; t < (int64_t) nthreads; t++) {
      const uint64_t per = (nw + (uint64_t) nthreads - 1) / (uint64_t) nthreads;
      const uint64_t a = (uint64_t) t * per, b = (a + per < nw) ? a + per : nw;
      uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, i;
      for (i = a; i + 32 <= b; i += 32) {
        uint64_t j;
        for (j = 0; j < 32; j += 4) { s0 ^= p[i + j]; s1 ^= p[i + j + 1]; s2 ^= p[i + j + 2]; s3 ^= p[i + j + 3]; }
      }
      sink ^= s0 ^ s1 ^ s2 ^ s3;
    }
    {
      const double dt = omp_get_wtime() - t0;
      const double gbs = dt > 0 ? (double) bytes / dt / 1e9 : 0.0;
      if (sink == 0x9E3779B97F4A7C15ull) gbs > 0 ? (void) 0 : (void) 0;   /* keeps the loads alive */
      if (gbs > best) best = gbs;
    }
  }
  return best;
}

/*
 * fm_build.cu -- GPU construction of the reference's tag-100 ".fmi" index
 * image, byte-identical to what genFMindex writes, plus device-side synthetic
 * inputs (fm_synth.h).
 *
 * SURVEY.md 8(f) rows 1-2.  The reference builder needs 28 min and 18.6 GB for
 * the 2 Gbp benchmark index (64-bit divsufsort, then n SERIAL LF steps to derive
 * BWT layers 1..k-1, src/genFMindex.c:327-400).  Here:
 *
 *   suffix order   radix sort of (32-base prefix key, position) pairs with
 *                  cub::DeviceRadixSort, then a fix-up pass that orders the
 *                  (rare) runs of equal keys by full suffix comparison with
 *                  the end of text as the smallest symbol -- the order
 *                  divbwt64 (resources/divsufsort.c:337-370) produces;
 *   BWT layers     layer s of row r is simply T$[(SA[r]-1-s) mod (n+1)]
 *                  (what generateOthersBWTs computes by walking LF), one gather;
 *   '$' rows       dollarPositionBWT[s] = row whose suffix starts at text
 *                  position s (src/genFMindex.c:352-361), stored as 'A' in the
 *                  planes (:506-509), dollarBaseBWT[s] = symbol read there (:518-520);
 *   planes         bit 31-p of word w of plane (s,b) = bit b of the code of
 *                  BWT_s[entry*d + 32w + p]   (substring2bitmap/bwt2bin, :402-455);
 *   counters       cnt[e][sigma] = acc[sigma] + #{rows < e*d with symbol sigma,
 *                  '$' rows excluded}, acc = C table + '$' adjustments
 *                  (precalculateBasesKSteps, :184-260).
 *
 * Runs of suffixes sharing their first 32 bases are rare in random text (ordered by
 * one thread each, fmb_fix_runs_kernel); repetitive texts (runs > 64) take the general
 * path, prefix doubling over the tied rows (fmb_prefix_doubling), so any text works.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include "../../include/fmindex_b200.h"
#include "fm_synth.h"

#define FMB_MAX_RUN 64u        /* longer runs of equal 32-base keys go to the prefix-doubling path */
#define FMB_MAX_DEPTH 8u       /* ... and so do runs whose members share more than 32*8 further bases */

struct fmgpu_build {
  int       device;
  uint32_t  tag, ncounters;   /* 100 from the builder; 101/200/201 after fmgpu_build_transform */
  uint32_t  k, d, bwtsize, nentries, entry_words;
  uint32_t  dpos[4], dbase[4];
  uint32_t *d_image;     /* header (6 + 2k words) + entries, as in the file */
  uint64_t  image_words;
};

static thread_local char g_berr[512] = "no error";
static int32_t fmb_fail(cudaError_t e, const char *what, const char *file, int line)
{
  snprintf(g_berr, sizeof g_berr, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  fprintf(stderr, "fm_build: %s\n", g_berr);
  return FM_E_CUDA;
}
#define CU_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fmb_fail(e_, #call, __FILE__, __LINE__); } while (0)

/* ---- packed text: base j in bits [63-2(j%32)-1, 63-2(j%32)] of word j/32 (big-endian, so that
 *      integer order of a word = lexicographic order of its 32 bases) ---- */
__device__ __forceinline__ uint32_t fmb_code_at(const uint64_t *__restrict__ pt, uint64_t j)
{
  return (uint32_t)(pt[j >> 5] >> (62 - 2 * (j & 31))) & 3u;
}

/* 32 bases starting at i (bases beyond the padded end read as A) */
__device__ __forceinline__ uint64_t fmb_key_at(const uint64_t *__restrict__ pt, uint64_t i)
{
  const uint64_t w = i >> 5, o = i & 31;
  const uint64_t hi = pt[w];
  if (o == 0) return hi;
  return (hi << (2 * o)) | (pt[w + 1] >> (64 - 2 * o));
}

__device__ __forceinline__ uint32_t fmb_ascii_code(uint32_t c)
{
  const uint32_t hi = (c >> 2) & 1u, mid = (c >> 1) & 1u;
  return (hi << 1) | (hi ^ mid);
}

/* one thread packs one 64-bit word; src = ASCII text or NULL for the synthetic generator */
__global__ void fmb_pack_text_kernel(const char *__restrict__ ascii, uint64_t n, uint64_t seed, uint64_t nwords,
                                     uint64_t *__restrict__ pt)
{
  const uint64_t w = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  uint64_t v = 0;
  for (uint32_t j = 0; j < 32; j++) {
    const uint64_t i = w * 32 + j;
    uint32_t c = 0;
    if (i < n) c = ascii ? fmb_ascii_code((uint32_t)(unsigned char) ascii[i]) : fm_synth_code(seed, i);
    v |= (uint64_t) c << (62 - 2 * j);
  }
  pt[w] = v;
}

__global__ void fmb_keys_kernel(const uint64_t *__restrict__ pt, uint64_t n, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ vals)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = fmb_key_at(pt, i);
  vals[i] = (uint32_t) i;
}

/* suffix a < suffix b, both known to agree on their first 32 (padded) bases; end of text is smallest */
__device__ bool fmb_suffix_less(const uint64_t *__restrict__ pt, uint64_t n, uint64_t a, uint64_t b, bool *too_deep)
{
  for (uint64_t off = 0;; off += 32) {
    if (off > 32 * FMB_MAX_DEPTH) { *too_deep = true; return false; }   /* long common prefix: leave it to prefix doubling */
    const uint64_t ra = a + off, rb = b + off;
    if (ra >= n || rb >= n) return ra > rb;           /* the one that ended (larger position) is smaller */
    const uint64_t ka = fmb_key_at(pt, ra), kb = fmb_key_at(pt, rb);
    if (ka != kb || ra + 32 > n || rb + 32 > n) {
      /* first differing base decides, unless a suffix ends before it */
      const uint64_t x = ka ^ kb;
      const uint64_t lim_a = n - ra, lim_b = n - rb;  /* real bases left */
      const uint64_t same = x ? (uint64_t)(__clzll((long long) x) >> 1) : 32;
      const uint64_t lim = lim_a < lim_b ? lim_a : lim_b;
      if (same >= lim && lim < 32) return lim_a < lim_b; /* shorter one hit '$' first */
      if (x) return ka < kb;
    }
  }
}

/* orders every run of equal keys in place; *status = longest run seen (0 = none) */
__global__ void fmb_fix_runs_kernel(const uint64_t *__restrict__ pt, uint64_t n, const uint64_t *__restrict__ keys,
                                    uint32_t *__restrict__ vals, uint32_t *status)
{
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r + 1 >= n) return;
  const uint64_t key = keys[r];
  if (keys[r + 1] != key || (r > 0 && keys[r - 1] == key)) return;    /* not the head of a run */
  uint64_t end = r + 1;
  while (end < n && keys[end] == key && end - r <= FMB_MAX_RUN) end++;
  const uint32_t len = (uint32_t)(end - r);
  atomicMax(status, len);
  if (len > FMB_MAX_RUN) return;
  bool too_deep = false;
  for (uint64_t i = r + 1; i < end && !too_deep; i++) {               /* insertion sort by full comparison */
    const uint32_t v = vals[i];
    uint64_t j = i;
    while (j > r && fmb_suffix_less(pt, n, v, vals[j - 1], &too_deep)) { vals[j] = vals[j - 1]; j--; }
    vals[j] = v;
  }
  if (too_deep) atomicMax(status, 0xFFFFFFFFu);                       /* still a permutation of the run; doubling finishes it */
}

/* SA of row r (row 0 is the '$' suffix) */
__device__ __forceinline__ uint64_t fmb_sa(const uint32_t *__restrict__ sorted, uint64_t n, uint64_t r)
{
  return r == 0 ? n : (uint64_t) sorted[r - 1];
}

/* code of BWT layer s at a row whose suffix starts at text position sa; '$' reads as A */
__device__ __forceinline__ uint32_t fmb_layer_code(const uint64_t *__restrict__ pt, uint64_t n, uint64_t sa, uint32_t s)
{
  const uint64_t m = n + 1;
  const uint64_t j = (sa + m - 1 - s) % m;
  return j == n ? 0u : fmb_code_at(pt, j);
}

__global__ void fmb_find_dollars_kernel(const uint32_t *__restrict__ sorted, uint64_t n, uint32_t k, uint32_t *dpos)
{
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  const uint64_t sa = fmb_sa(sorted, n, r);
  if (sa < k) dpos[sa] = (uint32_t) r;
}

/* one thread = one 32-row word: writes the 2k plane words (tag-100 order) and per-word symbol counts */
__global__ void fmb_planes_kernel(const uint64_t *__restrict__ pt, const uint32_t *__restrict__ sorted, uint64_t n,
                                  uint32_t k, uint32_t d, uint32_t entry_words, uint32_t *__restrict__ entries,
                                  uint64_t nwords32)
{
  const uint64_t w = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords32) return;
  const uint64_t row0 = w * 32, rows = n + 1;
  uint32_t p[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };               /* 2 planes per BWT layer, k <= 4 */
  for (uint32_t i = 0; i < 32; i++) {
    const uint64_t r = row0 + i;
    if (r >= rows) break;
    const uint64_t sa = fmb_sa(sorted, n, r);
    for (uint32_t s = 0; s < k; s++) {
      const uint32_t c = fmb_layer_code(pt, n, sa, s);
      p[2 * s]     |= (c & 1u) << (31 - i);
      p[2 * s + 1] |= (c >> 1) << (31 - i);
    }
  }
  const uint32_t W = d / 32;
  const uint64_t e = row0 / d;
  const uint32_t wn = (uint32_t)(row0 % d) / 32;
  uint32_t *ent = entries + e * entry_words;
  for (uint32_t s = 0; s < k; s++) {
    ent[2 * W * s + wn]     = p[2 * s];
    ent[2 * W * s + W + wn] = p[2 * s + 1];
  }
}

/* per entry and symbol: rows carrying the symbol, '$' rows excluded -> hist[sigma][e] */
__global__ void fmb_hist_kernel(const uint32_t *__restrict__ entries, uint32_t k, uint32_t d, uint32_t entry_words,
                                uint32_t nentries, uint32_t bwtsize, const uint32_t *__restrict__ dpos,
                                uint32_t *__restrict__ hist)
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nentries) return;
  const uint32_t W = d / 32, nsym = 1u << (2 * k);
  const uint32_t *ent = entries + (size_t) e * entry_words;
  for (uint32_t sigma = 0; sigma < nsym; sigma++) hist[(size_t) sigma * nentries + e] = 0;
  for (uint32_t wn = 0; wn < W; wn++) {
    const uint64_t row0 = (uint64_t) e * d + 32u * wn;
    const int64_t nvalid = (int64_t) bwtsize - (int64_t) row0;
    uint32_t keep = nvalid >= 32 ? 0xFFFFFFFFu : (nvalid <= 0 ? 0u : ~(0xFFFFFFFFu >> nvalid));
    for (uint32_t s = 0; s < k; s++)
      if (dpos[s] >= row0 && dpos[s] < row0 + 32) keep &= ~(0x80000000u >> (dpos[s] - row0));
    if (!keep) continue;
    for (uint32_t sigma = 0; sigma < nsym; sigma++) {
      uint32_t m = keep;
      for (uint32_t s = 0; s < k; s++) {
        const uint32_t c = (sigma >> (2 * s)) & 3u;
        const uint32_t p0 = ent[2 * W * s + wn], p1 = ent[2 * W * s + W + wn];
        m &= ((c & 1u) ? p0 : ~p0) & ((c & 2u) ? p1 : ~p1);
      }
      hist[(size_t) sigma * nentries + e] += __popc(m);
    }
  }
}

__global__ void fmb_counters_kernel(const uint32_t *__restrict__ scan, const uint32_t *__restrict__ acc, uint32_t k,
                                    uint32_t d, uint32_t entry_words, uint32_t nentries, uint32_t *__restrict__ entries)
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nentries) return;
  const uint32_t nsym = 1u << (2 * k);
  uint32_t *cnt = entries + (size_t) e * entry_words + 2 * (d / 32) * k;
  for (uint32_t sigma = 0; sigma < nsym; sigma++) cnt[sigma] = scan[(size_t) sigma * nentries + e] + acc[sigma];
}

/* Counter stage shared by the builder and by the k >= 3 -> 2-step projection of fm_index.cu: fills cnt[] of every
 * tag-100 entry from the planes already in `entries` (device), k <= 4 (src/genFMindex.c:184-260 for any K_STEPS).
 * dpos / dbase are host arrays. */
cudaError_t fmb_counter_stage(uint32_t *entries, uint32_t k, uint32_t d, uint32_t entry_words, uint32_t nentries, uint32_t bwtsize,
                              const uint32_t *dpos, const uint32_t *dbase)
{
  const uint32_t nsym = 1u << (2 * k);
  uint32_t *hist = NULL, *d_acc = NULL, *d_dpos = NULL;
  void *d_temp = NULL;
  cudaError_t e = cudaMalloc((void **) &hist, (size_t) nsym * nentries * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_acc, nsym * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_dpos, 4 * 4);
  if (e == cudaSuccess) e = cudaMemcpy(d_dpos, dpos, k * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    fmb_hist_kernel<<<(nentries + 127) / 128, 128>>>(entries, k, d, entry_words, nentries, bwtsize, d_dpos, hist);
    e = cudaGetLastError();
  }
  uint32_t totals[256], acc[256];
  for (uint32_t sigma = 0; sigma < nsym && e == cudaSuccess; sigma++) {
    uint32_t *h = hist + (size_t) sigma * nentries, last_in = 0, last_out = 0;
    e = cudaMemcpy(&last_in, h + nentries - 1, 4, cudaMemcpyDeviceToHost);
    size_t tb = 0;
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(NULL, tb, h, h, (int) nentries);
    if (e == cudaSuccess) e = cudaMalloc(&d_temp, tb ? tb : 16);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_temp, tb, h, h, (int) nentries);
    if (e == cudaSuccess) e = cudaMemcpy(&last_out, h + nentries - 1, 4, cudaMemcpyDeviceToHost);
    cudaFree(d_temp); d_temp = NULL;
    totals[sigma] = last_in + last_out;
  }
  if (e == cudaSuccess) {
    uint32_t run = 0;
    for (uint32_t sigma = 0; sigma < nsym; sigma++) { acc[sigma] = run; run += totals[sigma]; }
    /* '$' adjustments of src/genFMindex.c:245-250: the suffix that starts with the '$'-row's symbol
     * (its layers below s cleared) gets one extra predecessor */
    for (uint32_t s = 0; s < k; s++) {
      const uint32_t from = dbase[s] & (0xFFFFFFFFu << (2 * s));
      for (uint32_t sigma = from; sigma < nsym; sigma++) acc[sigma]++;
    }
    e = cudaMemcpy(d_acc, acc, nsym * 4, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    fmb_counters_kernel<<<(nentries + 127) / 128, 128>>>(hist, d_acc, k, d, entry_words, nentries, entries);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(hist); cudaFree(d_acc); cudaFree(d_dpos);
  return e;
}

/* k-step symbol stored at a row ('$' as A) */
__global__ void fmb_row_symbols_kernel(const uint64_t *__restrict__ pt, const uint32_t *__restrict__ sorted, uint64_t n,
                                       uint32_t k, const uint32_t *__restrict__ rows, uint32_t *__restrict__ syms)
{
  const uint32_t s = threadIdx.x;
  if (s >= k) return;
  const uint64_t sa = fmb_sa(sorted, n, rows[s]);
  uint32_t sym = 0;
  for (uint32_t l = 0; l < k; l++) sym |= fmb_layer_code(pt, n, sa, l) << (2 * l);
  syms[s] = sym;
}

/* ------------------------------------------------------------------------ *
 * General path for repetitive texts: prefix doubling over the rows that are still tied after the 32-base sort.
 * rank[p] (ISA) = first row of the group of suffix p; a tied suffix p gets the secondary key rank[p+h] (or, when
 * p+h is past the end, "ended": before every live suffix, shorter first), groups are re-sorted with a segmented
 * sort, split, and h doubles -- O(log n) rounds whatever the repeat structure.  Zero padding of the first key
 * never contradicts the true order (end of text < A), it only leaves ties, which the rounds resolve.
 * ------------------------------------------------------------------------ */
__global__ void fmb_pd_heads_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ head)
{
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  head[r] = (r == 0 || keys[r] != keys[r - 1]) ? (uint32_t) r : 0u;      /* max-scan turns this into the group's first row */
}

__global__ void fmb_pd_isa_kernel(const uint32_t *__restrict__ sa, const uint32_t *__restrict__ grp, uint64_t n, uint32_t *__restrict__ isa,
                                  uint8_t *__restrict__ tied)
{
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  isa[sa[r]] = grp[r];
  tied[r] = (uint8_t)((r > 0 && grp[r - 1] == grp[r]) || (r + 1 < n && grp[r + 1] == grp[r]));
}

__global__ void fmb_pd_gather_kernel(const uint32_t *__restrict__ rows, uint32_t m, const uint32_t *__restrict__ sa,
                                     const uint32_t *__restrict__ grp, uint32_t *__restrict__ t_sa, uint32_t *__restrict__ t_grp)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  t_sa[i] = sa[rows[i]]; t_grp[i] = grp[rows[i]];
}

__global__ void fmb_pd_sec_kernel(const uint32_t *__restrict__ t_sa, const uint32_t *__restrict__ t_grp, uint32_t m, uint64_t n,
                                  uint64_t h, const uint32_t *__restrict__ isa, uint64_t *__restrict__ sec, uint32_t *__restrict__ seg_flag)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t p = (uint64_t) t_sa[i] + h;
  sec[i] = p < n ? (1ull << 32) + isa[p] : (n - t_sa[i]);              /* ended suffixes first, shorter first */
  seg_flag[i] = (i == 0 || t_grp[i] != t_grp[i - 1]) ? 1u : 0u;
}

__global__ void fmb_pd_split_kernel(const uint64_t *__restrict__ sec_sorted, const uint32_t *__restrict__ seg_flag,
                                    const uint32_t *__restrict__ rows, uint32_t m, uint32_t *__restrict__ newhead_row)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const bool head = seg_flag[i] || sec_sorted[i] != sec_sorted[i - 1];
  newhead_row[i] = head ? rows[i] : 0u;
}

__global__ void fmb_pd_commit_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ sa_sorted,
                                     const uint32_t *__restrict__ newgrp, uint32_t m, uint32_t *__restrict__ sa,
                                     uint32_t *__restrict__ grp, uint32_t *__restrict__ isa, uint8_t *__restrict__ still)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  sa[rows[i]] = sa_sorted[i];
  grp[rows[i]] = newgrp[i];
  isa[sa_sorted[i]] = newgrp[i];
  still[i] = (uint8_t)((i > 0 && newgrp[i - 1] == newgrp[i]) || (i + 1 < m && newgrp[i + 1] == newgrp[i]));
}

struct FmbMaxOp { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

#define PD_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fmb_fail(e_, #call, __FILE__, __LINE__); goto pd_done; } } while (0)

static int32_t fmb_prefix_doubling(uint64_t n, const uint64_t *keys_sorted, uint32_t *sa)
{
  int32_t rc = FM_SUCCESS;
  uint32_t *grp = NULL, *isa = NULL, *rows = NULL, *rows2 = NULL, *t_sa = NULL, *t_sa2 = NULL, *t_grp = NULL, *seg_flag = NULL;
  uint32_t *seg_off = NULL, *newhead = NULL, *newgrp = NULL, *d_count = NULL;
  uint64_t *sec = NULL, *sec2 = NULL;
  uint8_t *tied = NULL, *still = NULL;
  void *tmp = NULL; size_t tmp_bytes = 0, need = 0;
  uint32_t m = 0, nseg = 0;
  const unsigned gb = (unsigned)((n + 255) / 256);
  {
    PD_TRY(cudaMalloc((void **) &grp, n * 4)); PD_TRY(cudaMalloc((void **) &isa, n * 4)); PD_TRY(cudaMalloc((void **) &tied, n));
    PD_TRY(cudaMalloc((void **) &d_count, 8));
    fmb_pd_heads_kernel<<<gb, 256>>>(keys_sorted, n, grp);
    PD_TRY(cudaGetLastError());
    PD_TRY(cub::DeviceScan::InclusiveScan(NULL, need, grp, grp, FmbMaxOp(), (int64_t) n));
    tmp_bytes = need; PD_TRY(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    PD_TRY(cub::DeviceScan::InclusiveScan(tmp, need, grp, grp, FmbMaxOp(), (int64_t) n));
    fmb_pd_isa_kernel<<<gb, 256>>>(sa, grp, n, isa, tied);
    PD_TRY(cudaGetLastError());
    /* rows still tied, ascending */
    PD_TRY(cudaMalloc((void **) &rows, n * 4));
    {
      cub::CountingInputIterator<uint32_t> iota(0);
      PD_TRY(cub::DeviceSelect::Flagged(NULL, need, iota, tied, rows, d_count, (int64_t) n));
      if (need > tmp_bytes) { cudaFree(tmp); tmp = NULL; tmp_bytes = need; PD_TRY(cudaMalloc(&tmp, tmp_bytes)); }
      PD_TRY(cub::DeviceSelect::Flagged(tmp, need, iota, tied, rows, d_count, (int64_t) n));
    }
    PD_TRY(cudaMemcpy(&m, d_count, 4, cudaMemcpyDeviceToHost));
    cudaFree(tied); tied = NULL;
    if (m == 0) goto pd_done;
    PD_TRY(cudaMalloc((void **) &rows2, (size_t) m * 4)); PD_TRY(cudaMalloc((void **) &t_sa, (size_t) m * 4)); PD_TRY(cudaMalloc((void **) &t_sa2, (size_t) m * 4));
    PD_TRY(cudaMalloc((void **) &t_grp, (size_t) m * 4)); PD_TRY(cudaMalloc((void **) &seg_flag, (size_t) m * 4)); PD_TRY(cudaMalloc((void **) &seg_off, ((size_t) m + 1) * 4));
    PD_TRY(cudaMalloc((void **) &newhead, (size_t) m * 4)); PD_TRY(cudaMalloc((void **) &newgrp, (size_t) m * 4));
    PD_TRY(cudaMalloc((void **) &sec, (size_t) m * 8)); PD_TRY(cudaMalloc((void **) &sec2, (size_t) m * 8)); PD_TRY(cudaMalloc((void **) &still, m));
    for (uint64_t h = 32; m > 0; h *= 2) {
      const unsigned mb = (m + 255) / 256;
      fmb_pd_gather_kernel<<<mb, 256>>>(rows, m, sa, grp, t_sa, t_grp);
      fmb_pd_sec_kernel<<<mb, 256>>>(t_sa, t_grp, m, n, h, isa, sec, seg_flag);
      PD_TRY(cudaGetLastError());
      /* segment offsets = positions of the group heads inside the tied list */
      {
        cub::CountingInputIterator<uint32_t> iota(0);
        PD_TRY(cub::DeviceSelect::Flagged(NULL, need, iota, seg_flag, seg_off, d_count, (int) m));
        if (need > tmp_bytes) { cudaFree(tmp); tmp = NULL; tmp_bytes = need; PD_TRY(cudaMalloc(&tmp, tmp_bytes)); }
        PD_TRY(cub::DeviceSelect::Flagged(tmp, need, iota, seg_flag, seg_off, d_count, (int) m));
      }
      PD_TRY(cudaMemcpy(&nseg, d_count, 4, cudaMemcpyDeviceToHost));
      PD_TRY(cudaMemcpy(seg_off + nseg, &m, 4, cudaMemcpyHostToDevice));
      PD_TRY(cub::DeviceSegmentedSort::SortPairs(NULL, need, sec, sec2, t_sa, t_sa2, (int) m, (int) nseg, seg_off, seg_off + 1));
      if (need > tmp_bytes) { cudaFree(tmp); tmp = NULL; tmp_bytes = need; PD_TRY(cudaMalloc(&tmp, tmp_bytes)); }
      PD_TRY(cub::DeviceSegmentedSort::SortPairs(tmp, need, sec, sec2, t_sa, t_sa2, (int) m, (int) nseg, seg_off, seg_off + 1));
      fmb_pd_split_kernel<<<mb, 256>>>(sec2, seg_flag, rows, m, newhead);
      PD_TRY(cudaGetLastError());
      PD_TRY(cub::DeviceScan::InclusiveScan(NULL, need, newhead, newgrp, FmbMaxOp(), (int) m));
      if (need > tmp_bytes) { cudaFree(tmp); tmp = NULL; tmp_bytes = need; PD_TRY(cudaMalloc(&tmp, tmp_bytes)); }
      PD_TRY(cub::DeviceScan::InclusiveScan(tmp, need, newhead, newgrp, FmbMaxOp(), (int) m));
      fmb_pd_commit_kernel<<<mb, 256>>>(rows, t_sa2, newgrp, m, sa, grp, isa, still);
      PD_TRY(cudaGetLastError());
      PD_TRY(cub::DeviceSelect::Flagged(NULL, need, rows, still, rows2, d_count, (int) m));
      if (need > tmp_bytes) { cudaFree(tmp); tmp = NULL; tmp_bytes = need; PD_TRY(cudaMalloc(&tmp, tmp_bytes)); }
      PD_TRY(cub::DeviceSelect::Flagged(tmp, need, rows, still, rows2, d_count, (int) m));
      PD_TRY(cudaMemcpy(&m, d_count, 4, cudaMemcpyDeviceToHost));
      { uint32_t *t = rows; rows = rows2; rows2 = t; }
      if (h > 2 * n) break;                                             /* cannot happen: every suffix has ended by then */
    }
  }
pd_done:
  cudaFree(grp); cudaFree(isa); cudaFree(tied); cudaFree(rows); cudaFree(rows2); cudaFree(t_sa); cudaFree(t_sa2); cudaFree(t_grp);
  cudaFree(seg_flag); cudaFree(seg_off); cudaFree(newhead); cudaFree(newgrp); cudaFree(sec); cudaFree(sec2); cudaFree(still);
  cudaFree(d_count); cudaFree(tmp);
  return rc;
}
#undef PD_TRY

/* ------------------------------------------------------------------------ */
static int32_t fmb_build(int device, const char *h_ascii, uint64_t n, uint64_t seed, uint32_t k, uint32_t d,
                         fmgpu_build_t **out)
{
  if (!out || k < 1 || k > 4 || d == 0 || d % 32 || n < 2 * k || n >= 0xFFFFFFFEull) {
    snprintf(g_berr, sizeof g_berr, "fmgpu_build: need k in {1,2,3,4}, d multiple of 32, 2k <= n < 2^32-2");
    return FM_E_BAD_ARGUMENT;
  }
  CU_TRY(cudaSetDevice(device));
  const uint64_t nwords = (n + 31) / 32 + 2;                           /* + zero padding for key reads past the end */
  uint64_t *pt = NULL, *keys_a = NULL, *keys_b = NULL;
  uint32_t *vals_a = NULL, *vals_b = NULL, *d_status = NULL, *d_dpos = NULL, *d_dbase = NULL;
  char *d_ascii = NULL;
  void *d_temp = NULL;
  int32_t rc = FM_SUCCESS;
  fmgpu_build_t *b = (fmgpu_build_t *) calloc(1, sizeof(*b));
  if (!b) return FM_E_ALLOCATING_FMI;
  b->device = device; b->k = k; b->d = d; b->bwtsize = (uint32_t)(n + 1); b->tag = 100; b->ncounters = 1u << (2 * k);
  b->nentries = (uint32_t)((n + 1 + d - 1) / d);
  b->entry_words = 2 * (d / 32) * k + (1u << (2 * k));
  b->image_words = 6 + 2 * k + (uint64_t) b->nentries * b->entry_words;
  const uint32_t nsym = 1u << (2 * k);

#define BTRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fmb_fail(e_, #call, __FILE__, __LINE__); goto done; } } while (0)
  {
    /* 1. packed text */
    BTRY(cudaMalloc((void **) &pt, nwords * 8));
    if (h_ascii) {
      BTRY(cudaMalloc((void **) &d_ascii, n));
      BTRY(cudaMemcpy(d_ascii, h_ascii, n, cudaMemcpyHostToDevice));
    }
    fmb_pack_text_kernel<<<(unsigned)((nwords + 255) / 256), 256>>>(d_ascii, n, seed, nwords, pt);
    BTRY(cudaGetLastError());
    if (d_ascii) { BTRY(cudaDeviceSynchronize()); cudaFree(d_ascii); d_ascii = NULL; }

    /* 2. suffix order */
    BTRY(cudaMalloc((void **) &keys_a, n * 8)); BTRY(cudaMalloc((void **) &keys_b, n * 8));
    BTRY(cudaMalloc((void **) &vals_a, n * 4)); BTRY(cudaMalloc((void **) &vals_b, n * 4));
    BTRY(cudaMalloc((void **) &d_status, 4));   BTRY(cudaMemset(d_status, 0, 4));
    fmb_keys_kernel<<<(unsigned)((n + 255) / 256), 256>>>(pt, n, keys_a, vals_a);
    BTRY(cudaGetLastError());
    size_t temp_bytes = 0;
    BTRY(cub::DeviceRadixSort::SortPairs(NULL, temp_bytes, keys_a, keys_b, vals_a, vals_b, (int64_t) n, 0, 64));
    BTRY(cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 16));
    BTRY(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_a, keys_b, vals_a, vals_b, (int64_t) n, 0, 64));
    cudaFree(d_temp); d_temp = NULL;
    fmb_fix_runs_kernel<<<(unsigned)((n + 127) / 128), 128>>>(pt, n, keys_b, vals_b, d_status);
    BTRY(cudaGetLastError());
    uint32_t longest = 0;
    BTRY(cudaMemcpy(&longest, d_status, 4, cudaMemcpyDeviceToHost));
    cudaFree(keys_a); keys_a = NULL; cudaFree(vals_a); vals_a = NULL;
    if (longest > FMB_MAX_RUN) {
      /* repetitive text: runs longer than FMB_MAX_RUN were left alone; order all tied rows by prefix doubling */
      rc = fmb_prefix_doubling(n, keys_b, vals_b);
      if (rc != FM_SUCCESS) goto done;
    }
    cudaFree(keys_b); keys_b = NULL;

    /* 3. '$' rows */
    BTRY(cudaMalloc((void **) &d_dpos, 4 * 4)); BTRY(cudaMalloc((void **) &d_dbase, 4 * 4));
    BTRY(cudaMemset(d_dpos, 0xFF, 8));
    fmb_find_dollars_kernel<<<(unsigned)((n + 1 + 255) / 256), 256>>>(vals_b, n, k, d_dpos);
    BTRY(cudaGetLastError());
    fmb_row_symbols_kernel<<<1, 32>>>(pt, vals_b, n, k, d_dpos, d_dbase);
    BTRY(cudaGetLastError());
    BTRY(cudaMemcpy(b->dpos, d_dpos, k * 4, cudaMemcpyDeviceToHost));
    BTRY(cudaMemcpy(b->dbase, d_dbase, k * 4, cudaMemcpyDeviceToHost));

    /* 4. planes */
    BTRY(cudaMalloc((void **) &b->d_image, b->image_words * 4));
    BTRY(cudaMemset(b->d_image, 0, b->image_words * 4));
    uint32_t *entries = b->d_image + 6 + 2 * k;
    const uint64_t nwords32 = ((uint64_t) b->bwtsize + 31) / 32;
    fmb_planes_kernel<<<(unsigned)((nwords32 + 127) / 128), 128>>>(pt, vals_b, n, k, d, b->entry_words, entries, nwords32);
    BTRY(cudaGetLastError());
    BTRY(cudaDeviceSynchronize());
    cudaFree(vals_b); vals_b = NULL; cudaFree(pt); pt = NULL;

    /* 5. counters: per-entry histogram, exclusive scan per symbol, C table */
    cudaError_t e = fmb_counter_stage(entries, k, d, b->entry_words, b->nentries, b->bwtsize, b->dpos, b->dbase);
    if (e != cudaSuccess) { rc = fmb_fail(e, "counter stage", __FILE__, __LINE__); goto done; }

    /* 6. header (src/genFMindex.c:167-176) */
    uint32_t head[14];
    head[0] = 100; head[1] = k; head[2] = b->bwtsize; head[3] = nsym; head[4] = b->nentries; head[5] = d;
    for (uint32_t s = 0; s < k; s++) { head[6 + s] = b->dpos[s]; head[6 + k + s] = b->dbase[s]; }
    BTRY(cudaMemcpy(b->d_image, head, (6 + 2 * k) * 4, cudaMemcpyHostToDevice));
  }
done:
#undef BTRY
  cudaFree(d_ascii); cudaFree(pt); cudaFree(keys_a); cudaFree(keys_b); cudaFree(vals_a); cudaFree(vals_b);
  cudaFree(d_status); cudaFree(d_dpos); cudaFree(d_dbase); cudaFree(d_temp);
  if (rc != FM_SUCCESS) { if (b->d_image) cudaFree(b->d_image); free(b); cudaGetLastError(); return rc; }
  *out = b;
  return FM_SUCCESS;
}

extern "C" const char *fmgpu_build_last_error(void) { return g_berr; }

extern "C" int32_t fmgpu_build_from_text(int32_t device, const char *h_ascii, uint64_t n, uint32_t k, uint32_t d,
                                         fmgpu_build_t **out)
{
  if (!h_ascii) return FM_E_BAD_ARGUMENT;
  return fmb_build(device, h_ascii, n, 0, k, d, out);
}

extern "C" int32_t fmgpu_build_from_synth(int32_t device, uint64_t n, uint64_t seed, uint32_t k, uint32_t d,
                                          fmgpu_build_t **out)
{
  return fmb_build(device, NULL, n, seed, k, d, out);
}

extern "C" uint64_t fmgpu_build_image_words(const fmgpu_build_t *b) { return b ? b->image_words : 0; }
extern "C" void *fmgpu_build_image_device(const fmgpu_build_t *b) { return b ? (void *) b->d_image : NULL; }

extern "C" int32_t fmgpu_build_download(const fmgpu_build_t *b, uint32_t *h_image)
{
  if (!b || !h_image) return FM_E_BAD_ARGUMENT;
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaMemcpy(h_image, b->d_image, b->image_words * 4, cudaMemcpyDeviceToHost));
  return FM_SUCCESS;
}

/* writes the image as an index FILE the reference tools (and loadIndex) read: same bytes as saveIndex of
 * src/genFMindex.c:155-181 / src/transformIndexBitmaps.c:96-123 / src/transformIndexAlternateCounters.c:163-217 */
extern "C" int32_t fmgpu_build_save(const fmgpu_build_t *b, const char *path)
{
  if (!b || !path) return FM_E_BAD_ARGUMENT;
  CU_TRY(cudaSetDevice(b->device));
  FILE *fp = fopen(path, "wb");
  if (!fp) return FM_E_SAVING_INDEX_FILE;
  const uint64_t slice = 64ull << 20;                       /* words per staging slice (256 MB) */
  uint32_t *h = (uint32_t *) malloc((size_t)((b->image_words < slice ? b->image_words : slice) * 4));
  if (!h) { fclose(fp); return FM_E_ALLOCATING_FMI; }
  int32_t rc = FM_SUCCESS;
  for (uint64_t w0 = 0; w0 < b->image_words && rc == FM_SUCCESS; w0 += slice) {
    const uint64_t n = (b->image_words - w0 < slice) ? b->image_words - w0 : slice;
    if (cudaMemcpy(h, b->d_image + w0, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); rc = FM_E_CUDA; break; }
    if (fwrite(h, 4, n, fp) != n) rc = FM_E_SAVING_INDEX_FILE;
  }
  free(h);
  if (fclose(fp) != 0 && rc == FM_SUCCESS) rc = FM_E_SAVING_INDEX_FILE;
  return rc;
}

/* re-block the built image into the searchable device layout (no host round trip) */
extern "C" int32_t fmgpu_build_to_index(const fmgpu_build_t *b, fmgpu_index_t **out)
{
  if (!b || !out) return FM_E_BAD_ARGUMENT;
  return fmgpu_index_create_from_device(b->device, b->tag, b->k, b->d, b->bwtsize, b->ncounters, b->nentries,
                                        b->dpos, b->dbase, b->d_image + 6 + 2 * b->k, out);
}

/* ------------------------------------------------------------------------ *
 * The reference's two layout transformers on the GPU, byte-identical outputs:
 *   tag 101  per-32-row word interleave of the planes      src/transformIndexBitmaps.c:269-295
 *   tag 200  half the counters per entry + padding entry   src/transformIndexAlternateCounters.c:434-479
 *   tag 201  both                                          src/transformIndexAlternateCounters.c:387-432
 * ------------------------------------------------------------------------ */
__global__ void fmb_transform_kernel(const uint32_t *__restrict__ src, uint32_t k, uint32_t d, uint32_t nent_src,
                                     uint32_t ew_src, uint32_t tag_dst, uint32_t *__restrict__ dst, uint32_t ew_dst)
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nent_src) return;
  const uint32_t W = d / 32, nsym = 1u << (2 * k), H = nsym / 2;
  const bool ac = tag_dst >= 200, il = (tag_dst & 1u) != 0;
  const uint32_t *se = src + (size_t) e * ew_src;
  uint32_t *de = dst + (size_t) e * ew_dst;
  uint32_t *dplanes = de + (ac ? H : 0u);
  for (uint32_t s = 0; s < k; s++)
    for (uint32_t bit = 0; bit < 2; bit++)
      for (uint32_t n = 0; n < W; n++) {
        const uint32_t w = se[2 * W * s + W * bit + n];
        dplanes[il ? (2 * k * n + 2 * s + bit) : (2 * W * s + W * bit + n)] = w;
      }
  const uint32_t *scnt = se + 2 * W * k;
  if (!ac) { for (uint32_t i = 0; i < nsym; i++) de[2 * W * k + i] = scnt[i]; }
  else     { for (uint32_t i = 0; i < H; i++) de[i] = scnt[((e & 1u) ? H : 0u) + i]; }
}

/* AltCounters padding entry: zero planes; counters = last real entry's counters + rows of the last chunk
 * carrying the symbol ('$' rows and the zero tail INCLUDED -- transformIndexAlternateCounters.c:420-431) */
__global__ void fmb_ac_padding_kernel(const uint32_t *__restrict__ src, uint32_t k, uint32_t d, uint32_t nent_src,
                                      uint32_t ew_src, uint32_t bwtsize, uint32_t *__restrict__ dst, uint32_t ew_dst)
{
  const uint32_t i = threadIdx.x;
  const uint32_t W = d / 32, nsym = 1u << (2 * k), H = nsym / 2;
  if (i >= H) return;
  const uint32_t pad = nent_src, elast = bwtsize / d, r = bwtsize % d;
  const uint32_t sigma = ((pad & 1u) ? H : 0u) + i;
  const uint32_t *se = src + (size_t) elast * ew_src;
  uint32_t cnt = (sigma == 0) ? d - r : 0u;
  for (uint32_t n = 0; n < W; n++) {
    const int32_t rem = (int32_t) r - (int32_t)(32 * n);
    uint32_t m = rem <= 0 ? 0u : (rem >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> rem));
    for (uint32_t s = 0; s < k; s++) {
      const uint32_t c = (sigma >> (2 * s)) & 3u;
      const uint32_t p0 = se[2 * W * s + n], p1 = se[2 * W * s + W + n];
      m &= ((c & 1u) ? p0 : ~p0) & ((c & 2u) ? p1 : ~p1);
    }
    cnt += __popc(m);
  }
  uint32_t *de = dst + (size_t) pad * ew_dst;
  de[i] = src[(size_t)(nent_src - 1) * ew_src + 2 * W * k + sigma] + cnt;
  if (i == 0) for (uint32_t w = 0; w < 2 * W * k; w++) de[H + w] = 0u;
}

extern "C" int32_t fmgpu_build_transform(const fmgpu_build_t *src, uint32_t tag, fmgpu_build_t **out)
{
  if (!src || !out || src->tag != 100 || !(tag == 101 || tag == 200 || tag == 201)) return FM_E_BAD_ARGUMENT;
  const bool ac = tag >= 200;
  if (ac && src->bwtsize % src->d == 0) {
    snprintf(g_berr, sizeof g_berr, "fmgpu_build_transform: bwtsize %% d == 0 is undefined in the reference AltCounters transformer");
    return FM_E_NOT_IMPLEMENTED;
  }
  CU_TRY(cudaSetDevice(src->device));
  fmgpu_build_t *b = (fmgpu_build_t *) calloc(1, sizeof(*b));
  if (!b) return FM_E_ALLOCATING_FMI;
  *b = *src;
  const uint32_t k = src->k, nsym = 1u << (2 * k);
  b->tag = tag; b->ncounters = ac ? nsym / 2 : nsym; b->nentries = src->nentries + (ac ? 1u : 0u);
  b->entry_words = 2 * (src->d / 32) * k + b->ncounters;
  b->image_words = 6 + 2 * k + (uint64_t) b->nentries * b->entry_words;
  b->d_image = NULL;
  cudaError_t e = cudaMalloc((void **) &b->d_image, b->image_words * 4);
  if (e != cudaSuccess) { free(b); return fmb_fail(e, "cudaMalloc(transformed image)", __FILE__, __LINE__); }
  const uint32_t *se = src->d_image + 6 + 2 * k;
  uint32_t *de = b->d_image + 6 + 2 * k;
  fmb_transform_kernel<<<(src->nentries + 127) / 128, 128>>>(se, k, src->d, src->nentries, src->entry_words, tag, de, b->entry_words);
  e = cudaGetLastError();
  if (e == cudaSuccess && ac) {
    fmb_ac_padding_kernel<<<1, 128>>>(se, k, src->d, src->nentries, src->entry_words, src->bwtsize, de, b->entry_words);
    e = cudaGetLastError();
  }
  uint32_t head[14];
  head[0] = tag; head[1] = k; head[2] = b->bwtsize; head[3] = b->ncounters; head[4] = b->nentries; head[5] = b->d;
  for (uint32_t s = 0; s < k; s++) { head[6 + s] = b->dpos[s]; head[6 + k + s] = b->dbase[s]; }
  if (e == cudaSuccess) e = cudaMemcpy(b->d_image, head, (6 + 2 * k) * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(b->d_image); free(b); return fmb_fail(e, "fmgpu_build_transform", __FILE__, __LINE__); }
  *out = b;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_build_free(fmgpu_build_t **pb)
{
  if (!pb || !*pb) return FM_SUCCESS;
  cudaSetDevice((*pb)->device);
  cudaFree((*pb)->d_image);
  free(*pb);
  *pb = NULL;
  return FM_SUCCESS;
}

/* ---- synthetic reads written straight into device memory (ASCII, plain order) ---- */
__global__ void fmb_synth_reads_kernel(uint64_t n, uint64_t seed_ref, uint64_t nq, uint32_t len, uint64_t seed_reads,
                                       uint64_t first, char *__restrict__ out)
{
  const uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nq * len) return;
  const uint64_t q = idx / len;
  const uint32_t c = (uint32_t)(idx - q * len);
  const uint64_t start = fm_synth_read_start(seed_reads, first + q, n, len);
  out[idx] = fm_synth_base(seed_ref, start + c);
}

extern "C" int32_t fmgpu_synth_reads_device(int32_t device, uint64_t n, uint64_t seed_ref, uint64_t nq, uint32_t len,
                                            uint64_t seed_reads, uint64_t first, char *d_ascii, void *stream)
{
  if (!d_ascii || len == 0 || len > n) return FM_E_BAD_ARGUMENT;
  CU_TRY(cudaSetDevice(device));
  const uint64_t total = nq * len;
  if (total == 0) return FM_SUCCESS;
  if ((total + 255) / 256 >= (1ull << 31)) return FM_E_BAD_ARGUMENT;
  fmb_synth_reads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t) stream>>>(n, seed_ref, nq, len, seed_reads, first, d_ascii);
  CU_TRY(cudaGetLastError());
  return FM_SUCCESS;
}

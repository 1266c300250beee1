/*
 * fm_kernels.cuh -- device code of the B200 k-step FM-index search path.
 *
 * Device layout "SB96" (symbol blocks of 96 BWT rows), produced by
 * fm_reblock_kernel from any of the reference's on-disk layouts:
 *
 *     blocks[sigma * nblocks + b] = uint4 { rank, w0, w1, w2 }
 *
 *   rank       = value the reference searcher returns for symbol sigma at row
 *                boundary X = 96*b  (counter + popcount - '$' fix, i.e.
 *                src/fmIndexCPUBaseline.c:227-257 evaluated at X)
 *   w0..w2     = indicator bits of rows 96*b .. 96*b+95: bit i of the 96-bit
 *                little-endian value is 1 iff row 96*b+i carries k-step symbol
 *                sigma, is < bwtsize and is not one of the k '$' rows.
 *
 * One rank query = ONE aligned 16-byte load (ld.global.nc.v4.u32) that brings
 * both the sampled counter and the bitmap ("interleaved bitmaps plus
 * counters"), i.e. one 32-byte DRAM sector; the two blocks of a sector are
 * consecutive row ranges of the same symbol, so L and R share the sector
 * whenever they fall in the same 192-row window.  The '$' corrections of the
 * reference (:252-256) are folded into the layout, so the hot loop is
 *     X' = rank + popc(w & prefixmask(X - 96*b))
 * with no branches.  AltCounters files (tags 200/201) are re-derived into the
 * same block format with the AltCounters searcher's semantics
 * (src/fmIndexCPUBaseline-AltCounters.c:218-266); its padding-entry quirk
 *  is reproduced by the QUIRK template flag.
 */
#ifndef FM_KERNELS_CUH_
#define FM_KERNELS_CUH_

#include <stdint.h>
#include <cuda_runtime.h>

#define FM_SB_ROWS 96u

struct FmSearchParams {
  const uint4    *blocks;     /* [nsymbols][nblocks]                              */
  const uint32_t *packed;     /* [nq][wpq] reversed 2-bit reads                   */
  uint32_t       *results;    /* [2*nq]                                           */
  unsigned long long *fetch_counters; /* COUNT only: [0]=blocks, [1]=sectors      */
  uint32_t nblocks;
  uint32_t nq;
  uint32_t nsteps;            /* len / k                                          */
  uint32_t wpq;               /* 32-bit words per packed read                     */
  uint32_t wpq_pad;           /* smem stride (odd)                                */
  uint32_t bwtsize;
  uint32_t quirk_start;
  uint32_t quirk_mask;
  /* odd read length on a 2-step index: after nsteps 2-base steps one base is left; it is consumed by a 1-step
   * rank DERIVED from the 2-step table (fm_tail_rank) -- the result is what a 1-step index of the same text gives */
  uint32_t has_tail;
  uint32_t tail_row, tail_base;   /* row whose layer-1 char is '$' (it has a layer-0 char but no 2-step symbol), and that char */
  uint32_t tail_const[4];         /* C1[c] - sum_c1 rank2(c | c1<<2, 0) */
  const uint4 *tail1;             /* [4][nblocks] tail table (fm_tail_table_kernel): that rank in ONE block fetch, or NULL */
};

/* raw (file-order) index as uploaded, for the re-blocker */
struct FmRawIndex {
  const uint32_t *entries;
  uint32_t tag, k, d, ncounters, nentries, entry_words, bwtsize, nentries_std;
  uint32_t dpos[2], dbase[2];
  uint32_t quirk_start, quirk_mask;
};

__device__ __forceinline__ uint4 fm_ldg16(const uint4 *p)
{
  uint4 v;
  /* no .L2::64B here: the 128-byte fill an L2 miss triggers by default brings the 7 neighbouring blocks along, which
   * the narrowing (L,R) interval of the next steps hits (profiles/r01_prefetch_variants.md) */
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

/* mask of the `width` low bits; width >= 32 gives all ones (BMSK.clamp) */
__device__ __forceinline__ uint32_t fm_lowmask(uint32_t width)
{
  uint32_t m;
  asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0u), "r"(width));
  return m;
}

__device__ __forceinline__ uint32_t fm_div96(uint32_t x) { return __umulhi(x, 0xAAAAAAABu) >> 6; }

/* rank inside one SB96 block: rows [96b, 96b + r), 0 <= r < 96 */
__device__ __forceinline__ uint32_t fm_block_rank(const uint4 v, uint32_t r)
{
  const uint32_t r1 = (uint32_t) max((int) r - 32, 0);
  const uint32_t r2 = (uint32_t) max((int) r - 64, 0);
  return v.x + __popc(v.y & fm_lowmask(r)) + __popc(v.z & fm_lowmask(r1)) + __popc(v.w & fm_lowmask(r2));
}

/* 1-step LF of row boundary X for base c on a 2-step table: rows below X whose layer-0 char is c are those
 * carrying one of the four 2-step symbols (c1, c), plus the row whose layer-1 char is '$' when it lies below X
 * and has layer-0 char c.  tail_const[c] folds the 1-step C table and the four block ranks at X = 0. */
__device__ __forceinline__ uint32_t fm_tail_rank(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t c, uint32_t X,
                                                 uint32_t tail_const, uint32_t tail_row, uint32_t tail_base)
{
  const uint32_t b = fm_div96(X), r = X - b * FM_SB_ROWS;
  uint4 v[4];
  #pragma unroll
  for (int c1 = 0; c1 < 4; c1++) v[c1] = fm_ldg16(blocks + (size_t)(c | (c1 << 2)) * nblocks + b);
  uint32_t sum = tail_const + ((X > tail_row && c == tail_base) ? 1u : 0u);
  #pragma unroll
  for (int c1 = 0; c1 < 4; c1++) sum += fm_block_rank(v[c1], r);
  return sum;
}

/* Tail table: the derived 1-step rank re-blocked like SB96, tail1[c * nblocks + b] = { fm_tail_rank(c, 96 b), the 96
 * indicator bits "layer-0 char of the row is c" } -- the OR of the four 2-step indicators (c1, c), which are disjoint,
 * plus the bit of the row whose layer-1 char is '$' (it carries no 2-step symbol).  For X = 96 b + r,
 * fm_block_rank(tail1[c][b], r) == fm_tail_rank(c, X) term by term, with one block fetch instead of four. */
__global__ void fm_tail_table_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t tc0, uint32_t tc1, uint32_t tc2,
                                     uint32_t tc3, uint32_t tail_row, uint32_t tail_base, uint4 *__restrict__ tail1)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint32_t tc[4] = { tc0, tc1, tc2, tc3 };
  const uint32_t tb = tail_row / FM_SB_ROWS, to = tail_row - tb * FM_SB_ROWS;
  #pragma unroll
  for (uint32_t c = 0; c < 4; c++) {
    uint4 o = make_uint4(tc[c] + ((b * FM_SB_ROWS > tail_row && c == tail_base) ? 1u : 0u), 0u, 0u, 0u);
    #pragma unroll
    for (uint32_t c1 = 0; c1 < 4; c1++) {
      const uint4 v = blocks[(size_t)(c | (c1 << 2)) * nblocks + b];
      o.x += v.x; o.y |= v.y; o.z |= v.z; o.w |= v.w;
    }
    if (b == tb && c == tail_base) {
      if (to < 32u) o.y |= 1u << to; else if (to < 64u) o.z |= 1u << (to - 32u); else o.w |= 1u << (to - 64u);
    }
    tail1[(size_t) c * nblocks + b] = o;
  }
}

/* last base of an odd-length read for both interval ends: one fetch from the tail table (the second only when R lies
 * in another block), or the four-fetch derivation when the table could not be allocated */
__device__ __forceinline__ void fm_tail_step(const uint4 *__restrict__ tail1, const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t c,
                                             uint32_t &L, uint32_t &R, uint32_t tail_const, uint32_t tail_row, uint32_t tail_base)
{
  if (tail1) {
    const uint32_t bL = fm_div96(L), bR = fm_div96(R);
    const uint4 *base = tail1 + (size_t) c * nblocks;
    const uint4 vL = fm_ldg16(base + bL);
    const uint4 vR = (bL == bR) ? vL : fm_ldg16(base + bR);
    L = fm_block_rank(vL, L - bL * FM_SB_ROWS);
    R = fm_block_rank(vR, R - bR * FM_SB_ROWS);
  } else {
    L = fm_tail_rank(blocks, nblocks, c, L, tail_const, tail_row, tail_base);
    R = fm_tail_rank(blocks, nblocks, c, R, tail_const, tail_row, tail_base);
  }
}

/* cooperative staging of this CTA's packed reads: global [q][wpq] -> smem [q][wpq_pad] */
__device__ __forceinline__ void fm_stage_queries(uint32_t *sq, const FmSearchParams &p, uint32_t q0, uint32_t nqb, int nthreads)
{
  const uint32_t total = nqb * p.wpq;
  const uint32_t *src = p.packed + (size_t) q0 * p.wpq;
  if (p.wpq == p.wpq_pad) {
    for (uint32_t i = threadIdx.x; i < total; i += nthreads) sq[i] = __ldg(src + i);
  } else {
    for (uint32_t i = threadIdx.x; i < total; i += nthreads) {
      const uint32_t q = i / p.wpq, w = i - q * p.wpq;
      sq[q * p.wpq_pad + w] = __ldg(src + i);
    }
  }
}

/* ------------------------------------------------------------------------ *
 * Task kernel: one thread owns QPT reads and both interval endpoints of each.
 * Replaces searchIndexKernel of src/fmIndexGPU-Task-{1Step,2Step,2Step-AltCounters}.cu.
 * All 2*QPT block fetches of a step are issued before any is consumed; the R
 * fetch is predicated off when R falls in L's block (about 5 steps in 6).
 * ------------------------------------------------------------------------ */
template <int K, int QPT, int THREADS, int MINB, bool QUIRK, bool COUNT>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_task_kernel(const FmSearchParams p)
{
  extern __shared__ uint32_t sq[];
  constexpr uint32_t SYMBITS = 2 * K, SYMMASK = (1u << SYMBITS) - 1u, STEPS_PER_WORD = 32 / SYMBITS;
  const uint32_t q0 = blockIdx.x * (THREADS * QPT);
  const uint32_t nqb = min((uint32_t)(THREADS * QPT), p.nq - q0);

  fm_stage_queries(sq, p, q0, nqb, THREADS);
  __syncthreads();

  uint32_t L[QPT], R[QPT], word[QPT];
  const uint32_t *myq[QPT];
  bool live[QPT];
  #pragma unroll
  for (int i = 0; i < QPT; i++) {
    const uint32_t lq = i * THREADS + threadIdx.x;
    live[i] = lq < nqb;
    myq[i] = sq + (live[i] ? lq : 0u) * p.wpq_pad;
    L[i] = 0u; R[i] = p.bwtsize; word[i] = 0u;
  }
  unsigned long long nblk_fetch = 0, nsec_fetch = 0;

  for (uint32_t step = 0; step < p.nsteps; step++) {
    if ((step % STEPS_PER_WORD) == 0) {
      #pragma unroll
      for (int i = 0; i < QPT; i++) word[i] = myq[i][step / STEPS_PER_WORD];
    }
    uint4 vL[QPT], vR[QPT];
    uint32_t rL[QPT], rR[QPT], sig[QPT];
    bool same[QPT];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      sig[i] = word[i] & SYMMASK; word[i] >>= SYMBITS;
      const uint32_t bL = fm_div96(L[i]), bR = fm_div96(R[i]);
      rL[i] = L[i] - bL * FM_SB_ROWS; rR[i] = R[i] - bR * FM_SB_ROWS;
      const uint4 *base = p.blocks + (size_t) sig[i] * p.nblocks;
      same[i] = (bL == bR);
      vL[i] = fm_ldg16(base + bL);
      if (!same[i]) vR[i] = fm_ldg16(base + bR);
      if (COUNT && live[i]) { nblk_fetch += same[i] ? 1 : 2; nsec_fetch += ((bL >> 1) == (bR >> 1)) ? 1 : 2; }
    }
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      if (same[i]) vR[i] = vL[i];
      uint32_t nL = fm_block_rank(vL[i], rL[i]);
      uint32_t nR = fm_block_rank(vR[i], rR[i]);
      if (QUIRK) {
        const uint32_t dq = (p.quirk_mask >> (2u * sig[i])) & 3u;
        if (L[i] >= p.quirk_start) nL += dq;
        if (R[i] >= p.quirk_start) nR += dq;
      }
      L[i] = nL; R[i] = nR;
    }
  }

  if (K == 2 && p.has_tail) {                     /* last base of an odd-length read */
    const uint32_t pos = 4u * p.nsteps;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint32_t c = (myq[i][pos >> 5] >> (pos & 31u)) & 3u;
      fm_tail_step(p.tail1, p.blocks, p.nblocks, c, L[i], R[i], p.tail_const[c], p.tail_row, p.tail_base);
    }
  }

  #pragma unroll
  for (int i = 0; i < QPT; i++)
    if (live[i]) {
      const uint32_t q = q0 + i * THREADS + threadIdx.x;
      reinterpret_cast<uint2 *>(p.results)[q] = make_uint2(L[i], R[i]);
    }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      nblk_fetch += __shfl_xor_sync(0xFFFFFFFFu, nblk_fetch, o);
      nsec_fetch += __shfl_xor_sync(0xFFFFFFFFu, nsec_fetch, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.fetch_counters, nblk_fetch); atomicAdd(p.fetch_counters + 1, nsec_fetch); }
  }
}

/* ------------------------------------------------------------------------ *
 * Coop kernel: a lane PAIR owns QPT reads; the even lane carries L, the odd
 * lane R (the reference's endpoint-per-thread mapping, e.g.
 * src/fmIndexGPU-Coop-2Step.cu:160-176).  Each step the pair compares block
 * ids with one shuffle; when both endpoints fall in the same block only the L
 * lane fetches and the 16 bytes are handed to the R lane with warp shuffles.
 * ------------------------------------------------------------------------ */
template <int K, int QPT, int THREADS, int MINB, bool QUIRK>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_coop_kernel(const FmSearchParams p)
{
  extern __shared__ uint32_t sq[];
  constexpr uint32_t SYMBITS = 2 * K, SYMMASK = (1u << SYMBITS) - 1u, STEPS_PER_WORD = 32 / SYMBITS;
  constexpr int PAIRS = THREADS / 2;
  const uint32_t q0 = blockIdx.x * (PAIRS * QPT);
  const uint32_t nqb = min((uint32_t)(PAIRS * QPT), p.nq - q0);
  const uint32_t side = threadIdx.x & 1u, pair = threadIdx.x >> 1;
  const int src_lane = (threadIdx.x & 31) & ~1;

  fm_stage_queries(sq, p, q0, nqb, THREADS);
  __syncthreads();

  uint32_t X[QPT], word[QPT];
  const uint32_t *myq[QPT];
  bool live[QPT];
  #pragma unroll
  for (int i = 0; i < QPT; i++) {
    const uint32_t lq = i * PAIRS + pair;
    live[i] = lq < nqb;
    myq[i] = sq + (live[i] ? lq : 0u) * p.wpq_pad;
    X[i] = side ? p.bwtsize : 0u; word[i] = 0u;
  }

  for (uint32_t step = 0; step < p.nsteps; step++) {
    if ((step % STEPS_PER_WORD) == 0) {
      #pragma unroll
      for (int i = 0; i < QPT; i++) word[i] = myq[i][step / STEPS_PER_WORD];
    }
    uint4 v[QPT];
    uint32_t r[QPT], sig[QPT];
    bool take[QPT];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      sig[i] = word[i] & SYMMASK; word[i] >>= SYMBITS;
      const uint32_t b = fm_div96(X[i]);
      r[i] = X[i] - b * FM_SB_ROWS;
      const uint32_t bp = __shfl_xor_sync(0xFFFFFFFFu, b, 1);
      take[i] = side && (b == bp);          /* R lane rides on the L lane's fetch */
      if (!take[i]) v[i] = fm_ldg16(p.blocks + (size_t) sig[i] * p.nblocks + b);
    }
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      uint4 t;
      t.x = __shfl_sync(0xFFFFFFFFu, v[i].x, src_lane);
      t.y = __shfl_sync(0xFFFFFFFFu, v[i].y, src_lane);
      t.z = __shfl_sync(0xFFFFFFFFu, v[i].z, src_lane);
      t.w = __shfl_sync(0xFFFFFFFFu, v[i].w, src_lane);
      if (take[i]) v[i] = t;
      uint32_t nX = fm_block_rank(v[i], r[i]);
      if (QUIRK) { if (X[i] >= p.quirk_start) nX += (p.quirk_mask >> (2u * sig[i])) & 3u; }
      X[i] = nX;
    }
  }

  if (K == 2 && p.has_tail) {                     /* last base of an odd-length read */
    const uint32_t pos = 4u * p.nsteps;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint32_t c = (myq[i][pos >> 5] >> (pos & 31u)) & 3u;
      if (p.tail1) {                                  /* the lanes of a pair share the fetch when L and R fall in one block */
        const uint32_t b = fm_div96(X[i]);
        X[i] = fm_block_rank(fm_ldg16(p.tail1 + (size_t) c * p.nblocks + b), X[i] - b * FM_SB_ROWS);
      } else {
        X[i] = fm_tail_rank(p.blocks, p.nblocks, c, X[i], p.tail_const[c], p.tail_row, p.tail_base);
      }
    }
  }

  #pragma unroll
  for (int i = 0; i < QPT; i++)
    if (live[i]) p.results[2 * (size_t)(q0 + i * PAIRS + pair) + side] = X[i];
}

/* ------------------------------------------------------------------------ *
 * ASCII -> reversed 2-bit packing.  Packed position t of a read is base
 * len-1-t, so LF step s of a k-step search consumes bit field
 * [2k*s, 2k*s+2k) and the field value IS the k-step symbol
 * code(q[j]) | code(q[j-1])<<2 ... of src/fmIndexCPUBaseline.c:213-226.
 * Replaces the host-side warp interleave of common/common.c:175-194.
 * ------------------------------------------------------------------------ */
__global__ void fm_pack_kernel(const char *__restrict__ ascii, uint64_t nq, uint32_t len, uint32_t wpq,
                               uint32_t *__restrict__ packed)
{
  const uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nq * wpq) return;
  const uint64_t q = idx / wpq;
  const uint32_t w = (uint32_t)(idx - q * wpq);
  const char *rd = ascii + q * len;
  uint32_t out = 0;
  #pragma unroll
  for (uint32_t i = 0; i < 16; i++) {
    const uint32_t t = 16 * w + i;
    if (t < len) {
      const uint32_t c = (uint32_t)(unsigned char) rd[len - 1 - t];
      const uint32_t hi = (c >> 2) & 1u, mid = (c >> 1) & 1u;
      out |= ((hi << 1) | (hi ^ mid)) << (2 * i);
    }
  }
  packed[idx] = out;
}

/* Host-packed stream -> per-read reversed words.  The host packer (fm_hostpack_stream) emits the whole chunk as one
 * 2-bit sequence, base g = q*len + i at bits [2(g%16), 2(g%16)+2) of 32-bit word g/16, doing no per-read work; this
 * kernel cuts it into reads, reverses each and aligns it to words: out field i of word w = base len-1-(16w+i). */
__global__ void fm_unstream_kernel(const uint32_t *__restrict__ stream, uint64_t nq, uint32_t len, uint32_t wpq,
                                   uint32_t *__restrict__ packed)
{
  const uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nq * wpq) return;
  const uint64_t q = idx / wpq;
  const uint32_t w = (uint32_t)(idx - q * wpq);
  const uint64_t first = q * len, hi = first + len - 16ull * w;        /* bases [lo, hi) of the stream feed this word */
  const uint64_t lo = (hi >= first + 16) ? hi - 16 : first;
  const uint32_t nv = (uint32_t)(hi - lo);                             /* 1..16 valid bases */
  const uint64_t bit0 = 2 * lo;
  const uint32_t a = stream[bit0 >> 5], b = stream[(bit0 >> 5) + 1];
  uint32_t x = __funnelshift_r(a, b, (uint32_t)(bit0 & 31));
  if (nv < 16) x &= (1u << (2 * nv)) - 1u;
  x <<= 2 * (16 - nv);
  x = __brev(x);                                                        /* reverses the 16 fields and the 2 bits inside each */
  packed[idx] = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);    /* put the 2 bits of every field back in order */
}

/* ------------------------------------------------------------------------ *
 * Re-blocker: raw file entries (tags 100/101/200/201) -> SB96.
 * ------------------------------------------------------------------------ */
__device__ __forceinline__ bool fm_raw_is_ac(const FmRawIndex &x)  { return x.tag >= 200; }
__device__ __forceinline__ bool fm_raw_is_il(const FmRawIndex &x)  { return (x.tag & 1u) != 0; }

/* word n of plane `bit` of BWT layer s (App. A of SURVEY.md) */
__device__ __forceinline__ uint32_t fm_raw_plane(const FmRawIndex &x, uint32_t entry, uint32_t s, uint32_t bit, uint32_t n)
{
  const uint32_t W = x.d / 32;
  const uint32_t *e = x.entries + (size_t) entry * x.entry_words + (fm_raw_is_ac(x) ? x.ncounters : 0u);
  return fm_raw_is_il(x) ? e[2 * x.k * n + 2 * s + bit] : e[2 * W * s + W * bit + n];
}

__device__ __forceinline__ uint32_t fm_raw_counter(const FmRawIndex &x, uint32_t entry, uint32_t slot)
{
  const uint32_t *e = x.entries + (size_t) entry * x.entry_words;
  return fm_raw_is_ac(x) ? e[slot] : e[2 * (x.d / 32) * x.k + slot];
}

/* rows (MSB-first, as stored) of word n of `entry` whose symbol is sigma */
__device__ __forceinline__ uint32_t fm_raw_match(const FmRawIndex &x, uint32_t entry, uint32_t n, uint32_t sigma)
{
  uint32_t m = 0xFFFFFFFFu;
  for (uint32_t s = 0; s < x.k; s++) {
    const uint32_t c = (sigma >> (2 * s)) & 3u;
    const uint32_t p0 = fm_raw_plane(x, entry, s, 0, n), p1 = fm_raw_plane(x, entry, s, 1, n);
    m &= ((c & 1u) ? p0 : ~p0) & ((c & 2u) ? p1 : ~p1);
  }
  return m;
}

/* Value the matching reference CPU searcher yields for (sigma, X), X a
 * multiple of 32 with X <= bwtsize: literal counter + popcount - '$' fix. */
__device__ uint32_t fm_raw_rank(const FmRawIndex &x, uint32_t sigma, uint32_t X)
{
  const uint32_t d = x.d, W = d / 32;
  uint32_t e = X / d;
  if (e >= x.nentries_std) e = x.nentries_std - 1;     /* X == bwtsize on a chunk boundary: count the whole last chunk */
  const uint32_t r = X - e * d;                        /* 0..d, multiple of 32 */
  const uint32_t full = r / 32;
  bool next = false;
  uint32_t cnt = 0, fix = 0;
  if (fm_raw_is_ac(x)) {
    const uint32_t H = x.ncounters;
    next = ((e & 1u) && sigma < H) || (!(e & 1u) && sigma >= H);
  }
  if (!next) { for (uint32_t n = 0; n < full; n++) cnt += __popc(fm_raw_match(x, e, n, sigma)); }
  else       { for (uint32_t n = full; n < W; n++) cnt += __popc(fm_raw_match(x, e, n, sigma)); }
  for (uint32_t s = 0; s < x.k; s++)
    if (x.dpos[s] / d == e && sigma == x.dbase[s]) {
      if (!next && X >  x.dpos[s]) fix++;
      if ( next && X <= x.dpos[s]) fix++;
    }
  if (!next) return fm_raw_counter(x, e, fm_raw_is_ac(x) ? (sigma & (x.ncounters - 1)) : sigma) + (cnt - fix);
  uint32_t v = fm_raw_counter(x, e + 1, sigma & (x.ncounters - 1)) - (cnt - fix);
  /* block counters hold the quirk-free value; the kernel adds the quirk back for X >= quirk_start */
  if (X >= x.quirk_start) v -= (x.quirk_mask >> (2u * sigma)) & 3u;
  return v;
}

__global__ void fm_reblock_kernel(const FmRawIndex x, uint4 *__restrict__ blocks, uint32_t nblocks)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint64_t p = (uint64_t) b * FM_SB_ROWS;
  const uint32_t nsym = 1u << (2 * x.k);
  if (p > x.bwtsize) {
    for (uint32_t sigma = 0; sigma < nsym; sigma++) blocks[(size_t) sigma * nblocks + b] = make_uint4(0, 0, 0, 0);
    return;
  }
  uint32_t ent[3], wn[3], keep[3];
  for (int j = 0; j < 3; j++) {
    const uint64_t pos = p + 32u * j;
    ent[j] = (uint32_t)(pos / x.d);
    wn[j]  = (uint32_t)(pos % x.d) / 32;
    const int64_t nvalid = (int64_t) x.bwtsize - (int64_t) pos;          /* rows of this word below bwtsize */
    keep[j] = nvalid >= 32 ? 0xFFFFFFFFu : (nvalid <= 0 ? 0u : ~(0xFFFFFFFFu >> nvalid));
    if (ent[j] >= x.nentries_std) keep[j] = 0u;
  }
  for (uint32_t sigma = 0; sigma < nsym; sigma++) {
    uint32_t w[3];
    for (int j = 0; j < 3; j++) {
      uint32_t m = keep[j] ? (fm_raw_match(x, ent[j], wn[j], sigma) & keep[j]) : 0u;
      const uint64_t pos = p + 32u * j;
      for (uint32_t s = 0; s < x.k; s++)
        if (sigma == x.dbase[s] && x.dpos[s] >= pos && x.dpos[s] < pos + 32) m &= ~(0x80000000u >> (x.dpos[s] - pos));
      w[j] = __brev(m);                                                   /* row i of the word -> bit i */
    }
    blocks[(size_t) sigma * nblocks + b] = make_uint4(fm_raw_rank(x, sigma, (uint32_t) p), w[0], w[1], w[2]);
  }
}

/* ------------------------------------------------------------------------ *
 * Gather roofline probe: independent uniformly random aligned accesses of
 * WIDTH consecutive 16-byte loads (16, 32, 64 or 128 bytes per access).
 * ------------------------------------------------------------------------ */
template <int UNROLL, int WIDTH>
__global__ void __launch_bounds__(256, 8) fm_gather_probe_kernel(const uint4 *__restrict__ table, uint64_t naccess,
                                                                   uint32_t loads_per_thread, uint32_t *sink)
{
  uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  uint32_t acc = 0;
  for (uint32_t it = 0; it < loads_per_thread; it += UNROLL) {
    uint4 v[UNROLL][WIDTH];
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;                        /* xorshift64 */
      const uint64_t idx = __umul64hi(s, naccess);                     /* uniform in [0, naccess) */
      #pragma unroll
      for (int w = 0; w < WIDTH; w++) v[u][w] = fm_ldg16(table + idx * WIDTH + w);
    }
    #pragma unroll
    for (int u = 0; u < UNROLL; u++)
      #pragma unroll
      for (int w = 0; w < WIDTH; w++) acc += v[u][w].x ^ v[u][w].y ^ v[u][w].z ^ v[u][w].w;
  }
  if (acc == 0x9E3779B9u) *sink = acc;                                 /* keeps the loads alive */
}

/* Locality probe: every warp-level load picks ONE random window of `window16` 16-byte blocks (the same for
 * its 32 lanes) and each lane a random block inside it.  Separates address-translation cost (one 2 MB page
 * per warp instruction) from DRAM sector cost (32 distinct sectors per warp instruction either way). */
template <int UNROLL>
__global__ void __launch_bounds__(256, 8) fm_gather_probe_local_kernel(const uint4 *__restrict__ table, uint64_t nwindows,
                                                                         uint32_t window16, uint32_t loads_per_thread,
                                                                         uint32_t *sink)
{
  const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t sw = (tid >> 5) * 0x9E3779B97F4A7C15ull + 0x7654321ull;    /* warp-uniform stream */
  uint64_t sl = tid * 0xD1B54A32D192ED03ull + 0x1234567ull;           /* per-lane stream */
  uint32_t acc = 0;
  for (uint32_t it = 0; it < loads_per_thread; it += UNROLL) {
    uint4 v[UNROLL];
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      sw ^= sw << 13; sw ^= sw >> 7; sw ^= sw << 17;
      sl ^= sl << 13; sl ^= sl >> 7; sl ^= sl << 17;
      const uint64_t win = __umul64hi(sw, nwindows);
      const uint32_t off = __umulhi((uint32_t)(sl >> 32), window16);
      v[u] = fm_ldg16(table + win * window16 + off);
    }
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x9E3779B9u) *sink = acc;
}

#endif /* FM_KERNELS_CUH_ */

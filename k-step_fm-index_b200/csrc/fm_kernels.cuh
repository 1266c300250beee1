/*
 * fm_kernels.cuh -- plain k-step search kernels on the SB96 table (Task and Coop), read packing.  Included by
 * fm_search.cu only.  Layout and shared helpers: fm_device.cuh.
 */
#ifndef FM_KERNELS_CUH_
#define FM_KERNELS_CUH_

#include "fm_device.cuh"

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
      FM_BOUND(bL, p.nblocks, "task: SB96 block (L)"); FM_BOUND(bR, p.nblocks, "task: SB96 block (R)");
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
      FM_BOUND(b, p.nblocks, "coop: SB96 block");
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
        FM_BOUND(b, p.nblocks, "coop: tail table block");
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

#endif /* FM_KERNELS_CUH_ */

/*
 * fm_fused.cuh -- "fused-step" device layout and search kernel.
 *
 * Measured on B200 (profiles/r01_line_variants.md): every L2 miss moves a whole 128-byte line, and a random
 * access costs the same (~43-46 G/s) whether ONE warp-level instruction fetches 16, 32, 64 or 128 bytes of it.
 * The plain SB96 kernels use 16 of those bytes.  This layout spends the rest on MORE QUERY BASES PER FETCH:
 *
 *   fused symbol of row i   F(i) = s(i) | s(LF(i)) << 2k | s(LF(LF(i))) << 4k ...   (m = KF/k hops)
 *     where s() is the index's own k-step symbol and LF its own k-step mapping, both read from the SB96
 *     table that was derived from the reference file -- so one fused step IS m consecutive reference LF steps:
 *         rank_F(sigma_F, X) = rank(sigma_{m-1}, ... rank(sigma_1, rank(sigma_0, X)))      for every X,
 *     exactly (increments of the composed function are the indicator [F(i) = sigma_F]; rows whose chain meets
 *     a '$' row carry no symbol).  Bit-exactness of the search therefore follows from that of SB96.
 *
 *   fused block  = LANES x 32 bytes  = { u32 rank_F at block start, (256*LANES - 32) indicator bits }
 *   table        = fblocks[(sigma_F * nfblocks + b) * 2*LANES .. ),  sigma_F < 4^KF
 *   one rank     = ONE 256-bit load (ld.global.v8.b32, new on sm_100) per lane of a LANES-lane group, the
 *                  group's loads coalescing into one 32/64/128-byte request; popcounts are split over the
 *                  lanes and summed with warp shuffles.
 *
 * A 100-bp read takes 25 fused steps of 4 bases instead of 50 two-base steps: half the dependent DRAM line
 * fetches.  Table size is 4^KF * LANES*32 B per (256*LANES-32) rows: KF=4, LANES=2: 34.1 B/base (68 GB for
 * 2 Gbp, the edge of the flat part of the footprint curve); KF=3: 8.5 B/base; KF=2: 2.1 B/base.
 * Read lengths that are not a multiple of KF do their (len/k) % m leading steps on the SB96 table.
 *
 * AltCounters files that carry the padding-entry quirk: the composed function then jumps by 2 at a few rows, which a
 * bitmap cannot hold -- those phantom occurrences (at most FM_MAX_FUSED_PHANTOMS) ride in the kernel parameters.
 */
#ifndef FM_FUSED_CUH_
#define FM_FUSED_CUH_

#include "fm_device.cuh"

#define FM_MAX_FUSED_PHANTOMS 16u

struct FmFusedParams {
  const uint4    *fblocks;    /* fused table                                                       */
  const uint4    *blocks;     /* SB96, for the leading base steps                                  */
  const uint32_t *packed;
  uint32_t       *results;
  uint32_t nfblocks;          /* fused blocks per fused symbol                                     */
  uint32_t nblocks;           /* SB96 stride                                                       */
  uint32_t nq;
  uint32_t nlead;             /* leading base-k steps                                              */
  uint32_t nfused;            /* fused steps                                                       */
  uint32_t wpq, wpq_pad;      /* words per packed read (wpq_pad unused: reads keep their natural stride) */
  uint32_t bwtsize;
  unsigned long long *fetch_counters;  /* COUNT only: [0] = fused blocks fetched, [1] = SB96 blocks of the leading steps */
  uint32_t has_tail, tail_row, tail_base, tail_const[4];   /* odd read length on a 2-step index, see fm_tail_rank */
  const uint4 *tail1;         /* tail table (fm_tail_table_kernel) or NULL */
  /* start table: (L,R) after the first FM_START_BASES bases, indexed by their 24 packed bits -- the intervals this
   * very kernel computes for all 4^12 12-mers, so looking them up instead of stepping changes nothing in the result;
   * it replaces the L2-resident steps and the first DRAM step (two fetches) by one mostly-L2 lookup */
  const uint2 *start;
  uint32_t start_steps;       /* fused steps the table stands for (FM_START_BASES / KF), 0 = no table */
  /* AltCounters padding-entry quirk (fm_device.cuh: fm_quirk_phantoms_kernel): the composed rank function gains one for every
   * phantom occurrence (fused symbol, row) with row < X -- a handful of them, kept beside the bitmaps (a bit cannot be set
   * twice); the SB96 leading steps add the per-symbol constant.  nph = 0 and quirk_mask = 0 for every other index. */
  uint32_t quirk_start, quirk_mask;
  uint32_t nph, ph_sym[FM_MAX_FUSED_PHANTOMS], ph_row[FM_MAX_FUSED_PHANTOMS];
};

#define FM_START_BASES 12u


/* rows per fused block = 32 * (8*LANES - 1); exact division of X < 2^32 by it */
template <int LANES> struct FmFusedGeom {
  static constexpr uint32_t ROWS = 32u * (8u * LANES - 1u);
  static constexpr uint32_t MAGIC = LANES == 1 ? 613566757u : (LANES == 2 ? 286331154u : 138547333u);  /* ceil(2^32/(8*LANES-1)) */
  __host__ __device__ static uint32_t div(uint32_t x)
  {
#ifdef __CUDA_ARCH__
    return __umulhi(x >> 5, MAGIC);
#else
    return (uint32_t)(((uint64_t)(x >> 5) * MAGIC) >> 32);
#endif
  }
};

/* this lane's share of rank_F: set bits of its 256-bit chunk that lie below row offset r of the block
 * (bit string of a block: 32 counter bits, then the indicator bits; lane l holds bits [256 l, 256 l + 256)) */
__device__ __forceinline__ uint32_t fm_fused_partial(const uint32_t (&w)[8], uint32_t r, uint32_t lane_in_group)
{
  const int t = (int) r + 32 - 256 * (int) lane_in_group;          /* prefix length inside this chunk (may be <0 or >256) */
  uint32_t sum = 0;
  #pragma unroll
  for (int c = 0; c < 8; c++) {
    const uint32_t n = (uint32_t) min(max(t - 32 * c, 0), 32);
    sum += __popc(w[c] & fm_lowmask(n));
  }
  return sum;
}


/* ------------------------------------------------------------------------ *
 * Fused search: a group of LANES lanes owns QPT reads (both endpoints of each).  Per fused step the group
 * issues ONE coalesced fetch of the block of L (and a second one only when R lies in another block).
 * ------------------------------------------------------------------------ */
template <int KF, int K, int LANES, int QPT, int THREADS, int MINB, bool COUNT>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_fused_kernel(const FmFusedParams p)
{
  extern __shared__ __align__(16) uint32_t fsm[];             /* [0..3]: mbarrier (8 B) + pad; [4..): packed reads, natural stride */
  uint32_t *sq = fsm + 4;
  constexpr uint32_t FBITS = 2 * KF, FMASK = (1u << FBITS) - 1u, BBITS = 2 * K, BMASK = (1u << BBITS) - 1u;
  constexpr uint32_t ROWS = FmFusedGeom<LANES>::ROWS;
  constexpr int GROUPS = THREADS / LANES;
  const uint32_t q0 = blockIdx.x * (GROUPS * QPT);
  const uint32_t nqb = min((uint32_t)(GROUPS * QPT), p.nq - q0);
  const uint32_t lg = threadIdx.x % LANES, group = threadIdx.x / LANES;

  /* Stage this CTA's packed reads (one contiguous piece of the batch) in shared memory with ONE TMA bulk copy
   * (cp.async.bulk global -> shared, completion on an mbarrier); a tail CTA whose piece is not a multiple of
   * 16 bytes copies it with plain loads.  Reads keep their natural stride: a bit field never straddles out of a
   * read, so the word after a read may be anything. */
  {
    const uint32_t bytes = nqb * p.wpq * 4u;
    const uint32_t *src = p.packed + (size_t) q0 * p.wpq;
    const bool bulk = (bytes % 16u) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    const uint32_t mbar = (uint32_t) __cvta_generic_to_shared(fsm);
    if (bulk) {
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((uint32_t) __cvta_generic_to_shared(sq)), "l"(src), "r"(bytes), "r"(mbar) : "memory");
      }
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar) : "memory");
    } else {
      for (uint32_t i = threadIdx.x; i < nqb * p.wpq; i += THREADS) sq[i] = __ldg(src + i);
      __syncthreads();
    }
  }

  uint32_t L[QPT], R[QPT];
  const uint32_t *myq[QPT];
  bool live[QPT];
  #pragma unroll
  for (int i = 0; i < QPT; i++) {
    const uint32_t lq = i * GROUPS + group;
    live[i] = lq < nqb;
    myq[i] = sq + (live[i] ? lq : 0u) * p.wpq;
    L[i] = 0u; R[i] = p.bwtsize;
  }

  unsigned long long nf_fetch = 0, nl_fetch = 0;
  /* leading base-k steps on SB96 (every lane of the group computes them redundantly: same address, one request) */
  uint32_t pos = 0;
  for (uint32_t step = 0; step < p.nlead; step++, pos += BBITS) {
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint32_t sig = fm_read_field(myq[i], pos, BMASK);
      const uint32_t bL = fm_div96(L[i]), bR = fm_div96(R[i]);
      const uint4 *base = p.blocks + (size_t) sig * p.nblocks;
      FM_BOUND(bL, p.nblocks, "fused: SB96 block (L)"); FM_BOUND(bR, p.nblocks, "fused: SB96 block (R)");
      const uint4 vL = fm_ldg16(base + bL);
      const uint4 vR = (bL == bR) ? vL : fm_ldg16(base + bR);
      if (COUNT && live[i] && lg == 0) nl_fetch += (bL == bR) ? 1 : 2;
      const uint32_t nL = fm_block_rank(vL, L[i] - bL * FM_SB_ROWS) + fm_quirk_delta(p.quirk_mask, p.quirk_start, sig, L[i]);
      const uint32_t nR = fm_block_rank(vR, R[i] - bR * FM_SB_ROWS) + fm_quirk_delta(p.quirk_mask, p.quirk_start, sig, R[i]);
      L[i] = nL; R[i] = nR;
    }
  }

  uint32_t step0 = 0;
  if (p.start_steps && p.nlead == 0 && p.nfused >= p.start_steps) {
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint2 lr = __ldg(p.start + (myq[i][0] & 0xFFFFFFu));
      L[i] = lr.x; R[i] = lr.y;
    }
    step0 = p.start_steps; pos = 2u * FM_START_BASES;
  }

  for (uint32_t step = step0; step < p.nfused; step++, pos += FBITS) {
    uint32_t wL[QPT][8], wR[QPT][8], rL[QPT], rR[QPT], phL[QPT], phR[QPT];
    bool same[QPT];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint32_t sig = fm_read_field(myq[i], pos, FMASK);
      phL[i] = 0u; phR[i] = 0u;
      if (p.nph) {                                             /* quirk indexes only: phantom occurrences below L / R */
        for (uint32_t j = 0; j < p.nph; j++)
          if (p.ph_sym[j] == sig) { phL[i] += p.ph_row[j] < L[i] ? 1u : 0u; phR[i] += p.ph_row[j] < R[i] ? 1u : 0u; }
      }
      const uint32_t bL = FmFusedGeom<LANES>::div(L[i]), bR = FmFusedGeom<LANES>::div(R[i]);
      rL[i] = L[i] - bL * ROWS; rR[i] = R[i] - bR * ROWS;
      const uint4 *base = p.fblocks + ((size_t) sig * p.nfblocks) * (2 * LANES) + 2 * lg;
      same[i] = (bL == bR);
      if (COUNT && live[i] && lg == 0) nf_fetch += same[i] ? 1 : 2;
      FM_BOUND(bL, p.nfblocks, "fused block (L)"); FM_BOUND(bR, p.nfblocks, "fused block (R)");
      fm_ldg32(base + (size_t) bL * (2 * LANES), wL[i]);
      if (!same[i]) fm_ldg32(base + (size_t) bR * (2 * LANES), wR[i]);
    }
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      if (same[i]) {
        #pragma unroll
        for (int c = 0; c < 8; c++) wR[i][c] = wL[i][c];
      }
      uint32_t cL = 0, cR = 0;
      if (lg == 0) { cL = wL[i][0]; cR = wR[i][0]; wL[i][0] = 0u; wR[i][0] = 0u; }   /* word 0 of the block is the sampled rank */
      cL += fm_fused_partial(wL[i], rL[i], lg);
      cR += fm_fused_partial(wR[i], rR[i], lg);
      L[i] = fm_group_sum<LANES>(cL) + phL[i];
      R[i] = fm_group_sum<LANES>(cR) + phR[i];
    }
  }

  if (K == 2 && p.has_tail) {                     /* last base of an odd-length read (every lane of the group, same addresses) */
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint32_t c = fm_read_field(myq[i], pos, 3u);
      fm_tail_step(p.tail1, p.blocks, p.nblocks, c, L[i], R[i], p.tail_const[c], p.tail_row, p.tail_base);
    }
  }

  if (lg == 0) {
    #pragma unroll
    for (int i = 0; i < QPT; i++)
      if (live[i]) reinterpret_cast<uint2 *>(p.results)[q0 + i * GROUPS + group] = make_uint2(L[i], R[i]);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      nf_fetch += __shfl_xor_sync(0xFFFFFFFFu, nf_fetch, o);
      nl_fetch += __shfl_xor_sync(0xFFFFFFFFu, nl_fetch, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.fetch_counters, nf_fetch); atomicAdd(p.fetch_counters + 1, nl_fetch); }
  }
}

/* ------------------------------------------------------------------------ *
 * Construction of the fused table from SB96 (all on the device)
 * ------------------------------------------------------------------------ */

/* SB96 -> k-step symbol of every row (FM_SYM_NONE for '$' rows and rows >= bwtsize); one thread per SB96 block */
__global__ void fm_fuse_symbols_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t nsym, uint64_t nrows_alloc,
                                       uint8_t *__restrict__ sym)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint64_t row0 = (uint64_t) b * FM_SB_ROWS;
  for (uint32_t i = 0; i < FM_SB_ROWS && row0 + i < nrows_alloc; i++) sym[row0 + i] = FM_SYM_NONE;
  for (uint32_t s = 0; s < nsym; s++) {
    const uint4 v = blocks[(size_t) s * nblocks + b];
    uint32_t w[3] = { v.y, v.z, v.w };
    for (int j = 0; j < 3; j++)
      while (w[j]) {
        const uint32_t bit = __ffs(w[j]) - 1;
        w[j] &= w[j] - 1;
        const uint64_t row = row0 + 32u * j + bit;
        if (row < nrows_alloc) sym[row] = (uint8_t) s;
      }
  }
}


/* fused symbol of every row: follow the index's own LF mapping hops-1 times */
__global__ void fm_fuse_compose_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, const uint8_t *__restrict__ sym,
                                       uint32_t bwtsize, uint32_t kbits, uint32_t hops, uint64_t nrows_alloc,
                                       uint32_t quirk_start, uint32_t quirk_mask, FmQuirkVisit *__restrict__ visits, uint32_t *__restrict__ nvisits,
                                       uint32_t max_visits, uint16_t *__restrict__ fsym)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows_alloc) return;
  uint32_t f = FM_FSYM_NONE;
  if (i < bwtsize) {
    uint32_t row = (uint32_t) i, acc = 0;
    bool ok = true;
    for (uint32_t h = 0; h < hops; h++) {
      if (quirk_mask && quirk_start != 0u && row == quirk_start - 1u) {   /* phantom copies may branch off here (fm_quirk_phantoms_kernel) */
        const uint32_t slot = atomicAdd(nvisits, 1u);
        if (slot < max_visits) { visits[slot].origin = (uint32_t) i; visits[slot].hop = h; visits[slot].acc = acc; }
      }
      const uint32_t s = sym[row];
      if (s == FM_SYM_NONE) { ok = false; break; }
      acc |= s << (kbits * h);
      if (h + 1 < hops) row = fm_sb96_rank_q(blocks, nblocks, s, row, quirk_start, quirk_mask);       /* LF(row), as the file's searcher computes it */
    }
    if (ok) f = acc;
  }
  fsym[i] = (uint16_t) f;
}

/* one CTA per fused block: indicator words of all fused symbols in shared memory, then one block per symbol
 * is written; word 0 receives the number of rows of the block carrying the symbol (turned into ranks by the scan) */
template <int LANES>
__global__ void __launch_bounds__(256) fm_fuse_write_kernel(const uint16_t *__restrict__ fsym, uint32_t nfsym, uint32_t nfblocks,
                                                            uint4 *__restrict__ fblocks)
{
  extern __shared__ uint32_t bits[];                        /* [chunk symbols][8*LANES words] */
  constexpr uint32_t ROWS = FmFusedGeom<LANES>::ROWS, WORDS = 8 * LANES;
  const uint32_t fb = blockIdx.x;
  const uint64_t row0 = (uint64_t) fb * ROWS;
  /* symbols are processed in chunks of 256 so that shared memory stays at 256*WORDS*4 bytes */
  for (uint32_t s0 = 0; s0 < nfsym; s0 += 256) {
    for (uint32_t i = threadIdx.x; i < 256 * WORDS; i += blockDim.x) bits[i] = 0u;
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < ROWS; r += blockDim.x) {
      const uint32_t f = fsym[row0 + r];
      if (f != FM_FSYM_NONE && f >= s0 && f < s0 + 256) atomicOr(&bits[(f - s0) * WORDS + 1 + (r >> 5)], 1u << (r & 31));
    }
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < 256 && s0 + s < nfsym; s += blockDim.x) {
      uint32_t *w = bits + s * WORDS;
      uint32_t cnt = 0;
      for (uint32_t c = 1; c < WORDS; c++) cnt += __popc(w[c]);
      w[0] = cnt;
      uint4 *dst = fblocks + ((size_t)(s0 + s) * nfblocks + fb) * (2 * LANES);
      for (uint32_t c = 0; c < 2 * LANES; c++) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
    }
    __syncthreads();
  }
}

/* one CTA per fused symbol: word 0 of its blocks := rank_F(sigma_F, block start) = composed rank at X = 0 plus
 * the exclusive prefix sum of the per-block counts */
template <int LANES>
__global__ void __launch_bounds__(1024) fm_fuse_scan_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t kbits,
                                                            uint32_t hops, uint32_t nfblocks, uint32_t quirk_start, uint32_t quirk_mask,
                                                            uint4 *__restrict__ fblocks)
{
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  const uint32_t sigma = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    uint32_t x = 0;
    for (uint32_t h = 0; h < hops; h++) x = fm_sb96_rank_q(blocks, nblocks, (sigma >> (kbits * h)) & ((1u << kbits) - 1u), x, quirk_start, quirk_mask);
    carry = x;
  }
  __syncthreads();
  uint32_t *w0 = reinterpret_cast<uint32_t *>(fblocks + (size_t) sigma * nfblocks * (2 * LANES));
  const uint32_t stride = 8 * LANES;                          /* words per block */
  for (uint32_t base = 0; base < nfblocks; base += blockDim.x) {
    const uint32_t b = base + threadIdx.x;
    const uint32_t v = b < nfblocks ? w0[(size_t) b * stride] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      uint32_t t = warp_tot[lane];
      for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += u; }
      warp_tot[lane] = t;                                     /* inclusive totals of the warps */
    }
    __syncthreads();
    const uint32_t before = carry + (wid ? warp_tot[wid - 1] : 0u) + (inc - v);
    if (b < nfblocks) w0[(size_t) b * stride] = before;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
}

#endif /* FM_FUSED_CUH_ */


/*
 * fm_sparse.cuh -- "sparse-step" device layout and search kernel: KS (up to 14) query bases per block fetch.
 *
 * Measured on B200 (profiles/r01_miss_ceiling.md): the memory system serves ~46 G random block fetches per second
 * whatever their size up to a 128-byte line, and nothing else limits the search.  The fused-step layout
 * (fm_fused.cuh) spends one fetch on 4 bases and cannot go further: its per-symbol indicator bitmaps cost
 * 4^KF bits per row.  But the indicator of ONE wide symbol is almost empty (density 4^-KS), so this layout stores
 * the set bits themselves:
 *
 *   wide symbol of row i    F(i) = s(i) | s(LF(i)) << 2k | ...   (hops = KS/k hops, exactly as in fm_fused.cuh, so
 *                           rank_F(sigma, X) = rank_F(sigma, 0) + #{ i < X : F(i) = sigma } IS `hops` consecutive
 *                           reference LF steps, for every X; rows whose chain meets a '$' row carry no symbol)
 *   occurrence list         rows i with F(i) = sigma, ascending  (a stable radix sort of (F(i), i))
 *   buckets                 symbol sigma owns nb(sigma) = max(1, ceil(count(sigma) / lambda)) blocks;
 *                           row X belongs to bucket umulhi(X, scale(sigma)) -- a monotone map of [0, bwtsize] onto
 *                           [0, nb) -- so a bucket holds ~lambda occurrences whatever the symbol's frequency
 *   block (32*LANES bytes)  word 0      = rank_F(sigma, first row of the bucket)
 *                           words 1..   = the bucket's occurrence rows, ascending, padded with 0xFFFFFFFF
 *                           (31 slots in a 128-byte block, LANES = 4; 15 in a 64-byte block, LANES = 2)
 *                           a bucket with more occurrences than slots stores 0xFFFFFFFE in the last word instead
 *   directory               dir[sigma] = { first block, scale }     (8 bytes x 4^KS: L2-resident for KS <= 10)
 *   uniform grid            when every symbol occurs about equally often (a uniformly random text: no count above 1.6 x the
 *                           mean or below 0.4 x of it) all symbols get the SAME number of blocks, first block = sigma * nb, one scale:
 *                           the directory lookup -- an L2 request per step that costs 15 % of the fetch rate -- is
 *                           replaced by a multiplication.  Skewed texts (any real genome) keep per-symbol block counts.
 *
 *   one rank     = dir[sigma] (L2 hit; sigma is known in advance, the lookup is issued one step ahead)
 *                  + ONE block fetch by LANES lanes x 256 bits;  rank = word 0 + #{ entries < X }
 *   L and R      almost always share the block (a bucket spans bwtsize / nb rows), so a step is one fetch
 *   overflow     a block marked 0xFFFFFFFE sends that read through `hops` ordinary steps on the SB96 table for
 *                this symbol -- exact, and rare by construction (Poisson tail for lambda = 16: 2e-4 per fetch on a
 *                random text; repeats of a real genome cost speed, never correctness)
 *
 * Table size is ~128 / lambda bytes per text base for ANY KS (8 B/base at lambda = 16: 16 GB for 2 Gbp, a quarter
 * of the fused table) and a 100-bp read needs 10 fetches at KS = 10 instead of 25 (9 after the start table).
 * Not available for AltCounters files carrying the padding-entry quirk, like the fused table.
 */
#ifndef FM_SPARSE_CUH_
#define FM_SPARSE_CUH_

#include "fm_device.cuh"

#define FM_SP_PAD   0xFFFFFFFFu
#define FM_SP_OVF   0xFFFFFFFEu
#define FM_SP_NONE  0xFFFFFFFFu            /* compose output of a row without a wide symbol (sorted last) */

struct FmSparseParams {
  const uint4    *sblocks;    /* sparse table, 2*LANES uint4 per block                             */
  const uint2    *dir;        /* per wide symbol: { first block, scale }                           */
  const uint4    *blocks;     /* SB96: leading steps, overflow fallback, odd tail                  */
  const uint32_t *packed;
  uint32_t       *results;
  uint32_t nblocks;           /* SB96 stride                                                       */
  uint32_t nq;
  uint32_t nfront, nback;     /* base-k steps on SB96 in front of / behind the sparse steps       */
  uint32_t nsteps;            /* sparse steps (after the start table's bases, if one is used)     */
  uint32_t wpq;
  uint32_t bwtsize;
  uint32_t sbits;             /* 2 * KS                                                            */
  uint32_t hops;              /* KS / k                                                            */
  unsigned long long *fetch_counters;  /* COUNT only: [0] sparse blocks, [1] SB96 blocks (leading + fallback), [2] overflow fallbacks */
  uint32_t has_tail, tail_row, tail_base, tail_const[4];
  const uint4 *tail1;         /* tail table (fm_tail_table_kernel) or NULL */
  uint32_t uni_nb, uni_scale; /* uniform grid (every symbol owns uni_nb blocks, first block = sigma * uni_nb): no directory
                                 lookup; 0 = per-symbol block counts, read from dir                                  */
  const uint2 *start;         /* (L,R) after the first start_bits / 2 bases, indexed by those packed bits, or NULL: the start
                                 table (a whole number of sparse steps) or a lead table (the leftover bases, taken first) */
  uint32_t start_bits;
};

/* directory entry { first block, scale } of a wide symbol: computed when the table is a uniform grid, else one L2-resident
 * lookup (which costs request slots beside the block fetches: profiles/r01_hit_miss_mix.md) */
__device__ __forceinline__ uint2 fm_sparse_dir(const FmSparseParams &p, uint32_t sig)
{
  return p.uni_nb ? make_uint2(sig * p.uni_nb, p.uni_scale) : __ldg(p.dir + sig);
}

__device__ __forceinline__ void fm_ldg32_line(const uint4 *p, uint32_t (&w)[8])
{
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

/* this lane's share of #{ entries < X }: lane 0 skips word 0 (the sampled rank) */
__device__ __forceinline__ uint32_t fm_sparse_partial(const uint32_t (&w)[8], uint32_t X, uint32_t lg)
{
  uint32_t c = (lg != 0u && w[0] < X) ? 1u : 0u;
  #pragma unroll
  for (int j = 1; j < 8; j++) c += (w[j] < X) ? 1u : 0u;
  return c;
}

template <int LANES> __device__ __forceinline__ void fm_sparse_load(const uint4 *p, uint32_t (&w)[8])
{
  if (LANES == 2) fm_ldg32(p, w);          /* 64-byte block: 64-byte L2 fill */
  else            fm_ldg32_line(p, w);
}

template <int K, int LANES, int QPT, int THREADS, int MINB, bool COUNT>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_sparse_kernel(const FmSparseParams p)
{
  extern __shared__ __align__(16) uint32_t fsm[];             /* [0..3]: mbarrier + pad; [4..): packed reads, natural stride */
  uint32_t *sq = fsm + 4;
  constexpr uint32_t BBITS = 2 * K, BMASK = (1u << BBITS) - 1u;
  constexpr int GROUPS = THREADS / LANES;
  constexpr uint32_t BU4 = 2u * LANES;                        /* uint4 per block */
  const uint32_t smask = (p.sbits >= 32u) ? 0xFFFFFFFFu : ((1u << p.sbits) - 1u);
  const uint32_t q0 = blockIdx.x * (GROUPS * QPT);
  const uint32_t nqb = min((uint32_t)(GROUPS * QPT), p.nq - q0);
  const uint32_t lg = threadIdx.x % LANES, group = threadIdx.x / LANES;

  /* packed reads of this CTA -> shared memory: one TMA bulk copy (see fm_search_fused_kernel) */
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

  unsigned long long n_sp = 0, n_sb = 0, n_ovf = 0;
  uint32_t pos = 0;
  /* the (len/k) % hops base-k steps that do not fill a sparse step run on SB96: in FRONT of the sparse steps (wide
   * interval, upper levels of the table, L2 hits) when there is no table to start from, BEHIND them when the start
   * table replaces the first sparse step (one DRAM block per step there); with 6 or more leftover bases a lead
   * table takes them first instead and every sparse step runs (fm_launch_sparse) */
  auto base_steps = [&](uint32_t count) {
    for (uint32_t step = 0; step < count; step++, pos += BBITS) {
      #pragma unroll
      for (int i = 0; i < QPT; i++) {
        const uint32_t sg = fm_read_field(myq[i], pos, BMASK);
        const uint32_t bL = fm_div96(L[i]), bR = fm_div96(R[i]);
        const uint4 *base = p.blocks + (size_t) sg * p.nblocks;
        const uint4 vL = fm_ldg16(base + bL);
        const uint4 vR = (bL == bR) ? vL : fm_ldg16(base + bR);
        if (COUNT && live[i] && lg == 0) n_sb += (bL == bR) ? 1 : 2;
        L[i] = fm_block_rank(vL, L[i] - bL * FM_SB_ROWS);
        R[i] = fm_block_rank(vR, R[i] - bR * FM_SB_ROWS);
      }
    }
  };
  if (p.start) {                                               /* the host picked the plan (fm_launch_sparse) */
    const uint32_t kmask = (p.start_bits >= 32u) ? 0xFFFFFFFFu : ((1u << p.start_bits) - 1u);
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint2 lr = __ldg(p.start + (myq[i][0] & kmask));
      L[i] = lr.x; R[i] = lr.y;
    }
    pos = p.start_bits;
  }
  base_steps(p.nfront);
  const uint32_t step0 = 0;

  uint32_t sig[QPT];
  uint2 d[QPT];
  if (step0 < p.nsteps) {
    #pragma unroll
    for (int i = 0; i < QPT; i++) { sig[i] = fm_read_field(myq[i], pos, smask); d[i] = fm_sparse_dir(p, sig[i]); }
  }
  for (uint32_t step = step0; step < p.nsteps; step++) {
    uint32_t w[QPT][8];
    const uint4 *aR[QPT];
    bool same[QPT];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const uint32_t bL = __umulhi(L[i], d[i].y), bR = __umulhi(R[i], d[i].y);
      const uint4 *base = p.sblocks + (size_t) d[i].x * BU4 + 2u * lg;
      same[i] = (bL == bR);
      aR[i] = base + (size_t) bR * BU4;
      if (COUNT && live[i] && lg == 0) n_sp += same[i] ? 1 : 2;
      fm_sparse_load<LANES>(base + (size_t) bL * BU4, w[i]);
    }
    /* directory entries of the NEXT step (independent of L,R): in flight together with the block fetches */
    uint32_t sig_now[QPT];
    pos += p.sbits;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      sig_now[i] = sig[i];
      if (step + 1 < p.nsteps) { sig[i] = fm_read_field(myq[i], pos, smask); d[i] = fm_sparse_dir(p, sig[i]); }
    }
    uint32_t cL[QPT], cR[QPT];
    bool ovf[QPT];
    bool any_far = false;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      cL[i] = fm_sparse_partial(w[i], L[i], lg);
      cR[i] = fm_sparse_partial(w[i], R[i], lg);
      if (lg == 0) { cL[i] += w[i][0]; cR[i] += w[i][0]; }
      ovf[i] = (lg == LANES - 1) && (w[i][7] == FM_SP_OVF);
      any_far |= !same[i];
    }
    if (__any_sync(0xFFFFFFFFu, any_far)) {                   /* rare: R lies in another bucket than L */
      #pragma unroll
      for (int i = 0; i < QPT; i++) {
        if (!same[i]) {
          uint32_t v[8];
          fm_sparse_load<LANES>(aR[i], v);
          cR[i] = fm_sparse_partial(v, R[i], lg) + (lg == 0 ? v[0] : 0u);
          ovf[i] |= (lg == LANES - 1) && (v[7] == FM_SP_OVF);
        }
      }
      __syncwarp();
    }
    uint32_t nL[QPT], nR[QPT], ob[QPT];
    bool any_ovf = false;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      nL[i] = fm_group_sum<LANES>(cL[i]);
      nR[i] = fm_group_sum<LANES>(cR[i]);
      ob[i] = (__ballot_sync(0xFFFFFFFFu, ovf[i]) >> ((threadIdx.x & 31u) & ~(uint32_t)(LANES - 1))) & ((1u << LANES) - 1u);
      any_ovf |= (ob[i] != 0u);
    }
    if (__any_sync(0xFFFFFFFFu, any_ovf)) {                   /* rare: overfull bucket -> the same `hops` steps on SB96 */
      #pragma unroll
      for (int i = 0; i < QPT; i++) {
        if (ob[i]) {
          uint32_t xl = L[i], xr = R[i];
          for (uint32_t h = 0; h < p.hops; h++) {
            const uint32_t s = (sig_now[i] >> (BBITS * h)) & BMASK;
            const uint32_t bL = fm_div96(xl), bR = fm_div96(xr);
            const uint4 *base = p.blocks + (size_t) s * p.nblocks;
            const uint4 vL = fm_ldg16(base + bL);
            const uint4 vR = (bL == bR) ? vL : fm_ldg16(base + bR);
            if (COUNT && live[i] && lg == 0) n_sb += (bL == bR) ? 1 : 2;
            xl = fm_block_rank(vL, xl - bL * FM_SB_ROWS);
            xr = fm_block_rank(vR, xr - bR * FM_SB_ROWS);
          }
          nL[i] = xl; nR[i] = xr;
          if (COUNT && live[i] && lg == 0) n_ovf += 1;
        }
      }
      __syncwarp();
    }
    #pragma unroll
    for (int i = 0; i < QPT; i++) { L[i] = nL[i]; R[i] = nR[i]; }
  }

  base_steps(p.nback);

  if (K == 2 && p.has_tail) {                     /* last base of an odd-length read */
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
      n_sp += __shfl_xor_sync(0xFFFFFFFFu, n_sp, o);
      n_sb += __shfl_xor_sync(0xFFFFFFFFu, n_sb, o);
      n_ovf += __shfl_xor_sync(0xFFFFFFFFu, n_ovf, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.fetch_counters, n_sp); atomicAdd(p.fetch_counters + 1, n_sb); atomicAdd(p.fetch_counters + 2, n_ovf); }
  }
}

/* ------------------------------------------------------------------------ *
 * Construction from SB96 (all on the device)
 * ------------------------------------------------------------------------ */

/* wide symbol of every row < bwtsize (the LF chain of fm_fuse_compose_kernel, 32-bit output) and its row number;
 * rows without a symbol get the key nsym, which sorts behind every symbol */
__global__ void fm_sparse_compose_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, const uint8_t *__restrict__ sym,
                                         uint32_t bwtsize, uint32_t kbits, uint32_t hops, uint32_t nsym,
                                         uint32_t *__restrict__ keys, uint32_t *__restrict__ rows)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bwtsize) return;
  uint32_t row = (uint32_t) i, acc = 0;
  bool ok = true;
  for (uint32_t h = 0; h < hops; h++) {
    const uint32_t s = sym[row];
    if (s == FM_SYM_NONE) { ok = false; break; }
    acc |= s << (kbits * h);
    if (h + 1 < hops) row = fm_sb96_rank(blocks, nblocks, s, row);
  }
  keys[i] = ok ? acc : nsym;
  rows[i] = (uint32_t) i;
}

/* symstart[s] = first position of key >= s in the sorted key array, s = 0..nsym (symstart[nsym] = rows carrying a symbol) */
__global__ void fm_sparse_symstart_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t nsym, uint32_t *__restrict__ symstart)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nsym) return;
  uint32_t lo = 0, hi = n;
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < s) lo = mid + 1; else hi = mid; }
  symstart[s] = lo;
}

/* smallest and largest occurrence count over the symbols: range[0] (preset to 0xFFFFFFFF) and range[1] (preset to 0) */
__global__ void fm_sparse_count_range_kernel(const uint32_t *__restrict__ symstart, uint32_t nsym, uint32_t *__restrict__ range)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  if (s < nsym) lo = hi = symstart[s + 1] - symstart[s];
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
  }
  if ((threadIdx.x & 31u) == 0) { atomicMin(range, lo); atomicMax(range + 1, hi); }
}

/* blocks per symbol: ceil(count / lambda), or the same `uniform_nb` for every symbol (uniform grid) */
__global__ void fm_sparse_nblocks_kernel(const uint32_t *__restrict__ symstart, uint32_t nsym, uint32_t lambda, uint32_t uniform_nb,
                                         uint32_t *__restrict__ nb)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsym) return;
  const uint32_t cnt = symstart[s + 1] - symstart[s];
  nb[s] = uniform_nb ? uniform_nb : (cnt ? (cnt + lambda - 1) / lambda : 1u);
}

/* directory entry and rank_F(sigma, 0) of every symbol */
__global__ void fm_sparse_dir_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t kbits, uint32_t hops, uint32_t nsym,
                                     uint32_t bwtsize, const uint32_t *__restrict__ nb, const uint32_t *__restrict__ first,
                                     uint2 *__restrict__ dir, uint32_t *__restrict__ rank0)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsym) return;
  /* largest scale with umulhi(bwtsize, scale) <= nb - 1, i.e. bwtsize * scale < nb * 2^32 */
  unsigned long long sc = ((((unsigned long long) nb[s]) << 32) - 1ull) / bwtsize;
  if (sc > 0xFFFFFFFFull) sc = 0xFFFFFFFFull;
  dir[s] = make_uint2(first[s], (uint32_t) sc);
  uint32_t x = 0;
  for (uint32_t h = 0; h < hops; h++) x = fm_sb96_rank(blocks, nblocks, (s >> (kbits * h)) & ((1u << kbits) - 1u), x);
  rank0[s] = x;
}

/* one CTA per symbol, one thread per block of the symbol */
template <int LANES>
__global__ void __launch_bounds__(128) fm_sparse_fill_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ symstart,
                                                             const uint2 *__restrict__ dir, const uint32_t *__restrict__ nb,
                                                             const uint32_t *__restrict__ rank0, uint4 *__restrict__ sblocks,
                                                             unsigned long long *__restrict__ novf)
{
  const uint32_t s = blockIdx.x;
  const uint32_t s0 = symstart[s], s1 = symstart[s + 1], scale = dir[s].y, first = dir[s].x, n = nb[s], r0 = rank0[s];
  for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
    uint32_t lo = s0, hi = s1;                                 /* first occurrence whose bucket is >= j */
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__umulhi(rows[mid], scale) < j) lo = mid + 1; else hi = mid; }
    const uint32_t t0 = lo;
    hi = s1;                                                   /* first occurrence whose bucket is > j */
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__umulhi(rows[mid], scale) <= j) lo = mid + 1; else hi = mid; }
    const uint32_t cnt = lo - t0;
    constexpr uint32_t WORDS = 8u * LANES, SLOTS = WORDS - 1u;
    uint32_t w[WORDS];
    w[0] = r0 + (t0 - s0);
    #pragma unroll
    for (uint32_t c = 1; c < WORDS; c++) w[c] = (c - 1 < cnt && cnt <= SLOTS) ? rows[t0 + c - 1] : FM_SP_PAD;
    if (cnt > SLOTS) { w[WORDS - 1] = FM_SP_OVF; atomicAdd(novf, 1ull); }
    uint4 *dst = sblocks + (size_t)(first + j) * (2u * LANES);
    #pragma unroll
    for (uint32_t c = 0; c < 2u * LANES; c++) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  }
}

/* the same for a uniform grid with many symbols and few blocks each (14 bases: 2^28 symbols x 2 blocks): one THREAD per block */
template <int LANES>
__global__ void __launch_bounds__(256) fm_sparse_fill_uniform_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ symstart,
                                                                     uint32_t nsym, uint32_t nbu, uint32_t scale, const uint32_t *__restrict__ rank0,
                                                                     uint4 *__restrict__ sblocks, unsigned long long *__restrict__ novf)
{
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t) nsym * nbu) return;
  const uint32_t s = (uint32_t)(g / nbu), j = (uint32_t)(g - (uint64_t) s * nbu);
  const uint32_t s0 = symstart[s], s1 = symstart[s + 1], r0 = rank0[s];
  uint32_t lo = s0, hi = s1;
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__umulhi(rows[mid], scale) < j) lo = mid + 1; else hi = mid; }
  const uint32_t t0 = lo;
  hi = s1;
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__umulhi(rows[mid], scale) <= j) lo = mid + 1; else hi = mid; }
  const uint32_t cnt = lo - t0;
  constexpr uint32_t WORDS = 8u * LANES, SLOTS = WORDS - 1u;
  uint32_t w[WORDS];
  w[0] = r0 + (t0 - s0);
  #pragma unroll
  for (uint32_t c = 1; c < WORDS; c++) w[c] = (c - 1 < cnt && cnt <= SLOTS) ? rows[t0 + c - 1] : FM_SP_PAD;
  if (cnt > SLOTS) { w[WORDS - 1] = FM_SP_OVF; atomicAdd(novf, 1ull); }
  uint4 *dst = sblocks + (size_t) g * (2u * LANES);
  #pragma unroll
  for (uint32_t c = 0; c < 2u * LANES; c++) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

/* rows that live in symbols a uniform grid of nbu blocks per symbol cannot hold (count > slots * nbu): out[0] += count */
__global__ void fm_sparse_heavy_rows_kernel(const uint32_t *__restrict__ symstart, uint32_t nsym, uint32_t limit, unsigned long long *__restrict__ out)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long c = 0;
  if (s < nsym) { const uint32_t cnt = symstart[s + 1] - symstart[s]; if (cnt > limit) c = cnt; }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31u) == 0 && c) atomicAdd(out, c);
}

#endif /* FM_SPARSE_CUH_ */

/*
 * fm_sparse.cuh -- "sparse-step" device layout and search kernel: KS (up to 14) query bases per block fetch, on ANY text.
 *
 * Measured on B200 (profiles/r01_miss_ceiling.md, r02_ceiling_counters.md): the memory system serves ~46 G random block
 * fetches per second whatever their size up to a 128-byte line, and nothing else limits the search.  The fused-step
 * layout (fm_fused.cuh) spends one fetch on 4 bases and cannot go further: its per-symbol indicator bitmaps cost 4^KF
 * bits per row.  But the indicator of ONE wide symbol is almost empty (density 4^-KS), so this layout stores the set
 * bits themselves:
 *
 *   wide symbol of row i    F(i) = s(i) | s(LF(i)) << 2k | ...   (hops = KS/k hops, exactly as in fm_fused.cuh, so
 *                           rank_F(sigma, X) = rank_F(sigma, 0) + #{ i < X : F(i) = sigma } IS `hops` consecutive
 *                           reference LF steps, for every X; rows whose chain meets a '$' row carry no symbol)
 *   occurrence list         rows i with F(i) = sigma, ascending  (a stable radix sort of (F(i), i)).  An AltCounters
 *                           file with an active padding-entry quirk adds a few "phantom" occurrences -- the composed
 *                           rank function then jumps by 2 at one row -- which are just list entries (a row may repeat).
 *   node (32*LANES bytes)   word 0, then SLOTS = 8*LANES - 1 ascending u32 entries padded with 0xFFFFFFFF:
 *        leaf               word 0 = rank_F(sigma, .) before the first entry; entries = occurrence rows.
 *                           rank_F(sigma, X) = word 0 + #{ entries < X }
 *        inner node         last entry = 0xFFFFFFFE (marker); the other SLOTS - 1 entries are separators = the first
 *                           occurrence row of children 1 .. SLOTS - 1; word 0 = block number of child 0.
 *                           child of X = word 0 + #{ entries < X }  -- the SAME arithmetic as a leaf's rank
 *   uniform grid (roots)    every wide symbol owns nb blocks; row X belongs to bucket umulhi(X, scale), a monotone map
 *                           of [0, bwtsize] onto [0, nb); root of (sigma, X) = block sigma * nb + umulhi(X, scale):
 *                           computed, never looked up (a directory lookup per step costs 15 % of the fetch rate,
 *                           profiles/r01_hit_miss_mix.md).  A bucket with <= SLOTS occurrences is a leaf: on a uniformly
 *                           random text 99.99 % of them, so a step is ONE fetch.
 *   search tree (skew)      a bucket with more occurrences is the root of a tree over ITS occurrence list: leaves of
 *                           SLOTS consecutive occurrences, fan-out SLOTS, all leaves at the same depth, nodes stored
 *                           behind the grid (level by level).  Subdividing by occurrence RANK (not by row range) keeps
 *                           the depth at ceil(log_SLOTS(count / SLOTS)) however the occurrences cluster -- repeats
 *                           of a genome put thousands of them in a few BWT runs.  Heavy symbols cost depth + 1 fetches
 *                           instead of `hops` fetches of the plain table; nothing falls back to SB96 any more.
 *
 * Kernel: every lane group runs its OWN state machine per read -- one block fetch per iteration, whatever that read
 * needs next (root of the next step, a child, the other interval end's node) -- so a read that walks a deep tree only
 * delays itself, not the 63 other reads of its warp (r01_repeat_text.md: in lockstep, fewer fetches ran slower).
 *
 * Table size: grid = ~32*LANES/lambda bytes per text base for ANY KS (12.8 B/base at the default lambda 5, LANES 2),
 * plus 32*LANES/(SLOTS-1) bytes per occurrence living in an overfull bucket (at most 4.6 B/base).
 */
#ifndef FM_SPARSE_CUH_
#define FM_SPARSE_CUH_

#include "fm_device.cuh"

#define FM_SP_PAD   0xFFFFFFFFu
#define FM_SP_INNER 0xFFFFFFFEu            /* last word of an inner node (bwtsize < this, so no row collides) */
#define FM_SP_NONE  0xFFFFFFFFu            /* compose output of a row without a wide symbol (sorted last) */
#define FM_SP_DONE  0xFFFFFFFFu            /* state machine: this interval end has its new value */
#define FM_SP_MAXDEPTH 9

struct FmSparseParams {
  const uint4    *sblocks;    /* grid + tree nodes, 2*LANES uint4 per block                        */
  const uint4    *blocks;     /* SB96: leftover base-k steps, odd tail                             */
  const uint32_t *packed;
  uint32_t       *results;
  uint32_t nblocks;           /* SB96 stride                                                       */
  uint32_t nq;
  uint32_t nfront, nback;     /* base-k steps on SB96 in front of / behind the sparse steps       */
  uint32_t nsteps;            /* sparse steps (after the start table's bases, if one is used)     */
  uint32_t wpq;
  uint32_t bwtsize;
  uint32_t sbits;             /* 2 * KS                                                            */
  unsigned long long *fetch_counters;  /* COUNT only: [0] root fetches, [1] SB96 blocks (leftover steps), [2] tree-node fetches below the roots */
  uint32_t has_tail, tail_row, tail_base, tail_const[4];
  const uint4 *tail1;         /* tail table (fm_tail_table_kernel) or NULL */
  uint32_t nb, scale;         /* grid: blocks per wide symbol, bucket = umulhi(X, scale)          */
  uint32_t nroots;            /* nb * 4^KS: blocks at or beyond it are tree nodes                  */
  uint32_t total_blocks;      /* grid + tree nodes (extent of sblocks, checked by the -DFM_DEBUG_BOUNDS build) */
  const uint2 *start;         /* (L,R) after the first start_bits / 2 bases, indexed by those packed bits, or NULL: the start
                                 table (a whole number of sparse steps) or a lead table (the leftover bases, taken first) */
  uint32_t start_bits;
  uint32_t quirk_start, quirk_mask;   /* AltCounters padding quirk, for the SB96 base steps (0xFFFFFFFF / 0 = none) */
};

__device__ __forceinline__ void fm_ldg32_line(const uint4 *p, uint32_t (&w)[8])
{
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

/* this lane's share of #{ entries < X }: lane 0 skips word 0 */
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

  unsigned long long n_root = 0, n_sb = 0, n_tree = 0;
  uint32_t pos = 0;
  /* the (len/k) % hops base-k steps that do not fill a sparse step run on SB96: in FRONT of the sparse steps (wide
   * interval, upper levels of the table, L2 hits) when there is no table to start from, BEHIND them when the start
   * table replaces the first sparse step (one DRAM block per step there); with 6 or more leftover bases a lead
   * table takes them first instead and every sparse step runs (fm_launch_sparse).  All reads of the CTA do the same
   * number of these, so they stay in lockstep. */
  auto base_steps = [&](uint32_t count) {
    for (uint32_t step = 0; step < count; step++, pos += BBITS) {
      #pragma unroll
      for (int i = 0; i < QPT; i++) {
        const uint32_t sg = fm_read_field(myq[i], pos, BMASK);
        const uint32_t bL = fm_div96(L[i]), bR = fm_div96(R[i]);
        const uint4 *base = p.blocks + (size_t) sg * p.nblocks;
        FM_BOUND(bL, p.nblocks, "sparse: SB96 block (L)"); FM_BOUND(bR, p.nblocks, "sparse: SB96 block (R)");
        const uint4 vL = fm_ldg16(base + bL);
        const uint4 vR = (bL == bR) ? vL : fm_ldg16(base + bR);
        if (COUNT && live[i] && lg == 0) n_sb += (bL == bR) ? 1 : 2;
        const uint32_t nL = fm_block_rank(vL, L[i] - bL * FM_SB_ROWS) + fm_quirk_delta(p.quirk_mask, p.quirk_start, sg, L[i]);
        const uint32_t nR = fm_block_rank(vR, R[i] - bR * FM_SB_ROWS) + fm_quirk_delta(p.quirk_mask, p.quirk_start, sg, R[i]);
        L[i] = nL; R[i] = nR;
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

  /* ---- sparse steps: one state machine per read.  aL / aR = block to fetch next for that interval end, FM_SP_DONE
   * once the end holds its value for the NEXT step; rem = sparse steps not yet finished.  Every iteration fetches
   * exactly one block per unfinished read: the node both ends share, else L's node, else R's. */
  uint32_t aL[QPT], aR[QPT], rem[QPT];
  bool busy = false;
  #pragma unroll
  for (int i = 0; i < QPT; i++) {
    rem[i] = live[i] ? p.nsteps : 0u;
    aL[i] = aR[i] = FM_SP_DONE;
    if (rem[i]) {
      const uint32_t first = fm_read_field(myq[i], pos, smask) * p.nb;
      aL[i] = first + __umulhi(L[i], p.scale);
      aR[i] = first + __umulhi(R[i], p.scale);
    }
    busy |= rem[i] != 0u;
  }
  while (__any_sync(0xFFFFFFFFu, busy)) {
    uint32_t w[QPT][8];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      if (rem[i]) {
        const uint32_t a = (aL[i] != FM_SP_DONE) ? aL[i] : aR[i];
        FM_BOUND(a, p.total_blocks, "sparse: grid / tree block");
        fm_sparse_load<LANES>(p.sblocks + (size_t) a * BU4 + 2u * lg, w[i]);
        if (COUNT && lg == 0) { if (a < p.nroots) n_root++; else n_tree++; }
      }
    }
    busy = false;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      /* the lanes of a group hold the same state, so they take the same branches */
      const bool act = rem[i] != 0u;
      const bool doL = act && aL[i] != FM_SP_DONE;
      const bool doR = act && (!doL || aR[i] == aL[i]);                  /* R alone, or riding on the node it shares with L */
      uint32_t cL = 0, cR = 0, inner = 0;
      if (act) {
        cL = fm_sparse_partial(w[i], L[i], lg);
        cR = fm_sparse_partial(w[i], R[i], lg);
        if (lg == 0) { cL += w[i][0]; cR += w[i][0]; }
        inner = (lg == LANES - 1 && w[i][7] == FM_SP_INNER) ? 1u : 0u;
      }
      const uint32_t vL = fm_group_sum<LANES>(cL), vR = fm_group_sum<LANES>(cR);
      const bool is_inner = fm_group_sum<LANES>(inner) != 0u;
      if (doL) { if (is_inner) aL[i] = vL; else { L[i] = vL; aL[i] = FM_SP_DONE; } }
      if (doR) { if (is_inner) aR[i] = vR; else { R[i] = vR; aR[i] = FM_SP_DONE; } }
      if (act && aL[i] == FM_SP_DONE && aR[i] == FM_SP_DONE) {           /* step complete: next step's roots */
        rem[i] -= 1u;
        if (rem[i]) {
          const uint32_t first = fm_read_field(myq[i], pos + (p.nsteps - rem[i]) * p.sbits, smask) * p.nb;
          aL[i] = first + __umulhi(L[i], p.scale);
          aR[i] = first + __umulhi(R[i], p.scale);
        }
      }
      busy |= rem[i] != 0u;
    }
  }
  pos += p.nsteps * p.sbits;

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
      n_root += __shfl_xor_sync(0xFFFFFFFFu, n_root, o);
      n_sb += __shfl_xor_sync(0xFFFFFFFFu, n_sb, o);
      n_tree += __shfl_xor_sync(0xFFFFFFFFu, n_tree, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.fetch_counters, n_root); atomicAdd(p.fetch_counters + 1, n_sb); atomicAdd(p.fetch_counters + 2, n_tree); }
  }
}

/* ------------------------------------------------------------------------ *
 * The same search with DYNAMIC read assignment, for plans that consist of sparse steps only (start / lead table or
 * nothing in front, no SB96 steps, no odd tail -- fm_launch_sparse picks it then): the CTA stages ROUNDS x as many
 * reads as it has slots, and a slot (one of QPT per lane group) that finishes a read pulls the next one from a
 * shared-memory counter.  On a skewed text reads need 7 to 40 fetches; with static assignment every warp lasts as long
 * as its slowest read and half of the fetch slots idle (profiles/r02_skewed_text.md), here only the CTA's last reads do.
 * The start-table lookup of a new read is one more state of the machine (aL = FM_SP_START), so every iteration still
 * issues exactly one load per busy slot.
 * ------------------------------------------------------------------------ */
#define FM_SP_START 0xFFFFFFFEu

template <int K, int LANES, int QPT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_sparse_dyn_kernel(const FmSparseParams p, uint32_t reads_per_cta)
{
  extern __shared__ __align__(16) uint32_t fsm[];             /* [0..1]: mbarrier; [2]: next read; [4..): packed reads */
  uint32_t *sq = fsm + 4;
  constexpr uint32_t BU4 = 2u * LANES;
  const uint32_t smask = (p.sbits >= 32u) ? 0xFFFFFFFFu : ((1u << p.sbits) - 1u);
  const uint32_t kmask = (p.start_bits >= 32u) ? 0xFFFFFFFFu : ((1u << p.start_bits) - 1u);
  const uint32_t q0 = blockIdx.x * reads_per_cta;
  const uint32_t nqb = min(reads_per_cta, p.nq - q0);
  const uint32_t lg = threadIdx.x % LANES;
  {
    const uint32_t bytes = nqb * p.wpq * 4u;
    const uint32_t *src = p.packed + (size_t) q0 * p.wpq;
    const bool bulk = (bytes % 16u) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    const uint32_t mbar = (uint32_t) __cvta_generic_to_shared(fsm);
    if (threadIdx.x == 0) fsm[2] = 0u;
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

  uint32_t L[QPT], R[QPT], aL[QPT], aR[QPT], rem[QPT], rd[QPT];
  /* takes the next read of the CTA for slot i (or parks the slot): all lanes of the warp call it together */
  auto take = [&](int i, bool need) {
    uint32_t r = 0;
    if (need && lg == 0) r = atomicAdd(&fsm[2], 1u);
    r = __shfl_sync(0xFFFFFFFFu, r, 0, LANES);
    if (need) {
      rd[i] = r;
      if (r < nqb) {
        rem[i] = p.nsteps; L[i] = 0u; R[i] = p.bwtsize;
        if (p.start) { aL[i] = FM_SP_START; aR[i] = FM_SP_DONE; }
        else if (rem[i]) {
          const uint32_t first = fm_read_field(sq + r * p.wpq, 0u, smask) * p.nb;
          aL[i] = first; aR[i] = first + __umulhi(p.bwtsize, p.scale);
        } else { aL[i] = aR[i] = FM_SP_DONE; }
      } else { rem[i] = 0u; aL[i] = aR[i] = FM_SP_DONE; rd[i] = 0xFFFFFFFFu; }
    }
  };
  #pragma unroll
  for (int i = 0; i < QPT; i++) { rd[i] = 0xFFFFFFFFu; rem[i] = 0u; aL[i] = aR[i] = FM_SP_DONE; L[i] = R[i] = 0u; take(i, true); }

  bool busy = false;
  #pragma unroll
  for (int i = 0; i < QPT; i++) busy |= rd[i] != 0xFFFFFFFFu;
  while (__any_sync(0xFFFFFFFFu, busy)) {
    uint32_t w[QPT][8];
    uint2 lr[QPT];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      if (rd[i] != 0xFFFFFFFFu) {
        if (aL[i] == FM_SP_START) lr[i] = __ldg(p.start + (sq[rd[i] * p.wpq] & kmask));
        else if (rem[i]) {
          const uint32_t a = (aL[i] != FM_SP_DONE) ? aL[i] : aR[i];
          FM_BOUND(a, p.total_blocks, "sparse (dynamic): grid / tree block"); FM_BOUND(rd[i], nqb, "sparse (dynamic): read slot");
          fm_sparse_load<LANES>(p.sblocks + (size_t) a * BU4 + 2u * lg, w[i]);
        }
      }
    }
    busy = false;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const bool have = rd[i] != 0xFFFFFFFFu;
      const bool starting = have && aL[i] == FM_SP_START;
      const bool act = have && !starting && rem[i] != 0u;
      const bool doL = act && aL[i] != FM_SP_DONE;
      const bool doR = act && (!doL || aR[i] == aL[i]);
      uint32_t cL = 0, cR = 0, inner = 0;
      if (act) {
        cL = fm_sparse_partial(w[i], L[i], lg);
        cR = fm_sparse_partial(w[i], R[i], lg);
        if (lg == 0) { cL += w[i][0]; cR += w[i][0]; }
        inner = (lg == LANES - 1 && w[i][7] == FM_SP_INNER) ? 1u : 0u;
      }
      const uint32_t vL = fm_group_sum<LANES>(cL), vR = fm_group_sum<LANES>(cR);
      const bool is_inner = fm_group_sum<LANES>(inner) != 0u;
      if (doL) { if (is_inner) aL[i] = vL; else { L[i] = vL; aL[i] = FM_SP_DONE; } }
      if (doR) { if (is_inner) aR[i] = vR; else { R[i] = vR; aR[i] = FM_SP_DONE; } }
      bool next_roots = false;
      if (starting) { L[i] = lr[i].x; R[i] = lr[i].y; aL[i] = aR[i] = FM_SP_DONE; next_roots = rem[i] != 0u; }
      else if (act && aL[i] == FM_SP_DONE && aR[i] == FM_SP_DONE) { rem[i] -= 1u; next_roots = rem[i] != 0u; }
      if (next_roots) {
        const uint32_t first = fm_read_field(sq + rd[i] * p.wpq, p.start_bits + (p.nsteps - rem[i]) * p.sbits, smask) * p.nb;
        aL[i] = first + __umulhi(L[i], p.scale);
        aR[i] = first + __umulhi(R[i], p.scale);
      }
      const bool finished = have && rem[i] == 0u && aL[i] == FM_SP_DONE;
      if (finished && lg == 0) reinterpret_cast<uint2 *>(p.results)[q0 + rd[i]] = make_uint2(L[i], R[i]);
      take(i, finished);
      busy |= rd[i] != 0xFFFFFFFFu;
    }
  }
}

/* ------------------------------------------------------------------------ *
 * Construction from SB96 (all on the device)
 * ------------------------------------------------------------------------ */

/* wide symbol of every row < bwtsize (the LF chain of fm_fuse_compose_kernel, 32-bit output) and its row number;
 * rows without a symbol get the key nsym, which sorts behind every symbol.  With an active quirk the chain follows the
 * quirked rank, and every visit of row quirk_start - 1 (where the quirked rank functions have their extra jump) is
 * reported in `visits` = { origin row, hop, symbols so far }: fm_quirk_phantoms_kernel adds the phantom occurrences. */
__global__ void fm_sparse_compose_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, const uint8_t *__restrict__ sym,
                                         uint32_t bwtsize, uint32_t kbits, uint32_t hops, uint32_t nsym,
                                         uint32_t quirk_start, uint32_t quirk_mask, FmQuirkVisit *__restrict__ visits, uint32_t *__restrict__ nvisits,
                                         uint32_t max_visits, uint32_t *__restrict__ keys, uint32_t *__restrict__ rows)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bwtsize) return;
  uint32_t row = (uint32_t) i, acc = 0;
  bool ok = true;
  for (uint32_t h = 0; h < hops; h++) {
    if (quirk_mask && quirk_start != 0u && row == quirk_start - 1u) {
      const uint32_t slot = atomicAdd(nvisits, 1u);
      if (slot < max_visits) { visits[slot].origin = (uint32_t) i; visits[slot].hop = h; visits[slot].acc = acc; }
    }
    const uint32_t s = sym[row];
    if (s == FM_SYM_NONE) { ok = false; break; }
    acc |= s << (kbits * h);
    if (h + 1 < hops) row = fm_sb96_rank_q(blocks, nblocks, s, row, quirk_start, quirk_mask);
  }
  keys[i] = ok ? acc : nsym;
  rows[i] = (uint32_t) i;
}

/* symstart[s] = first position of key >= s in the sorted key array, s = 0..nsym (symstart[nsym] = rows carrying a symbol) */
__global__ void fm_sparse_symstart_kernel(const uint32_t *__restrict__ keys, uint64_t n, uint32_t nsym, uint32_t *__restrict__ symstart)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nsym) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (keys[mid] < s) lo = mid + 1; else hi = mid; }
  symstart[s] = (uint32_t) lo;
}

/* rank_F(sigma, 0) of every wide symbol: the composed (quirked) rank at X = 0 */
__global__ void fm_sparse_rank0_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t kbits, uint32_t hops, uint32_t nsym,
                                       uint32_t quirk_start, uint32_t quirk_mask, uint32_t *__restrict__ rank0)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsym) return;
  uint32_t x = 0;
  for (uint32_t h = 0; h < hops; h++) x = fm_sb96_rank_q(blocks, nblocks, (s >> (kbits * h)) & ((1u << kbits) - 1u), x, quirk_start, quirk_mask);
  rank0[s] = x;
}

/* occurrences [t0, t0 + cnt) of the sorted row list that fall in bucket j of symbol s */
__device__ __forceinline__ void fm_sparse_bucket_range(const uint32_t *__restrict__ rows, uint32_t s0, uint32_t s1, uint32_t scale, uint32_t j,
                                                       uint32_t &t0, uint32_t &cnt)
{
  uint32_t lo = s0, hi = s1;                                   /* first occurrence whose bucket is >= j */
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__umulhi(rows[mid], scale) < j) lo = mid + 1; else hi = mid; }
  t0 = lo;
  hi = s1;                                                     /* first occurrence whose bucket is > j */
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__umulhi(rows[mid], scale) <= j) lo = mid + 1; else hi = mid; }
  cnt = lo - t0;
}

/* shape of the tree over cnt > SLOTS occurrences: N[v] = nodes of level v (0 = leaves), depth D with N[D] = 1 (the root,
 * which lives in the grid); nodes below the root = sum of N[0 .. D-1] */
template <int LANES> struct FmSparseTree {
  static constexpr uint32_t SLOTS = 8u * LANES - 1u, FAN = SLOTS;
  uint32_t N[FM_SP_MAXDEPTH + 1];
  uint32_t D;
  __host__ __device__ explicit FmSparseTree(uint32_t cnt)
  {
    N[0] = (cnt + SLOTS - 1) / SLOTS;
    D = 0;
    while (N[D] > 1 && D < FM_SP_MAXDEPTH) { N[D + 1] = (N[D] + FAN - 1) / FAN; D++; }
  }
  __host__ __device__ uint32_t below_root() const { uint32_t t = 0; for (uint32_t v = 0; v < D; v++) t += N[v]; return t; }
  /* offset of level v's first node inside the root's extension area (levels stored top-down: D-1 first, leaves last) */
  __host__ __device__ uint32_t level_offset(uint32_t v) const { uint32_t t = 0; for (uint32_t u = v + 1; u < D; u++) t += N[u]; return t; }
};

/* pass 1: extension nodes every root needs (0 for a bucket that fits its block); one thread per root */
template <int LANES>
__global__ void __launch_bounds__(256) fm_sparse_count_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ symstart,
                                                              uint32_t nsym, uint32_t nb, uint32_t scale, uint32_t *__restrict__ ext,
                                                              unsigned long long *__restrict__ stats)
{
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t) nsym * nb) return;
  const uint32_t s = (uint32_t)(g / nb), j = (uint32_t)(g - (uint64_t) s * nb);
  uint32_t t0, cnt;
  fm_sparse_bucket_range(rows, symstart[s], symstart[s + 1], scale, j, t0, cnt);
  uint32_t e = 0;
  if (cnt > FmSparseTree<LANES>::SLOTS) {
    const FmSparseTree<LANES> t(cnt);
    e = t.below_root();
    atomicAdd(stats, 1ull);                                    /* overfull buckets */
    atomicAdd(stats + 1, (unsigned long long) cnt);            /* occurrences living in them */
    atomicMax(stats + 2, (unsigned long long) t.D);            /* deepest tree */
  }
  ext[g] = e;
}

/* one node of a tree: level v, index m, over occurrences occ[0 .. cnt) of its root; `area` = first block of the root's
 * extension area, rank_before = rank_F before occ[0] */
template <int LANES>
__device__ __forceinline__ void fm_sparse_write_node(const FmSparseTree<LANES> &t, uint32_t v, uint32_t m, const uint32_t *__restrict__ occ,
                                                     uint32_t cnt, uint32_t area, uint32_t rank_before, uint4 *__restrict__ dst)
{
  constexpr uint32_t WORDS = 8u * LANES, SLOTS = WORDS - 1u, FAN = SLOTS;
  uint32_t w[WORDS];
  if (v == 0) {
    const uint64_t first = (uint64_t) m * SLOTS;
    w[0] = rank_before + (uint32_t) first;
    #pragma unroll
    for (uint32_t c = 1; c < WORDS; c++) w[c] = (first + c - 1 < cnt) ? occ[first + c - 1] : FM_SP_PAD;
  } else {
    uint64_t span = SLOTS;                                     /* occurrences under one child: SLOTS * FAN^(v-1) */
    for (uint32_t u = 1; u < v; u++) span *= FAN;
    w[0] = area + t.level_offset(v - 1) + m * FAN;             /* block of child 0 */
    #pragma unroll
    for (uint32_t c = 1; c < WORDS - 1; c++) {
      const uint64_t child = (uint64_t) m * FAN + c, at = child * span;
      w[c] = (child < t.N[v - 1] && at < cnt) ? occ[at] : FM_SP_PAD;
    }
    w[WORDS - 1] = FM_SP_INNER;
  }
  #pragma unroll
  for (uint32_t c = 0; c < 2u * LANES; c++) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

/* pass 2a: the grid -- a leaf, or the root of a tree; one thread per root.  extoff = exclusive scan of ext. */
template <int LANES>
__global__ void __launch_bounds__(256) fm_sparse_fill_roots_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ symstart,
                                                                   uint32_t nsym, uint32_t nb, uint32_t scale, const uint32_t *__restrict__ rank0,
                                                                   const uint32_t *__restrict__ extoff, uint4 *__restrict__ sblocks)
{
  const uint64_t g = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nroots = (uint64_t) nsym * nb;
  if (g >= nroots) return;
  const uint32_t s = (uint32_t)(g / nb), j = (uint32_t)(g - (uint64_t) s * nb);
  const uint32_t s0 = symstart[s];
  uint32_t t0, cnt;
  fm_sparse_bucket_range(rows, s0, symstart[s + 1], scale, j, t0, cnt);
  const FmSparseTree<LANES> t(cnt);
  fm_sparse_write_node<LANES>(t, t.D, 0u, rows + t0, cnt, (uint32_t) nroots + extoff[g], rank0[s] + (t0 - s0), sblocks + g * (2u * LANES));
}

/* pass 2b: the tree nodes below the roots; one thread per node.  The owning root is found by binary search in extoff. */
template <int LANES>
__global__ void __launch_bounds__(256) fm_sparse_fill_ext_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ symstart,
                                                                 uint32_t nsym, uint32_t nb, uint32_t scale, const uint32_t *__restrict__ rank0,
                                                                 const uint32_t *__restrict__ extoff, uint32_t total_ext, uint4 *__restrict__ sblocks)
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total_ext) return;
  const uint64_t nroots = (uint64_t) nsym * nb;
  uint64_t lo = 0, hi = nroots;                                /* last root with extoff <= e: the owner (roots without extension before it share its offset) */
  while (hi - lo > 1) { const uint64_t mid = lo + ((hi - lo) >> 1); if (extoff[mid] <= e) lo = mid; else hi = mid; }
  const uint64_t g = lo;
  const uint32_t s = (uint32_t)(g / nb), j = (uint32_t)(g - (uint64_t) s * nb);
  const uint32_t s0 = symstart[s];
  uint32_t t0, cnt;
  fm_sparse_bucket_range(rows, s0, symstart[s + 1], scale, j, t0, cnt);
  const FmSparseTree<LANES> t(cnt);
  uint32_t local = e - extoff[g], v = t.D;                     /* levels are stored top-down */
  while (v > 0) { v--; if (local < t.N[v]) break; local -= t.N[v]; }
  FM_BOUND(v, t.D, "sparse build: level of an extension node"); FM_BOUND(local, t.N[v], "sparse build: node index in its level");
  FM_BOUND(t.level_offset(v) + local, t.below_root(), "sparse build: node offset in its root's area");
  fm_sparse_write_node<LANES>(t, v, local, rows + t0, cnt, (uint32_t) nroots + extoff[g], rank0[s] + (t0 - s0),
                              sblocks + (nroots + e) * (2u * LANES));
}

#endif /* FM_SPARSE_CUH_ */

/*
 * fm_wide.cuh -- "wide-step" device layout and search kernels: up to 30 query bases per 64- or 128-byte block fetch.
 *
 * Why.  profiles/r02_ceiling_counters.md: the memory system accepts ~46 G miss-bound requests per second whatever
 * they carry up to one 128-byte line, and the search is bound by nothing else.  The sparse-step table (fm_sparse.cuh)
 * spends a request on 14 bases because it gives every wide symbol its own blocks: 4^KS symbols must stay below the
 * number of blocks the memory can hold.  Here the blocks belong to a PREFIX of the wide symbol and every entry carries
 * the rest of it, so the width of a step is bounded by the entry size (64 bits), not by the table:
 *
 *   wide symbol of row i    F(i) = s(i) | s(LF(i)) << 2k | ...   (hops = W/k hops, W <= 30 bases: the composition of
 *                           fm_fused.cuh / fm_sparse.cuh, so that rank_F(sigma, X) = G(sigma) + #{ i < X : F(i) = sigma }
 *                           IS `hops` consecutive reference LF steps for every X; rows whose chain meets a '$' row
 *                           carry no symbol).  The hop consumed last sits in the TOP bits: numeric order of F is the
 *                           lexicographic order of the W text bases in front of the suffix.
 *   bucket                  the top prefix_bits of sigma (all 15-mers for a 2 Gbp text: 1.9 rows per bucket).
 *                           Block of (sigma, X) = sigma >> sub_bits: computed from the READ alone, never looked up and
 *                           independent of X -- both interval ends of a step always share their fetch.
 *   entry (64 bits)         (sigma & submask) << row_bits | row, for every row of the bucket, ascending: that is the
 *                           order of (F(i), i), i.e. one stable radix sort of the composed keys.
 *   block (32 * LANES B)    4 * LANES x u64: word 0 = { value (low half), kind (high half) }, then SLOTS = 4 * LANES - 1
 *                           ascending entries padded with ~0 (LANES = 2: 64 bytes, 7 entries, the default; LANES = 4:
 *                           128 bytes, 15 entries).  Steps wider than 30 bases need 96-bit entries: five to a 64-byte block
 *                           behind a one-word header (EW = 5, see FmWideKey below), else two per lane (EW = 3):
 *        leaf               value = G(first symbol of the bucket) + entries of the bucket in front of this block;
 *                           rank_F(sigma, X) = value + #{ entries < (sub(sigma) << row_bits | X) }
 *        inner node         value = block number of child 0, entries = SLOTS separators (first entry of children 1..);
 *                           child = value + #{ separators < key } -- the same arithmetic; fan-out SLOTS + 1
 *        exceptional        kind = 2: the step runs as `hops` plain SB96 steps instead (see below)
 *   search tree             a bucket with more than SLOTS rows is the root of a tree over its sorted entries (leaves of
 *                           SLOTS consecutive entries, all at the same depth, stored behind the grid level by level), as
 *                           in the sparse-step table: the Poisson tail on a random text (0.3 % of the steps at 1.9 rows per
 *                           7-entry bucket, 8 % at 3.7), repeats on a real one.
 *
 * Exactness.  A leaf answers with ONE base value for all symbols of its bucket, which is right iff
 *   G(sigma') - G(sigma) = #{ rows with sigma <= F < sigma' }  for the symbols of one bucket.
 * That is the suffix-array order of the W-base contexts, except where one of the (at most W) suffixes shorter than W
 * sorts into the bucket.  Nothing is assumed: the builder carries y(i) = the row the chain of i ends in (= rank_F(F(i), i))
 * through the sort and checks for EVERY entry that y = G(bucket's smallest symbol) + position in the bucket, and for every
 * bucket that the next bucket's G continues the count (both composed from the SB96 table itself).  With G monotone in
 * sigma this pins G for the absent symbols of the bucket too.  A bucket that fails is marked exceptional and the kernel
 * takes `hops` SB96 steps there -- 28 buckets out of 2^30 on the benchmark text.  AltCounters files with an active padding
 * quirk (their composed rank is not a plain counting function) are refused; the sparse-step table serves them.
 *
 * Kernels.  fm_search_wide_kernel: the per-read state machine of fm_sparse.cuh -- one block fetch per iteration and
 * unfinished read -- on LANES-lane groups (one 256-bit load per lane), 64- or 96-bit compares; the timed kernel of bench.py
 * (LANES = 2, EW = 5: 46 bases per step, one read per lane pair, 32 registers; 100 bp = an 8-base lead table + 2 fetches).  Per read the kernel's instructions scale with the lanes it occupies:
 * the 128-byte form is issue-bound (72 % of the issue slots, profiles/r02w_*), the 64-byte form is not (43 %) and runs
 * at 0.96 of the request-rate ceiling (30 bases per step; 0.77 at 46, where the per-read work around two fetches weighs more).
 * fm_search_wide_dyn_kernel hands reads to the lane groups from a CTA queue (repeat-rich texts).  fm_search_wide_burst_kernel issues all block loads of a read at once (their
 * addresses do not depend on the interval); it holds fewer reads per SM and measured slower ($FMGPU_WIDE_BURST=1).
 */
#ifndef FM_WIDE_CUH_
#define FM_WIDE_CUH_

#include "fm_device.cuh"

#define FM_WD_PAD      0xFFFFFFFFFFFFFFFFull
#define FM_WD_LEAF     0u
#define FM_WD_INNER    1u
#define FM_WD_EXC      2u
#define FM_WD_DONE     0xFFFFFFFFu
#define FM_WD_MAXDEPTH 14                  /* 4 * 5^14 entries > 2^32 rows (64-byte blocks of 96-bit entries); 7 * 8^11, 15 * 16^8 for the others */
/* block = 32 * LANES bytes = 4 * LANES u64: a header + SLOTS = 4 * LANES - 1 entries; inner nodes have SLOTS + 1 children.
 * LANES = 4: 128-byte blocks (15 entries), LANES = 2: 64-byte blocks (7 entries, 64-byte L2 fills): half the lanes,
 * instructions and DRAM bytes per read for the same number of requests, but twice the buckets for the same occupancy. */

struct FmWideParams {
  const uint4    *wblocks;    /* grid + tree nodes, 2 * LANES uint4 (64 or 128 bytes) per block            */
  const uint4    *blocks;     /* SB96: the steps of exceptional buckets                                     */
  const uint32_t *packed;
  uint32_t       *results;
  uint32_t nblocks;           /* SB96 stride                                                               */
  uint32_t nq;
  uint32_t nsteps;            /* wide steps (after the lead table's bases, if one is used)                 */
  uint32_t wpq;
  uint32_t bwtsize;
  uint32_t wbits;             /* 2 * W                                                                     */
  uint32_t sub_bits;          /* bits of the wide symbol kept in the entries: wbits - prefix_bits          */
  uint32_t row_bits;          /* bits of a row number in an entry                                          */
  uint32_t hops, kbits;       /* W / k, 2 * k                                                              */
  uint32_t nroots;            /* 2^prefix_bits: blocks at or beyond it are tree nodes                      */
  uint32_t total_blocks;      /* grid + tree nodes (extent of wblocks, checked by the -DFM_DEBUG_BOUNDS build) */
  const uint2 *start;         /* lead table: (L,R) after the first start_bits / 2 bases, or NULL = (0, bwtsize) */
  uint32_t start_bits;
  unsigned long long *fetch_counters;  /* COUNT only: [0] grid blocks, [1] SB96 blocks (exceptional buckets), [2] tree nodes below the grid */
};

/* one 256-bit load per lane.  128-byte blocks use the whole line (default fill); 64-byte blocks ask for a 64-byte fill */
template <int LANES> __device__ __forceinline__ void fm_wide_load(const uint4 *p, uint32_t (&w)[8])
{
  if (LANES == 2) fm_ldg32(p, w);
  else asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

/* bits [pos, pos + nbits) of a packed read in shared memory, nbits <= 64 (one readable spare word behind the read) */
__device__ __forceinline__ uint64_t fm_read_field64(const uint32_t *q, uint32_t pos, uint32_t nbits)
{
  const uint32_t lo = fm_read_field(q, pos, 0xFFFFFFFFu);
  const uint32_t hi = nbits > 32u ? fm_read_field(q, pos + 32u, 0xFFFFFFFFu) : 0u;
  const uint64_t v = (uint64_t) lo | ((uint64_t) hi << 32);
  return nbits >= 64u ? v : (v & ((1ull << nbits) - 1ull));
}

/* Entry width.  EW = 2: 64-bit entries, steps up to 30 bases.  EW = 3: 96-bit entries { lo, mid, hi } -- steps up to 46 bases
 * (100 bp = 8 + 2 x 46: two fetches instead of three), two entries per lane behind a two-word lane header (the block header
 * in lane 0), i.e. 2 * LANES entries per block and a fan-out of 2 * LANES + 1. */
typedef unsigned __int128 fm_u128;
/* EW = 5: 96-bit entries PACKED five to a 64-byte block (LANES = 2 only): a one-word header { inner flag : 1, value : 31 } --
 * 0xFFFFFFFF = exceptional -- then the entries back to back, the third one straddling the two lanes (one shuffle).  4 % instead
 * of 12 % of the steps continue into a search tree on a random text.  Needs bwtsize and the block count below 2^31. */
template <int EW> struct FmWideKey { typedef uint64_t type; };
template <> struct FmWideKey<3> { typedef fm_u128 type; };
template <> struct FmWideKey<5> { typedef fm_u128 type; };
template <int LANES, int EW> struct FmWideSlots { static constexpr uint32_t value = EW == 5 ? 5u : EW == 3 ? 2u * LANES : 4u * LANES - 1u; };

/* bits [pos, pos + nbits) of a packed read in shared memory, nbits <= 96 */
__device__ __forceinline__ fm_u128 fm_read_field96(const uint32_t *q, uint32_t pos, uint32_t nbits)
{
  const uint32_t w0 = fm_read_field(q, pos, 0xFFFFFFFFu);
  const uint32_t w1 = nbits > 32u ? fm_read_field(q, pos + 32u, 0xFFFFFFFFu) : 0u;
  const uint32_t w2 = nbits > 64u ? fm_read_field(q, pos + 64u, 0xFFFFFFFFu) : 0u;
  const fm_u128 v = (fm_u128) w0 | ((fm_u128) w1 << 32) | ((fm_u128) w2 << 64);
  return v & ((((fm_u128) 1) << nbits) - 1);
}
template <int EW> __device__ __forceinline__ typename FmWideKey<EW>::type fm_wide_read_key(const uint32_t *q, uint32_t pos, uint32_t nbits)
{
  if constexpr (EW != 2) return fm_read_field96(q, pos, nbits);
  else return fm_read_field64(q, pos, nbits);
}

/* this lane's share of #{ entries < key }, 96-bit entries: words [2..4] and [5..7] of the lane */
__device__ __forceinline__ uint32_t fm_wide_partial3(const uint32_t (&w)[8], fm_u128 key)
{
  const uint32_t klo = (uint32_t) key;
  const uint64_t khi = (uint64_t)(key >> 32);
  uint32_t c = 0;
  #pragma unroll
  for (int j = 0; j < 2; j++) {
    const uint32_t elo = w[2 + 3 * j];
    const uint64_t ehi = (uint64_t) w[3 + 3 * j] | ((uint64_t) w[4 + 3 * j] << 32);
    c += (ehi < khi || (ehi == khi && elo < klo)) ? 1u : 0u;
  }
  return c;
}

__device__ __forceinline__ uint32_t fm_wide_lt96(uint32_t elo, uint32_t emid, uint32_t ehi, uint64_t khi, uint32_t klo)
{
  const uint64_t eh = (uint64_t) emid | ((uint64_t) ehi << 32);
  return (eh < khi || (eh == khi && elo < klo)) ? 1u : 0u;
}
/* packed form: lane 0 holds the header word, entries 1, 2 and the low word of entry 3 (handed over in xw); lane 1 the rest */
__device__ __forceinline__ uint32_t fm_wide_partial5(const uint32_t (&w)[8], fm_u128 key, uint32_t lg, uint32_t xw)
{
  const uint32_t klo = (uint32_t) key;
  const uint64_t khi = (uint64_t)(key >> 32);
  const bool hi = lg != 0u;
  uint32_t c = fm_wide_lt96(hi ? xw : w[1], hi ? w[0] : w[2], hi ? w[1] : w[3], khi, klo);
  c += fm_wide_lt96(hi ? w[2] : w[4], hi ? w[3] : w[5], hi ? w[4] : w[6], khi, klo);
  c += hi ? fm_wide_lt96(w[5], w[6], w[7], khi, klo) : 0u;
  return c;
}

/* this lane's share of #{ entries < key }: lane 0 skips word 0 (the header) */
__device__ __forceinline__ uint32_t fm_wide_partial(const uint32_t (&w)[8], uint64_t key, uint32_t lg)
{
  uint32_t c = 0;
  #pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint64_t e = (uint64_t) w[2 * j] | ((uint64_t) w[2 * j + 1] << 32);
    if (j == 0) c += (lg != 0u && e < key) ? 1u : 0u;
    else        c += (e < key) ? 1u : 0u;
  }
  return c;
}

/* packed reads of a CTA -> shared memory: one TMA bulk copy when the piece is 16-byte granular (fm_search_fused_kernel) */
template <int THREADS>
__device__ __forceinline__ void fm_stage_reads(uint32_t *fsm, uint32_t *sq, const uint32_t *src, uint32_t words)
{
  const uint32_t bytes = words * 4u;
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
    for (uint32_t i = threadIdx.x; i < words; i += THREADS) sq[i] = __ldg(src + i);
    __syncthreads();
  }
}

template <int EW> __device__ __forceinline__ uint32_t fm_wide_count(const uint32_t (&w)[8], typename FmWideKey<EW>::type key, uint32_t lg, uint32_t xw)
{
  if constexpr (EW == 5) return fm_wide_partial5(w, key, lg, xw);
  else if constexpr (EW == 3) return fm_wide_partial3(w, key);
  else return fm_wide_partial(w, key, lg);
}
/* header of a block from lane 0's first two words */
template <int EW> __device__ __forceinline__ void fm_wide_header(uint32_t w0, uint32_t w1, uint32_t &value, uint32_t &kind)
{
  if constexpr (EW == 5) { value = w0 & 0x7FFFFFFFu; kind = w0 == 0xFFFFFFFFu ? FM_WD_EXC : (w0 >> 31); }
  else { value = w0; kind = w1; }
}

/* one wide step of an exceptional bucket: `hops` base-k steps on SB96 for both interval ends */
template <typename KeyT>
__device__ __forceinline__ void fm_wide_plain_step(const FmWideParams &p, KeyT key, uint32_t &L, uint32_t &R)
{
  const uint32_t kmask = (1u << p.kbits) - 1u;
  for (uint32_t h = 0; h < p.hops; h++) {
    const uint32_t s = (uint32_t)(key >> (p.kbits * h)) & kmask;
    const uint32_t bL = fm_div96(L), bR = fm_div96(R);
    const uint4 *base = p.blocks + (size_t) s * p.nblocks;
    FM_BOUND(bL, p.nblocks, "wide: SB96 block (L)"); FM_BOUND(bR, p.nblocks, "wide: SB96 block (R)");
    const uint4 vL = fm_ldg16(base + bL);
    const uint4 vR = (bL == bR) ? vL : fm_ldg16(base + bR);
    L = fm_block_rank(vL, L - bL * FM_SB_ROWS);
    R = fm_block_rank(vR, R - bR * FM_SB_ROWS);
  }
}

template <int LANES, int EW, int QPT, int THREADS, int MINB, bool COUNT>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_wide_kernel(const FmWideParams p)
{
  extern __shared__ __align__(16) uint32_t fsm[];             /* [0..3]: mbarrier + pad; [4..): packed reads, natural stride */
  uint32_t *sq = fsm + 4;
  typedef typename FmWideKey<EW>::type KeyT;
  constexpr int GROUPS = THREADS / LANES;
  const uint32_t q0 = blockIdx.x * (GROUPS * QPT);
  const uint32_t nqb = min((uint32_t)(GROUPS * QPT), p.nq - q0);
  const uint32_t lg = threadIdx.x % LANES, group = threadIdx.x / LANES;
  const KeyT submask = p.sub_bits >= 8u * sizeof(KeyT) ? ~(KeyT) 0 : ((((KeyT) 1) << p.sub_bits) - 1);

  fm_stage_reads<THREADS>(fsm, sq, p.packed + (size_t) q0 * p.wpq, nqb * p.wpq);

  uint32_t L[QPT], R[QPT], aL[QPT], aR[QPT], rem[QPT];
  KeyT key[QPT];
  const uint32_t *myq[QPT];
  bool live[QPT];
  const uint32_t kmask = (p.start_bits >= 32u) ? 0xFFFFFFFFu : ((1u << p.start_bits) - 1u);
  bool busy = false;
  #pragma unroll
  for (int i = 0; i < QPT; i++) {
    const uint32_t lq = i * GROUPS + group;
    live[i] = lq < nqb;
    myq[i] = sq + (live[i] ? lq : 0u) * p.wpq;
    L[i] = 0u; R[i] = p.bwtsize;
    if (p.start) {
      const uint2 lr = __ldg(p.start + (myq[i][0] & kmask));
      L[i] = lr.x; R[i] = lr.y;
    }
    rem[i] = live[i] ? p.nsteps : 0u;
    aL[i] = aR[i] = FM_WD_DONE; key[i] = 0;
    if (rem[i]) {
      key[i] = fm_wide_read_key<EW>(myq[i], p.start_bits, p.wbits);
      aL[i] = aR[i] = (uint32_t)(key[i] >> p.sub_bits);
    }
    busy |= rem[i] != 0u;
  }

  unsigned long long n_root = 0, n_sb = 0, n_tree = 0;
  /* one state machine per read: aL / aR = block to fetch next for that interval end, FM_WD_DONE once the end holds its
   * value for the next step.  Every iteration fetches one block per unfinished read: the node both ends share (always,
   * at the grid level), else L's node, else R's. */
  while (__any_sync(0xFFFFFFFFu, busy)) {
    uint32_t w[QPT][8];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      if (rem[i]) {
        const uint32_t a = (aL[i] != FM_WD_DONE) ? aL[i] : aR[i];
        FM_BOUND(a, p.total_blocks, "wide: grid / tree block");
        fm_wide_load<LANES>(p.wblocks + (size_t) a * (2u * LANES) + 2u * lg, w[i]);
        if (COUNT && lg == 0) { if (a < p.nroots) n_root++; else n_tree++; }
      }
    }
    busy = false;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      /* the lanes of a group hold the same state, so they take the same branches */
      const bool act = rem[i] != 0u;
      const bool doL = act && aL[i] != FM_WD_DONE;
      const bool doR = act && (!doL || aR[i] == aL[i]);                  /* R alone, or riding on the node it shares with L */
      const KeyT ksub = (key[i] & submask) << p.row_bits;
      uint32_t cL = 0, cR = 0, xw = 0;
      if constexpr (EW == 5) xw = __shfl_xor_sync(0xFFFFFFFFu, w[i][7], 1);
      if (act) {
        cL = fm_wide_count<EW>(w[i], ksub | L[i], lg, xw);
        cR = fm_wide_count<EW>(w[i], ksub | R[i], lg, xw);
      }
      uint32_t hval, hkind;
      fm_wide_header<EW>(__shfl_sync(0xFFFFFFFFu, w[i][0], 0, LANES), EW == 5 ? 0u : __shfl_sync(0xFFFFFFFFu, w[i][1], 0, LANES), hval, hkind);
      const uint32_t vL = hval + fm_group_sum<LANES>(cL), vR = hval + fm_group_sum<LANES>(cR);
      if (act && hkind == FM_WD_EXC) {                                   /* (only ever at a grid block: both ends are here) */
        fm_wide_plain_step<KeyT>(p, key[i], L[i], R[i]);
        if (COUNT && lg == 0) n_sb += 2ull * p.hops;
        aL[i] = aR[i] = FM_WD_DONE;
      } else {
        const bool is_inner = hkind == FM_WD_INNER;
        if (doL) { if (is_inner) aL[i] = vL; else { L[i] = vL; aL[i] = FM_WD_DONE; } }
        if (doR) { if (is_inner) aR[i] = vR; else { R[i] = vR; aR[i] = FM_WD_DONE; } }
      }
      if (act && aL[i] == FM_WD_DONE && aR[i] == FM_WD_DONE) {           /* step complete: next step's block */
        rem[i] -= 1u;
        if (rem[i]) {
          key[i] = fm_wide_read_key<EW>(myq[i], p.start_bits + (p.nsteps - rem[i]) * p.wbits, p.wbits);
          aL[i] = aR[i] = (uint32_t)(key[i] >> p.sub_bits);
        }
      }
      busy |= rem[i] != 0u;
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
 * The same search with DYNAMIC read assignment (fm_search_sparse_dyn_kernel's scheme): the CTA stages `reads_per_cta` reads
 * (several rounds of its slots) with one bulk copy and a lane group that finishes a read pulls the next one from a
 * shared-memory counter.  The lead-table lookup of a new read is issued TOGETHER with the read's first block fetch -- both
 * addresses are functions of the read alone -- so a read costs exactly its block fetches in iterations, as in the static
 * kernel.  With static assignment a warp lasts as long as the slowest of its reads:
 * that is nothing on the roomy 64-bit grid (0.3 % of the steps meet a tree), but with 96-bit entries (four per block,
 * 12 % of the steps continue into a tree) and on repeat-rich texts half of the fetch slots would idle.
 * ------------------------------------------------------------------------ */
template <int LANES, int EW, int QPT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_wide_dyn_kernel(const FmWideParams p, uint32_t reads_per_cta)
{
  extern __shared__ __align__(16) uint32_t fsm[];             /* [0..1]: mbarrier; [2]: next read; [4..): packed reads */
  uint32_t *sq = fsm + 4;
  typedef typename FmWideKey<EW>::type KeyT;
  const KeyT submask = p.sub_bits >= 8u * sizeof(KeyT) ? ~(KeyT) 0 : ((((KeyT) 1) << p.sub_bits) - 1);
  const uint32_t kmask = (p.start_bits >= 32u) ? 0xFFFFFFFFu : ((1u << p.start_bits) - 1u);
  const uint32_t q0 = blockIdx.x * reads_per_cta;
  const uint32_t nqb = min(reads_per_cta, p.nq - q0);
  const uint32_t lg = threadIdx.x % LANES;
  if (threadIdx.x == 0) fsm[2] = 0u;
  fm_stage_reads<THREADS>(fsm, sq, p.packed + (size_t) q0 * p.wpq, nqb * p.wpq);   /* (its barrier also publishes fsm[2]) */

  uint32_t L[QPT], R[QPT], aL[QPT], aR[QPT], rem[QPT], rd[QPT];
  KeyT key[QPT];
  bool fresh[QPT];                                             /* the read was just taken: its (L,R) still come from the lead table */
  /* takes the next read of the CTA for slot i (or parks the slot): all lanes of the warp call it together */
  auto take = [&](int i, bool need) {
    uint32_t r = 0;
    if (need && lg == 0) r = atomicAdd(&fsm[2], 1u);
    r = __shfl_sync(0xFFFFFFFFu, r, 0, LANES);
    if (need) {
      rd[i] = r;
      if (r < nqb) {
        rem[i] = p.nsteps; L[i] = 0u; R[i] = p.bwtsize;
        fresh[i] = p.start != NULL;
        if (rem[i]) {
          key[i] = fm_wide_read_key<EW>(sq + r * p.wpq, p.start_bits, p.wbits);
          aL[i] = aR[i] = (uint32_t)(key[i] >> p.sub_bits);
        } else { aL[i] = aR[i] = FM_WD_DONE; }
      } else { rem[i] = 0u; aL[i] = aR[i] = FM_WD_DONE; rd[i] = 0xFFFFFFFFu; fresh[i] = false; }
    }
  };
  #pragma unroll
  for (int i = 0; i < QPT; i++) { rd[i] = 0xFFFFFFFFu; rem[i] = 0u; aL[i] = aR[i] = FM_WD_DONE; L[i] = R[i] = 0u; key[i] = 0; fresh[i] = false; take(i, true); }

  bool busy = false;
  #pragma unroll
  for (int i = 0; i < QPT; i++) busy |= rd[i] != 0xFFFFFFFFu;
  while (__any_sync(0xFFFFFFFFu, busy)) {
    uint32_t w[QPT][8];
    uint2 lr[QPT];
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      if (rd[i] != 0xFFFFFFFFu) {
        if (fresh[i]) lr[i] = __ldg(p.start + (sq[rd[i] * p.wpq] & kmask));   /* in flight together with the first block */
        if (rem[i]) {
          const uint32_t a = (aL[i] != FM_WD_DONE) ? aL[i] : aR[i];
          FM_BOUND(a, p.total_blocks, "wide (dynamic): grid / tree block"); FM_BOUND(rd[i], nqb, "wide (dynamic): read slot");
          fm_wide_load<LANES>(p.wblocks + (size_t) a * (2u * LANES) + 2u * lg, w[i]);
        }
      }
    }
    busy = false;
    #pragma unroll
    for (int i = 0; i < QPT; i++) {
      const bool have = rd[i] != 0xFFFFFFFFu;
      if (have && fresh[i]) { L[i] = lr[i].x; R[i] = lr[i].y; fresh[i] = false; }
      const bool act = have && rem[i] != 0u;
      const bool doL = act && aL[i] != FM_WD_DONE;
      const bool doR = act && (!doL || aR[i] == aL[i]);
      const KeyT ksub = (key[i] & submask) << p.row_bits;
      uint32_t cL = 0, cR = 0, xw = 0;
      if constexpr (EW == 5) xw = __shfl_xor_sync(0xFFFFFFFFu, w[i][7], 1);
      if (act) {
        cL = fm_wide_count<EW>(w[i], ksub | L[i], lg, xw);
        cR = fm_wide_count<EW>(w[i], ksub | R[i], lg, xw);
      }
      uint32_t hval, hkind;
      fm_wide_header<EW>(__shfl_sync(0xFFFFFFFFu, w[i][0], 0, LANES), EW == 5 ? 0u : __shfl_sync(0xFFFFFFFFu, w[i][1], 0, LANES), hval, hkind);
      const uint32_t vL = hval + fm_group_sum<LANES>(cL), vR = hval + fm_group_sum<LANES>(cR);
      if (act && hkind == FM_WD_EXC) {
        fm_wide_plain_step<KeyT>(p, key[i], L[i], R[i]);
        aL[i] = aR[i] = FM_WD_DONE;
      } else {
        const bool is_inner = hkind == FM_WD_INNER;
        if (doL) { if (is_inner) aL[i] = vL; else { L[i] = vL; aL[i] = FM_WD_DONE; } }
        if (doR) { if (is_inner) aR[i] = vR; else { R[i] = vR; aR[i] = FM_WD_DONE; } }
      }
      bool next_block = false;
      if (act && aL[i] == FM_WD_DONE && aR[i] == FM_WD_DONE) { rem[i] -= 1u; next_block = rem[i] != 0u; }
      if (next_block) {
        key[i] = fm_wide_read_key<EW>(sq + rd[i] * p.wpq, p.start_bits + (p.nsteps - rem[i]) * p.wbits, p.wbits);
        aL[i] = aR[i] = (uint32_t)(key[i] >> p.sub_bits);
      }
      const bool finished = have && rem[i] == 0u && aL[i] == FM_WD_DONE;
      if (finished && lg == 0) reinterpret_cast<uint2 *>(p.results)[q0 + rd[i]] = make_uint2(L[i], R[i]);
      take(i, finished);
      busy |= rd[i] != 0xFFFFFFFFu;
    }
  }
}

/* ------------------------------------------------------------------------ *
 * Burst kernel.  The block of a wide step is a function of the READ alone (the top bits of that step's 2W-bit field),
 * not of the interval: the S grid blocks a read needs are known before its search starts.  So a lane group issues its
 * lead-table lookup and up to PF block loads of each of its QPT reads back to back -- S independent requests in flight
 * per read instead of a chain of S dependent ones -- and then evaluates the steps in order from registers.  Only a step
 * whose grid block is the root of a search tree (or exceptional) continues with dependent fetches (fm_wide_slow_step),
 * 1 % of the steps on a random text.  Reads with more than PF steps run in chunks of PF.
 * ------------------------------------------------------------------------ */
template <int LANES> __device__ __forceinline__ uint32_t fm_wide_group_sum(uint32_t v, uint32_t gmask)
{
  #pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

/* a step whose grid block (w, already loaded) is not a leaf: exceptional -> plain steps; inner -> walk the tree, the two
 * interval ends apart once they leave a node.  Called by the lanes of ONE group (others may be elsewhere): shuffles are
 * confined to the group's mask. */
template <int LANES, bool COUNT>
__device__ __forceinline__ void fm_wide_slow_step(const FmWideParams &p, const uint32_t (&w)[8], uint32_t hval, uint32_t hkind, uint64_t key,
                                                  uint64_t ksub, uint32_t lg, uint32_t gmask, uint32_t &L, uint32_t &R,
                                                  unsigned long long &n_sb, unsigned long long &n_tree)
{
  if (hkind == FM_WD_EXC) {
    fm_wide_plain_step<uint64_t>(p, key, L, R);
    if (COUNT && lg == 0) n_sb += 2ull * p.hops;
    return;
  }
  uint32_t aL = hval + fm_wide_group_sum<LANES>(fm_wide_partial(w, ksub | L, lg), gmask);
  uint32_t aR = hval + fm_wide_group_sum<LANES>(fm_wide_partial(w, ksub | R, lg), gmask);
  while (aL != FM_WD_DONE || aR != FM_WD_DONE) {
    const uint32_t a = (aL != FM_WD_DONE) ? aL : aR;
    uint32_t t[8];
    FM_BOUND(a, p.total_blocks, "wide (burst): tree block");
    fm_wide_load<LANES>(p.wblocks + (size_t) a * (2u * LANES) + 2u * lg, t);
    if (COUNT && lg == 0) n_tree++;
    const bool doL = aL != FM_WD_DONE;
    const bool doR = !doL || aR == aL;
    const uint32_t hv = __shfl_sync(gmask, t[0], 0, LANES), hk = __shfl_sync(gmask, t[1], 0, LANES);
    const uint32_t vL = hv + fm_wide_group_sum<LANES>(fm_wide_partial(t, ksub | L, lg), gmask);
    const uint32_t vR = hv + fm_wide_group_sum<LANES>(fm_wide_partial(t, ksub | R, lg), gmask);
    const bool is_inner = hk == FM_WD_INNER;
    if (doL) { if (is_inner) aL = vL; else { L = vL; aL = FM_WD_DONE; } }
    if (doR) { if (is_inner) aR = vR; else { R = vR; aR = FM_WD_DONE; } }
  }
}

template <int LANES, int QPT, int PF, int THREADS, int MINB, bool COUNT>
__global__ void __launch_bounds__(THREADS, MINB) fm_search_wide_burst_kernel(const FmWideParams p)
{
  extern __shared__ __align__(16) uint32_t fsm[];             /* [0..3]: mbarrier + pad; [4..): packed reads, natural stride */
  uint32_t *sq = fsm + 4;
  constexpr int GROUPS = THREADS / LANES;
  const uint32_t q0 = blockIdx.x * (GROUPS * QPT);
  const uint32_t nqb = min((uint32_t)(GROUPS * QPT), p.nq - q0);
  const uint32_t lg = threadIdx.x % LANES, group = threadIdx.x / LANES;
  const uint32_t gmask = ((1u << LANES) - 1u) << ((threadIdx.x & 31u) & ~(uint32_t)(LANES - 1));
  const uint64_t submask = p.sub_bits >= 64u ? ~0ull : ((1ull << p.sub_bits) - 1ull);

  fm_stage_reads<THREADS>(fsm, sq, p.packed + (size_t) q0 * p.wpq, nqb * p.wpq);

  uint32_t L[QPT], R[QPT];
  const uint32_t *myq[QPT];
  bool live[QPT];
  const uint32_t kmask = (p.start_bits >= 32u) ? 0xFFFFFFFFu : ((1u << p.start_bits) - 1u);
  #pragma unroll
  for (int i = 0; i < QPT; i++) {
    const uint32_t lq = i * GROUPS + group;
    live[i] = lq < nqb;
    myq[i] = sq + (live[i] ? lq : 0u) * p.wpq;                 /* a dead slot repeats the CTA's first read (valid addresses, no store) */
    uint2 lr = make_uint2(0u, p.bwtsize);
    if (p.start) lr = __ldg(p.start + (myq[i][0] & kmask));    /* (in flight together with the block loads below: first used by the first compare) */
    L[i] = lr.x; R[i] = lr.y;
  }
  unsigned long long n_root = 0, n_sb = 0, n_tree = 0;
  for (uint32_t base = 0; base < p.nsteps; base += PF) {
    uint32_t w[QPT][PF][8];
    #pragma unroll
    for (int s = 0; s < PF; s++) {
      if (base + s < p.nsteps) {
        #pragma unroll
        for (int i = 0; i < QPT; i++) {
          const uint32_t a = (uint32_t)(fm_read_field64(myq[i], p.start_bits + (base + s) * p.wbits, p.wbits) >> p.sub_bits);
          FM_BOUND(a, p.nroots, "wide (burst): grid block");
          fm_wide_load<LANES>(p.wblocks + (size_t) a * (2u * LANES) + 2u * lg, w[i][s]);
        }
      }
    }
    #pragma unroll
    for (int s = 0; s < PF; s++) {
      if (base + s < p.nsteps) {                                /* uniform over the grid: all lanes of a warp take it together */
        #pragma unroll
        for (int i = 0; i < QPT; i++) {
          const uint64_t key = fm_read_field64(myq[i], p.start_bits + (base + s) * p.wbits, p.wbits);
          const uint64_t ksub = (key & submask) << p.row_bits;
          const uint32_t cL = fm_wide_partial(w[i][s], ksub | L[i], lg), cR = fm_wide_partial(w[i][s], ksub | R[i], lg);
          const uint32_t hval  = __shfl_sync(0xFFFFFFFFu, w[i][s][0], 0, LANES);
          const uint32_t hkind = __shfl_sync(0xFFFFFFFFu, w[i][s][1], 0, LANES);
          const uint32_t sL = fm_group_sum<LANES>(cL), sR = fm_group_sum<LANES>(cR);
          if (COUNT && lg == 0 && live[i]) n_root++;
          if (hkind == FM_WD_LEAF) { L[i] = hval + sL; R[i] = hval + sR; }
          else fm_wide_slow_step<LANES, COUNT>(p, w[i][s], hval, hkind, key, ksub, lg, gmask, L[i], R[i], n_sb, n_tree);
        }
      }
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
 * Construction from SB96 (all on the device).  KeyT = uint64_t (steps up to 30 bases, 64-bit entries) or fm_u128 (up to 46
 * bases, 96-bit entries).
 * ------------------------------------------------------------------------ */

/* key[i] = wide symbol of row i (none_key for a row whose chain meets a '$' row: sorts behind every symbol),
 * val[i] = i | y(i) << 32 with y(i) = the row the chain ends in = rank_F(F(i), i) */
template <typename KeyT>
__global__ void fm_wide_compose_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, const uint8_t *__restrict__ sym,
                                       uint32_t bwtsize, uint32_t kbits, uint32_t hops, uint32_t wbits,
                                       KeyT *__restrict__ keys, uint64_t *__restrict__ vals)
{
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bwtsize) return;
  uint32_t row = (uint32_t) i;
  KeyT acc = 0;
  bool ok = true;
  for (uint32_t h = 0; h < hops; h++) {
    const uint32_t s = sym[row];
    if (s == FM_SYM_NONE) { ok = false; break; }
    acc |= (KeyT) s << (kbits * h);
    row = fm_sb96_rank(blocks, nblocks, s, row);
  }
  keys[i] = ok ? acc : (((KeyT) 1) << wbits);
  vals[i] = i | ((uint64_t) row << 32);
}

/* bstart[b] = first position of the sorted keys whose bucket is >= b, b = 0 .. nroots (bstart[nroots] = rows carrying a symbol) */
template <typename KeyT>
__global__ void fm_wide_bstart_kernel(const KeyT *__restrict__ keys, uint64_t n, uint32_t sub_bits, uint32_t nroots, uint32_t *__restrict__ bstart)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nroots) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if ((uint64_t)(keys[mid] >> sub_bits) < (uint64_t) b) lo = mid + 1; else hi = mid; }
  bstart[b] = (uint32_t) lo;
}

/* g0[b] = G(smallest symbol of bucket b): the composed rank at X = 0 */
template <typename KeyT>
__global__ void fm_wide_g0_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t kbits, uint32_t hops, uint32_t sub_bits,
                                  uint32_t nroots, uint32_t *__restrict__ g0)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nroots) return;
  const KeyT sigma = (KeyT) b << sub_bits;
  uint32_t x = 0;
  for (uint32_t h = 0; h < hops; h++) x = fm_sb96_rank(blocks, nblocks, (uint32_t)(sigma >> (kbits * h)) & ((1u << kbits) - 1u), x);
  g0[b] = x;
}

/* every entry must sit where the composed LF walk says: y == g0[bucket] + position in the bucket; else the bucket is exceptional */
template <typename KeyT>
__global__ void fm_wide_verify_entries_kernel(const KeyT *__restrict__ keys, const uint64_t *__restrict__ vals, uint64_t nvalid,
                                              uint32_t sub_bits, const uint32_t *__restrict__ bstart, const uint32_t *__restrict__ g0,
                                              uint32_t *__restrict__ exc)
{
  const uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nvalid) return;
  const uint32_t b = (uint32_t)(keys[j] >> sub_bits);
  const uint32_t expect = g0[b] + ((uint32_t) j - bstart[b]);
  if ((uint32_t)(vals[j] >> 32) != expect) atomicOr(exc + (b >> 5), 1u << (b & 31u));
}

/* the next bucket's G must continue the count (a suffix shorter than W sorting into or behind the bucket breaks it); behind
 * the last bucket the count must have reached every row of the BWT */
__global__ void fm_wide_verify_buckets_kernel(const uint32_t *__restrict__ bstart, const uint32_t *__restrict__ g0, uint32_t nroots,
                                              uint32_t bwtsize, uint32_t force_every, uint32_t *__restrict__ exc)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nroots) return;
  bool bad = g0[b] + (bstart[b + 1] - bstart[b]) != (b + 1u < nroots ? g0[b + 1] : bwtsize);
  if (force_every && (b % force_every) == force_every - 1u) bad = true;   /* tests: exercises the exceptional path */
  if (bad) atomicOr(exc + (b >> 5), 1u << (b & 31u));
}

/* shape of the tree over cnt > SLOTS entries: N[v] = nodes of level v (0 = leaves), depth D with N[D] = 1 (the root, which
 * lives in the grid); nodes below the root = sum of N[0 .. D-1] */
template <int LANES, int EW> struct FmWideTree {
  static constexpr uint32_t SLOTS = FmWideSlots<LANES, EW>::value, FAN = SLOTS + 1u;
  uint32_t N[FM_WD_MAXDEPTH + 1];
  uint32_t D;
  __host__ __device__ explicit FmWideTree(uint32_t cnt)
  {
    N[0] = (cnt + SLOTS - 1) / SLOTS;
    D = 0;
    while (N[D] > 1 && D < FM_WD_MAXDEPTH) { N[D + 1] = (N[D] + FAN - 1) / FAN; D++; }
  }
  __host__ __device__ uint32_t below_root() const { uint32_t t = 0; for (uint32_t v = 0; v < D; v++) t += N[v]; return t; }
  /* offset of level v's first node inside the root's extension area (levels stored top-down: D-1 first, leaves last) */
  __host__ __device__ uint32_t level_offset(uint32_t v) const { uint32_t t = 0; for (uint32_t u = v + 1; u < D; u++) t += N[u]; return t; }
};

template <typename KeyT> struct FmWideBuild {
  const KeyT *keys;               /* sorted */
  const uint64_t *vals;
  const uint32_t *bstart, *g0, *exc, *extoff;
  uint32_t nroots, sub_bits, row_bits;
};

template <typename KeyT> __device__ __forceinline__ bool fm_wide_is_exc(const FmWideBuild<KeyT> &x, uint32_t b) { return (x.exc[b >> 5] >> (b & 31u)) & 1u; }
template <typename KeyT> __device__ __forceinline__ KeyT fm_wide_entry(const FmWideBuild<KeyT> &x, uint64_t j)
{
  const KeyT submask = x.sub_bits >= 8u * sizeof(KeyT) ? ~(KeyT) 0 : ((((KeyT) 1) << x.sub_bits) - 1);
  return ((x.keys[j] & submask) << x.row_bits) | (KeyT)(x.vals[j] & 0xFFFFFFFFull);
}

/* pass 1: extension nodes every bucket needs (0 for a bucket that fits its block or is exceptional) */
template <int LANES, int EW>
__global__ void __launch_bounds__(256) fm_wide_count_kernel(const FmWideBuild<typename FmWideKey<EW>::type> x, uint32_t *__restrict__ ext,
                                                            unsigned long long *__restrict__ stats)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= x.nroots) return;
  const uint32_t cnt = x.bstart[b + 1] - x.bstart[b];
  uint32_t e = 0;
  if (fm_wide_is_exc(x, b)) atomicAdd(stats + 3, 1ull);
  else if (cnt > FmWideTree<LANES, EW>::SLOTS) {
    const FmWideTree<LANES, EW> t(cnt);
    e = t.below_root();
    atomicAdd(stats, 1ull);                                    /* overfull buckets */
    atomicAdd(stats + 1, (unsigned long long) cnt);            /* rows living in them */
    atomicMax(stats + 2, (unsigned long long) t.D);            /* deepest tree */
  }
  ext[b] = e;
}

/* slot j (1 .. SLOTS) of a node image: 64-bit entries follow the header word pair; 96-bit entries sit two per lane behind
 * each lane's two-word header */
template <int LANES, int EW, typename KeyT>
__device__ __forceinline__ void fm_wide_put(uint32_t (&w)[8 * LANES], uint32_t j, KeyT e, bool pad)
{
  if constexpr (EW == 5) {
    const uint32_t at = 1u + 3u * (j - 1u);
    w[at] = pad ? 0xFFFFFFFFu : (uint32_t) e;
    w[at + 1] = pad ? 0xFFFFFFFFu : (uint32_t)(e >> 32);
    w[at + 2] = pad ? 0xFFFFFFFFu : (uint32_t)((fm_u128) e >> 64);
  } else if constexpr (EW == 3) {
    const uint32_t at = 8u * ((j - 1u) >> 1) + 2u + 3u * ((j - 1u) & 1u);
    w[at] = pad ? 0xFFFFFFFFu : (uint32_t) e;
    w[at + 1] = pad ? 0xFFFFFFFFu : (uint32_t)(e >> 32);
    w[at + 2] = pad ? 0xFFFFFFFFu : (uint32_t)((fm_u128) e >> 64);
  } else {
    w[2 * j] = pad ? 0xFFFFFFFFu : (uint32_t) e;
    w[2 * j + 1] = pad ? 0xFFFFFFFFu : (uint32_t)(e >> 32);
  }
}

/* one node: level v, index m, of the tree over entries [j0, j0 + cnt) of a bucket; `area` = first block of the bucket's
 * extension area, base = G before the bucket's first entry */
template <int LANES, int EW>
__device__ __forceinline__ void fm_wide_write_node(const FmWideBuild<typename FmWideKey<EW>::type> &x, const FmWideTree<LANES, EW> &t, uint32_t v, uint32_t m,
                                                   uint64_t j0, uint32_t cnt, uint32_t area, uint32_t base, uint4 *__restrict__ dst)
{
  typedef typename FmWideKey<EW>::type KeyT;
  constexpr uint32_t SLOTS = FmWideTree<LANES, EW>::SLOTS, FAN = FmWideTree<LANES, EW>::FAN;
  uint32_t w[8 * LANES];
  #pragma unroll
  for (uint32_t c = 0; c < 8u * LANES; c++) w[c] = 0xFFFFFFFFu;
  if (v == 0) {
    const uint64_t first = (uint64_t) m * SLOTS;
    if (EW == 5) w[0] = base + (uint32_t) first; else { w[0] = base + (uint32_t) first; w[1] = FM_WD_LEAF; }
    #pragma unroll
    for (uint32_t c = 1; c <= SLOTS; c++) {
      const bool have = first + c - 1 < cnt;
      fm_wide_put<LANES, EW, KeyT>(w, c, have ? fm_wide_entry(x, j0 + first + c - 1) : (KeyT) 0, !have);
    }
  } else {
    uint64_t span = SLOTS;                                     /* entries under one child: SLOTS * FAN^(v-1) */
    for (uint32_t u = 1; u < v; u++) span *= FAN;
    if (EW == 5) w[0] = (area + t.level_offset(v - 1) + m * FAN) | 0x80000000u; else { w[0] = area + t.level_offset(v - 1) + m * FAN; w[1] = FM_WD_INNER; }
    #pragma unroll
    for (uint32_t c = 1; c <= SLOTS; c++) {
      const uint64_t child = (uint64_t) m * FAN + c, at = child * span;
      const bool have = child < t.N[v - 1] && at < cnt;
      fm_wide_put<LANES, EW, KeyT>(w, c, have ? fm_wide_entry(x, j0 + at) : (KeyT) 0, !have);
    }
  }
  #pragma unroll
  for (uint32_t c = 0; c < 2u * LANES; c++) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

/* pass 2a: the grid -- a leaf, the root of a tree, or an exceptional marker; one thread per bucket */
template <int LANES, int EW>
__global__ void __launch_bounds__(256) fm_wide_fill_roots_kernel(const FmWideBuild<typename FmWideKey<EW>::type> x, uint4 *__restrict__ wblocks)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= x.nroots) return;
  uint4 *dst = wblocks + (size_t) b * (2u * LANES);
  if (fm_wide_is_exc(x, b)) {
    dst[0] = EW == 5 ? make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu) : make_uint4(0u, FM_WD_EXC, 0xFFFFFFFFu, 0xFFFFFFFFu);
    #pragma unroll
    for (uint32_t c = 1; c < 2u * LANES; c++) dst[c] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    return;
  }
  const uint32_t j0 = x.bstart[b], cnt = x.bstart[b + 1] - j0;
  const FmWideTree<LANES, EW> t(cnt);
  fm_wide_write_node<LANES, EW>(x, t, t.D, 0u, j0, cnt, x.nroots + x.extoff[b], x.g0[b], dst);
}

/* pass 2b: the tree nodes below the grid; one thread per node.  The owning bucket is found by binary search in extoff. */
template <int LANES, int EW>
__global__ void __launch_bounds__(256) fm_wide_fill_ext_kernel(const FmWideBuild<typename FmWideKey<EW>::type> x, uint32_t total_ext, uint4 *__restrict__ wblocks)
{
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total_ext) return;
  uint32_t lo = 0, hi = x.nroots;                              /* last bucket with extoff <= e: the owner (buckets without extension before it share its offset) */
  while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (x.extoff[mid] <= e) lo = mid; else hi = mid; }
  const uint32_t b = lo;
  const uint32_t j0 = x.bstart[b], cnt = x.bstart[b + 1] - j0;
  const FmWideTree<LANES, EW> t(cnt);
  uint32_t local = e - x.extoff[b], v = t.D;                   /* levels are stored top-down */
  while (v > 0) { v--; if (local < t.N[v]) break; local -= t.N[v]; }
  FM_BOUND(v, t.D, "wide build: level of an extension node"); FM_BOUND(local, t.N[v], "wide build: node index in its level");
  fm_wide_write_node<LANES, EW>(x, t, v, local, j0, cnt, x.nroots + x.extoff[b], x.g0[b], wblocks + ((size_t) x.nroots + e) * (2u * LANES));
}

#endif /* FM_WIDE_CUH_ */

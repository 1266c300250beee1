/*
 * fm_device.cuh -- device helpers shared by every kernel family of the search path (inline only: this header is
 * included by several translation units).
 *
 * Device layout "SB96" (symbol blocks of 96 BWT rows), produced by fm_reblock_kernel (fm_reblock.cuh) from any of the
 * reference's on-disk layouts:
 *
 *     blocks[sigma * nblocks + b] = uint4 { rank, w0, w1, w2 }
 *
 *   rank       = value the reference searcher returns for symbol sigma at row boundary X = 96*b (counter + popcount -
 *                '$' fix, i.e. src/fmIndexCPUBaseline.c:227-257 evaluated at X)
 *   w0..w2     = indicator bits of rows 96*b .. 96*b+95: bit i of the 96-bit little-endian value is 1 iff row 96*b+i
 *                carries k-step symbol sigma, is < bwtsize and is not one of the k '$' rows.
 *
 * One rank query = ONE aligned 16-byte load that brings both the sampled counter and the bitmap, i.e. one 32-byte DRAM
 * sector; the '$' corrections of the reference (:252-256) are folded into the layout, so the hot loop is
 *     X' = rank + popc(w & prefixmask(X - 96*b))
 * with no branches.  AltCounters files (tags 200/201) are re-derived into the same block format with the AltCounters
 * searcher's semantics (src/fmIndexCPUBaseline-AltCounters.c:218-266); its padding-entry quirk is
 * a per-symbol constant added for X >= quirk_start (fm_quirk_delta).
 */
#ifndef FM_DEVICE_CUH_
#define FM_DEVICE_CUH_

#include <stdint.h>
#include <cuda_runtime.h>

#define FM_SB_ROWS 96u

/* `make debug` (-DFM_DEBUG_BOUNDS): every table index a kernel computes is checked against the table's extent before the
 * load; a violation prints the site and traps (the launch then fails with an error, the tests with it).  This is the
 * stand-in for compute-sanitizer memcheck, which is closed on this GPU pool (profiles/r02_sanitizers.txt). */
#ifdef FM_DEBUG_BOUNDS
#include <stdio.h>
#define FM_BOUND(index, limit, what) do { if ((unsigned long long)(index) >= (unsigned long long)(limit)) { \
  printf("fm bounds: %s: %llu >= %llu (block %u thread %u)\n", what, (unsigned long long)(index), (unsigned long long)(limit), blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define FM_BOUND(index, limit, what) do { } while (0)
#endif
#define FM_SYM_NONE  0xFFu
#define FM_FSYM_NONE 0xFFFFu

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
  FM_BOUND(b, nblocks, "tail rank: SB96 block");
  for (int c1 = 0; c1 < 4; c1++) v[c1] = fm_ldg16(blocks + (size_t)(c | (c1 << 2)) * nblocks + b);
  uint32_t sum = tail_const + ((X > tail_row && c == tail_base) ? 1u : 0u);
  #pragma unroll
  for (int c1 = 0; c1 < 4; c1++) sum += fm_block_rank(v[c1], r);
  return sum;
}

/* last base of an odd-length read for both interval ends: one fetch from the tail table (the second only when R lies
 * in another block), or the four-fetch derivation when the table could not be allocated */
__device__ __forceinline__ void fm_tail_step(const uint4 *__restrict__ tail1, const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t c,
                                             uint32_t &L, uint32_t &R, uint32_t tail_const, uint32_t tail_row, uint32_t tail_base)
{
  if (tail1) {
    const uint32_t bL = fm_div96(L), bR = fm_div96(R);
    const uint4 *base = tail1 + (size_t) c * nblocks;
    FM_BOUND(bL, nblocks, "tail table block (L)"); FM_BOUND(bR, nblocks, "tail table block (R)"); FM_BOUND(c, 4u, "tail base");
    const uint4 vL = fm_ldg16(base + bL);
    const uint4 vR = (bL == bR) ? vL : fm_ldg16(base + bR);
    L = fm_block_rank(vL, L - bL * FM_SB_ROWS);
    R = fm_block_rank(vR, R - bR * FM_SB_ROWS);
  } else {
    L = fm_tail_rank(blocks, nblocks, c, L, tail_const, tail_row, tail_base);
    R = fm_tail_rank(blocks, nblocks, c, R, tail_const, tail_row, tail_base);
  }
}

/* .L2::64B: an L2 miss then fills 64 bytes (the whole LANES=2 block) instead of the 128-byte line every other
 * flavour pulls from HBM (profiles/r01_prefetch_variants.md) -- same fetch rate, half the DRAM traffic */
__device__ __forceinline__ void fm_ldg32(const uint4 *p, uint32_t (&w)[8])
{
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

template <int LANES> __device__ __forceinline__ uint32_t fm_group_sum(uint32_t v)
{
  #pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

/* bits [pos, pos+nbits) of a packed read kept in shared memory (one readable spare word after the read) */
__device__ __forceinline__ uint32_t fm_read_field(const uint32_t *q, uint32_t pos, uint32_t mask)
{
  const uint32_t i = pos >> 5;
  return __funnelshift_r(q[i], q[i + 1], pos & 31u) & mask;
}

__device__ __forceinline__ uint32_t fm_sb96_rank(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t s, uint32_t X)
{
  const uint32_t b = fm_div96(X);
  return fm_block_rank(blocks[(size_t) s * nblocks + b], X - b * FM_SB_ROWS);
}

/* out[i] = i: the keys of a start / lead table (a packed b-mer is its own key) */
static __global__ void fm_iota_kernel(uint32_t *out, uint32_t n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

/* AltCounters padding-entry quirk: what the reference AC searcher adds to rank(sigma, X) for X >= quirk_start
 * (2 bits per k-step symbol in quirk_mask; the block table holds the quirk-free ranks) */
__device__ __forceinline__ uint32_t fm_quirk_delta(uint32_t quirk_mask, uint32_t quirk_start, uint32_t sigma, uint32_t X)
{
  return X >= quirk_start ? ((quirk_mask >> (2u * sigma)) & 3u) : 0u;
}

/* rank of the reference searcher this table reproduces: SB96 value + the AltCounters padding-quirk constant */
__device__ __forceinline__ uint32_t fm_sb96_rank_q(const uint4 *__restrict__ blocks, uint32_t nblocks, uint32_t s, uint32_t X,
                                                   uint32_t quirk_start, uint32_t quirk_mask)
{
  return fm_sb96_rank(blocks, nblocks, s, X) + fm_quirk_delta(quirk_mask, quirk_start, s, X);
}


/* a compose kernel standing on row quirk_start - 1 at hop `hop` with the symbols `acc` so far, for the chain of row `origin` */
struct FmQuirkVisit { uint32_t origin, hop, acc; };

/* Phantom occurrences of an active AltCounters quirk (one thread; a handful of chains).
 * The quirked rank of symbol s is rank'(s, X) = rank(s, X) + delta_s [X >= Q]: as a counting function it is
 * "base + #{ m in M_s : m < X }" with M_s = occ(s) plus delta_s extra copies of row Q - 1.  Composing such functions
 * gives base + #{ elements < X } again, where an element is a chain  row -> s0 -> rank'-image -> s1 -> ...  and a chain
 * may take, at any hop where it stands on row Q - 1, a phantom copy instead of the row's real symbol.  Chains through
 * real symbols only are what the compose kernels (fm_sparse_compose_kernel, fm_fuse_compose_kernel) emit; this kernel enumerates every chain that takes at least
 * one phantom copy (depth-first from the recorded visits of row Q - 1) and writes its (key, origin row) pair.
 * Image of copy c of phantom symbol s at row Q - 1:  rank(s, Q - 1) + [real symbol of the row is s] + c. */
static __global__ void fm_quirk_phantoms_kernel(const uint4 *__restrict__ blocks, uint32_t nblocks, const uint8_t *__restrict__ sym,
                                          uint32_t kbits, uint32_t hops, uint32_t quirk_start, uint32_t quirk_mask,
                                          const FmQuirkVisit *__restrict__ visits, uint32_t nvisits,
                                          uint32_t *__restrict__ keys, uint32_t *__restrict__ rows, uint32_t max_out, uint32_t *__restrict__ nout)
{
  if (blockIdx.x || threadIdx.x) return;
  const uint32_t Q1 = quirk_start - 1u, nsym_k = 1u << kbits;
  struct Item { uint32_t origin, hop, acc, row, from_visit; };
  Item stack[64];
  uint32_t out = 0;
  for (uint32_t v = 0; v < nvisits; v++) {
    int sp = 0;
    stack[sp++] = Item{ visits[v].origin, visits[v].hop, visits[v].acc, Q1, 1u };
    while (sp) {
      const Item it = stack[--sp];
      if (it.hop == hops) {                                     /* a complete chain that used a phantom copy */
        if (out < max_out) { keys[out] = it.acc; rows[out] = it.origin; }
        out++;
        continue;
      }
      /* the row's real symbol: only for chains that already took a phantom copy (the all-real chain of a recorded
       * visit is the compose kernel's own output) */
      if (!it.from_visit) {
        const uint32_t s = sym[it.row];
        if (s != FM_SYM_NONE && sp < 63) {
          const uint32_t nrow = it.hop + 1 < hops ? fm_sb96_rank_q(blocks, nblocks, s, it.row, quirk_start, quirk_mask) : 0u;
          stack[sp++] = Item{ it.origin, it.hop + 1, it.acc | (s << (kbits * it.hop)), nrow, 0u };
        }
      }
      if (it.row != Q1) continue;
      /* phantom copies at row Q - 1 */
      const uint32_t real = sym[Q1];
      for (uint32_t s = 0; s < nsym_k; s++) {
        const uint32_t d = (quirk_mask >> (2u * s)) & 3u;
        for (uint32_t c = 0; c < d && sp < 63; c++) {
          const uint32_t nrow = fm_sb96_rank(blocks, nblocks, s, Q1) + (real == s ? 1u : 0u) + c;
          stack[sp++] = Item{ it.origin, it.hop + 1, it.acc | (s << (kbits * it.hop)), nrow, 0u };
        }
      }
    }
  }
  *nout = out;
}


#endif /* FM_DEVICE_CUH_ */

/*
 * fm_locate.cuh -- from (L,R) intervals to text positions (the follow-on of interval search; the
 * reference stops at (L,R), src/fmIndexCPUBaseline.c:288-290, so there is no reference code for this -- the checker is
 * a brute-force scan of the text in tests/).
 *
 * B200 has room for the whole suffix array next to the index (4 bytes per row: 8 GB for 2 Gbp), so nothing is
 * sampled: locate = one gather per occurrence, SA[L .. R).
 *
 * The suffix array is DERIVED FROM THE INDEX ITSELF, so it exists for index files too (they carry no SA):
 *   LF(r)    = rank1(c(r), r), c(r) = layer-0 BWT char of row r      -- one 1-step LF per row, read from the 4-symbol
 *              block table (k = 1: the SB96 table; k = 2: the tail table of fm_kernels.cuh, whose blocks hold exactly
 *              rank1 at block starts + the "row has char c" bits)
 *   SA[LF(r)] = SA[r] - 1, and the row whose char is '$' (dollarPositionBWT[0], the only row without a bit) has SA = 0,
 *   so SA[r] = number of LF steps from r to that row: list ranking over the LF permutation by pointer jumping
 *   (Wyllie): node[r] = { next, dist }, 32 rounds of one random 8-byte read per row.
 */
#ifndef FM_LOCATE_CUH_
#define FM_LOCATE_CUH_

#include "fm_device.cuh"

/* node[r] = { LF(r), 1 } for every row carrying a char, { r, 0 } for the row that carries none (the '$' row: the end of
 * the list); term[0] = that row, term[1] = how many such rows were seen (must be 1).  One thread per 96-row block. */
__global__ void fm_locate_lf_kernel(const uint4 *__restrict__ t1, uint32_t nblocks, uint32_t bwtsize, uint2 *__restrict__ node,
                                    uint32_t *__restrict__ term)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint64_t row0 = (uint64_t) b * FM_SB_ROWS;
  if (row0 >= bwtsize) return;
  uint32_t have[3] = { 0u, 0u, 0u };
  #pragma unroll
  for (uint32_t c = 0; c < 4; c++) {
    const uint4 v = t1[(size_t) c * nblocks + b];
    uint32_t w[3] = { v.y, v.z, v.w };
    uint32_t rank = v.x;
    for (int j = 0; j < 3; j++) {
      have[j] |= w[j];
      while (w[j]) {
        const uint32_t bit = __ffs(w[j]) - 1;
        w[j] &= w[j] - 1;
        const uint64_t row = row0 + 32u * j + bit;
        if (row < bwtsize) node[row] = make_uint2(rank, 1u);
        rank++;
      }
    }
  }
  for (int j = 0; j < 3; j++) {
    uint32_t miss = ~have[j];
    while (miss) {
      const uint32_t bit = __ffs(miss) - 1;
      miss &= miss - 1;
      const uint64_t row = row0 + 32u * j + bit;
      if (row < bwtsize) { node[row] = make_uint2((uint32_t) row, 0u); term[0] = (uint32_t) row; atomicAdd(term + 1, 1u); }
    }
  }
}

/* one pointer-jumping round: next <- next[next], dist <- dist + dist[next] (the terminal row points at itself, dist 0) */
__global__ void fm_locate_jump_kernel(const uint2 *__restrict__ in, uint2 *__restrict__ out, uint32_t bwtsize)
{
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= bwtsize) return;
  const uint2 a = in[r];
  if (a.x >= bwtsize) { out[r] = a; return; }                    /* a row the table did not cover: reported by the extract pass */
  const uint2 nx = __ldg(in + a.x);
  out[r] = make_uint2(nx.x, a.y + nx.y);
}

/* SA[r] = dist; status[0] counts rows that did not reach the terminal row (0 for a consistent index) */
__global__ void fm_locate_extract_kernel(const uint2 *__restrict__ node, uint32_t bwtsize, uint32_t terminal, uint32_t *__restrict__ sa,
                                         unsigned long long *__restrict__ status)
{
  const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= bwtsize) return;
  const uint2 a = node[r];
  sa[r] = a.y;
  if (a.x != terminal || a.y >= bwtsize) atomicAdd(status, 1ull);
}

/* positions[q * max_hits + j] = SA[L + j] for j < min(R - L, max_hits), 0xFFFFFFFF beyond; nhits[q] = R - L (0 for an
 * empty interval, whatever (L,R) the search left there).  One thread per (q, j). */
__global__ void fm_locate_kernel(const uint32_t *__restrict__ sa, const uint2 *__restrict__ lr, uint64_t nq, uint32_t max_hits,
                                 uint32_t *__restrict__ positions, uint32_t *__restrict__ nhits)
{
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * max_hits) return;
  const uint64_t q = t / max_hits;
  const uint32_t j = (uint32_t)(t - q * max_hits);
  const uint2 x = lr[q];
  const uint32_t cnt = x.y > x.x ? x.y - x.x : 0u;
  if (j == 0 && nhits) nhits[q] = cnt;
  if (j < cnt) FM_BOUND((uint64_t) x.x + j, 0x100000000ull, "locate: row");
  positions[t] = j < cnt ? __ldg(sa + x.x + j) : 0xFFFFFFFFu;
}


/* ------------------------------------------------------------------------ *
 * Sampled suffix array ("SA sampling"): only the rows whose text position is a multiple of `rate`
 * keep their SA value; any other row walks the 1-step LF mapping until it meets a marked row:
 *     SA[r] = SA[LF^t(r)] + t.
 * The walk runs on a table made for it -- one 64-byte block per 128 rows,
 *     words 0..3   rank1(c, 128 b) for c = A, C, G, T  (exact: the '$' correction is folded in, as in SB96)
 *     words 4..7   low bit of the row's BWT character, words 8..11 its high bit   (row 128 b + i = bit i)
 *     words 12..15 marks: the row's SA value is a multiple of `rate` (the row without a character -- text position 0 --
 *                  is always marked, so the walk never needs its character)
 * -- so that one LF step is ONE 64-byte fetch by a pair of lanes (character, rank and mark together; the four per-symbol
 * blocks of the tail table would cost four requests).  0.5 bytes per row + 4 bytes per `rate` rows + 4 bytes per 128 rows:
 * 1.3 GB instead of 8 GB for 2 Gbp at rate 32; a located occurrence costs (rate - 1) / 2 fetches on average + 2.
 * ------------------------------------------------------------------------ */
#define FM_LOC_ROWS 128u

/* one thread per 128-row block; sym[row] in {0..3, FM_SYM_NONE}, t1 = the 4-symbol SB96-shaped table */
__global__ void fm_locate_pack_kernel(const uint4 *__restrict__ t1, uint32_t nblocks96, const uint8_t *__restrict__ sym, const uint32_t *__restrict__ sa,
                                      uint32_t bwtsize, uint32_t rate, uint32_t nlb, uint4 *__restrict__ out, uint32_t *__restrict__ nmarks)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nlb) return;
  const uint64_t row0 = (uint64_t) b * FM_LOC_ROWS;
  uint32_t w[16];
  #pragma unroll
  for (int i = 0; i < 16; i++) w[i] = 0u;
  const uint32_t x0 = row0 < bwtsize ? (uint32_t) row0 : bwtsize;
  for (uint32_t c = 0; c < 4; c++) w[c] = fm_sb96_rank(t1, nblocks96, c, x0);
  uint32_t marked = 0;
  for (uint32_t i = 0; i < FM_LOC_ROWS; i++) {
    const uint64_t r = row0 + i;
    if (r >= bwtsize) break;
    const uint32_t c = sym[r];
    if (c != FM_SYM_NONE) { w[4 + (i >> 5)] |= (c & 1u) << (i & 31u); w[8 + (i >> 5)] |= (c >> 1) << (i & 31u); }
    if (c == FM_SYM_NONE || sa[r] % rate == 0u) { w[12 + (i >> 5)] |= 1u << (i & 31u); marked++; }
  }
  #pragma unroll
  for (int i = 0; i < 4; i++) out[(size_t) b * 4 + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  nmarks[b] = marked;
}

/* samples[markrank[b] + (marked rows of block b before row r)] = SA[r] for every marked row; one thread per block */
__global__ void fm_locate_samples_kernel(const uint4 *__restrict__ lblocks, const uint32_t *__restrict__ markrank, const uint32_t *__restrict__ sa,
                                         uint32_t bwtsize, uint32_t nlb, uint32_t *__restrict__ samples)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nlb) return;
  const uint4 m = lblocks[(size_t) b * 4 + 3];
  const uint32_t mw[4] = { m.x, m.y, m.z, m.w };
  uint32_t at = markrank[b];
  for (uint32_t j = 0; j < 4; j++) {
    uint32_t x = mw[j];
    while (x) {
      const uint32_t bit = __ffs(x) - 1;
      x &= x - 1;
      const uint64_t r = (uint64_t) b * FM_LOC_ROWS + 32u * j + bit;
      if (r < bwtsize) samples[at] = sa[r];
      at++;
    }
  }
}

/* 4-way select without dynamic register indexing */
__device__ __forceinline__ uint32_t fm_pick4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t i)
{
  return i == 0u ? a : (i == 1u ? b : (i == 2u ? c : d));
}

/* positions[q * max_hits + j] = SA[L + j] by LF walks; a PAIR of lanes per (q, j): lane 0 holds words 0..7 of the block
 * (ranks, low bits), lane 1 words 8..15 (high bits, marks) */
__global__ void __launch_bounds__(256) fm_locate_sampled_kernel(const uint4 *__restrict__ lblocks, const uint32_t *__restrict__ markrank,
                                                                const uint32_t *__restrict__ samples, const uint2 *__restrict__ lr, uint64_t nq,
                                                                uint32_t max_hits, uint32_t bwtsize, uint32_t rate, uint32_t norow,
                                                                uint32_t *__restrict__ positions, uint32_t *__restrict__ nhits)
{
  const uint64_t t = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const uint32_t lg = threadIdx.x & 1u;
  const bool in_range = t < nq * max_hits;
  const uint64_t q = in_range ? t / max_hits : 0;
  const uint32_t j = in_range ? (uint32_t)(t - q * max_hits) : 0u;
  const uint2 x = in_range ? lr[q] : make_uint2(0u, 0u);
  const uint32_t cnt = x.y > x.x ? x.y - x.x : 0u;
  if (in_range && j == 0 && lg == 0 && nhits) nhits[q] = cnt;
  bool walking = in_range && j < cnt;
  uint32_t r = x.x + j, steps = 0, result = 0xFFFFFFFFu;
  while (__any_sync(0xFFFFFFFFu, walking)) {
    uint32_t w[8];
    const uint32_t b = r / FM_LOC_ROWS, o = r % FM_LOC_ROWS, wi = o >> 5, bit = o & 31u;
    if (walking) { FM_BOUND(r, bwtsize, "sampled locate: row"); fm_ldg32(lblocks + (size_t) b * 4 + 2u * lg, w); }
    else { for (int i = 0; i < 8; i++) w[i] = 0u; }
    /* lane 0: w[0..3] ranks, w[4..7] low bits; lane 1: w[0..3] high bits, w[4..7] marks.  Exchange what the other needs. */
    uint32_t lo[4], hi[4];
    #pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t mine = lg == 0 ? w[4 + i] : w[i];               /* my plane's word i */
      const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, mine, 1);
      lo[i] = lg == 0 ? mine : other; hi[i] = lg == 0 ? other : mine;
    }
    uint32_t mk[4];                                                  /* the mark words, from lane 1 (all shuffles are executed by every lane) */
    #pragma unroll
    for (int i = 0; i < 4; i++) mk[i] = __shfl_sync(0xFFFFFFFFu, w[4 + i], 1, 2);
    const bool marked = (fm_pick4(mk[0], mk[1], mk[2], mk[3], wi) >> bit) & 1u;
    const uint32_t c = ((fm_pick4(lo[0], lo[1], lo[2], lo[3], wi) >> bit) & 1u) | (((fm_pick4(hi[0], hi[1], hi[2], hi[3], wi) >> bit) & 1u) << 1);
    const uint32_t rank_c = __shfl_sync(0xFFFFFFFFu, fm_pick4(w[0], w[1], w[2], w[3], c), 0, 2);    /* from lane 0 */
    if (walking && marked) {
      /* sample index = marks before this block + marks of the block below this row */
      uint32_t below = 0;
      #pragma unroll
      for (int i = 0; i < 4; i++) below += __popc(mk[i] & fm_lowmask((uint32_t) max((int) o - 32 * i, 0)));
      result = __ldg(samples + __ldg(markrank + b) + below) + steps;
      walking = false;
    } else if (walking) {
      uint32_t below = 0;
      #pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t m0 = (c & 1u) ? lo[i] : ~lo[i], m1 = (c & 2u) ? hi[i] : ~hi[i];
        below += __popc(m0 & m1 & fm_lowmask((uint32_t) max((int) o - 32 * i, 0)));
      }
      /* the row without a character stores character bits 00: rows behind it in its block must not count it as an 'A' */
      if (c == 0u && norow / FM_LOC_ROWS == b && norow < r) below -= 1u;
      r = rank_c + below;
      steps++;
    }
  }
  if (in_range && lg == 0) positions[t] = result;
}

#endif /* FM_LOCATE_CUH_ */

/*
 * fm_reblock.cuh -- index residency / layout stage: raw file entries (tags 100/101/200/201) -> SB96,
 * and the tail table derived from a 2-step SB96 table.  Included by fm_index.cu only.
 */
#ifndef FM_REBLOCK_CUH_
#define FM_REBLOCK_CUH_

#include "fm_device.cuh"

/* raw (file-order) index as uploaded, for the re-blocker */
struct FmRawIndex {
  const uint32_t *entries;
  uint32_t tag, k, d, ncounters, nentries, entry_words, bwtsize, nentries_std;
  uint32_t dpos[2], dbase[2];
  uint32_t quirk_start, quirk_mask;
};

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

#endif /* FM_REBLOCK_CUH_ */

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
  positions[t] = j < cnt ? __ldg(sa + x.x + j) : 0xFFFFFFFFu;
}

#endif /* FM_LOCATE_CUH_ */

/*
 * fm_sparsetab.cu -- sparse-step table (fm_sparse.cuh): construction on the GPU, start / lead tables, launch plan, fetch counter.
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 */
#include "fm_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <mutex>
#include "fm_sparse.cuh"

/* ------------------------------------------------------------------------ *
 * sparse-step table (fm_sparse.cuh)
 * ------------------------------------------------------------------------ */
static const uint2 *fm_build_lead(fmgpu_index_t *idx, uint32_t b);

extern "C" int32_t fmgpu_index_unsparsify(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sblocks || idx->sstart) {
    CU_TRY(cudaSetDevice(idx->device));
    cudaFree(idx->sblocks); cudaFree(idx->sstart);
    idx->sblocks = NULL; idx->sstart = NULL;
    for (int b = 0; b < 16; b++) { cudaFree(idx->slead[b]); idx->slead[b] = NULL; }
    idx->slead_tried = 0;
  }
  idx->stables = 0;
  idx->s_uni_nb = 0; idx->s_uni_scale = 0; idx->meta.sparse_uniform_nb = 0;
  idx->meta.sparse_bases = 0; idx->meta.sparse_lambda = 0; idx->meta.sparse_bytes = 0; idx->meta.sparse_blocks = 0;
  idx->meta.sparse_overflow = 0; idx->meta.sparse_start_bases = 0; idx->meta.sparse_lanes = 0;
  idx->meta.sparse_tree_nodes = 0; idx->meta.sparse_tree_rows = 0; idx->meta.sparse_tree_depth = 0;
  fm_budget_account(idx);
  return FM_SUCCESS;
}

template <int LANES>
static cudaError_t fm_sparse_fill(const uint32_t *rows, const uint32_t *symstart, uint32_t nsym, uint32_t nb, uint32_t scale,
                                  const uint32_t *rank0, const uint32_t *extoff, uint32_t total_ext, uint4 *sblocks)
{
  const uint64_t nroots = (uint64_t) nsym * nb;
  fm_sparse_fill_roots_kernel<LANES><<<(unsigned)((nroots + 255) / 256), 256>>>(rows, symstart, nsym, nb, scale, rank0, extoff, sblocks);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && total_ext) {
    fm_sparse_fill_ext_kernel<LANES><<<(total_ext + 255) / 256, 256>>>(rows, symstart, nsym, nb, scale, rank0, extoff, total_ext, sblocks);
    e = cudaGetLastError();
  }
  return e;
}

/* Widest step the automatic choice takes: the widest multiple of k up to 14 bases that still leaves lambda rows per wide
 * symbol on average (the grid then costs ~32*lanes/lambda bytes per text base whatever the width; wider would only add
 * empty blocks): 14 bases from 1.34 Gbp, 12 from 84 Mbp, 10 from 5.2 Mbp ... at lambda 5. */
static uint32_t fm_sparse_auto_bases(uint32_t k, uint32_t n, uint32_t lambda)
{
  for (uint32_t cand = 14; cand >= 2 * k; cand--)
    if (cand % k == 0 && (((uint64_t) lambda) << (2 * cand)) <= n) return cand;
  return 2 * k;
}

extern "C" int32_t fmgpu_index_sparsify(fmgpu_index_t *idx, uint32_t sparse_bases, uint32_t lambda, uint32_t lanes)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sblocks) return FM_SUCCESS;
  if (idx->meta.bwtsize >= FM_SP_INNER - 4096u) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "text too long for the sparse-step table");
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t k = idx->meta.steps, n = idx->meta.bwtsize;
  if (lanes == 0) lanes = 2;
  if (lanes != 2 && lanes != 4) return fm_fail_msg(FM_E_BAD_ARGUMENT, "sparse block lanes must be 2 (64-byte blocks) or 4 (128-byte blocks)");
  const uint32_t slots = 8 * lanes - 1, bbytes = 32 * lanes;
  if (lambda == 0) lambda = lanes == 4 ? 12 : 5;
  if (lambda > slots) return fm_fail_msg(FM_E_BAD_ARGUMENT, "lambda must not exceed the slots of a block (15 or 31)");
  const uint32_t ks = sparse_bases ? sparse_bases : fm_sparse_auto_bases(k, n, lambda);
  if (ks % k || ks <= k || ks > 14) return fm_fail_msg(FM_E_BAD_ARGUMENT, "sparse bases must be a multiple of k, larger than k and at most 14");
  const uint32_t nsym = 1u << (2 * ks), hops = ks / k, kbits = 2 * k;
  const uint32_t qstart = idx->meta.quirk_start, qmask = idx->meta.quirk_mask;
  /* grid: the same block count for every wide symbol, ~lambda rows per block on average */
  uint64_t per = ((uint64_t) n + (uint64_t) nsym * lambda - 1) / ((uint64_t) nsym * lambda);
  if (per < 1) per = 1;
  const uint64_t nroots = per * nsym;
  if (nroots >= (1ull << 32) - (1ull << 29)) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "too many blocks for the sparse-step table");
  const uint32_t nbu = (uint32_t) per;
  unsigned long long sc = ((((unsigned long long) nbu) << 32) - 1ull) / n;          /* largest scale with umulhi(bwtsize, scale) <= nb - 1 */
  const uint32_t scale = sc > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) sc;

  const uint64_t nrows = (uint64_t) idx->meta.nblocks * FM_SB_ROWS;
  const uint64_t nkeys_cap = (uint64_t) n + 4096;                                    /* rows + phantom occurrences */
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  const uint64_t worst_tree = ((uint64_t) n / (slots - 1) + 2) * bbytes;            /* every row in an overfull bucket */
  const uint64_t build_peak = 16ull * nkeys_cap + nrows + 16ull * nsym + 8ull * nroots + (1ull << 30);
  const uint64_t final_peak = 4ull * nkeys_cap + 16ull * nsym + 8ull * nroots + nroots * bbytes + (1ull << 30);
  if ((build_peak > final_peak ? build_peak : final_peak) > free_b)
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough free device memory to build the sparse-step table");
  if (!fm_budget_allows(idx, nroots * bbytes)) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "the sparse-step table would exceed the derived-table budget");
  (void) worst_tree;

  uint8_t *sym = NULL; uint32_t *keys = NULL, *rows = NULL, *keys2 = NULL, *rows2 = NULL, *symstart = NULL, *rank0 = NULL, *ext = NULL, *extoff = NULL;
  uint4 *sblocks = NULL; void *tmp = NULL; unsigned long long *d_stats = NULL; FmQuirkVisit *visits = NULL; uint32_t *d_cnt = NULL;
  size_t tmp_bytes = 0, tmp2 = 0;
  unsigned long long stats[3] = { 0, 0, 0 };
  uint32_t hcnt[2] = { 0, 0 }, total_ext = 0;
  uint64_t nkeys = n;
  const uint32_t max_visits = 1024, max_ph = 4096;
  cudaError_t e = cudaMalloc((void **) &sym, nrows);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys, 4ull * nkeys_cap);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rows, 4ull * nkeys_cap);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys2, 4ull * nkeys_cap);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rows2, 4ull * nkeys_cap);
  if (e == cudaSuccess) e = cudaMalloc((void **) &symstart, 4ull * (nsym + 1));
  if (e == cudaSuccess) e = cudaMalloc((void **) &rank0, 4ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_stats, 24);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_cnt, 8);
  if (e == cudaSuccess) e = cudaMalloc((void **) &visits, sizeof(FmQuirkVisit) * max_visits);
  if (e == cudaSuccess) e = cudaMemset(d_stats, 0, 24);
  if (e == cudaSuccess) e = cudaMemset(d_cnt, 0, 8);
  if (e == cudaSuccess) e = fm_row_symbols(idx, nrows, sym);
  if (e == cudaSuccess) {
    fm_sparse_compose_kernel<<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(idx->blocks, idx->meta.nblocks, sym, n, kbits, hops, nsym,
                                                                              qstart, qmask, visits, d_cnt, max_visits, keys, rows);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && qmask && qstart != 0u) {
    /* AltCounters padding quirk: the few phantom occurrences go behind the n real (key, row) pairs, before the sort */
    e = cudaMemcpy(hcnt, d_cnt, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && hcnt[0] > max_visits) e = cudaErrorInvalidValue;
    if (e == cudaSuccess && hcnt[0]) {
      fm_quirk_phantoms_kernel<<<1, 1>>>(idx->blocks, idx->meta.nblocks, sym, kbits, hops, qstart, qmask, visits, hcnt[0],
                                          keys + n, rows + n, max_ph, d_cnt + 1);
      e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaMemcpy(hcnt + 1, d_cnt + 1, 4, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess && hcnt[1] > max_ph) e = cudaErrorInvalidValue;
      nkeys = (uint64_t) n + hcnt[1];
    }
  }
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, keys, keys2, rows, rows2, (int64_t) nkeys, 0, (int)(2 * ks + 1));
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(NULL, tmp2, (uint32_t *) NULL, (uint32_t *) NULL, (int64_t) nroots);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
  /* two-key order (symbol, then row): the sort is stable and the input rows ascend, but phantom pairs sit at the end, so
   * rows are sorted first when there are any */
  if (e == cudaSuccess && nkeys > n) {
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, rows, rows2, keys, keys2, (int64_t) nkeys, 0, 32);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys2, keys, rows2, rows, (int64_t) nkeys, 0, (int)(2 * ks + 1));
    if (e == cudaSuccess) { uint32_t *t = keys; keys = keys2; keys2 = t; t = rows; rows = rows2; rows2 = t; }     /* result in keys2 / rows2 */
  } else if (e == cudaSuccess) {
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, rows, rows2, (int64_t) nkeys, 0, (int)(2 * ks + 1));
  }
  if (e == cudaSuccess) {
    fm_sparse_symstart_kernel<<<(nsym + 1 + 255) / 256, 256>>>(keys2, nkeys, nsym, symstart);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    fm_sparse_rank0_kernel<<<(nsym + 255) / 256, 256>>>(idx->blocks, idx->meta.nblocks, kbits, hops, nsym, qstart, qmask, rank0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  /* the sort's other buffers are dead now: release them before the per-root arrays and the table are allocated */
  cudaFree(keys); keys = NULL; cudaFree(rows); rows = NULL; cudaFree(sym); sym = NULL; cudaFree(keys2); keys2 = NULL;
  if (e == cudaSuccess) e = cudaMalloc((void **) &ext, 4ull * nroots);
  if (e == cudaSuccess) e = cudaMalloc((void **) &extoff, 4ull * nroots);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)((nroots + 255) / 256);
    if (lanes == 4) fm_sparse_count_kernel<4><<<grid, 256>>>(rows2, symstart, nsym, nbu, scale, ext, d_stats);
    else            fm_sparse_count_kernel<2><<<grid, 256>>>(rows2, symstart, nsym, nbu, scale, ext, d_stats);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, ext, extoff, (int64_t) nroots);
  if (e == cudaSuccess) {
    uint32_t last_off = 0, last_ext = 0;
    e = cudaMemcpy(&last_off, extoff + (nroots - 1), 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&last_ext, ext + (nroots - 1), 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(stats, d_stats, 24, cudaMemcpyDeviceToHost);
    total_ext = last_off + last_ext;                             /* <= n / (slots - 1) + roots: far below 2^32 */
  }
  cudaFree(ext); ext = NULL;
  const uint64_t total_blocks = nroots + total_ext;
  if (e == cudaSuccess && total_blocks >= 0xFFFFFFF0ull) e = cudaErrorInvalidValue;
  if (e == cudaSuccess && !fm_budget_allows(idx, total_blocks * bbytes)) {
    cudaFree(rows2); cudaFree(symstart); cudaFree(rank0); cudaFree(extoff); cudaFree(tmp); cudaFree(d_stats); cudaFree(d_cnt); cudaFree(visits);
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "the sparse-step table would exceed the derived-table budget");
  }
  if (e == cudaSuccess) e = cudaMalloc((void **) &sblocks, total_blocks * bbytes);
  if (e == cudaSuccess) e = lanes == 4 ? fm_sparse_fill<4>(rows2, symstart, nsym, nbu, scale, rank0, extoff, total_ext, sblocks)
                                       : fm_sparse_fill<2>(rows2, symstart, nsym, nbu, scale, rank0, extoff, total_ext, sblocks);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(sym); cudaFree(keys); cudaFree(rows); cudaFree(keys2); cudaFree(rows2); cudaFree(symstart); cudaFree(rank0); cudaFree(ext); cudaFree(extoff);
  cudaFree(tmp); cudaFree(d_stats); cudaFree(d_cnt); cudaFree(visits);
  if (e != cudaSuccess) {
    cudaFree(sblocks);
    cudaGetLastError();                                          /* a failed cudaMalloc stays "last error" otherwise and fails the next attempt's first check */
    if (e == cudaErrorMemoryAllocation) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough device memory for the sparse-step table (the plain kernels still serve this index)");
    return fm_fail(e, "fmgpu_index_sparsify", __FILE__, __LINE__);
  }
  idx->sblocks = sblocks; idx->s_uni_nb = nbu; idx->s_uni_scale = scale; idx->meta.sparse_uniform_nb = nbu;
  idx->meta.sparse_bases = ks; idx->meta.sparse_lambda = lambda; idx->meta.sparse_blocks = total_blocks;
  idx->meta.sparse_overflow = stats[0]; idx->meta.sparse_bytes = total_blocks * bbytes; idx->meta.sparse_lanes = lanes;
  idx->meta.sparse_tree_nodes = total_ext; idx->meta.sparse_tree_rows = stats[1]; idx->meta.sparse_tree_depth = (uint32_t) stats[2];

  /* start table: the sparse kernel itself searches every SB-mer once (a packed SB-mer IS its key); SB = the
   * largest whole number of sparse steps within 12 bases */
  {
    const char *env = getenv("FMGPU_START_TABLE");
    const bool want = env && *env ? atoi(env) != 0 : idx->meta.nbytes >= (1ull << 30);
    const uint32_t ssteps = 12 / ks, sb = ssteps * ks;
    if (want && ssteps && ((uint64_t) 1 << (2 * sb)) < n) {
      const uint32_t nk = 1u << (2 * sb);
      uint32_t *skeys = NULL; uint2 *table = NULL;
      e = cudaMalloc((void **) &skeys, (size_t) nk * 4);
      if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nk * 8);
      if (e == cudaSuccess) { fm_iota_kernel<<<(nk + 255) / 256, 256>>>(skeys, nk); e = cudaGetLastError(); }
      int32_t rc = FM_SUCCESS;
      if (e == cudaSuccess) rc = fm_launch_sparse(idx, skeys, nk, sb, (uint32_t *) table, FM_DEFAULT_VARIANT, 0, NULL, false);
      if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
      cudaFree(skeys);
      if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); }   /* the table is optional */
      else { idx->sstart = table; idx->meta.sparse_start_bases = sb; idx->meta.sparse_bytes += (uint64_t) nk * 8; }
    }
    /* lead tables (fm_build_lead): the small ones now, the wide ones (12 .. 15 bases, up to 8.6 GB) when fmgpu_index_prepare asks */
    idx->stables = want ? 1 : 0;
    if (want)
      for (uint32_t b = 1; b <= ks + 1 && b < 12; b++) fm_build_lead(idx, b);
  }
  fm_budget_account(idx);
  return FM_SUCCESS;
}

/* Lead table of width b: (L,R) of every b-mer, computed by the search kernel itself (a packed b-mer is its own key).
 * A read whose length leaves b bases over -- or b - KS, giving up one sparse step -- starts from it and then runs only
 * whole sparse steps: no SB96 fetches behind them and no tail fetch.  Any parity: on a 2-step index an odd width ends
 * with the derived 1-step rank while the TABLE is computed (bases may be grouped into steps in any way; every grouping
 * composes the same LF steps).  Widths 6 .. 11 are built with the sparse table, 12 .. 15 (134 MB .. 8.6 GB) by the first
 * search that needs them.  NULL when tables are off for this index, the width is not representable or memory is short. */
static std::mutex g_lead_mutex;
static const uint2 *fm_build_lead(fmgpu_index_t *idx, uint32_t b)
{
  if (b < 1 || b >= 16 || !idx->stables) return NULL;
  std::lock_guard<std::mutex> lock(g_lead_mutex);
  if (idx->slead[b]) return idx->slead[b];
  if (idx->slead_tried & (1u << b)) return NULL;
  idx->slead_tried |= 1u << b;
  const uint32_t k = idx->meta.steps, n = idx->meta.bwtsize;
  if ((idx->sstart && b == idx->meta.sparse_start_bases) || (b % k && !(k == 2 && idx->meta.tail_valid)) || ((uint64_t) 1 << (2 * b)) >= n) return NULL;
  if (cudaSetDevice(idx->device) != cudaSuccess) { cudaGetLastError(); return NULL; }
  const uint32_t nkeys = 1u << (2 * b);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return NULL; }
  if ((uint64_t) nkeys * 12 + (2ull << 30) > free_b || !fm_budget_allows(idx, (uint64_t) nkeys * 8)) return NULL;
  if (b % k) fm_build_tail(idx);                               /* an odd width ends with the derived 1-step rank */
  uint32_t *skeys = NULL; uint2 *table = NULL;
  cudaError_t e = cudaMalloc((void **) &skeys, (size_t) nkeys * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nkeys * 8);
  if (e == cudaSuccess) { fm_iota_kernel<<<(nkeys + 255) / 256, 256>>>(skeys, nkeys); e = cudaGetLastError(); }
  int32_t rc = FM_SUCCESS;
  if (e == cudaSuccess) rc = fm_launch_sparse(idx, skeys, nkeys, b, (uint32_t *) table, FM_DEFAULT_VARIANT, 0, NULL, false);   /* (a lead table is computed without lead tables) */
  if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
  cudaFree(skeys);
  if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); return NULL; }
  idx->slead[b] = table; idx->meta.sparse_bytes += (uint64_t) nkeys * 8;
  return table;
}

/* The plan of a sparse search of `len`-base reads: S whole sparse steps + lb leftover bases (rem base-k steps + an odd
 * tail base).  (B) the leftover bases -- or leftover + one sparse step's bases -- are taken first from their lead table
 * and only whole sparse steps follow: S (or S - 1) block fetches plus one lookup, no tail fetch; (A) no such table: the
 * start table replaces the first sparse step(s) and the rem steps follow the sparse ones, S - m + rem block fetches from
 * DRAM (+ tail); (C) no tables at all (small indexes): the rem steps run first on the upper (L2-resident) levels of SB96.
 * want[] lists, in order of preference, the lead-table widths plan (B) could use; `have` tells which exist. */
struct fm_sparse_plan { uint32_t S, rem, m, lb; uint32_t want[2]; };
static fm_sparse_plan fm_sparse_plan_for(const fmgpu_index_t *idx, uint32_t len)
{
  const uint32_t k = idx->meta.steps, ks = idx->meta.sparse_bases, hops = ks / k;
  fm_sparse_plan pl;
  pl.S = (len / k) / hops; pl.rem = (len / k) % hops;
  pl.m = idx->sstart ? idx->meta.sparse_start_bases / ks : 0u;
  pl.lb = len - pl.S * ks;                                     /* leftover bases, the odd one included */
  pl.want[0] = pl.want[1] = 0;
  if (idx->stables && (pl.lb >= 1 || !pl.m)) {
    /* the leftover bases themselves when the interval they leave is much narrower than a bucket (4^lb >= 8 x blocks per
     * symbol: the first sparse step then rarely needs two fetches), else leftover + one sparse step's bases (12 .. 15;
     * with no leftover and no start table -- 14 bases per step -- the table of all 14-mers IS the start table) */
    const uint64_t nb_mean = idx->meta.sparse_blocks >> (2 * ks);
    const uint32_t lb = pl.lb, S = pl.S;
    const bool narrow = lb >= 1 && lb < 16 && ((uint64_t) 1 << (2 * lb)) >= 8 * (nb_mean ? nb_mean : 1);
    if (narrow && S >= 1) pl.want[0] = lb;
    else if (!narrow && S >= 2 && lb + ks < 16 && lb + ks > idx->meta.sparse_start_bases) { pl.want[0] = lb + ks; if (lb >= 1 && !pl.m) pl.want[1] = lb; }
    else if (!narrow && lb >= 1 && S >= 1 && !pl.m) pl.want[0] = lb;            /* better two fetches in the first step than SB96 steps */
    else if (S == 0 && lb >= 1 && lb < 16) pl.want[0] = lb;                     /* a read shorter than one sparse step: one lookup */
  }
  return pl;
}

void fm_sparse_prepare(fmgpu_index_t *idx, uint32_t len)
{
  if (!idx->sblocks) return;
  const fm_sparse_plan pl = fm_sparse_plan_for(idx, len);
  for (int c = 0; c < 2; c++)
    if (pl.want[c] && fm_build_lead(idx, pl.want[c])) break;
}

typedef void (*fm_sparse_fn)(const FmSparseParams);
template <int K, int LANES>
static fm_sparse_fn fm_pick_sparse(int qpt)
{
  if (qpt == 0) return fm_search_sparse_kernel<K, LANES, 1, 256, 6, true>;        /* instrumented */
  if (qpt == 1) return fm_search_sparse_kernel<K, LANES, 1, 256, 6, false>;
  if (qpt == 2) return fm_search_sparse_kernel<K, LANES, 2, 256, 4, false>;
  if (qpt == 3) return fm_search_sparse_kernel<K, LANES, 3, 256, 4, false>;
  if (qpt == 4) return fm_search_sparse_kernel<K, LANES, 4, 256, 3, false>;
  return NULL;
}

typedef void (*fm_sparse_dyn_fn_t)(const FmSparseParams, uint32_t);
template <int K, int LANES>
static fm_sparse_dyn_fn_t fm_pick_sparse_dyn(int qpt)
{
  if (qpt == 1) return fm_search_sparse_dyn_kernel<K, LANES, 1, 256, 6>;
  if (qpt == 2) return fm_search_sparse_dyn_kernel<K, LANES, 2, 256, 4>;
  if (qpt == 3) return fm_search_sparse_dyn_kernel<K, LANES, 3, 256, 4>;
  if (qpt == 4) return fm_search_sparse_dyn_kernel<K, LANES, 4, 256, 3>;
  return NULL;
}

/* dynamic read assignment pays when reads differ in length of their walk, i.e. when a visible part of the text lives in
 * search trees; on an even text the static kernel is the same speed with fewer instructions.  $FMGPU_SPARSE_DYNAMIC=0/1 forces. */
static bool fm_sparse_dynamic_enabled(const fmgpu_index_t *idx)
{
  const char *env = getenv("FMGPU_SPARSE_DYNAMIC");
  if (env && *env) return atoi(env) != 0;
  return idx->meta.sparse_tree_rows * 100ull > idx->meta.bwtsize;      /* more than 1 % of the rows in trees */
}

int32_t fm_launch_sparse(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                         uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters, bool use_lead_tables)
{
  if (!idx->sblocks) return fm_fail_msg(FM_E_BAD_ARGUMENT, "FMGPU_MODE_SPARSE needs fmgpu_index_sparsify() on this replica first");
  const uint32_t k = idx->meta.steps, ks = idx->meta.sparse_bases, lanes = idx->meta.sparse_lanes;
  const int vq_asked = v.queries_per_thread;
  if (v.queries_per_thread < 1 || v.queries_per_thread > 4) v.queries_per_thread = 3;   /* profiles/r02_sparse_qpt_sweep.jsonl */
  FmSparseParams p;
  p.sblocks = idx->sblocks; p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results;
  p.nblocks = idx->meta.nblocks; p.nq = (uint32_t) nq;
  /* plan (fm_sparse_plan_for); only tables that exist are used -- fmgpu_index_prepare builds the ones a length wants */
  const fm_sparse_plan pl = fm_sparse_plan_for(idx, len);
  const uint32_t S = pl.S, rem = pl.rem, m = pl.m, lb = pl.lb;
  uint32_t lead = 0;
  for (int c = 0; c < 2 && use_lead_tables && !lead; c++)
    if (pl.want[c] && idx->slead[pl.want[c]]) lead = pl.want[c];
  p.nfront = 0; p.nback = 0; p.nsteps = S; p.start = NULL; p.start_bits = 0;
  if (lead) { p.start = idx->slead[lead]; p.start_bits = 2 * lead; p.nsteps = S - (lead > lb ? 1u : 0u); }
  else if (m && S >= m) { p.start = idx->sstart; p.start_bits = 2 * ks * m; p.nsteps = S - m; p.nback = rem; }
  else p.nfront = rem;
  p.wpq = fmgpu_words_per_query(len); p.bwtsize = idx->meta.bwtsize;
  p.sbits = 2 * ks;
  p.nb = idx->s_uni_nb; p.scale = idx->s_uni_scale; p.nroots = idx->s_uni_nb << (2 * ks);
  p.total_blocks = (uint32_t) idx->meta.sparse_blocks;
  p.quirk_start = idx->meta.quirk_start; p.quirk_mask = idx->meta.quirk_mask;
  p.fetch_counters = d_counters;
  p.has_tail = lead ? 0u : len % k;                            /* a lead table already holds the odd base */
  p.tail_row = idx->meta.tail_row; p.tail_base = idx->meta.tail_base;
  for (int c = 0; c < 4; c++) p.tail_const[c] = idx->meta.tail_const[c];
  p.tail1 = p.has_tail ? idx->tail1 : NULL;
  if (d_counters) v.queries_per_thread = 1;
  uint32_t qper; size_t smem;
  for (;;) {
    qper = (256 / lanes) * v.queries_per_thread;
    smem = 16 + ((size_t) qper * p.wpq + 4) * 4;
    if (smem <= 200 * 1024) break;
    if (v.queries_per_thread > 1) v.queries_per_thread -= 1;
    else return fm_fail_msg(FM_E_QUERY_SHAPE, "reads too long to stage in shared memory");
  }
  /* a plan of sparse steps only (the benchmark's; most lengths on a large index thanks to the lead tables): reads are handed
   * to the lane groups dynamically, so that on a skewed text a read that walks deep trees does not hold its warp back
   * ($FMGPU_SPARSE_DYNAMIC=0 forces the static kernel).  queries_per_thread 0 = 1 here. */
  if (!d_counters && p.nfront == 0 && p.nback == 0 && !p.has_tail && p.nsteps >= 1 && fm_sparse_dynamic_enabled(idx)) {
    typedef void (*fm_sparse_dyn_fn)(const FmSparseParams, uint32_t);
    const int q = vq_asked >= 1 && vq_asked <= 4 ? vq_asked : 1;       /* one read per lane group, 6 CTAs per SM: the fastest on skewed texts */
    fm_sparse_dyn_fn dfn = NULL;
    if (k == 2) dfn = lanes == 4 ? fm_pick_sparse_dyn<2, 4>(q) : fm_pick_sparse_dyn<2, 2>(q);
    else        dfn = lanes == 4 ? fm_pick_sparse_dyn<1, 4>(q) : fm_pick_sparse_dyn<1, 2>(q);
    const char *renv = getenv("FMGPU_SPARSE_ROUNDS");
    uint32_t rounds = renv && *renv && atoi(renv) >= 1 ? (uint32_t) atoi(renv) : (q == 1 ? 8u : 4u), rpc; size_t dsmem;   /* profiles/r02_skewed_text.md, "rounds" */
    for (;;) {
      rpc = (256 / lanes) * q * rounds;
      dsmem = 16 + ((size_t) rpc * p.wpq + 4) * 4;
      if (dsmem <= 48 * 1024 || rounds == 1) break;
      rounds--;
    }
    if (dfn && dsmem <= 200 * 1024) {
      if (dsmem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) dfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dsmem));
      const uint32_t dgrid = (uint32_t)((nq + rpc - 1) / rpc);
      void *dargs[] = { (void *) &p, (void *) &rpc };
      CU_TRY(cudaLaunchKernel((const void *) dfn, dim3(dgrid), dim3(256), dargs, dsmem, stream));
      return FM_SUCCESS;
    }
  }
  const int qsel = d_counters ? 0 : v.queries_per_thread;
  fm_sparse_fn fn = k == 2 ? (lanes == 4 ? fm_pick_sparse<2, 4>(qsel) : fm_pick_sparse<2, 2>(qsel))
                           : (lanes == 4 ? fm_pick_sparse<1, 4>(qsel) : fm_pick_sparse<1, 2>(qsel));
  if (!fn) return fm_fail_msg(FM_E_BAD_ARGUMENT, "no sparse kernel for this variant");
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const uint32_t grid = (uint32_t)((nq + qper - 1) / qper);
  void *args[] = { (void *) &p };
  CU_TRY(cudaLaunchKernel((const void *) fn, dim3(grid), dim3(256), args, smem, stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_count_fetches_sparse_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                                     uint32_t *d_results, void *stream, uint64_t *nsparse_blocks, uint64_t *nsb96_blocks,
                                                     uint64_t *noverflows)
{
  if (!idx || !d_packed || !d_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (len == 0 || (len % idx->meta.steps && !idx->meta.tail_valid)) return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a positive multiple of k");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long *d_c = NULL, h[3] = { 0, 0, 0 };
  CU_TRY(cudaMalloc((void **) &d_c, 24));
  CU_TRY(cudaMemsetAsync(d_c, 0, 24, (cudaStream_t) stream));
  int32_t rc = nq ? fm_launch_sparse(idx, d_packed, nq, len, d_results, FM_DEFAULT_VARIANT, (cudaStream_t) stream, d_c, true) : FM_SUCCESS;
  if (rc == FM_SUCCESS) {
    cudaError_t e = cudaMemcpyAsync(h, d_c, 24, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) rc = fm_fail(e, "fetch counters D2H", __FILE__, __LINE__);
  }
  cudaFree(d_c);
  if (nsparse_blocks) *nsparse_blocks = h[0];
  if (nsb96_blocks) *nsb96_blocks = h[1];
  if (noverflows) *noverflows = h[2];
  return rc;
}

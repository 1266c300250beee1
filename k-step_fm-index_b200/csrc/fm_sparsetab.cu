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
static int32_t fm_sparsify(fmgpu_index_t *idx, uint32_t sparse_bases, uint32_t lambda, uint32_t lanes, bool require_uniform);

extern "C" int32_t fmgpu_index_unsparsify(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sblocks || idx->sdir || idx->sstart) {
    CU_TRY(cudaSetDevice(idx->device));
    cudaFree(idx->sblocks); cudaFree(idx->sdir); cudaFree(idx->sstart);
    idx->sblocks = NULL; idx->sdir = NULL; idx->sstart = NULL;
    for (int b = 0; b < 16; b++) { cudaFree(idx->slead[b]); idx->slead[b] = NULL; }
    idx->slead_tried = 0;
  }
  idx->stables = 0;
  idx->s_uni_nb = 0; idx->s_uni_scale = 0; idx->meta.sparse_uniform_nb = 0;
  idx->meta.sparse_bases = 0; idx->meta.sparse_lambda = 0; idx->meta.sparse_bytes = 0; idx->meta.sparse_blocks = 0;
  idx->meta.sparse_overflow = 0; idx->meta.sparse_start_bases = 0; idx->meta.sparse_lanes = 0;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_sparsify(fmgpu_index_t *idx, uint32_t sparse_bases, uint32_t lambda, uint32_t lanes)
{
  const int32_t rc = fm_sparsify(idx, sparse_bases, lambda, lanes, false);
  if (idx) fm_budget_account(idx);
  return rc;
}

/* require_uniform: the automatic choice is trying a wide (14 / 12 bases) table, which only exists as a uniform grid */
static int32_t fm_sparsify(fmgpu_index_t *idx, uint32_t sparse_bases, uint32_t lambda, uint32_t lanes, bool require_uniform)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sblocks) return FM_SUCCESS;
  if (idx->meta.quirk_mask) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "sparse steps are unavailable for an AltCounters index carrying the padding-entry quirk");
  if (idx->meta.bwtsize >= FM_SP_OVF) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "text too long for the sparse-step table");
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t k = idx->meta.steps, n = idx->meta.bwtsize;
  if (lanes == 0) lanes = 2;
  if (lanes != 2 && lanes != 4) return fm_fail_msg(FM_E_BAD_ARGUMENT, "sparse block lanes must be 2 (64-byte blocks) or 4 (128-byte blocks)");
  const uint32_t slots = 8 * lanes - 1, bbytes = 32 * lanes;
  if (lambda == 0) lambda = lanes == 4 ? 12 : 5;
  if (lambda > slots) return fm_fail_msg(FM_E_BAD_ARGUMENT, "lambda must not exceed the slots of a block (15 or 31)");
  uint32_t ks = sparse_bases;
  if (ks == 0) {
    /* 12 bases per step when that table can be a uniform grid (a directory of 4^12 entries would not stay in L2): tried first,
     * given up as soon as the symbol counts turn out uneven.  Else the widest multiple of k up to 10 with >= 64 rows per symbol. */
    const char *env = getenv("FMGPU_SPARSE_UNIFORM");
    if (!require_uniform && !(env && *env && atoi(env) == 0)) {
      /* 14 bases need >= lambda rows per 14-mer on average, 12 bases >= 64 rows per 12-mer */
      const uint32_t wide[2] = { 14, 12 };
      const uint64_t least[2] = { (uint64_t) lambda << 28, (uint64_t) 64 << 24 };
      for (int c = 0; c < 2; c++) {
        if (wide[c] % k || least[c] > n) continue;
        const int32_t rcw = fm_sparsify(idx, wide[c], lambda, lanes, true);
        if (rcw == FM_SUCCESS) return FM_SUCCESS;
      }
    }
    for (uint32_t cand = 10; cand >= 2 * k; cand--)
      if (cand % k == 0 && (((uint64_t) 64) << (2 * cand)) <= n) { ks = cand; break; }
    if (ks == 0) ks = 2 * k;
  }
  if (ks % k || ks <= k || ks > 14) return fm_fail_msg(FM_E_BAD_ARGUMENT, "sparse bases must be a multiple of k, larger than k and at most 14");
  const uint32_t nsym = 1u << (2 * ks), hops = ks / k, kbits = 2 * k;
  const uint64_t nrows = (uint64_t) idx->meta.nblocks * FM_SB_ROWS;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  const uint64_t est_blocks = (uint64_t) n / lambda + nsym;
  const uint64_t need = 16ull * n + nrows + est_blocks * bbytes + 32ull * nsym + (1ull << 30);
  if (need > free_b) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough free device memory to build the sparse-step table");
  if (!fm_budget_allows(idx, est_blocks * bbytes + 8ull * nsym)) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "the sparse-step table would exceed the derived-table budget");

  uint8_t *sym = NULL; uint32_t *keys = NULL, *rows = NULL, *keys2 = NULL, *rows2 = NULL, *symstart = NULL, *nb = NULL, *first = NULL, *rank0 = NULL;
  uint2 *dir = NULL; uint4 *sblocks = NULL; void *tmp = NULL; unsigned long long *d_novf = NULL;
  size_t tmp_bytes = 0, tmp2 = 0;
  uint64_t total_blocks = 0; unsigned long long novf = 0;
  cudaError_t e = cudaMalloc((void **) &sym, nrows);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rows, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys2, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rows2, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &symstart, 4ull * (nsym + 1));
  if (e == cudaSuccess) e = cudaMalloc((void **) &nb, 4ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &first, 4ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rank0, 4ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &dir, 8ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_novf, 8);
  if (e == cudaSuccess) e = cudaMemset(d_novf, 0, 8);
  if (e == cudaSuccess) {
    e = fm_row_symbols(idx, nrows, sym);
  }
  if (e == cudaSuccess) {
    fm_sparse_compose_kernel<<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(idx->blocks, idx->meta.nblocks, sym, n, kbits, hops, nsym, keys, rows);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, keys, keys2, rows, rows2, (int64_t) n, 0, (int)(2 * ks + 1));
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(NULL, tmp2, nb, first, (int) nsym);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, rows, rows2, (int64_t) n, 0, (int)(2 * ks + 1));
  if (e == cudaSuccess) {
    fm_sparse_symstart_kernel<<<(nsym + 1 + 255) / 256, 256>>>(keys2, n, nsym, symstart);
    e = cudaGetLastError();
  }
  /* uniform grid or per-symbol block counts?  $FMGPU_SPARSE_UNIFORM = 0 / 1 forces; default: uniform when no symbol occurs
   * more than 1.6 x as often as the mean (its blocks then expect <= 8 rows at lambda 5: < 1 % of THOSE symbols' blocks
   * overflow 15 slots, far cheaper than a directory lookup in every step) nor less than 0.4 x as often (wasted blocks) */
  uint32_t uni_nb = 0, uni_scale = 0;
  if (e == cudaSuccess) {
    uint32_t range[2] = { 0xFFFFFFFFu, 0u }, carrying = 0;
    uint32_t *d_range = first;                                 /* scratch: `first` is written by the scan below */
    e = cudaMemcpy(d_range, range, 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { fm_sparse_count_range_kernel<<<(nsym + 255) / 256, 256>>>(symstart, nsym, d_range); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(range, d_range, 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&carrying, symstart + nsym, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
      const double mean = (double) carrying / nsym;
      const char *env = getenv("FMGPU_SPARSE_UNIFORM");
      const uint64_t per = ((uint64_t) carrying + (uint64_t) nsym * lambda - 1) / ((uint64_t) nsym * lambda);
      /* rows living in symbols that cannot fit their share of the grid even if spread perfectly (count > slots x blocks) */
      unsigned long long heavy = 0, *d_heavy = d_novf;            /* (d_novf is zero here: the fill kernel runs later) */
      fm_sparse_heavy_rows_kernel<<<(nsym + 255) / 256, 256>>>(symstart, nsym, (uint32_t)(slots * (per ? per : 1)), d_heavy);
      e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaMemcpy(&heavy, d_heavy, 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemset(d_heavy, 0, 8);
      const bool even = mean >= lambda && range[1] <= 1.6 * mean && range[0] >= 0.4 * mean;          /* many rows per symbol: tight counts */
      const bool sparse_even = mean >= lambda && mean < 64 && heavy * 1000ull <= carrying;           /* few rows per symbol (Poisson scatter): no heavy tail */
      const bool want = env && *env ? atoi(env) != 0 : (even || sparse_even);
      if (want && per >= 1 && per * nsym < (1ull << 32)) {
        uni_nb = (uint32_t) per;
        unsigned long long sc = ((((unsigned long long) uni_nb) << 32) - 1ull) / n;    /* as fm_sparse_dir_kernel */
        uni_scale = sc > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) sc;
      }
    }
  }
  if (e == cudaSuccess && require_uniform && !uni_nb) {      /* the 12-base attempt of the automatic choice: counts are uneven */
    cudaFree(sym); cudaFree(keys); cudaFree(rows); cudaFree(keys2); cudaFree(rows2); cudaFree(symstart); cudaFree(nb); cudaFree(first);
    cudaFree(rank0); cudaFree(tmp); cudaFree(d_novf); cudaFree(dir);
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "symbol counts too uneven for a uniform grid");
  }
  if (e == cudaSuccess) {
    fm_sparse_nblocks_kernel<<<(nsym + 255) / 256, 256>>>(symstart, nsym, lambda, uni_nb, nb);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, nb, first, (int) nsym);
  if (e == cudaSuccess) {
    uint32_t last_first = 0, last_nb = 0;
    e = cudaMemcpy(&last_first, first + (nsym - 1), 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&last_nb, nb + (nsym - 1), 4, cudaMemcpyDeviceToHost);
    total_blocks = (uint64_t) last_first + last_nb;
  }
  /* the sort's input buffers are dead now: release them before the table is allocated */
  cudaFree(keys); keys = NULL; cudaFree(rows); rows = NULL; cudaFree(sym); sym = NULL;
  if (e == cudaSuccess && total_blocks >= (1ull << 32)) e = cudaErrorInvalidValue;
  if (e == cudaSuccess) e = cudaMalloc((void **) &sblocks, total_blocks * bbytes);
  if (e == cudaSuccess) {
    fm_sparse_dir_kernel<<<(nsym + 255) / 256, 256>>>(idx->blocks, idx->meta.nblocks, kbits, hops, nsym, n, nb, first, dir, rank0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    if (uni_nb && uni_nb < 32) {                               /* many symbols, few blocks each: one thread per block */
      const unsigned grid = (unsigned)((total_blocks + 255) / 256);
      if (lanes == 4) fm_sparse_fill_uniform_kernel<4><<<grid, 256>>>(rows2, symstart, nsym, uni_nb, uni_scale, rank0, sblocks, d_novf);
      else            fm_sparse_fill_uniform_kernel<2><<<grid, 256>>>(rows2, symstart, nsym, uni_nb, uni_scale, rank0, sblocks, d_novf);
    } else if (lanes == 4) fm_sparse_fill_kernel<4><<<nsym, 128>>>(rows2, symstart, dir, nb, rank0, sblocks, d_novf);
    else                   fm_sparse_fill_kernel<2><<<nsym, 128>>>(rows2, symstart, dir, nb, rank0, sblocks, d_novf);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(&novf, d_novf, 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(sym); cudaFree(keys); cudaFree(rows); cudaFree(keys2); cudaFree(rows2); cudaFree(symstart); cudaFree(nb); cudaFree(first);
  cudaFree(rank0); cudaFree(tmp); cudaFree(d_novf);
  if (e != cudaSuccess) {
    cudaFree(sblocks); cudaFree(dir);
    cudaGetLastError();                                          /* a failed cudaMalloc stays "last error" otherwise and fails the next attempt's first check */
    if (e == cudaErrorMemoryAllocation) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough device memory for the sparse-step table (the plain kernels still serve this index)");
    return fm_fail(e, "fmgpu_index_sparsify", __FILE__, __LINE__);
  }
  idx->sblocks = sblocks; idx->sdir = dir; idx->s_uni_nb = uni_nb; idx->s_uni_scale = uni_scale; idx->meta.sparse_uniform_nb = uni_nb;
  idx->meta.sparse_bases = ks; idx->meta.sparse_lambda = lambda; idx->meta.sparse_blocks = total_blocks;
  idx->meta.sparse_overflow = novf; idx->meta.sparse_bytes = total_blocks * bbytes + 8ull * nsym; idx->meta.sparse_lanes = lanes;

  /* start table: the sparse kernel itself searches every SB-mer once (a packed SB-mer IS its key); SB = the
   * largest whole number of sparse steps within 12 bases */
  {
    const char *env = getenv("FMGPU_START_TABLE");
    const bool want = env && *env ? atoi(env) != 0 : idx->meta.nbytes >= (1ull << 30);
    const uint32_t ssteps = 12 / ks, sb = ssteps * ks;
    if (want && ssteps && ((uint64_t) 1 << (2 * sb)) < n) {
      const uint32_t nkeys = 1u << (2 * sb);
      uint32_t *skeys = NULL; uint2 *table = NULL;
      e = cudaMalloc((void **) &skeys, (size_t) nkeys * 4);
      if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nkeys * 8);
      if (e == cudaSuccess) { fm_iota_kernel<<<(nkeys + 255) / 256, 256>>>(skeys, nkeys); e = cudaGetLastError(); }
      int32_t rc = FM_SUCCESS;
      if (e == cudaSuccess) rc = fm_launch_sparse(idx, skeys, nkeys, sb, (uint32_t *) table, FM_DEFAULT_VARIANT, 0, NULL, false);
      if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
      cudaFree(skeys);
      if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); }   /* the table is optional */
      else { idx->sstart = table; idx->meta.sparse_start_bases = sb; idx->meta.sparse_bytes += (uint64_t) nkeys * 8; }
    }
    /* lead tables (fm_ensure_lead): the small ones now, the wide ones (12 .. 15 bases, up to 8.6 GB) when a read length asks */
    idx->stables = want ? 1 : 0;
    if (want)
      for (uint32_t b = 1; b <= ks + 1 && b < 12; b++) fm_build_lead(idx, b);
  }
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

int32_t fm_launch_sparse(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                         uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters, bool use_lead_tables)
{
  if (!idx->sblocks) return fm_fail_msg(FM_E_BAD_ARGUMENT, "FMGPU_MODE_SPARSE needs fmgpu_index_sparsify() on this replica first");
  const uint32_t k = idx->meta.steps, ks = idx->meta.sparse_bases, hops = ks / k, lanes = idx->meta.sparse_lanes;
  if (v.queries_per_thread < 1 || v.queries_per_thread > 4) v.queries_per_thread = 4;
  FmSparseParams p;
  p.sblocks = idx->sblocks; p.dir = idx->sdir; p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results;
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
  p.sbits = 2 * ks; p.hops = hops;
  p.uni_nb = idx->s_uni_nb; p.uni_scale = idx->s_uni_scale;
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

/*
 * fm_fusedtab.cu -- fused-step table (fm_fused.cuh): construction on the GPU, launch, fetch counter.
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 */
#include "fm_internal.h"
#include "fm_fused.cuh"

/* ------------------------------------------------------------------------ *
 * fused-step table (fm_fused.cuh)
 * ------------------------------------------------------------------------ */
static uint32_t fm_fused_rows(uint32_t lanes) { return 32u * (8u * lanes - 1u); }

cudaError_t fm_table_symbols(const uint4 *table, uint32_t nblocks, uint32_t nsym, uint64_t nrows_alloc, uint8_t *d_sym)
{
  fm_fuse_symbols_kernel<<<(nblocks + 127) / 128, 128>>>(table, nblocks, nsym, nrows_alloc, d_sym);
  return cudaGetLastError();
}
cudaError_t fm_row_symbols(const fmgpu_index_t *idx, uint64_t nrows_alloc, uint8_t *d_sym)
{
  return fm_table_symbols(idx->blocks, idx->meta.nblocks, idx->meta.nsymbols, nrows_alloc, d_sym);
}

template <int LANES>
static cudaError_t fm_fuse_build(const fmgpu_index_t *idx, const uint16_t *fsym, uint32_t nfsym, uint32_t nfb, uint32_t kbits,
                                 uint32_t hops, uint4 *fblocks)
{
  fm_fuse_write_kernel<LANES><<<nfb, 256, 256 * 8 * LANES * 4>>>(fsym, nfsym, nfb, fblocks);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fm_fuse_scan_kernel<LANES><<<nfsym, 1024>>>(idx->blocks, idx->meta.nblocks, kbits, hops, nfb, idx->meta.quirk_start, idx->meta.quirk_mask, fblocks);
  return cudaGetLastError();
}

extern "C" int32_t fmgpu_index_fuse(fmgpu_index_t *idx, uint32_t fused_bases, uint32_t lanes, uint64_t budget_bytes)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->fblocks) return FM_SUCCESS;
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t k = idx->meta.steps;
  if (lanes == 0) lanes = 2;
  if (!(lanes == 1 || lanes == 2 || lanes == 4)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "fused block lanes must be 1, 2 or 4");
  if (budget_bytes == 0) {
    const char *env = getenv("FMGPU_FUSE_BUDGET_GB");
    budget_bytes = (uint64_t)((env && *env ? atof(env) : 69.0) * 1e9);       /* flat part of the footprint curve ends near 68-70 GB */
  }
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  const uint32_t rows = fm_fused_rows(lanes);
  const uint32_t nfb = idx->meta.bwtsize / rows + 1;
  uint32_t kf = 0;
  for (uint32_t cand = 4; cand > k; cand--) {
    if (fused_bases && cand != fused_bases) continue;
    if (cand % k) continue;
    const uint64_t bytes = ((uint64_t) 1 << (2 * cand)) * nfb * 32ull * lanes;
    const uint64_t scratch = 3ull * idx->meta.bwtsize + (1ull << 30);       /* construction: 3 bytes per row, plus slack */
    if (bytes + scratch > free_b) continue;                                  /* does not fit in HBM right now */
    if (!fm_budget_allows(idx, bytes)) continue;                             /* derived-table budget of the replica */
    if (bytes <= budget_bytes || fused_bases) { kf = cand; break; }
  }
  if (!kf) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "no fused-step table fits the memory budget for this index (or k already is the requested width)");
  const uint32_t nfsym = 1u << (2 * kf), hops = kf / k, kbits = 2 * k;
  const uint64_t fbytes = (uint64_t) nfsym * nfb * 32ull * lanes;
  uint64_t nrows = (uint64_t) idx->meta.nblocks * FM_SB_ROWS;
  if ((uint64_t) nfb * rows > nrows) nrows = (uint64_t) nfb * rows;
  uint8_t *sym = NULL; uint16_t *fsym = NULL; uint4 *fblocks = NULL;
  cudaError_t e = cudaMalloc((void **) &fblocks, fbytes);
  if (e == cudaSuccess) e = cudaMalloc((void **) &sym, nrows);
  if (e == cudaSuccess) e = cudaMalloc((void **) &fsym, nrows * 2);
  if (e == cudaSuccess) {
    e = fm_row_symbols(idx, nrows, sym);
  }
  /* AltCounters padding quirk: the chains follow the quirked rank; the phantom occurrences are listed beside the bitmaps */
  const uint32_t qstart = idx->meta.quirk_start, qmask = idx->meta.quirk_mask, max_visits = 1024;
  FmQuirkVisit *visits = NULL; uint32_t *d_cnt = NULL, *ph_keys = NULL, *ph_rows = NULL, hcnt[2] = { 0, 0 };
  fm_phantoms ph; memset(&ph, 0, sizeof ph);
  if (e == cudaSuccess) e = cudaMalloc((void **) &visits, sizeof(FmQuirkVisit) * max_visits);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_cnt, 8);
  if (e == cudaSuccess) e = cudaMemset(d_cnt, 0, 8);
  if (e == cudaSuccess) {
    fm_fuse_compose_kernel<<<(unsigned)((nrows + 255) / 256), 256>>>(idx->blocks, idx->meta.nblocks, sym, idx->meta.bwtsize, kbits, hops, nrows,
                                                                       qstart, qmask, visits, d_cnt, max_visits, fsym);
    e = cudaGetLastError();
  }
  bool too_many = false;
  if (e == cudaSuccess && qmask && qstart != 0u) {
    e = cudaMemcpy(hcnt, d_cnt, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && hcnt[0] > max_visits) too_many = true;
    if (e == cudaSuccess && hcnt[0] && !too_many) {
      e = cudaMalloc((void **) &ph_keys, 4 * 256);
      if (e == cudaSuccess) e = cudaMalloc((void **) &ph_rows, 4 * 256);
      if (e == cudaSuccess) {
        fm_quirk_phantoms_kernel<<<1, 1>>>(idx->blocks, idx->meta.nblocks, sym, kbits, hops, qstart, qmask, visits, hcnt[0], ph_keys, ph_rows, 256, d_cnt + 1);
        e = cudaGetLastError();
      }
      if (e == cudaSuccess) e = cudaMemcpy(hcnt + 1, d_cnt + 1, 4, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess && hcnt[1] > FM_MAX_PHANTOMS) too_many = true;
      if (e == cudaSuccess && !too_many && hcnt[1]) {
        ph.n = hcnt[1];
        e = cudaMemcpy(ph.sym, ph_keys, 4 * ph.n, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(ph.row, ph_rows, 4 * ph.n, cudaMemcpyDeviceToHost);
      }
    }
  }
  cudaFree(visits); cudaFree(d_cnt); cudaFree(ph_keys); cudaFree(ph_rows);
  if (e == cudaSuccess && too_many) {
    cudaFree(sym); cudaFree(fsym); cudaFree(fblocks);
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "too many phantom occurrences for the fused-step table (AltCounters padding quirk); the sparse-step table serves this index");
  }
  if (e == cudaSuccess) e = lanes == 1 ? fm_fuse_build<1>(idx, fsym, nfsym, nfb, kbits, hops, fblocks)
                          : lanes == 2 ? fm_fuse_build<2>(idx, fsym, nfsym, nfb, kbits, hops, fblocks)
                                       : fm_fuse_build<4>(idx, fsym, nfsym, nfb, kbits, hops, fblocks);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(sym); cudaFree(fsym);
  if (e != cudaSuccess) {
    cudaFree(fblocks);
    cudaGetLastError();
    if (e == cudaErrorMemoryAllocation) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough device memory for the fused-step table (the plain kernels still serve this index)");
    return fm_fail(e, "fmgpu_index_fuse", __FILE__, __LINE__);
  }
  idx->fblocks = fblocks; idx->nfblocks = nfb; idx->fphantoms = ph;
  idx->meta.fused_bases = kf; idx->meta.fused_lanes = lanes; idx->meta.fused_bytes = fbytes;

  /* start table: the fused kernel itself searches all 4^12 12-mers once (a packed 12-mer IS its 24-bit key);
   * only worth it when the table (134 MB) is small next to the index, and FM_START_BASES must be whole fused steps */
  {
    const char *env = getenv("FMGPU_START_TABLE");
    const bool want = env && *env ? atoi(env) != 0 : idx->meta.nbytes >= (1ull << 30);
    if (want && FM_START_BASES % kf == 0 && idx->meta.bwtsize > (1u << 24)) {
      const uint32_t nkeys = 1u << (2 * FM_START_BASES);
      uint32_t *keys = NULL; uint2 *table = NULL;
      e = cudaMalloc((void **) &keys, (size_t) nkeys * 4);
      if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nkeys * 8);
      if (e == cudaSuccess) { fm_iota_kernel<<<(nkeys + 255) / 256, 256>>>(keys, nkeys); e = cudaGetLastError(); }
      int32_t rc = FM_SUCCESS;
      if (e == cudaSuccess) rc = fm_launch_fused(idx, keys, nkeys, FM_START_BASES, (uint32_t *) table, FM_DEFAULT_VARIANT, 0, NULL);
      if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
      cudaFree(keys);
      if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); }   /* the table is optional */
      else { idx->start = table; idx->meta.start_bases = FM_START_BASES; }
    }
  }
  fm_budget_account(idx);
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_unfuse(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->fblocks || idx->start) { CU_TRY(cudaSetDevice(idx->device)); cudaFree(idx->fblocks); cudaFree(idx->start); idx->fblocks = NULL; idx->start = NULL; }
  memset(&idx->fphantoms, 0, sizeof idx->fphantoms);
  idx->nfblocks = 0; idx->meta.fused_bases = 0; idx->meta.fused_lanes = 0; idx->meta.fused_bytes = 0; idx->meta.start_bases = 0;
  return FM_SUCCESS;
}

typedef void (*fm_fused_fn)(const FmFusedParams);

template <int KF, int K, int LANES>
static fm_fused_fn fm_pick_fused_q(int qpt)
{
  if (qpt == 0) return fm_search_fused_kernel<KF, K, LANES, 1, 256, 6, true>;      /* instrumented */
  if (qpt == 1) return fm_search_fused_kernel<KF, K, LANES, 1, 256, 6, false>;
  if (qpt == 2) return fm_search_fused_kernel<KF, K, LANES, 2, 256, 4, false>;
  return NULL;
}
template <int KF, int K>
static fm_fused_fn fm_pick_fused_l(int lanes, int qpt)
{
  if (lanes == 1) return fm_pick_fused_q<KF, K, 1>(qpt);
  if (lanes == 2) return fm_pick_fused_q<KF, K, 2>(qpt);
  if (lanes == 4) return fm_pick_fused_q<KF, K, 4>(qpt);
  return NULL;
}
static fm_fused_fn fm_pick_fused(uint32_t kf, uint32_t k, int lanes, int qpt)
{
  if (kf == 4 && k == 2) return fm_pick_fused_l<4, 2>(lanes, qpt);
  if (kf == 4 && k == 1) return fm_pick_fused_l<4, 1>(lanes, qpt);
  if (kf == 3 && k == 1) return fm_pick_fused_l<3, 1>(lanes, qpt);
  if (kf == 2 && k == 1) return fm_pick_fused_l<2, 1>(lanes, qpt);
  return NULL;
}

int32_t fm_launch_fused(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                        uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters)
{
  if (!idx->fblocks) return fm_fail_msg(FM_E_BAD_ARGUMENT, "FMGPU_MODE_FUSED needs fmgpu_index_fuse() on this replica first");
  const uint32_t k = idx->meta.steps, kf = idx->meta.fused_bases, lanes = idx->meta.fused_lanes, hops = kf / k;
  if (v.queries_per_thread != 1 && v.queries_per_thread != 2) v.queries_per_thread = 2;
  FmFusedParams p;
  p.fblocks = idx->fblocks; p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results;
  p.nfblocks = idx->nfblocks; p.nblocks = idx->meta.nblocks; p.nq = (uint32_t) nq;
  p.nlead = (len / k) % hops; p.nfused = (len / k) / hops;
  p.wpq = fmgpu_words_per_query(len); p.wpq_pad = (p.wpq + 1) | 1u; p.bwtsize = idx->meta.bwtsize;
  p.fetch_counters = d_counters;
  p.start = idx->start; p.start_steps = idx->start ? FM_START_BASES / kf : 0u;
  p.quirk_start = idx->meta.quirk_start; p.quirk_mask = idx->meta.quirk_mask; p.nph = idx->fphantoms.n;
  for (uint32_t j = 0; j < FM_MAX_FUSED_PHANTOMS; j++) { p.ph_sym[j] = j < p.nph ? idx->fphantoms.sym[j] : 0u; p.ph_row[j] = j < p.nph ? idx->fphantoms.row[j] : 0u; }
  p.has_tail = len % k; p.tail_row = idx->meta.tail_row; p.tail_base = idx->meta.tail_base;
  for (int c = 0; c < 4; c++) p.tail_const[c] = idx->meta.tail_const[c];
  p.tail1 = p.has_tail ? idx->tail1 : NULL;
  if (d_counters) v.queries_per_thread = 1;
  uint32_t qper; size_t smem;
  for (;;) {
    qper = (256 / lanes) * v.queries_per_thread;
    smem = 16 + ((size_t) qper * p.wpq + 4) * 4;             /* mbarrier + reads + one readable spare word */
    if (smem <= 200 * 1024) break;
    if (v.queries_per_thread > 1) v.queries_per_thread = 1;
    else return fm_fail_msg(FM_E_QUERY_SHAPE, "reads too long to stage in shared memory");
  }
  fm_fused_fn fn = fm_pick_fused(kf, k, (int) lanes, d_counters ? 0 : v.queries_per_thread);
  if (!fn) return fm_fail_msg(FM_E_BAD_ARGUMENT, "no fused kernel for this (fused bases, k, lanes)");
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const uint32_t grid = (uint32_t)((nq + qper - 1) / qper);
  void *args[] = { (void *) &p };
  CU_TRY(cudaLaunchKernel((const void *) fn, dim3(grid), dim3(256), args, smem, stream));
  return FM_SUCCESS;
}
/* fused-step fetch counter: blocks of the fused table and SB96 blocks of the leading steps that one search must fetch */
extern "C" int32_t fmgpu_count_fetches_fused_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                                    uint32_t *d_results, void *stream, uint64_t *nfused_blocks, uint64_t *nlead_blocks)
{
  if (!idx || !d_packed || !d_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (len == 0 || (len % idx->meta.steps && !idx->meta.tail_valid)) return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a positive multiple of k");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long *d_c = NULL, h[2] = { 0, 0 };
  CU_TRY(cudaMalloc((void **) &d_c, 16));
  CU_TRY(cudaMemsetAsync(d_c, 0, 16, (cudaStream_t) stream));
  int32_t rc = nq ? fm_launch_fused(idx, d_packed, nq, len, d_results, FM_DEFAULT_VARIANT, (cudaStream_t) stream, d_c) : FM_SUCCESS;
  if (rc == FM_SUCCESS) {
    cudaError_t e = cudaMemcpyAsync(h, d_c, 16, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) rc = fm_fail(e, "fetch counters D2H", __FILE__, __LINE__);
  }
  cudaFree(d_c);
  if (nfused_blocks) *nfused_blocks = h[0];
  if (nlead_blocks) *nlead_blocks = h[1];
  return rc;
}

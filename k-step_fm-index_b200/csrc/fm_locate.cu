/*
 * fm_locate.cu -- locate (fm_locate.cuh): suffix array derived from the replica's own table, SA[L..R) gathers.
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 */
#include "fm_internal.h"
#include <cub/device/device_scan.cuh>
#include "fm_locate.cuh"

/* ------------------------------------------------------------------------ *
 * locate (fm_locate.cuh): suffix array derived from the replica's own table, SA[L..R) gathers
 * ------------------------------------------------------------------------ */
/* the 4-symbol table that holds rank1 at block starts + per-row char bits: SB96 itself (k = 1) or the tail table (k = 2;
 * a private copy in *tmp when the replica keeps none).  The stored ranks are quirk-free, so this is the text's own 1-step
 * index also for AltCounters files with an active padding quirk. */
static int32_t fm_locate_table(fmgpu_index_t *idx, const uint4 **t1, uint4 **tmp)
{
  *t1 = idx->blocks; *tmp = NULL;
  if (idx->meta.steps == 2) {
    if (!idx->tail_consts_ok) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "this 2-step index has no derived 1-step rank");
    *t1 = idx->meta.tail_valid ? fm_build_tail(idx) : NULL;
    if (!*t1) {                                                 /* $FMGPU_TAIL_TABLE=0, no room, or a quirk file: a private copy for this build */
      CU_TRY(cudaMalloc((void **) tmp, (size_t) 4 * idx->meta.nblocks * sizeof(uint4)));
      if (fm_tail_table_into(idx, *tmp) != cudaSuccess) { cudaFree(*tmp); *tmp = NULL; return fm_fail_msg(FM_E_CUDA, "fm_tail_table_kernel"); }
      *t1 = *tmp;
    }
  }
  return FM_SUCCESS;
}

/* full suffix array of the indexed text, derived from the table (fm_locate.cuh): *sa_out holds bwtsize words; *norow =
 * the row without a BWT character (text position 0) */
static int32_t fm_derive_sa(fmgpu_index_t *idx, const uint4 *t1, uint32_t **sa_out, uint32_t *norow)
{
  const uint32_t n = idx->meta.bwtsize, nb = idx->meta.nblocks;
  uint2 *na = NULL, *nbuf = NULL; uint32_t *sa = NULL, *d_term = NULL; unsigned long long *d_status = NULL;
  uint32_t term[2] = { 0xFFFFFFFFu, 0u }; unsigned long long bad = 0;
  cudaError_t e = cudaMalloc((void **) &na, 8ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &nbuf, 8ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &sa, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_term, 8);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_status, 8);
  if (e == cudaSuccess) e = cudaMemcpy(d_term, term, 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(d_status, 0, 8);
  if (e == cudaSuccess) e = cudaMemset(na, 0xFF, 8ull * n);     /* a row the table does not cover would point nowhere valid */
  if (e == cudaSuccess) { fm_locate_lf_kernel<<<(nb + 127) / 128, 128>>>(t1, nb, n, na, d_term); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(term, d_term, 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && term[1] != 1u) {
    cudaFree(na); cudaFree(nbuf); cudaFree(sa); cudaFree(d_term); cudaFree(d_status);
    return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "index table is inconsistent: expected exactly one row without a BWT character");
  }
  if (e == cudaSuccess) {
    uint32_t rounds = 0;
    while (rounds < 32 && (1ull << rounds) < n) rounds++;
    for (uint32_t r = 0; r < rounds && e == cudaSuccess; r++) {
      fm_locate_jump_kernel<<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(na, nbuf, n);
      e = cudaGetLastError();
      uint2 *t = na; na = nbuf; nbuf = t;
    }
  }
  if (e == cudaSuccess) { fm_locate_extract_kernel<<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(na, n, term[0], sa, d_status); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(&bad, d_status, 8, cudaMemcpyDeviceToHost);
  cudaFree(na); cudaFree(nbuf); cudaFree(d_term); cudaFree(d_status);
  if (e != cudaSuccess) { cudaFree(sa); cudaGetLastError(); return fm_fail(e, "suffix array derivation", __FILE__, __LINE__); }
  if (bad) { cudaFree(sa); return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "index table is inconsistent: its LF mapping is not a single cycle"); }
  *sa_out = sa; *norow = term[0];
  return FM_SUCCESS;
}

static int32_t fm_build_sa_common(fmgpu_index_t *idx, uint32_t rate)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (rate < 1 || rate > 4096) return fm_fail_msg(FM_E_BAD_ARGUMENT, "suffix array sampling rate must be 1 .. 4096");
  if (idx->sa && idx->sa_rate == rate) return FM_SUCCESS;
  if (idx->sa) { const int32_t rc = fmgpu_index_drop_sa(idx); if (rc) return rc; }
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t n = idx->meta.bwtsize, nb = idx->meta.nblocks;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  if (20ull * n + 64ull * nb + (256ull << 20) > free_b) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough free device memory to derive the suffix array");
  const uint32_t nlb = (uint32_t)(((uint64_t) n + FM_LOC_ROWS - 1) / FM_LOC_ROWS);
  const uint64_t keep = rate == 1 ? 4ull * n : 64ull * nlb + 4ull * nlb + 4ull * ((uint64_t) n / rate + 2);
  if (!fm_budget_allows(idx, keep)) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "the suffix array would exceed the derived-table budget");
  const uint4 *t1 = NULL; uint4 *t1_tmp = NULL;
  int32_t rc = fm_locate_table(idx, &t1, &t1_tmp);
  if (rc) return rc;
  uint32_t *sa = NULL, norow = 0;
  rc = fm_derive_sa(idx, t1, &sa, &norow);
  if (rc) { cudaFree(t1_tmp); return rc; }
  if (rate == 1) {
    cudaFree(t1_tmp);
    idx->sa = sa; idx->sa_rate = 1; idx->sa_norow = norow; idx->meta.sa_bytes = 4ull * n; idx->meta.sa_rate = 1;
    fm_budget_account(idx);
    return FM_SUCCESS;
  }
  /* sampled: the walk table (64 bytes per 128 rows), the marks' prefix counts, the samples */
  uint8_t *sym = NULL; uint4 *lblocks = NULL; uint32_t *nmarks = NULL, *markrank = NULL, *samples = NULL; void *tmp = NULL; size_t tmp_bytes = 0;
  uint32_t last[2] = { 0, 0 };
  const uint64_t nrows = (uint64_t) nb * FM_SB_ROWS;
  cudaError_t e = cudaMalloc((void **) &sym, nrows);
  if (e == cudaSuccess) e = cudaMalloc((void **) &lblocks, 64ull * nlb);
  if (e == cudaSuccess) e = cudaMalloc((void **) &nmarks, 4ull * nlb);
  if (e == cudaSuccess) e = cudaMalloc((void **) &markrank, 4ull * nlb);
  if (e == cudaSuccess) e = fm_table_symbols(t1, nb, 4, nrows, sym);
  if (e == cudaSuccess) { fm_locate_pack_kernel<<<(nlb + 127) / 128, 128>>>(t1, nb, sym, sa, n, rate, nlb, lblocks, nmarks); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(NULL, tmp_bytes, nmarks, markrank, (int64_t) nlb);
  if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, nmarks, markrank, (int64_t) nlb);
  if (e == cudaSuccess) e = cudaMemcpy(&last[0], markrank + (nlb - 1), 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&last[1], nmarks + (nlb - 1), 4, cudaMemcpyDeviceToHost);
  const uint64_t nsamples = (uint64_t) last[0] + last[1];
  if (e == cudaSuccess) e = cudaMalloc((void **) &samples, 4ull * (nsamples ? nsamples : 1));
  if (e == cudaSuccess) { fm_locate_samples_kernel<<<(nlb + 127) / 128, 128>>>(lblocks, markrank, sa, n, nlb, samples); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(sym); cudaFree(nmarks); cudaFree(tmp); cudaFree(sa); cudaFree(t1_tmp);
  if (e != cudaSuccess) { cudaFree(lblocks); cudaFree(markrank); cudaFree(samples); cudaGetLastError(); return fm_fail(e, "fmgpu_index_build_sa_sampled", __FILE__, __LINE__); }
  idx->sa = samples; idx->sa_marks = (uint32_t *) lblocks; idx->sa_markrank = markrank; idx->sa_rate = rate; idx->sa_norow = norow; idx->sa_nlb = nlb;
  idx->meta.sa_bytes = 64ull * nlb + 4ull * nlb + 4ull * nsamples; idx->meta.sa_rate = rate;
  fm_budget_account(idx);
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_build_sa(fmgpu_index_t *idx) { return fm_build_sa_common(idx, 1); }
extern "C" int32_t fmgpu_index_build_sa_sampled(fmgpu_index_t *idx, uint32_t rate) { return fm_build_sa_common(idx, rate ? rate : 32u); }

extern "C" int32_t fmgpu_index_drop_sa(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sa) {
    CU_TRY(cudaSetDevice(idx->device));
    cudaFree(idx->sa); cudaFree(idx->sa_marks); cudaFree(idx->sa_markrank);
    idx->sa = NULL; idx->sa_marks = NULL; idx->sa_markrank = NULL;
  }
  idx->meta.sa_bytes = 0; idx->meta.sa_rate = 0; idx->sa_rate = 0;
  fm_budget_account(idx);
  return FM_SUCCESS;
}

extern "C" void *fmgpu_index_sa(const fmgpu_index_t *idx) { return idx && idx->sa_rate == 1 ? (void *) idx->sa : NULL; }

extern "C" int32_t fmgpu_locate_device(const fmgpu_index_t *idx, const uint32_t *d_results, uint64_t nq, uint32_t max_hits,
                                       uint32_t *d_positions, uint32_t *d_nhits, void *stream)
{
  if (!idx || !d_results || !d_positions || max_hits == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  if (!idx->sa) return fm_fail_msg(FM_E_BAD_ARGUMENT, "locate needs fmgpu_index_build_sa() on this replica first");
  if (nq == 0) return FM_SUCCESS;
  const uint64_t total = nq * max_hits;
  if (total >= (1ull << 39) - 256) return fm_fail_msg(FM_E_BAD_ARGUMENT, "too many (read, hit) slots in one launch; shard the batch");
  CU_TRY(cudaSetDevice(idx->device));
  if (idx->sa_rate == 1)
    fm_locate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t) stream>>>(idx->sa, (const uint2 *) d_results, nq, max_hits, d_positions, d_nhits);
  else                                                          /* sampled: a lane pair per (read, hit) walks LF to a marked row */
    fm_locate_sampled_kernel<<<(unsigned)((2 * total + 255) / 256), 256, 0, (cudaStream_t) stream>>>((const uint4 *) idx->sa_marks, idx->sa_markrank, idx->sa,
                                                                                                      (const uint2 *) d_results, nq, max_hits, idx->meta.bwtsize,
                                                                                                      idx->sa_rate, idx->sa_norow, d_positions, d_nhits);
  CU_TRY(cudaGetLastError());
  return FM_SUCCESS;
}
/* locate for a shard: its (L,R) -> positions and hit counts in host memory */
extern "C" int32_t fmgpu_batch_locate(const fmgpu_index_t *idx, fmgpu_batch_t *b, uint32_t max_hits, uint32_t *h_positions, uint32_t *h_nhits)
{
  if (!idx || !b || !h_positions || max_hits == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  if (b->nq == 0) return FM_SUCCESS;
  CU_TRY(cudaSetDevice(b->device));
  uint32_t *d_pos = NULL, *d_n = NULL;
  CU_TRY(cudaMalloc((void **) &d_pos, b->nq * max_hits * 4ull));
  cudaError_t e = cudaMalloc((void **) &d_n, b->nq * 4ull);
  int32_t rc = FM_SUCCESS;
  if (e != cudaSuccess) rc = fm_fail(e, "cudaMalloc(hit counts)", __FILE__, __LINE__);
  if (rc == FM_SUCCESS) rc = fmgpu_locate_device(idx, b->d_results, b->nq, max_hits, d_pos, d_n, b->stream);
  if (rc == FM_SUCCESS) {
    e = cudaMemcpyAsync(h_positions, d_pos, b->nq * max_hits * 4ull, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess && h_nhits) e = cudaMemcpyAsync(h_nhits, d_n, b->nq * 4ull, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    if (e != cudaSuccess) rc = fm_fail(e, "locate D2H", __FILE__, __LINE__);
  }
  cudaFree(d_pos); cudaFree(d_n);
  return rc;
}

extern "C" int32_t fmgpu_index_download_sa(const fmgpu_index_t *idx, uint32_t *h_sa)
{
  if (!idx || !h_sa) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (!idx->sa || idx->sa_rate != 1) return fm_fail_msg(FM_E_BAD_ARGUMENT, "no full suffix array on this replica (fmgpu_index_build_sa)");
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(cudaMemcpy(h_sa, idx->sa, 4ull * idx->meta.bwtsize, cudaMemcpyDeviceToHost));
  return FM_SUCCESS;
}

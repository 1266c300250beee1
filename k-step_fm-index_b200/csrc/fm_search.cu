/*
 * fm_search.cu -- kernel dispatch of the plain Task / Coop kernels, read packing, query shards (fmgpu_batch_*).
 * Replaces searchIndexGPU's launch code of the six reference .cu files (e.g. src/fmIndexGPU-Coop-2Step.cu:231-248).
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 */
#include "fm_internal.h"
#include "fm_kernels.cuh"

/* ------------------------------------------------------------------------ *
 * kernel dispatch
 * ------------------------------------------------------------------------ */
typedef void (*fm_kernel_fn)(const FmSearchParams);

template <int K, int QPT, int THREADS, int MINB>
static fm_kernel_fn fm_pick_task(bool quirk, bool count)
{
  if (count) return quirk ? fm_search_task_kernel<K, QPT, THREADS, MINB, true, true> : fm_search_task_kernel<K, QPT, THREADS, MINB, false, true>;
  return quirk ? fm_search_task_kernel<K, QPT, THREADS, MINB, true, false> : fm_search_task_kernel<K, QPT, THREADS, MINB, false, false>;
}
template <int K, int QPT, int THREADS, int MINB>
static fm_kernel_fn fm_pick_coop(bool quirk)
{
  return quirk ? fm_search_coop_kernel<K, QPT, THREADS, MINB, true> : fm_search_coop_kernel<K, QPT, THREADS, MINB, false>;
}

/* register budgets: QPT=1 -> 32 regs (2048 thr/SM), QPT=2 -> 40, QPT=4 -> 64 */
template <int K>
static fm_kernel_fn fm_pick(int mode, int qpt, int tpb, bool quirk, bool count)
{
  if (mode == FMGPU_MODE_TASK) {
    if (qpt == 1 && tpb == 128) return fm_pick_task<K, 1, 128, 16>(quirk, count);
    if (qpt == 1 && tpb == 256) return fm_pick_task<K, 1, 256, 8>(quirk, count);
    if (qpt == 1 && tpb == 512) return fm_pick_task<K, 1, 512, 4>(quirk, count);
    if (qpt == 2 && tpb == 128) return fm_pick_task<K, 2, 128, 12>(quirk, count);
    if (qpt == 2 && tpb == 256) return fm_pick_task<K, 2, 256, 6>(quirk, count);
    if (qpt == 2 && tpb == 512) return fm_pick_task<K, 2, 512, 3>(quirk, count);
    if (qpt == 4 && tpb == 128) return fm_pick_task<K, 4, 128, 8>(quirk, count);
    if (qpt == 4 && tpb == 256) return fm_pick_task<K, 4, 256, 4>(quirk, count);
    if (qpt == 4 && tpb == 512) return fm_pick_task<K, 4, 512, 2>(quirk, count);
  } else if (mode == FMGPU_MODE_COOP && !count) {
    if (qpt == 1 && tpb == 128) return fm_pick_coop<K, 1, 128, 16>(quirk);
    if (qpt == 1 && tpb == 256) return fm_pick_coop<K, 1, 256, 8>(quirk);
    if (qpt == 1 && tpb == 512) return fm_pick_coop<K, 1, 512, 4>(quirk);
    if (qpt == 2 && tpb == 128) return fm_pick_coop<K, 2, 128, 16>(quirk);
    if (qpt == 2 && tpb == 256) return fm_pick_coop<K, 2, 256, 8>(quirk);
    if (qpt == 2 && tpb == 512) return fm_pick_coop<K, 2, 512, 4>(quirk);
    if (qpt == 4 && tpb == 128) return fm_pick_coop<K, 4, 128, 12>(quirk);
    if (qpt == 4 && tpb == 256) return fm_pick_coop<K, 4, 256, 6>(quirk);
    if (qpt == 4 && tpb == 512) return fm_pick_coop<K, 4, 512, 3>(quirk);
  }
  return NULL;
}

int32_t fm_launch_search(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                         uint32_t *d_results, const fmgpu_variant_t *vin, cudaStream_t stream,
                         unsigned long long *d_counters)
{
  if (!idx || !d_packed || !d_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  const uint32_t k = idx->meta.steps;
  if (len == 0 || (len % k && !idx->meta.tail_valid))
    return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a positive multiple of k (undefined in the reference, SURVEY.md App. C-5; "
                                         "odd lengths are served on 2-step indexes without the AltCounters quirk only)");
  if (nq == 0) return FM_SUCCESS;
  if (nq >= (1ull << 31)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "more than 2^31 reads in one launch; shard the batch");
  fmgpu_variant_t v = vin ? *vin : FM_DEFAULT_VARIANT;
  if (v.queries_per_thread == 0) v.queries_per_thread = FM_DEFAULT_VARIANT.queries_per_thread;
  if (v.threads_per_block == 0) v.threads_per_block = FM_DEFAULT_VARIANT.threads_per_block;
  const bool count = d_counters != NULL;
  if (count) { v.mode = FMGPU_MODE_TASK; v.queries_per_thread = 1; v.threads_per_block = 256; }
  if (v.mode == FMGPU_MODE_FUSED) return fm_launch_fused(idx, d_packed, nq, len, d_results, vin ? *vin : FM_DEFAULT_VARIANT, stream, NULL);
  if (v.mode == FMGPU_MODE_SPARSE) return fm_launch_sparse(idx, d_packed, nq, len, d_results, vin ? *vin : FM_DEFAULT_VARIANT, stream, NULL, true);
  if (v.mode == FMGPU_MODE_WIDE) {
    fmgpu_variant_t w = vin ? *vin : FM_DEFAULT_VARIANT;
    if (!vin) w.queries_per_thread = 0;
    return fm_launch_wide(idx, d_packed, nq, len, d_results, w, stream, NULL);
  }

  FmSearchParams p;
  p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results; p.fetch_counters = d_counters;
  p.nblocks = idx->meta.nblocks; p.nq = (uint32_t) nq; p.nsteps = len / k;
  p.wpq = fmgpu_words_per_query(len); p.wpq_pad = p.wpq | 1u;
  p.bwtsize = idx->meta.bwtsize; p.quirk_start = idx->meta.quirk_start; p.quirk_mask = idx->meta.quirk_mask;
  p.has_tail = len % k; p.tail_row = idx->meta.tail_row; p.tail_base = idx->meta.tail_base;
  for (int c = 0; c < 4; c++) p.tail_const[c] = idx->meta.tail_const[c];
  p.tail1 = p.has_tail ? idx->tail1 : NULL;
  const bool quirk = idx->meta.quirk_mask != 0;

  /* shrink the CTA's read count until the staged reads fit in shared memory */
  uint32_t qper; size_t smem;
  for (;;) {
    qper = (v.mode == FMGPU_MODE_COOP ? v.threads_per_block / 2 : v.threads_per_block) * v.queries_per_thread;
    smem = (size_t) qper * p.wpq_pad * 4;
    if (smem <= 200 * 1024) break;
    if (v.queries_per_thread > 1) v.queries_per_thread /= 2;
    else if (v.threads_per_block > 128) v.threads_per_block /= 2;
    else return fm_fail_msg(FM_E_QUERY_SHAPE, "reads too long to stage in shared memory");
  }
  fm_kernel_fn fn = (k == 1) ? fm_pick<1>(v.mode, v.queries_per_thread, v.threads_per_block, quirk, count)
                             : fm_pick<2>(v.mode, v.queries_per_thread, v.threads_per_block, quirk, count);
  if (!fn) return fm_fail_msg(FM_E_BAD_ARGUMENT, "unsupported kernel variant (mode 0/1, queries_per_thread 1/2/4, threads_per_block 128/256/512)");
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const uint32_t grid = (uint32_t)((nq + qper - 1) / qper);
  void *args[] = { (void *) &p };
  CU_TRY(cudaLaunchKernel((const void *) fn, dim3(grid), dim3(v.threads_per_block), args, smem, stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_search_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                       uint32_t *d_results, const fmgpu_variant_t *v, void *stream)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  CU_TRY(cudaSetDevice(idx->device));
  return fm_launch_search(idx, d_packed, nq, len, d_results, v, (cudaStream_t) stream, NULL);
}

extern "C" int32_t fmgpu_pack_queries_device(int32_t device, const char *d_ascii, uint64_t nq, uint32_t len,
                                             uint32_t *d_packed, void *stream)
{
  if (!d_ascii || !d_packed || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(device));
  if (nq == 0) return FM_SUCCESS;
  const uint32_t wpq = fmgpu_words_per_query(len);
  const uint64_t total = nq * wpq;
  if ((total + 255) / 256 >= (1ull << 31)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "batch too large for one pack launch");
  fm_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t) stream>>>(d_ascii, nq, len, wpq, d_packed);
  CU_TRY(cudaGetLastError());
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_unstream_device(int32_t device, const uint32_t *d_stream, uint64_t nq, uint32_t len,
                                         uint32_t *d_packed, void *stream)
{
  if (!d_stream || !d_packed || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(device));
  if (nq == 0) return FM_SUCCESS;
  const uint32_t wpq = fmgpu_words_per_query(len);
  if ((nq * wpq + 255) / 256 >= (1ull << 31)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "batch too large for one launch");
  fm_unstream_kernel<<<(unsigned)((nq * wpq + 255) / 256), 256, 0, (cudaStream_t) stream>>>(d_stream, nq, len, wpq, d_packed);
  CU_TRY(cudaGetLastError());
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ *
 * query shards
 * ------------------------------------------------------------------------ */
extern "C" int32_t fmgpu_batch_create(int32_t device, uint64_t nq, uint32_t len, uint32_t steps, fmgpu_batch_t **out)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!out || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  fmgpu_batch_t *b = (fmgpu_batch_t *) calloc(1, sizeof(*b));
  if (!b) return fm_fail_msg(FM_E_ALLOCATING_MFASTA, "host allocation failed");
  b->device = device; b->nq = nq; b->len = len; b->steps = steps; b->wpq = fmgpu_words_per_query(len);
  const size_t pw = (size_t)(nq ? nq : 1) * b->wpq * 4, rw = (size_t)(nq ? nq : 1) * 8;
  cudaError_t e = cudaMalloc((void **) &b->d_packed, pw);
  if (e == cudaSuccess) e = cudaMalloc((void **) &b->d_results, rw);
  if (e == cudaSuccess) e = cudaMemset(b->d_results, 0, rw);                 /* reference: cudaMemset of results */
  if (e == cudaSuccess) e = cudaMalloc((void **) &b->d_counters, 16);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&b->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
  if (e != cudaSuccess) { fmgpu_batch_free(&b); return fm_fail(e, "fmgpu_batch_create", __FILE__, __LINE__); }
  *out = b;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_upload_ascii(fmgpu_batch_t *b, const char *h_ascii)
{
  if (!b || !h_ascii) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(b->device));
  if (b->nq == 0) return FM_SUCCESS;
  /* staged in slices so the ASCII staging buffer stays small next to the packed shard */
  const uint64_t slice = 4ull << 20;                                         /* reads per slice */
  const uint64_t cap = b->nq < slice ? b->nq : slice;
  if (!b->d_ascii) CU_TRY(cudaMalloc((void **) &b->d_ascii, cap * b->len));
  for (uint64_t q0 = 0; q0 < b->nq; q0 += slice) {
    const uint64_t n = (b->nq - q0 < slice) ? b->nq - q0 : slice;
    CU_TRY(cudaMemcpyAsync(b->d_ascii, h_ascii + q0 * b->len, n * b->len, cudaMemcpyHostToDevice, b->stream));
    int32_t rc = fmgpu_pack_queries_device(b->device, b->d_ascii, n, b->len, b->d_packed + q0 * b->wpq, b->stream);
    if (rc) return rc;
  }
  CU_TRY(cudaStreamSynchronize(b->stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_search(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v)
{
  if (!idx || !b) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (idx->device != b->device) return fm_fail_msg(FM_E_BAD_ARGUMENT, "index replica and shard live on different devices");
  CU_TRY(cudaSetDevice(b->device));
  return fm_launch_search(idx, b->d_packed, b->nq, b->len, b->d_results, v, b->stream, NULL);
}

extern "C" int32_t fmgpu_batch_search_timed_async(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v)
{
  if (!idx || !b) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (idx->device != b->device) return fm_fail_msg(FM_E_BAD_ARGUMENT, "index replica and shard live on different devices");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaEventRecord(b->ev0, b->stream));
  const int32_t rc = fm_launch_search(idx, b->d_packed, b->nq, b->len, b->d_results, v, b->stream, NULL);
  if (rc) return rc;
  CU_TRY(cudaEventRecord(b->ev1, b->stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_last_ms(fmgpu_batch_t *b, float *ms)
{
  if (!b || !ms) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaEventSynchronize(b->ev1));
  CU_TRY(cudaEventElapsedTime(ms, b->ev0, b->ev1));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_sync(fmgpu_batch_t *b)
{
  if (!b) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaStreamSynchronize(b->stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_download(fmgpu_batch_t *b, uint32_t *h_results)
{
  if (!b || !h_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(b->device));
  if (b->nq) CU_TRY(cudaMemcpyAsync(h_results, b->d_results, b->nq * 8, cudaMemcpyDeviceToHost, b->stream));
  CU_TRY(cudaStreamSynchronize(b->stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_search_timed(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v,
                                            int32_t iters, float *ms_per_iter)
{
  if (!idx || !b || !ms_per_iter || iters < 1) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaEventRecord(b->ev0, b->stream));
  for (int i = 0; i < iters; i++) {
    int32_t rc = fmgpu_batch_search(idx, b, v);
    if (rc) return rc;
  }
  CU_TRY(cudaEventRecord(b->ev1, b->stream));
  CU_TRY(cudaEventSynchronize(b->ev1));
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  *ms_per_iter = ms / (float) iters;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_count_fetches(const fmgpu_index_t *idx, fmgpu_batch_t *b, uint64_t *nblocks, uint64_t *nsectors)
{
  if (!idx || !b) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (idx->device != b->device) return fm_fail_msg(FM_E_BAD_ARGUMENT, "index replica and shard live on different devices");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaMemsetAsync(b->d_counters, 0, 16, b->stream));
  int32_t rc = fm_launch_search(idx, b->d_packed, b->nq, b->len, b->d_results, NULL, b->stream, b->d_counters);
  if (rc) return rc;
  unsigned long long h[2] = { 0, 0 };
  CU_TRY(cudaMemcpyAsync(h, b->d_counters, 16, cudaMemcpyDeviceToHost, b->stream));
  CU_TRY(cudaStreamSynchronize(b->stream));
  if (nblocks) *nblocks = h[0];
  if (nsectors) *nsectors = h[1];
  return FM_SUCCESS;
}
extern "C" void *fmgpu_batch_packed(const fmgpu_batch_t *b)  { return b ? (void *) b->d_packed : NULL; }
extern "C" void *fmgpu_batch_results(const fmgpu_batch_t *b) { return b ? (void *) b->d_results : NULL; }
extern "C" void *fmgpu_batch_stream(const fmgpu_batch_t *b)  { return b ? (void *) b->stream : NULL; }

extern "C" int32_t fmgpu_batch_free(fmgpu_batch_t **pb)
{
  if (!pb || !*pb) return FM_SUCCESS;
  fmgpu_batch_t *b = *pb;
  cudaSetDevice(b->device);
  if (b->d_ascii) cudaFree(b->d_ascii);
  if (b->d_packed) cudaFree(b->d_packed);
  if (b->d_results) cudaFree(b->d_results);
  if (b->d_counters) cudaFree(b->d_counters);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream) cudaStreamDestroy(b->stream);
  free(b);
  *pb = NULL;
  return FM_SUCCESS;
}
/* fetch counter on caller-owned device memory: one instrumented search (results are written too) */
extern "C" int32_t fmgpu_count_fetches_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                              uint32_t *d_results, void *stream, uint64_t *nblocks, uint64_t *nsectors)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long *d_c = NULL, h[2] = { 0, 0 };
  CU_TRY(cudaMalloc((void **) &d_c, 16));
  CU_TRY(cudaMemsetAsync(d_c, 0, 16, (cudaStream_t) stream));
  int32_t rc = fm_launch_search(idx, d_packed, nq, len, d_results, NULL, (cudaStream_t) stream, d_c);
  if (rc == FM_SUCCESS) {
    cudaError_t e = cudaMemcpyAsync(h, d_c, 16, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) rc = fm_fail(e, "fetch counters D2H", __FILE__, __LINE__);
  }
  cudaFree(d_c);
  if (nblocks) *nblocks = h[0];
  if (nsectors) *nsectors = h[1];
  return rc;
}

/*
 * fm_probe.cu -- HBM random-access roofline probes (SURVEY.md 8(d) "Ceiling").
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 */
#include "fm_internal.h"
#include "fm_probe.cuh"

/* ------------------------------------------------------------------------ */
template <int WIDTH>
static cudaError_t fm_probe_launch(uint32_t grid, const uint4 *table, uint64_t naccess, uint32_t lpt, uint32_t *sink)
{
  fm_gather_probe_kernel<(WIDTH >= 4 ? 2 : 4), WIDTH><<<grid, 256>>>(table, naccess, lpt, sink);
  return cudaGetLastError();
}

extern "C" int32_t fmgpu_gather_probe_ex(int32_t device, uint64_t table_bytes, uint32_t access_bytes,
                                         uint64_t loads_per_thread, int32_t iters, double *accesses_per_second)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!accesses_per_second || table_bytes < 4096 || iters < 1) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  if (!(access_bytes == 16 || access_bytes == 32 || access_bytes == 64 || access_bytes == 128))
    return fm_fail_msg(FM_E_BAD_ARGUMENT, "access_bytes must be 16, 32, 64 or 128");
  const uint64_t naccess = table_bytes / access_bytes;
  uint4 *table = NULL; uint32_t *sink = NULL;
  CU_TRY(cudaMalloc((void **) &table, naccess * access_bytes));
  CU_TRY(cudaMalloc((void **) &sink, 4));
  CU_TRY(cudaMemset(table, 0x5A, naccess * access_bytes));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const uint32_t lpt = (uint32_t)((loads_per_thread + 3) & ~3ull);
  const uint32_t grid = (uint32_t) sms * 8 * 4;                    /* 4 waves of 8 CTAs per SM */
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int i = 0; i <= iters; i++) {                               /* i == 0 is the warm-up */
    CU_TRY(cudaEventRecord(e0));
    cudaError_t e = access_bytes == 16 ? fm_probe_launch<1>(grid, table, naccess, lpt, sink)
                  : access_bytes == 32 ? fm_probe_launch<2>(grid, table, naccess, lpt, sink)
                  : access_bytes == 64 ? fm_probe_launch<4>(grid, table, naccess, lpt, sink)
                                       : fm_probe_launch<8>(grid, table, naccess, lpt, sink);
    if (e != cudaSuccess) return fm_fail(e, "fm_gather_probe_kernel", __FILE__, __LINE__);
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (i > 0 && ms < best) best = ms;
  }
  *accesses_per_second = (double) grid * 256.0 * lpt / (best * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(table); cudaFree(sink);
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_gather_probe(int32_t device, uint64_t table_bytes, uint64_t loads_per_thread, int32_t iters,
                                      double *loads_per_second)
{
  return fmgpu_gather_probe_ex(device, table_bytes, 16, loads_per_thread, iters, loads_per_second);
}

/* locality probe: the 32 lanes of each warp-level load fall in one random window of `window_bytes` */
extern "C" int32_t fmgpu_gather_probe_local(int32_t device, uint64_t table_bytes, uint64_t window_bytes,
                                            uint64_t loads_per_thread, int32_t iters, double *loads_per_second)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!loads_per_second || window_bytes < 512 || table_bytes < window_bytes || iters < 1 || window_bytes > (1ull << 34))
    return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  const uint32_t window16 = (uint32_t)(window_bytes / 16);
  const uint64_t nwindows = table_bytes / window_bytes;
  uint4 *table = NULL; uint32_t *sink = NULL;
  CU_TRY(cudaMalloc((void **) &table, nwindows * window_bytes));
  CU_TRY(cudaMalloc((void **) &sink, 4));
  CU_TRY(cudaMemset(table, 0x5A, nwindows * window_bytes));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const uint32_t lpt = (uint32_t)((loads_per_thread + 3) & ~3ull);
  const uint32_t grid = (uint32_t) sms * 8 * 4;
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int i = 0; i <= iters; i++) {
    CU_TRY(cudaEventRecord(e0));
    fm_gather_probe_local_kernel<4><<<grid, 256>>>(table, nwindows, window16, lpt, sink);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (i > 0 && ms < best) best = ms;
  }
  *loads_per_second = (double) grid * 256.0 * lpt / (best * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(table); cudaFree(sink);
  return FM_SUCCESS;
}

/*
 * fm_probe.cuh -- HBM random-access roofline probes.  Included by fm_probe.cu only.
 */
#ifndef FM_PROBE_CUH_
#define FM_PROBE_CUH_

#include "fm_device.cuh"

/* ------------------------------------------------------------------------ *
 * Gather roofline probe: independent uniformly random aligned accesses of
 * WIDTH consecutive 16-byte loads (16, 32, 64 or 128 bytes per access).
 * ------------------------------------------------------------------------ */
template <int UNROLL, int WIDTH>
__global__ void __launch_bounds__(256, 8) fm_gather_probe_kernel(const uint4 *__restrict__ table, uint64_t naccess,
                                                                   uint32_t loads_per_thread, uint32_t *sink)
{
  uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  uint32_t acc = 0;
  for (uint32_t it = 0; it < loads_per_thread; it += UNROLL) {
    uint4 v[UNROLL][WIDTH];
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;                        /* xorshift64 */
      const uint64_t idx = __umul64hi(s, naccess);                     /* uniform in [0, naccess) */
      #pragma unroll
      for (int w = 0; w < WIDTH; w++) v[u][w] = fm_ldg16(table + idx * WIDTH + w);
    }
    #pragma unroll
    for (int u = 0; u < UNROLL; u++)
      #pragma unroll
      for (int w = 0; w < WIDTH; w++) acc += v[u][w].x ^ v[u][w].y ^ v[u][w].z ^ v[u][w].w;
  }
  if (acc == 0x9E3779B9u) *sink = acc;                                 /* keeps the loads alive */
}

/* Locality probe: every warp-level load picks ONE random window of `window16` 16-byte blocks (the same for
 * its 32 lanes) and each lane a random block inside it.  Separates address-translation cost (one 2 MB page
 * per warp instruction) from DRAM sector cost (32 distinct sectors per warp instruction either way). */
template <int UNROLL>
__global__ void __launch_bounds__(256, 8) fm_gather_probe_local_kernel(const uint4 *__restrict__ table, uint64_t nwindows,
                                                                         uint32_t window16, uint32_t loads_per_thread,
                                                                         uint32_t *sink)
{
  const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t sw = (tid >> 5) * 0x9E3779B97F4A7C15ull + 0x7654321ull;    /* warp-uniform stream */
  uint64_t sl = tid * 0xD1B54A32D192ED03ull + 0x1234567ull;           /* per-lane stream */
  uint32_t acc = 0;
  for (uint32_t it = 0; it < loads_per_thread; it += UNROLL) {
    uint4 v[UNROLL];
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      sw ^= sw << 13; sw ^= sw >> 7; sw ^= sw << 17;
      sl ^= sl << 13; sl ^= sl >> 7; sl ^= sl << 17;
      const uint64_t win = __umul64hi(sw, nwindows);
      const uint32_t off = __umulhi((uint32_t)(sl >> 32), window16);
      v[u] = fm_ldg16(table + win * window16 + off);
    }
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x9E3779B9u) *sink = acc;
}

#endif /* FM_PROBE_CUH_ */

// Standalone experiment: is the random-access ceiling an address-translation limit?  Every CTA (256 threads) confines
// its random 16-byte loads to ONE randomly placed window of W bytes for its whole life; the union of the windows still
// covers the table uniformly, so the L2 hit rate stays that of uniform random access (table >> L2).  If small windows
// run faster, the limit is translation reach per SM, not the DRAM/L2 miss path.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o window_cta window_cta.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 8) probe(const uint4 *__restrict__ table, uint64_t n16, uint64_t win16, uint32_t lpt, uint32_t *sink)
{
  uint64_t h = (blockIdx.x + 1) * 0xD6E8FEB86659FD93ull; h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
  const uint64_t base = (__umul64hi(h, n16 - win16 + 1)) & ~(uint64_t) 0x1FFFF;      /* 2 MB aligned (in 16-byte units) */
  uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  uint32_t acc = 0;
  for (uint32_t it = 0; it < lpt; it += 4) {
    uint4 v[4];
    #pragma unroll
    for (int u = 0; u < 4; u++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const uint4 *p = table + base + __umul64hi(s, win16);
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
    }
    #pragma unroll
    for (int u = 0; u < 4; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x9E3779B9u) *sink = acc;
}

int main(int argc, char **argv)
{
  const double gb = argc > 1 ? atof(argv[1]) : 5.4;
  const uint32_t lpt = 128;
  const uint64_t n16 = (uint64_t)(gb * (1ull << 30)) / 16;
  uint4 *table; uint32_t *sink;
  if (cudaMalloc(&table, n16 * 16) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(table, 0x5A, n16 * 16);
  const int grid = 148 * 8 * 4;
  const double wins_mb[] = { 2, 8, 32, 128, 512, 2048, 0 };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (double w : wins_mb) {
    uint64_t win16 = w > 0 ? (uint64_t)(w * (1 << 20)) / 16 : n16;
    if (win16 > n16) win16 = n16;
    float best = 1e30f;
    for (int i = 0; i < 3; i++) {
      cudaEventRecord(e0);
      probe<<<grid, 256>>>(table, n16, win16, lpt, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (i && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("{\"cuda_error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    printf("{\"table_gb\": %.1f, \"window_mb_per_cta\": %.0f, \"ms\": %.4f, \"gloads_per_s\": %.2f}\n", gb, w > 0 ? w : gb * 1024, best,
           (double) grid * 256 * lpt / (best * 1e-3) / 1e9);
  }
  return 0;
}

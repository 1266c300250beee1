// Standalone experiment: is the ~46 G/s random-miss ceiling a per-SM limit or a memory-side limit?
// One 1024-thread CTA per SM (forced by 120 KB of dynamic shared memory), S CTAs => S SMs issue random 16-byte loads
// (UNROLL independent loads in flight per thread).  If the rate grows linearly up to S = 148 the SMs are the limit;
// if it saturates earlier the memory system is.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o sm_scaling sm_scaling.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void __launch_bounds__(1024, 1) probe(const uint4 *__restrict__ table, uint64_t n16, uint32_t lpt, uint32_t *sink)
{
  extern __shared__ uint32_t pad[];
  uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  uint32_t acc = 0;
  for (uint32_t it = 0; it < lpt; it += UNROLL) {
    uint4 v[UNROLL];
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const uint4 *p = table + __umul64hi(s, n16);
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
    }
    #pragma unroll
    for (int u = 0; u < UNROLL; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x9E3779B9u) { *sink = acc; pad[threadIdx.x] = acc; }
}

template <int UNROLL> void sweep(const uint4 *table, uint64_t n16, uint32_t *sink, size_t smem, const char *what)
{
  cudaFuncSetAttribute(probe<UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const int grids[] = { 9, 18, 37, 74, 111, 148 };
  const uint32_t lpt = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int g : grids) {
    float best = 1e30f;
    for (int i = 0; i < 3; i++) {
      cudaEventRecord(e0);
      probe<UNROLL><<<g, 1024, smem>>>(table, n16, lpt, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (i && ms < best) best = ms;
    }
    const double rate = (double) g * 1024 * lpt / (best * 1e-3) / 1e9;
    printf("{\"config\": \"%s\", \"unroll\": %d, \"sms\": %d, \"ms\": %.3f, \"gloads_per_s\": %.2f, \"gloads_per_s_per_sm\": %.4f}\n", what, UNROLL, g, best, rate, rate / g);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("{\"cuda_error\": \"%s\"}\n", cudaGetErrorString(e)); exit(1); }
}

int main(int argc, char **argv)
{
  const double gb = argc > 1 ? atof(argv[1]) : 5.4;
  const uint64_t n16 = (uint64_t)(gb * (1ull << 30)) / 16;
  uint4 *table; uint32_t *sink;
  if (cudaMalloc(&table, n16 * 16) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(table, 0x5A, n16 * 16);
  sweep<4>(table, n16, sink, 120 * 1024, "1 CTA of 1024 threads per SM");
  sweep<8>(table, n16, sink, 120 * 1024, "1 CTA of 1024 threads per SM");
  sweep<2>(table, n16, sink, 120 * 1024, "1 CTA of 1024 threads per SM");
  return 0;
}

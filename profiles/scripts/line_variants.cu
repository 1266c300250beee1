// Experiment: how do B200's L2 / HBM treat (a) per-lane random 16-byte loads, (b) G lanes cooperatively
// loading one random aligned 16*G-byte chunk in ONE instruction, (c) one lane issuing several loads to one line?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o line_variants line_variants.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld16(const uint4 *p)
{
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// G lanes share one random chunk of G*16 bytes per load instruction (G = 1: independent lanes)
template <int G>
__global__ void __launch_bounds__(256, 8) coop_probe(const uint4 *__restrict__ table, uint64_t nchunks, uint32_t lpt, uint32_t *sink)
{
  const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = (tid / G) * 0x9E3779B97F4A7C15ull + 0x1234567ull;      // same stream for the G lanes of a group
  const uint32_t part = (uint32_t)(tid % G);
  uint32_t acc = 0;
  for (uint32_t it = 0; it < lpt; it += 4) {
    uint4 v[4];
    #pragma unroll
    for (int u = 0; u < 4; u++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      v[u] = ld16(table + __umul64hi(s, nchunks) * G + part);
    }
    #pragma unroll
    for (int u = 0; u < 4; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x9E3779B9u) *sink = acc;
}

// one lane issues W separate 16-byte loads to one random 16*W-byte chunk
template <int W>
__global__ void __launch_bounds__(256, 8) serial_probe(const uint4 *__restrict__ table, uint64_t nchunks, uint32_t lpt, uint32_t *sink)
{
  uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  uint32_t acc = 0;
  for (uint32_t it = 0; it < lpt; it += 2) {
    uint4 v[2][W];
    #pragma unroll
    for (int u = 0; u < 2; u++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const uint64_t c = __umul64hi(s, nchunks);
      #pragma unroll
      for (int w = 0; w < W; w++) v[u][w] = ld16(table + c * W + w);
    }
    #pragma unroll
    for (int u = 0; u < 2; u++)
      #pragma unroll
      for (int w = 0; w < W; w++) acc += v[u][w].x ^ v[u][w].y;
  }
  if (acc == 0x9E3779B9u) *sink = acc;
}

// one lane, one 256-bit load of a random 32-byte sector
__global__ void __launch_bounds__(256, 8) v8_probe(const uint4 *__restrict__ table, uint64_t nchunks, uint32_t lpt, uint32_t *sink)
{
  uint64_t s = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  uint32_t acc = 0;
  for (uint32_t it = 0; it < lpt; it += 4) {
    uint32_t r[4][8];
    #pragma unroll
    for (int u = 0; u < 4; u++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const uint4 *p = table + __umul64hi(s, nchunks) * 2;
      asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]), "=r"(r[u][6]), "=r"(r[u][7]) : "l"(p));
    }
    #pragma unroll
    for (int u = 0; u < 4; u++) acc += r[u][0] ^ r[u][3] ^ r[u][4] ^ r[u][7];
  }
  if (acc == 0x9E3779B9u) *sink = acc;
}

template <typename F> float timeit(F launch)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int i = 0; i < 3; i++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (i && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char **argv)
{
  const double gb = argc > 1 ? atof(argv[1]) : 5.4;
  const uint32_t lpt = argc > 2 ? atoi(argv[2]) : 128;
  const uint64_t bytes = (uint64_t)(gb * (1ull << 30)) & ~1023ull;
  uint4 *table; uint32_t *sink;
  if (cudaMalloc(&table, bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(table, 0x5A, bytes);
  const int grid = 148 * 8 * 4;
  const double nthreads = (double) grid * 256;
#define REPORT(name, chunk, per_thread_accesses, ms) \
  printf("{\"variant\": \"%s\", \"table_gb\": %.2f, \"chunk_bytes\": %d, \"ms\": %.4f, \"gchunks_per_s\": %.2f, \"useful_gbs\": %.1f}\n", name, gb, chunk, ms, \
         nthreads * (per_thread_accesses) / ((ms) * 1e-3) / 1e9, nthreads * (per_thread_accesses) * (chunk) / ((ms) * 1e-3) / 1e9); fflush(stdout);
  float ms;
  ms = timeit([&] { coop_probe<1><<<grid, 256>>>(table, bytes / 16, lpt, sink); });   REPORT("lanes independent, 16 B", 16, (double) lpt, ms);
  ms = timeit([&] { coop_probe<2><<<grid, 256>>>(table, bytes / 32, lpt, sink); });   REPORT("2 lanes share a 32 B sector (one instr)", 32, lpt / 2.0, ms);
  ms = timeit([&] { coop_probe<4><<<grid, 256>>>(table, bytes / 64, lpt, sink); });   REPORT("4 lanes share 64 B (one instr)", 64, lpt / 4.0, ms);
  ms = timeit([&] { coop_probe<8><<<grid, 256>>>(table, bytes / 128, lpt, sink); });  REPORT("8 lanes share a 128 B line (one instr)", 128, lpt / 8.0, ms);
  ms = timeit([&] { coop_probe<16><<<grid, 256>>>(table, bytes / 256, lpt, sink); }); REPORT("16 lanes share 256 B (one instr)", 256, lpt / 16.0, ms);
  ms = timeit([&] { coop_probe<32><<<grid, 256>>>(table, bytes / 512, lpt, sink); }); REPORT("32 lanes share 512 B (one instr)", 512, lpt / 32.0, ms);
  ms = timeit([&] { serial_probe<2><<<grid, 256>>>(table, bytes / 32, lpt, sink); }); REPORT("1 lane, 2 loads in one sector", 32, (double) lpt, ms);
  ms = timeit([&] { serial_probe<8><<<grid, 256>>>(table, bytes / 128, lpt, sink); });REPORT("1 lane, 8 loads in one line", 128, (double) lpt, ms);
  ms = timeit([&] { v8_probe<<<grid, 256>>>(table, bytes / 32, lpt, sink); });        REPORT("1 lane, one 256-bit load", 32, (double) lpt, ms);
  return 0;
}

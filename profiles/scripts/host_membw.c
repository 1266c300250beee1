// Host memory bandwidth of the GPU box, all cores: the end-to-end search is fed with 100 bytes of ASCII per read from host
// memory, so (read bandwidth / 100 B) bounds the reads/s any feed can reach (DMA and CPU packers read the same DRAM).
// build: gcc -O3 -march=native -fopenmp -o host_membw host_membw.c
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }
int main(int argc, char **argv)
{
  const size_t bytes = (argc > 1 ? (size_t) atof(argv[1]) : 4) * 1000000000ull;
  uint64_t *a = aligned_alloc(64, bytes), *b = aligned_alloc(64, bytes / 4);
  const size_t n = bytes / 8;
  #pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) a[i] = i * 0x9E3779B97F4A7C15ull;
  #pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n / 4; i++) b[i] = 0;
  double best_r = 0, best_p = 0;
  uint64_t sink = 0;
  for (int it = 0; it < 5; it++) {
    double t0 = now();
    uint64_t s = 0;
    #pragma omp parallel for schedule(static) reduction(+ : s)
    for (size_t i = 0; i < n; i++) s += a[i];
    double dt = now() - t0; sink += s;
    if (bytes / dt > best_r) best_r = bytes / dt;
    t0 = now();
    #pragma omp parallel for schedule(static)                       /* read 4 words, write 1: the traffic shape of 2-bit packing */
    for (size_t i = 0; i < n / 4; i++) b[i] = a[4 * i] ^ a[4 * i + 1] ^ a[4 * i + 2] ^ a[4 * i + 3];
    dt = now() - t0;
    if (bytes / dt > best_p) best_p = bytes / dt;
  }
  printf("{\"threads\": %d, \"buffer_gb\": %.1f, \"read_gb_per_s\": %.1f, \"read4_write1_input_gb_per_s\": %.1f, \"reads_per_s_bound_100B\": %.0f, \"sink\": %llu}\n",
         omp_get_max_threads(), bytes / 1e9, best_r / 1e9, best_p / 1e9, best_r / 100.0, (unsigned long long) (sink & 1));
  return 0;
}

/*
 * fm_synth.h -- deterministic synthetic inputs for the FM-index search path.
 *
 * Counter-based (stateless) generator, so the SAME text and reads can be
 * produced by plain C on the host and by a CUDA kernel on the device, at any
 * position, in any order, by any number of threads.
 *
 * Follows the reference's input conventions:
 *   - reference text: uniform i.i.d. A/C/G/T (the scripts feed a FASTA file to
 *     gfmiBaseLine_*; reader common/common.c:42-76);
 *   - reads: exact substrings with a uniformly random start in [0, n-len]
 *     (resources/genreads.py:71-76), FASTA header ">rid<i> <s+1>-<e+1>".
 */
#ifndef FM_SYNTH_H_
#define FM_SYNTH_H_

#include <stdint.h>

#if defined(__CUDACC__)
  #define FM_HD __host__ __device__ static __forceinline__
#else
  #define FM_HD static inline
#endif

/* splitmix64: Weyl sequence + Stafford mix13 finaliser, evaluated at counter i */
FM_HD uint64_t fm_synth_mix64(uint64_t seed, uint64_t i)
{
  uint64_t x = (seed + i + 1ull) * 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* 2-bit code of text position i: 0=A 1=C 2=G 3=T (the reference's coding,
 * src/genFMindex.c:71-84) */
FM_HD uint32_t fm_synth_code(uint64_t seed, uint64_t i)
{
  return (uint32_t)(fm_synth_mix64(seed, i) >> 62);
}

FM_HD char fm_synth_base(uint64_t seed, uint64_t i)
{
  return (char)((0x54474341u >> (8u * fm_synth_code(seed, i))) & 0xFFu); /* "ACGT" */
}

/* start of read j: uniform in [0, n-len] (resources/genreads.py:72) */
FM_HD uint64_t fm_synth_read_start(uint64_t seed, uint64_t j, uint64_t n, uint64_t len)
{
  return fm_synth_mix64(seed ^ 0xA5A5A5A55A5A5A5Aull, j) % (n - len + 1ull);
}

#endif /* FM_SYNTH_H_ */

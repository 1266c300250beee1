/*
 * fm_mismatch.cu -- one-mismatch search ("inexact extensions"; the reference stops at exact
 * matching, so the checker in tests/ is a brute-force scan of the text).
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 *
 * Built on the exact-match kernels, whichever family the caller picks: every read is searched as it is and in all its
 * 3 * len single-substitution variants, which are generated on the device in chunks (a variant is the packed read with one
 * 2-bit field changed), searched by ONE launch of the ordinary kernel per chunk, and reduced per read.  With 14 bases per
 * fetch a variant costs ~7 block fetches, so a 100-bp read costs ~2 100: ~20 M reads/s per GPU at the random-access
 * ceiling.  (Sharing the exact suffix steps between the variants of a read would save less than half of that, because a
 * substitution changes the whole 14-base symbol it falls into and every step after it.)
 */
#include "fm_internal.h"

/* variants[(q * nvar + j) * wpq + w]: read q with packed position j / 3 changed to (old + 1 + j % 3) & 3 */
__global__ void fm_mm1_variants_kernel(const uint32_t *__restrict__ packed, uint64_t nq, uint32_t len, uint32_t wpq, uint32_t *__restrict__ variants)
{
  const uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nvar = 3u * len;
  if (idx >= nq * nvar * wpq) return;
  const uint32_t w = (uint32_t)(idx % wpq);
  const uint64_t qj = idx / wpq;
  const uint32_t j = (uint32_t)(qj % nvar);
  const uint64_t q = qj / nvar;
  uint32_t word = packed[q * wpq + w];
  const uint32_t t = j / 3u, s = j - 3u * t;
  if (t / 16u == w) {
    const uint32_t sh = 2u * (t & 15u), old = (word >> sh) & 3u;
    word = (word & ~(3u << sh)) | (((old + 1u + s) & 3u) << sh);
  }
  variants[idx] = word;
}

/* one warp per read: variants with a non-empty interval, and their occurrences in total */
__global__ void fm_mm1_reduce_kernel(const uint2 *__restrict__ exact, const uint2 *__restrict__ vlr, uint64_t nq, uint32_t nvar, fmgpu_mm1_t *__restrict__ out)
{
  const uint64_t q = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (q >= nq) return;
  uint32_t found = 0;
  unsigned long long occ = 0;
  for (uint32_t j = lane; j < nvar; j += 32) {
    const uint2 x = vlr[q * nvar + j];
    if (x.y > x.x) { found++; occ += x.y - x.x; }
  }
  for (int o = 16; o > 0; o >>= 1) { found += __shfl_xor_sync(0xFFFFFFFFu, found, o); occ += __shfl_xor_sync(0xFFFFFFFFu, occ, o); }
  if (lane == 0) {
    fmgpu_mm1_t r;
    r.L = exact[q].x; r.R = exact[q].y; r.variants_found = found;
    r.occurrences_1mm = occ > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) occ;
    out[q] = r;
  }
}

extern "C" int32_t fmgpu_search_device_mm1(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len, fmgpu_mm1_t *d_out,
                                           uint32_t *d_variant_lr, const fmgpu_variant_t *v, void *stream_)
{
  if (!idx || !d_packed || !d_out || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  if (nq == 0) return FM_SUCCESS;
  CU_TRY(cudaSetDevice(idx->device));
  cudaStream_t stream = (cudaStream_t) stream_;
  const uint32_t wpq = fmgpu_words_per_query(len), nvar = 3u * len;
  /* chunk so that variants + their (L,R) stay near 1.5 GB */
  const uint64_t per_read = (uint64_t) nvar * (wpq * 4ull + 8ull);
  uint64_t chunk = (1536ull << 20) / per_read;
  if (chunk < 1) chunk = 1;
  if (chunk > nq) chunk = nq;
  uint32_t *vars = NULL; uint2 *vlr = NULL, *exact = NULL;
  cudaError_t e = cudaMalloc((void **) &exact, nq * 8);
  if (e == cudaSuccess && !d_variant_lr) e = cudaMalloc((void **) &vlr, chunk * nvar * 8ull);
  if (e == cudaSuccess) e = cudaMalloc((void **) &vars, chunk * nvar * wpq * 4ull);
  int32_t rc = e == cudaSuccess ? FM_SUCCESS : fm_fail(e, "cudaMalloc(one-mismatch scratch)", __FILE__, __LINE__);
  if (rc == FM_SUCCESS) rc = fm_launch_search(idx, d_packed, nq, len, (uint32_t *) exact, v, stream, NULL);
  for (uint64_t q0 = 0; q0 < nq && rc == FM_SUCCESS; q0 += chunk) {
    const uint64_t n = nq - q0 < chunk ? nq - q0 : chunk;
    uint2 *dst = d_variant_lr ? (uint2 *) d_variant_lr + q0 * nvar : vlr;
    const uint64_t total = n * nvar * wpq;
    fm_mm1_variants_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_packed + q0 * wpq, n, len, wpq, vars);
    if ((e = cudaGetLastError()) != cudaSuccess) { rc = fm_fail(e, "fm_mm1_variants_kernel", __FILE__, __LINE__); break; }
    rc = fm_launch_search(idx, vars, n * nvar, len, (uint32_t *) dst, v, stream, NULL);
    if (rc) break;
    fm_mm1_reduce_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(exact + q0, dst, n, nvar, d_out + q0);
    if ((e = cudaGetLastError()) != cudaSuccess) { rc = fm_fail(e, "fm_mm1_reduce_kernel", __FILE__, __LINE__); break; }
  }
  e = cudaStreamSynchronize(stream);                           /* the scratch is released below */
  if (rc == FM_SUCCESS && e != cudaSuccess) rc = fm_fail(e, "one-mismatch search", __FILE__, __LINE__);
  cudaFree(vars); cudaFree(vlr); cudaFree(exact);
  return rc;
}

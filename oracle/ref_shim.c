/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A few lines of glue compiled TOGETHER with the unmodified reference sources
 * (common/common.c + src/fmIndexCPUBaseline[-AltCounters].c, taken from where
 * they lie under /root/reference) into oracle/_ref/libref_search_*.so, so that
 * tests and bench.py's cpu_baseline / --impl reference legs can call the
 * reference's own loadIndex / searchIndexCPU in-process.
 *
 * Why a shim is needed: the reference's searchIndexCPU uses an ORPHANED
 * "#pragma omp for" (src/fmIndexCPUBaseline.c:195) and relies on its caller
 * (common/searchQueries.c:84-95) to open the parallel region.  Called from
 * ctypes it would run on one thread.  ref_search_parallel() opens the same
 * parallel region the reference main() opens, nothing more.
 *
 * It also lets a caller hand the reference searcher an index / query batch
 * that is already in memory (bench: the 2 Gbp index built on the GPU in the
 * reference's tag-100 byte layout), by filling the first fields of the
 * reference structs, whose layout is fixed by common/common.h:64-81 and
 * src/fmIndexCPUBaseline.c:54-69.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "common.h"
#include "interface.h"

double sampleTime();

/* Mirrors the field order of fmi_t (src/fmIndexCPUBaseline.c:54-69); the
 * entry type is opaque here because its size depends on -DK_STEPS/-DNUM_CHUNK. */
typedef struct {
  uint32_t steps, bwtsize, ncounters, nentries, chunk, nbitmaps;
  uint32_t *h_dollarPositionBWT, *h_dollarBaseBWT, *h_modposdollarBWT;
  void     *h_index;
  uint32_t *d_dollarPositionBWT, *d_dollarBaseBWT, *d_modposdollarBWT;
  void     *d_index;
} shim_fmi_t;

/* compile-time configuration of THIS .so, so a caller can check it */
int32_t ref_cfg_steps(void)   { return K_STEPS; }
int32_t ref_cfg_chunk(void)   { return NUM_CHUNK; }
int32_t ref_cfg_ac(void)
{
#ifdef REF_SHIM_AC
  return 1;
#else
  return 0;
#endif
}
int32_t ref_max_threads(void) { return omp_get_max_threads(); }

/* Runs `iters` passes of the reference search with the reference's own
 * threading (common/searchQueries.c:84-95) and returns seconds per pass. */
double ref_search_parallel(void *index, void *queries, void *results, int32_t iters, int32_t nthreads)
{
  double t0, t1;
  int32_t n;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  t0 = sampleTime();
  #pragma omp parallel private(n)
  {
    for (n = 0; n < iters; n++)
      searchIndexCPU(index, queries, results);
  }
  t1 = sampleTime();
  return (t1 - t0) / (double)(iters > 0 ? iters : 1);
}

/* Wraps an in-memory index image (the bytes of a reference index FILE: header
 * then entries, Appendix A of SURVEY.md) as a reference fmi_t without going
 * through the file system.  The image must stay alive while the handle is used. */
void *ref_wrap_index_image(const uint32_t *image)
{
  shim_fmi_t *f = (shim_fmi_t *) calloc(1, sizeof(shim_fmi_t));
  uint32_t k = image[1], i;
  f->steps = k; f->bwtsize = image[2]; f->ncounters = image[3];
  f->nentries = image[4]; f->chunk = image[5];
  f->h_dollarPositionBWT = (uint32_t *) malloc(k * sizeof(uint32_t));
  f->h_dollarBaseBWT     = (uint32_t *) malloc(k * sizeof(uint32_t));
  f->h_modposdollarBWT   = (uint32_t *) malloc(k * sizeof(uint32_t));
  for (i = 0; i < k; i++) {
    f->h_dollarPositionBWT[i] = image[6 + i];
    f->h_dollarBaseBWT[i]     = image[6 + k + i];
    f->h_modposdollarBWT[i]   = image[6 + i] / f->chunk;
  }
  f->h_index = (void *) (image + 6 + 2 * k);
  return f;
}

/* Wraps caller-owned query / result buffers in the reference's structs
 * (common/common.h:64-81). */
void *ref_wrap_queries(char *ascii, uint32_t num, uint32_t size)
{
  qrys_t *q = (qrys_t *) calloc(1, sizeof(qrys_t));
  q->num = num; q->size = size; q->h_queries = ascii;
  return q;
}

void *ref_wrap_results(uint32_t *buf, uint32_t num)
{
  res_t *r = (res_t *) calloc(1, sizeof(res_t));
  r->num = num; r->h_results = buf;
  return r;
}

uint32_t *ref_results_ptr(void *results) { return ((res_t *) results)->h_results; }
char     *ref_queries_ptr(void *queries) { return ((qrys_t *) queries)->h_queries; }

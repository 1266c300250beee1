/*
 * host_sanitize.c -- exercises the plain-C host layer (fm_host.c, fm_ingest.c, fm_hostpack.c) under
 * AddressSanitizer + UndefinedBehaviorSanitizer: `make -C k-step_fm-index_b200/csrc sanitize` compiles those three
 * files with -fsanitize=address,undefined into this program (the CUDA side comes from the ordinary library) and
 * tests/test_host.py runs it.  No GPU needed: loaders, writers, packers, error paths, handle lifetimes.
 *
 *   host_sanitize <index file> <k> <scratch dir>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/fmindex_b200.h"

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "host_sanitize: %s failed (%s:%d)\n", #cond, __FILE__, __LINE__); return 1; } } while (0)

static void write_reads(const char *fn, uint32_t num, uint32_t len, int crlf, int short_line)
{
  FILE *fp = fopen(fn, "wb");
  uint32_t q, i;
  for (q = 0; q < num; q++) {
    fprintf(fp, ">rid%u 1-%u%s", q + 1, len, crlf ? "\r\n" : "\n");
    for (i = 0; i < len - (short_line && q == num / 2 ? 1u : 0u); i++) fputc("ACGT"[(q * 7 + i * 13 + (i >> 2)) & 3], fp);
    if (q + 1 < num || !crlf) fputs(crlf ? "\r\n" : "\n", fp);       /* crlf file: no final newline */
  }
  fclose(fp);
}

int main(int argc, char **argv)
{
  void *index = NULL, *queries = NULL, *results = NULL, *loaded = NULL;
  char path[1024], out[1100];
  uint32_t num = 1001, len = 37, i;
  if (argc < 4) return 2;
  /* index loader: the real file, a truncated copy, a file that is no index, a missing file */
  CHECK(loadIndex(argv[1], &index) == FM_SUCCESS);
  {
    fmi_t *f = (fmi_t *) index;
    CHECK(f->steps == (uint32_t) atoi(argv[2]) && f->h_index && f->entry_words == f->nbitmaps * f->steps + f->ncounters);
    snprintf(path, sizeof path, "%s/trunc.fmi", argv[3]);
    {
      FILE *in = fopen(argv[1], "rb"), *o = fopen(path, "wb");
      char buf[4096];
      size_t n = fread(buf, 1, sizeof buf, in);
      fwrite(buf, 1, n / 2, o);
      fclose(in); fclose(o);
    }
    CHECK(loadIndex(path, &loaded) == FM_E_READING_FMI && loaded == NULL);
    snprintf(path, sizeof path, "%s/noindex.fmi", argv[3]);
    { FILE *o = fopen(path, "wb"); fputs("this is not an index file, not at all", o); fclose(o); }
    CHECK(loadIndex(path, &loaded) == FM_E_INDEX_VER_BASELINE);
    CHECK(loadIndex("/nonexistent/x.fmi", &loaded) == FM_E_OPENING_INDEX_FILE);
  }
  /* query loader: LF, CRLF without a final newline, a short line, fewer reads than asked */
  snprintf(path, sizeof path, "%s/reads.fa", argv[3]);
  write_reads(path, num, len, 0, 0);
  CHECK(loadQueries(path, len, num, &queries) == FM_SUCCESS);
  {
    qrys_t *q = (qrys_t *) queries;
    void *q2 = NULL;
    char *keep = (char *) malloc((size_t) num * len);
    memcpy(keep, q->h_queries, (size_t) num * len);
    snprintf(path, sizeof path, "%s/reads_crlf.fa", argv[3]);
    write_reads(path, num, len, 1, 0);
    CHECK(loadQueries(path, len, num, &q2) == FM_SUCCESS);
    CHECK(memcmp(((qrys_t *) q2)->h_queries, keep, (size_t) num * len) == 0);
    freeQueries(&q2); free(q2); q2 = NULL;
    CHECK(loadQueries(path, len, num + 1, &q2) == FM_E_READING_MFASTA_FILE);
    snprintf(path, sizeof path, "%s/reads_short.fa", argv[3]);
    write_reads(path, num, len, 0, 1);
    CHECK(loadQueries(path, len, num, &q2) == FM_E_READING_MFASTA_FILE);
    CHECK(loadQueries("/nonexistent/r.fa", len, num, &q2) == FM_E_OPENING_MFASTA_FILE);
    /* packers: AVX-512 / scalar agree; stream packer touches exactly (nbases + 3) / 4 bytes */
    {
      const uint32_t wpq = fmgpu_words_per_query(len);
      uint32_t *a = (uint32_t *) malloc((size_t) num * wpq * 4), *b = (uint32_t *) malloc((size_t) num * wpq * 4);
      unsigned char *s = (unsigned char *) malloc(((size_t) num * len + 3) / 4);
      fm_hostpack_reads(keep, num, len, a, 0);
      fm_hostpack_reads_scalar(keep, num, len, b);
      CHECK(memcmp(a, b, (size_t) num * wpq * 4) == 0);
      fm_hostpack_stream(keep, (uint64_t) num * len, s, 0);
      fm_hostpack_stream(keep, 5, s, 1);
      CHECK(fm_host_read_bandwidth(keep, (uint64_t) num * len, 2, 1) > 0.0);
      free(a); free(b); free(s);
    }
    free(keep);
  }
  /* results: init, write, load back, save under the reference's name */
  CHECK(initResults(num, &results) == FM_SUCCESS);
  {
    res_t *r = (res_t *) results;
    for (i = 0; i < 2 * num; i++) r->h_results[i] = i * 2654435761u;
    snprintf(path, sizeof path, "%s/res.txt", argv[3]);
    CHECK(writeResults(path, r->h_results, num) == FM_SUCCESS);
    CHECK(loadResults(path, &loaded) == FM_SUCCESS);
    CHECK(((res_t *) loaded)->num == num && memcmp(((res_t *) loaded)->h_results, r->h_results, (size_t) 2 * num * 4) == 0);
    freeResults(&loaded); free(loaded); loaded = NULL;
    snprintf(path, sizeof path, "%s/idx", argv[3]);
    CHECK(saveResults(path, results, index) == FM_SUCCESS);
    snprintf(out, sizeof out, "%s.res.gpu", path);
    CHECK(loadResults(out, &loaded) == FM_SUCCESS);
    freeResults(&loaded); free(loaded); loaded = NULL;
    CHECK(loadResults("/nonexistent/res", &loaded) == FM_E_OPENING_RESULTS_FILE);
  }
  /* the GPU entry points without a transfer / without a GPU: error codes, never a crash; free*GPU are idempotent */
  CHECK(transferGPUtoCPU(results) == FM_E_BAD_ARGUMENT);
  CHECK(fmgpu_search_index(index, queries, results) == FM_E_BAD_ARGUMENT);
  if (fmgpu_device_count() == 0) CHECK(transferCPUtoGPU(index, queries, results) != FM_SUCCESS);
  CHECK(freeIndexGPU(&index) == FM_SUCCESS && freeQueriesGPU(&queries) == FM_SUCCESS && freeResultsGPU(&results) == FM_SUCCESS);
  CHECK(freeIndexGPU(&index) == FM_SUCCESS && freeQueriesGPU(&queries) == FM_SUCCESS && freeResultsGPU(&results) == FM_SUCCESS);
  for (i = 0; i < 60; i++) CHECK(errorCommon((int32_t) i) != NULL);
  CHECK(errorCommon(100) && errorCommon(201) && errorCommon(-7));
  {
    int32_t devs[3] = { 0, 1, 2 };
    fmgpu_transfer_stats_t st;
    CHECK(fmgpu_set_devices(devs, 3) == FM_SUCCESS && fmgpu_set_devices(NULL, 0) == FM_SUCCESS && fmgpu_set_devices(devs, 99) == FM_E_BAD_ARGUMENT);
    CHECK(fmgpu_get_transfer_stats(&st) == FM_SUCCESS && fmgpu_get_transfer_stats(NULL) == FM_E_BAD_ARGUMENT);
  }
  /* like the reference, free* release the inner buffers and leave the handle to the caller */
  freeIndex(&index); freeQueries(&queries); freeResults(&results);
  freeIndex(&index); freeQueries(&queries); freeResults(&results);
  free(index); free(queries); free(results);
  printf("host_sanitize OK\n");
  return 0;
}

/*
 * fm_ingest.c -- fast host I/O for the two text formats of the path, plain C + OpenMP.
 *
 *   reads   : multi-FASTA, every line not starting with '>' is one read (reader common/common.c:167-173,
 *             which parses ~1 GB/s with fgets + memcpy; 100 M reads = 12 GB of text)
 *   results : "<num>\n" then "<L> <R>\n" per read (writer common/common.c:201-220, one sprintf + fputs per read)
 *
 * Same bytes in, same bytes out; the file is mmap'ed and cut at line boundaries into one slice per thread
 * (count pass, prefix sum, copy pass); results are formatted by all threads into per-thread buffers and
 * written in order.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <omp.h>
#include "../../include/fmindex_b200.h"

/* first byte of the line that starts at or after p (p == base, or the byte after the previous '\n') */
static const char *next_line_start(const char *base, const char *p, const char *end)
{
  if (p == base) return p;
  if (p >= end) return end;
  if (p[-1] == '\n') return p;
  {
    const char *nl = (const char *) memchr(p, '\n', (size_t)(end - p));
    return nl ? nl + 1 : end;
  }
}

/* walks the lines of [lo, hi); for every read line calls back with (line, length without \r\n) */
typedef struct { uint64_t count; uint64_t first_index; char *out; uint32_t len; uint64_t limit; int bad; } slice_t;

static void scan_slice(const char *lo, const char *hi, slice_t *s, int copy)
{
  const char *p = lo;
  uint64_t k = 0;
  while (p < hi) {
    const char *nl = (const char *) memchr(p, '\n', (size_t)(hi - p));
    const char *e = nl ? nl : hi;
    if (*p != '>') {
      size_t m = (size_t)(e - p);
      while (m && (p[m - 1] == '\r')) m--;
      if (copy) {
        const uint64_t idx = s->first_index + k;
        if (idx < s->limit) {
          if (m != s->len) s->bad = 1;
          else memcpy(s->out + idx * s->len, p, s->len);
        }
      }
      k++;
    }
    p = nl ? nl + 1 : hi;
  }
  s->count = k;
}

/* returns FM_SUCCESS, or an error of loadQueries' contract; *used_mmap tells whether the fast path ran */
int32_t fm_parse_queries_mmap(const char *fn, uint32_t len, uint64_t num, char *out)
{
  int fd = open(fn, O_RDONLY);
  struct stat st;
  const char *base, *end;
  int nth, t, bad = 0;
  slice_t *sl;
  uint64_t total = 0;
  if (fd < 0) return FM_E_OPENING_MFASTA_FILE;
  if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); return num == 0 ? FM_SUCCESS : FM_E_READING_MFASTA_FILE; }
  base = (const char *) mmap(NULL, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (base == MAP_FAILED) return FM_E_NOT_IMPLEMENTED;            /* caller falls back to the stdio reader */
  madvise((void *) base, (size_t) st.st_size, MADV_SEQUENTIAL);
  end = base + st.st_size;
  nth = omp_get_max_threads();
  if ((uint64_t) st.st_size < (1u << 20)) nth = 1;
  sl = (slice_t *) calloc((size_t) nth, sizeof(slice_t));
  if (!sl) { munmap((void *) base, (size_t) st.st_size); return FM_E_ALLOCATING_MFASTA; }

  #pragma omp parallel for schedule(static, 1) num_threads(nth)
  for (t = 0; t < nth; t++) {
    const char *lo = next_line_start(base, base + (uint64_t) st.st_size * t / nth, end);
    const char *hi = next_line_start(base, base + (uint64_t) st.st_size * (t + 1) / nth, end);
    sl[t].len = len; sl[t].out = out; sl[t].limit = num;
    scan_slice(lo, hi, &sl[t], 0);
  }
  for (t = 0; t < nth; t++) { sl[t].first_index = total; total += sl[t].count; }
  if (total < num) { free(sl); munmap((void *) base, (size_t) st.st_size); return FM_E_READING_MFASTA_FILE; }
  #pragma omp parallel for schedule(static, 1) num_threads(nth)
  for (t = 0; t < nth; t++) {
    const char *lo = next_line_start(base, base + (uint64_t) st.st_size * t / nth, end);
    const char *hi = next_line_start(base, base + (uint64_t) st.st_size * (t + 1) / nth, end);
    if (sl[t].first_index < num) scan_slice(lo, hi, &sl[t], 1);
  }
  for (t = 0; t < nth; t++) bad |= sl[t].bad;
  free(sl);
  munmap((void *) base, (size_t) st.st_size);
  return bad ? FM_E_READING_MFASTA_FILE : FM_SUCCESS;
}

/* decimal digits of v, returns the byte after the last digit */
static char *put_u32(char *p, uint32_t v)
{
  char tmp[10];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10u); v /= 10u; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

int32_t fm_write_results_fast(const char *fn, const uint32_t *results, uint32_t num)
{
  FILE *fp = fopen(fn, "w");
  const uint64_t block = 1u << 22;                                  /* reads formatted per round */
  const int nth = omp_get_max_threads();
  char **buf;
  size_t *used;
  uint64_t q0;
  int t, rc = FM_SUCCESS;
  if (fp == NULL) return FM_E_OPENING_RESULTS_FILE;
  buf = (char **) calloc((size_t) nth, sizeof(char *));
  used = (size_t *) calloc((size_t) nth, sizeof(size_t));
  if (!buf || !used) { fclose(fp); free(buf); free(used); return FM_E_ALLOCATING_RESULTS; }
  for (t = 0; t < nth; t++) {
    buf[t] = (char *) malloc((size_t)((block / (uint64_t) nth + 1) * 22 + 16));
    if (!buf[t]) rc = FM_E_ALLOCATING_RESULTS;
  }
  if (rc == FM_SUCCESS) {
    fprintf(fp, "%u\n", num);
    for (q0 = 0; q0 < num && rc == FM_SUCCESS; q0 += block) {
      const uint64_t n = (num - q0 < block) ? num - q0 : block;
      #pragma omp parallel for schedule(static, 1) num_threads(nth)
      for (t = 0; t < nth; t++) {
        const uint64_t a = q0 + n * (uint64_t) t / (uint64_t) nth, b = q0 + n * (uint64_t)(t + 1) / (uint64_t) nth;
        char *p = buf[t];
        uint64_t q;
        for (q = a; q < b; q++) {
          p = put_u32(p, results[2 * q]); *p++ = ' ';
          p = put_u32(p, results[2 * q + 1]); *p++ = '\n';
        }
        used[t] = (size_t)(p - buf[t]);
      }
      for (t = 0; t < nth; t++)
        if (used[t] && fwrite(buf[t], 1, used[t], fp) != used[t]) rc = FM_E_OPENING_RESULTS_FILE;
    }
  }
  for (t = 0; t < nth; t++) free(buf[t]);
  free(buf); free(used);
  if (fclose(fp) != 0 && rc == FM_SUCCESS) rc = FM_E_OPENING_RESULTS_FILE;
  return rc;
}

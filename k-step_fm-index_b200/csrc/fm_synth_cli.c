/*
 * fmsynth -- writes the synthetic reference / reads of fm_synth.h as FASTA
 * files in the formats the reference tools read:
 *   reference: header line "> <n>" then 70-column lines   (writer shape of
 *              common/common.c:95-123; reader common/common.c:42-76)
 *   reads    : ">rid<i> <s+1>-<e+1>" then one sequence line (resources/genreads.py:75-79;
 *              reader common/common.c:167-173)
 *
 * usage: fmsynth ref   <out.fa> <n> <seed>
 *        fmsynth reads <out.fa> <n> <seed_ref> <num> <len> <seed_reads> [first]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fm_synth.h"

static int write_ref(const char *fn, uint64_t n, uint64_t seed)
{
  FILE *fp = fopen(fn, "wb");
  char line[80];
  uint64_t i = 0;
  if (!fp) { perror(fn); return 1; }
  setvbuf(fp, NULL, _IOFBF, 1 << 22);
  fprintf(fp, "> %llu", (unsigned long long) n);
  while (i < n) {
    uint64_t m = (n - i < 70) ? n - i : 70, c;
    line[0] = '\n';
    for (c = 0; c < m; c++) line[1 + c] = fm_synth_base(seed, i + c);
    fwrite(line, 1, (size_t)(m + 1), fp);
    i += m;
  }
  fputc('\n', fp);
  return fclose(fp) != 0;
}

static int write_reads(const char *fn, uint64_t n, uint64_t seed_ref, uint64_t num, uint64_t len,
                       uint64_t seed_reads, uint64_t first)
{
  FILE *fp = fopen(fn, "wb");
  char *line;
  uint64_t j, c;
  if (!fp) { perror(fn); return 1; }
  if (len > n || len > 1000) { fprintf(stderr, "bad read length\n"); return 1; }
  setvbuf(fp, NULL, _IOFBF, 1 << 22);
  line = (char *) malloc((size_t) len + 2);
  for (j = first; j < first + num; j++) {
    uint64_t s = fm_synth_read_start(seed_reads, j, n, len);
    fprintf(fp, ">rid%llu %llu-%llu\n", (unsigned long long)(j + 1),
            (unsigned long long)(s + 1), (unsigned long long)(s + len + 1));
    for (c = 0; c < len; c++) line[c] = fm_synth_base(seed_ref, s + c);
    line[len] = '\n';
    fwrite(line, 1, (size_t)(len + 1), fp);
  }
  free(line);
  return fclose(fp) != 0;
}

int main(int argc, char **argv)
{
  if (argc >= 5 && !strcmp(argv[1], "ref"))
    return write_ref(argv[2], strtoull(argv[3], 0, 10), strtoull(argv[4], 0, 10));
  if (argc >= 8 && !strcmp(argv[1], "reads"))
    return write_reads(argv[2], strtoull(argv[3], 0, 10), strtoull(argv[4], 0, 10), strtoull(argv[5], 0, 10),
                       strtoull(argv[6], 0, 10), strtoull(argv[7], 0, 10), argc > 8 ? strtoull(argv[8], 0, 10) : 0);
  fprintf(stderr, "usage: fmsynth ref <out.fa> <n> <seed>\n"
                  "       fmsynth reads <out.fa> <n> <seed_ref> <num> <len> <seed_reads> [first]\n");
  return 2;
}

/*
 * gfmi_b200 <reference.fa> <n> <k> <d> [--all]
 *
 * The reference's index-build driver (common/generateIndex.c:30-55: loadRef -> buildIndex -> saveIndex) rebuilt on
 * the GPU builder of libfmindex_b200: reads the first <n> bases of a FASTA reference (header line starting with
 * '>', then sequence lines; reader semantics of common/common.c:42-76), builds the k-step index on the GPU and
 * writes "<reference.fa>.<n>.<d>fmi<k>steps.fmi" -- the file name and the bytes gfmiBaseLine_<d>bases_<k>step
 * writes (src/genFMindex.c:155-181).  With --all it also writes the three transformed layouts tfmiBMP_* / tfmiAC_*
 * would produce: ".interleaving", ".ac", ".interleaving.ac".
 *
 * gfmi_b200 --synth <out prefix> <n> <seed> <k> <d> [--all]
 *   the same for the synthetic text of fm_synth.h (what `fmsynth ref <out prefix> <n> <seed>` would write), generated on
 *   the GPU: no FASTA file is read.  Output name "<out prefix>.<n>.<d>fmi<k>steps.fmi".
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/fmindex_b200.h"

static int32_t read_reference(const char *fn, uint64_t n, char **out)
{
  FILE *fp = fopen(fn, "rb");
  char line[1 << 16];
  uint64_t got = 0;
  char *ref;
  if (!fp) return FM_E_OPENING_REFERENCE_FILE;
  ref = (char *) malloc(n ? n : 1);
  if (!ref) { fclose(fp); return FM_E_ALLOCATING_REFERENCE; }
  if (!fgets(line, sizeof line, fp)) { fclose(fp); free(ref); return FM_E_READING_REFERENCE_FILE; }
  if (line[0] != '>') { fclose(fp); free(ref); return FM_E_READING_MFASTA_FILE; }
  while (got < n && fgets(line, sizeof line, fp)) {
    size_t m = strlen(line);
    while (m && (line[m - 1] == '\n' || line[m - 1] == '\r')) m--;
    if (m > n - got) m = (size_t)(n - got);
    memcpy(ref + got, line, m);
    got += m;
  }
  fclose(fp);
  if (got != n) { free(ref); return FM_E_READING_REFERENCE_FILE; }
  *out = ref;
  return FM_SUCCESS;
}

#define CHECK(e) do { int32_t e_ = (e); if (e_) { fprintf(stderr, "%s (%s)\n", errorCommon(e_), fmgpu_build_last_error()); return EXIT_FAILURE; } } while (0)

int main(int argc, char **argv)
{
  char *ref = NULL, name[2048];
  fmgpu_build_t *b = NULL, *t = NULL;
  uint64_t n;
  uint32_t k, d, i;
  static const uint32_t tags[3] = { 101, 200, 201 };
  static const char *suffix[3] = { ".interleaving", ".ac", ".interleaving.ac" };
  double t0;
  const char *prefix;
  int all;
  if (argc >= 7 && !strcmp(argv[1], "--synth")) {
    const uint64_t seed = strtoull(argv[4], NULL, 10);
    n = strtoull(argv[3], NULL, 10); k = (uint32_t) atoi(argv[5]); d = (uint32_t) atoi(argv[6]);
    t0 = sampleTime();
    CHECK(fmgpu_build_from_synth(0, n, seed, k, d, &b));
    printf("BUILD TIME: \t %f \n", sampleTime() - t0);
    prefix = argv[2]; all = argc > 7 && !strcmp(argv[7], "--all");
    goto save;
  }
  if (argc < 5) {
    fprintf(stderr, "usage: %s <reference.fa> <n> <k> <d> [--all]\n       %s --synth <out prefix> <n> <seed> <k> <d> [--all]\n", argv[0], argv[0]);
    return EXIT_FAILURE;
  }
  n = strtoull(argv[2], NULL, 10); k = (uint32_t) atoi(argv[3]); d = (uint32_t) atoi(argv[4]);
  CHECK(read_reference(argv[1], n, &ref));
  t0 = sampleTime();
  CHECK(fmgpu_build_from_text(0, ref, n, k, d, &b));
  printf("BUILD TIME: \t %f \n", sampleTime() - t0);
  free(ref);
  prefix = argv[1]; all = argc > 5 && !strcmp(argv[5], "--all");
save:
  snprintf(name, sizeof name, "%s.%llu.%ufmi%usteps.fmi", prefix, (unsigned long long) n, d, k);
  CHECK(fmgpu_build_save(b, name));
  if (all)
    for (i = 0; i < 3; i++) {
      char tn[2100];
      CHECK(fmgpu_build_transform(b, tags[i], &t));
      snprintf(tn, sizeof tn, "%s%s", name, suffix[i]);
      CHECK(fmgpu_build_save(t, tn));
      fmgpu_build_free(&t);
    }
  fmgpu_build_free(&b);
  return FM_SUCCESS;
}

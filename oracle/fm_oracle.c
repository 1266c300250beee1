/*
 * oracle/fm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, runtime-configured restatement of the reference's k-step FM-index
 * backward search, used ONLY as a checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.  Nothing under k-step_fm-index_b200/ may
 * include, link or call it; the product has no CPU search path.
 *
 * Parity pinning: the reference tree holds no tests or golden vectors
 * (SURVEY.md section 4), so this restatement is pinned against OUTPUTS OF THE
 * REFERENCE ITSELF: tests/test_oracle.py runs the unmodified reference CPU
 * searchers (oracle/_ref, built by oracle/Makefile from /root/reference) on
 * seeded inputs and requires identical (L,R) for k in {1,2}, d in {32,64,128},
 * tags 100/101/200/201, including the '$'-row and AltCounters padding-entry
 * corner cases; md5s of those outputs are committed under tests/golden/.
 *
 * What each function follows (paths relative to /root/reference):
 *   fmo_load_index     header + entry read   src/fmIndexCPUBaseline.c:71-143
 *                                            src/fmIndexCPUBaseline-AltCounters.c:71-143
 *                      writer side           src/genFMindex.c:155-181
 *   fmo_plane_word     tag 100/200 planes    src/genFMindex.c:427-455 (bwt2bin)
 *                      tag 101/201 planes    src/transformIndexBitmaps.c:275-282
 *                                            src/transformIndexAlternateCounters.c:396-403
 *   fmo_counter        std counters last     src/fmIndexCPUBaseline.c:49-52
 *                      AC counters first     src/fmIndexCPUBaseline-AltCounters.c:49-52
 *   fmo_base_code      ASCII -> 2 bit        src/fmIndexCPUBaseline.c:213-226
 *   fmo_lf (std)       one rank + '$' fix    src/fmIndexCPUBaseline.c:227-257
 *   fmo_lf (AC)        fwd/bwd counting      src/fmIndexCPUBaseline-AltCounters.c:218-266
 *   fmo_search         query loop            src/fmIndexCPUBaseline.c:195-290
 *   fmo_load_queries   FASTA reader          common/common.c:167-173
 *   fmo_write_results  "(L R)" text          common/common.c:201-220
 *
 * Deliberate differences from the reference (none changes a defined result):
 *   - k, d, tag are read from the file header instead of -D macros;
 *   - sizes are 64-bit (the reference overflows at num*len >= 2^32, common/common.c:163);
 *   - tags 101/201 are searched directly (the reference CPU code only indexes
 *     the 100/200 word order); expected results are those of the 100/200 file
 *     of the same index, which is what tests check;
 *   - no shift by >= 32 is ever evaluated (the reference computes
 *     0xFFFFFFFF << (32-shift) unconditionally and discards it, :235-237);
 *   - X = bwtsize with bwtsize % d == 0 reads past the last entry in the
 *     reference; here it counts the whole last chunk.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fm_oracle.h"

enum { FMO_OK = 0, FMO_E_OPEN = 1, FMO_E_ALLOC = 3, FMO_E_READ = 5, FMO_E_FORMAT = 19 };

static int tag_is_ac(uint32_t tag)          { return tag == 200 || tag == 201; }
static int tag_is_interleaved(uint32_t tag) { return tag == 101 || tag == 201; }

static int32_t parse_header(const uint32_t *h, fmo_index_t *x)
{
  uint32_t k, i, nsym;
  x->tag = h[0]; x->steps = h[1]; x->bwtsize = h[2];
  x->ncounters = h[3]; x->nentries = h[4]; x->chunk = h[5];
  k = x->steps;
  if (!(x->tag == 100 || x->tag == 101 || x->tag == 200 || x->tag == 201)) return FMO_E_FORMAT;
  if (k < 1 || k > FMO_MAX_STEPS || x->chunk == 0 || (x->chunk % 32) != 0) return FMO_E_FORMAT;
  nsym = 1u << (2 * k);
  if (x->ncounters != (tag_is_ac(x->tag) ? nsym / 2 : nsym)) return FMO_E_FORMAT;
  for (i = 0; i < k; i++) {
    x->dollarPositionBWT[i] = h[6 + i];
    x->dollarBaseBWT[i]     = h[6 + k + i];
  }
  x->entry_words = 2 * (x->chunk / 32) * k + x->ncounters;
  return FMO_OK;
}

int32_t fmo_load_index(const char *fn, fmo_index_t **out)
{
  FILE *fp = fopen(fn, "rb");
  uint32_t head[6 + 2 * FMO_MAX_STEPS];
  fmo_index_t *x;
  size_t nwords;
  int32_t e;
  if (!fp) return FMO_E_OPEN;
  x = (fmo_index_t *) calloc(1, sizeof(*x));
  if (!x) { fclose(fp); return FMO_E_ALLOC; }
  if (fread(head, 4, 6, fp) != 6 || head[1] < 1 || head[1] > FMO_MAX_STEPS ||
      fread(head + 6, 4, 2 * head[1], fp) != 2 * head[1]) { fclose(fp); free(x); return FMO_E_READ; }
  if ((e = parse_header(head, x)) != FMO_OK) { fclose(fp); free(x); return e; }
  nwords = (size_t) x->nentries * x->entry_words;
  x->entries = (uint32_t *) malloc(nwords * 4 + 16);
  if (!x->entries) { fclose(fp); free(x); return FMO_E_ALLOC; }
  if (fread(x->entries, 4, nwords, fp) != nwords) { fclose(fp); free(x->entries); free(x); return FMO_E_READ; }
  x->owns_entries = 1;
  fclose(fp);
  *out = x;
  return FMO_OK;
}

/* index image already in memory (same bytes as the file) */
int32_t fmo_wrap_image(const uint32_t *image, uint64_t nwords, fmo_index_t **out)
{
  fmo_index_t *x = (fmo_index_t *) calloc(1, sizeof(*x));
  int32_t e;
  if (!x) return FMO_E_ALLOC;
  if (nwords < 8 || (e = parse_header(image, x)) != FMO_OK) { free(x); return FMO_E_FORMAT; }
  if (nwords < 6 + 2ull * x->steps + (uint64_t) x->nentries * x->entry_words) { free(x); return FMO_E_READ; }
  x->entries = (uint32_t *) (image + 6 + 2 * x->steps);
  x->owns_entries = 0;
  *out = x;
  return FMO_OK;
}

void fmo_free_index(fmo_index_t *x)
{
  if (!x) return;
  if (x->owns_entries) free(x->entries);
  free(x);
}

/* A=0 C=1 G=2 T=3 from ASCII bits 2 and 1; case-insensitive, other bytes alias */
uint32_t fmo_base_code(uint32_t c)
{
  uint32_t hi = (c >> 2) & 1u, mid = (c >> 1) & 1u;
  return (hi << 1) | (hi ^ mid);
}

/* word n (32 BWT rows) of plane `bit` (0 = low code bit) of BWT layer `step` */
uint32_t fmo_plane_word(const fmo_index_t *x, uint32_t entry, uint32_t step, uint32_t bit, uint32_t n)
{
  const uint32_t W = x->chunk / 32, k = x->steps;
  const uint32_t *e = x->entries + (size_t) entry * x->entry_words;
  const uint32_t *bm = tag_is_ac(x->tag) ? e + x->ncounters : e;
  if (tag_is_interleaved(x->tag)) return bm[2 * k * n + 2 * step + bit];
  return bm[2 * W * step + W * bit + n];
}

uint32_t fmo_counter(const fmo_index_t *x, uint32_t entry, uint32_t slot)
{
  const uint32_t *e = x->entries + (size_t) entry * x->entry_words;
  return tag_is_ac(x->tag) ? e[slot] : e[2 * (x->chunk / 32) * x->steps + slot];
}

/* rows of `entry` strictly before offset r (forward) or at/after r (backward)
 * whose k-step symbol, as stored in the planes, equals sigma */
static uint32_t count_matches(const fmo_index_t *x, uint32_t entry, uint32_t sigma, uint32_t r, int backward)
{
  const uint32_t W = x->chunk / 32;
  uint32_t n, s, total = 0;
  for (n = 0; n < W; n++) {
    int32_t rem = (int32_t) r - (int32_t)(32 * n);
    uint32_t sel = rem <= 0 ? 0u : (rem >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> rem));
    if (backward) sel = ~sel;
    for (s = 0; s < x->steps; s++) {
      uint32_t c  = (sigma >> (2 * s)) & 3u;
      uint32_t p0 = fmo_plane_word(x, entry, s, 0, n);
      uint32_t p1 = fmo_plane_word(x, entry, s, 1, n);
      sel &= ((c & 1u) ? p0 : ~p0) & ((c & 2u) ? p1 : ~p1);
    }
    total += (uint32_t) __builtin_popcount(sel);
  }
  return total;
}

/* one k-step LF mapping of row boundary X for symbol sigma */
uint32_t fmo_lf(const fmo_index_t *x, uint32_t sigma, uint32_t X)
{
  const uint32_t d = x->chunk;
  const uint32_t nstd = tag_is_ac(x->tag) ? x->nentries - 1 : x->nentries;
  uint32_t e = X / d, r = X % d;
  uint32_t s, cnt, fix = 0;
  /* bwtsize % d == 0: the reference reads one entry past the end for X = bwtsize
   * (undefined).  Defined here as "count the whole last chunk". */
  if (e >= nstd) { e = nstd - 1; r = X - e * d; }
  if (!tag_is_ac(x->tag)) {
    cnt = count_matches(x, e, sigma, r, 0);
    for (s = 0; s < x->steps; s++)
      if (x->dollarPositionBWT[s] / d == e && sigma == x->dollarBaseBWT[s] && X > x->dollarPositionBWT[s]) fix++;
    return fmo_counter(x, e, sigma) + (cnt - fix);
  } else {
    const uint32_t H = x->ncounters;               /* counters per entry = 4^k / 2 */
    const int next = ((e & 1u) && sigma < H) || (!(e & 1u) && sigma >= H);
    cnt = count_matches(x, e, sigma, r, next);
    for (s = 0; s < x->steps; s++)
      if (x->dollarPositionBWT[s] / d == e && sigma == x->dollarBaseBWT[s]) {
        if (!next && X >  x->dollarPositionBWT[s]) fix++;
        if ( next && X <= x->dollarPositionBWT[s]) fix++;
      }
    if (next) return fmo_counter(x, e + 1, sigma & (H - 1)) - (cnt - fix);
    return fmo_counter(x, e, sigma & (H - 1)) + (cnt - fix);
  }
}

/* k-step symbol consumed at query offset j: base j in the low 2 bits, j-1 above it, ... */
static uint32_t symbol_at(const char *q, int64_t j, uint32_t k)
{
  uint32_t s, sigma = 0;
  for (s = 0; s < k; s++) sigma |= fmo_base_code((uint32_t)(unsigned char) q[j - (int64_t) s]) << (2 * s);
  return sigma;
}

void fmo_search(const fmo_index_t *x, const char *queries, uint64_t num, uint32_t len, uint32_t *results)
{
  int64_t qi;
  #pragma omp parallel for schedule(static)
  for (qi = 0; qi < (int64_t) num; qi++) {
    const char *q = queries + (uint64_t) qi * len;
    uint32_t L = 0, R = x->bwtsize;
    int64_t j;
    for (j = (int64_t) len - 1; j >= 0; j -= x->steps) {
      uint32_t sigma = symbol_at(q, j, x->steps);
      L = fmo_lf(x, sigma, L);
      R = fmo_lf(x, sigma, R);
    }
    results[2 * qi] = L;
    results[2 * qi + 1] = R;
  }
}

/* Necessary 32-byte sectors of a one-load-per-rank layout:
 * sum over LF steps of |{sector(L), sector(R)}| where a sector holds
 * `blocks_per_sector` consecutive blocks of `block_positions` BWT rows of the
 * SAME symbol. */
uint64_t fmo_count_sectors(const fmo_index_t *x, const char *queries, uint64_t num, uint32_t len,
                           uint32_t block_positions, uint32_t blocks_per_sector)
{
  uint64_t total = 0;
  int64_t qi;
  const uint32_t span = block_positions * blocks_per_sector;
  #pragma omp parallel for schedule(static) reduction(+:total)
  for (qi = 0; qi < (int64_t) num; qi++) {
    const char *q = queries + (uint64_t) qi * len;
    uint32_t L = 0, R = x->bwtsize;
    int64_t j;
    for (j = (int64_t) len - 1; j >= 0; j -= x->steps) {
      uint32_t sigma = symbol_at(q, j, x->steps);
      total += (L / span == R / span) ? 1 : 2;
      L = fmo_lf(x, sigma, L);
      R = fmo_lf(x, sigma, R);
    }
  }
  return total;
}

/* multi-FASTA: header lines start with '>', every other line is one read and
 * its trailing newline is dropped; reads are concatenated without terminators */
int32_t fmo_load_queries(const char *fn, uint32_t len, uint64_t num, char *out)
{
  FILE *fp = fopen(fn, "rb");
  char line[2048];
  uint64_t got = 0;
  if (!fp) return FMO_E_OPEN;
  while (got < num && fgets(line, sizeof line, fp)) {
    size_t m;
    if (line[0] == '>') continue;
    m = strlen(line);
    while (m && (line[m - 1] == '\n' || line[m - 1] == '\r')) m--;
    if (m != len) { fclose(fp); return FMO_E_FORMAT; }
    memcpy(out + got * len, line, len);
    got++;
  }
  fclose(fp);
  return got == num ? FMO_OK : FMO_E_READ;
}

int32_t fmo_write_results(const char *fn, const uint32_t *results, uint32_t num)
{
  FILE *fp = fopen(fn, "w");
  uint32_t i;
  if (!fp) return FMO_E_OPEN;
  fprintf(fp, "%u\n", num);
  for (i = 0; i < num; i++) fprintf(fp, "%u %u\n", results[2 * i], results[2 * i + 1]);
  return fclose(fp) ? FMO_E_READ : FMO_OK;
}

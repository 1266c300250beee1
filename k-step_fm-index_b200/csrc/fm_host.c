/*
 * fm_host.c -- PART 1 of include/fmindex_b200.h: the reference's plugin
 * surface for the search path, host side, plain C.
 *
 * File readers/writers keep the reference's formats byte for byte
 * (SURVEY.md App. A); what is new is run-time configuration (k, d, flavour
 * from the file header), 64-bit sizes, pinned query/result buffers and a
 * multi-GPU driver (replicate the index, shard the batch) behind the same
 * six *GPU entry points.  Everything device-side goes through the fmgpu_*
 * functions of fm_index.cu, fm_search.cu, fm_sparsetab.cu, ... (fm_internal.h); this file contains no search arithmetic.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/fmindex_b200.h"

#define FM_MAX_GPUS 16

/* fm_ingest.c */
int32_t fm_parse_queries_mmap(const char *fn, uint32_t len, uint64_t num, char *out);
int32_t fm_write_results_fast(const char *fn, const uint32_t *results, uint32_t num);

/* device-side state hung off fmi_t.d_index */
typedef struct {
  int32_t        ndev;
  int32_t        dev[FM_MAX_GPUS];
  fmgpu_index_t *replica[FM_MAX_GPUS];
} fm_replica_set_t;

/* device-side state hung off qrys_t.d_queries and res_t.d_results: ONE shard set per (queries, results) pair, owned
 * by both handles; it is released when both have let go (free*GPU) or when either is transferred again */
typedef struct {
  int32_t        ndev;
  uint64_t       first[FM_MAX_GPUS + 1];   /* shard g = reads [first[g], first[g+1]) */
  fmgpu_batch_t *shard[FM_MAX_GPUS];
  fm_replica_set_t *index;                 /* replicas the shards were created for */
  qrys_t        *owner_q;                  /* the handles whose d_* fields point here (NULL once released) */
  res_t         *owner_r;
} fm_shard_set_t;

static int32_t g_ndev = 0;
static int32_t g_dev[FM_MAX_GPUS];
static fmgpu_variant_t g_variant = { 0, 0, 0, 0 };
static int g_variant_set = 0;
static fmgpu_transfer_stats_t g_stats;

/* common/common.c:28-33 */
double sampleTime(void)
{
  struct timespec tv;
  clock_gettime(CLOCK_REALTIME, &tv);
  return (double) tv.tv_sec + (double) tv.tv_nsec / 1e9;
}

/* messages of common/common.c:282-310, plus the new codes */
char *errorCommon(int32_t e)
{
  switch (e) {
    case FM_SUCCESS:                    return "No error";
    case FM_E_OPENING_INDEX_FILE:       return "Cannot open index file";
    case FM_E_ALLOCATING_BWT:           return "Cannot allocate memory for bwt";
    case FM_E_ALLOCATING_FMI:           return "Cannot allocate memory for counters";
    case FM_E_READING_BWT:              return "Error reading index bwt";
    case FM_E_READING_FMI:              return "Error reading index counters";
    case FM_E_SAVING_INDEX_FILE:        return "Cannot open index file for save";
    case FM_E_SAVING_BWT_FILE:          return "Cannot open bwt file for save";
    case FM_E_BUILDING_BWT:             return "Error building bwt";
    case FM_E_BUILDING_FMI:             return "Error building FMI, cannot allocate memory for bwt";
    case FM_E_OPENING_REFERENCE_FILE:   return "Cannot open reference file";
    case FM_E_ALLOCATING_REFERENCE:     return "Cannot allocate reference";
    case FM_E_READING_MFASTA_FILE:      return "Reference file isn't MFASTA format";
    case FM_E_READING_REFERENCE_FILE:   return "Error reading reference file";
    case FM_E_OPENING_MFASTA_FILE:      return "Cannot open MFASTS queries file";
    case FM_E_ALLOCATING_MFASTA:        return "Cannot allocate MFASTA queries";
    case FM_E_ALLOCATING_RESULTS:       return "Cannot allocate results";
    case FM_E_OPENING_RESULTS_FILE:     return "Cannot open results file for load intervals";
    case FM_E_READING_RESULTS_FILE:     return "Error reading results";
    case FM_E_NOT_IMPLEMENTED:          return "Not implemented";
    case FM_E_CUDA:                     return (char *) fmgpu_last_error();
    case FM_E_BAD_ARGUMENT:             return (char *) fmgpu_last_error();
    case FM_E_UNSUPPORTED_INDEX:        return "Unsupported index (needs k in {1,2,3,4}, d a multiple of 32, no active AltCounters quirk for k >= 3)";
    case FM_E_QUERY_SHAPE:              return "Read length must be a positive multiple of k";
    case FM_E_INDEX_VER_BASELINE:       return "Error in the index type, use gfmiBaseLine_*Bases_*Step to generate an index_name.fmi type";
    case FM_E_INDEX_VER_INTERLEAVE:     return "Error in the index type, use tfmiBMP_*Bases_*Step to generate an index_name.fmi.interleaving type";
    case FM_E_INDEX_VER_BASELINE_AC:    return "Error in the index type, use tfmiAC_*Bases_*Step to generate an index_name.fmi.ac type";
    case FM_E_INDEX_VER_INTERLEAVE_AC:  return "Error in the index type, use tfmiAC_*Bases_*Step to generate an index_name.fmi.interleaving.ac type";
    default:                            return "Unknown error";
  }
}

/* ------------------------------------------------------------------------ *
 * index file: u32 tag, k, bwtsize, ncounters, nentries, d, $pos[k], $base[k],
 * then nentries entries (writer: src/genFMindex.c:155-181)
 * ------------------------------------------------------------------------ */
int32_t loadIndex(const char *fn, void **index)
{
  FILE *fp;
  fmi_t *fmi;
  uint32_t head[6], i;
  size_t nwords;

  if (!fn || !index) return FM_E_OPENING_INDEX_FILE;
  fp = fopen(fn, "rb");
  if (fp == NULL) return FM_E_OPENING_INDEX_FILE;
  if (fread(head, sizeof(uint32_t), 6, fp) != 6) { fclose(fp); return FM_E_READING_FMI; }

  if (!(head[0] == 100 || head[0] == 101 || head[0] == 200 || head[0] == 201)) {
    fclose(fp);
    return FM_E_INDEX_VER_BASELINE;          /* not an index of this family: ask for a .fmi */
  }
  if (head[1] < 1 || head[1] > 4 || head[5] == 0 || (head[5] % 32) != 0) { fclose(fp); return FM_E_UNSUPPORTED_INDEX; }
  if (head[3] != (1u << (2 * head[1])) >> (head[0] >= 200 ? 1 : 0)) { fclose(fp); return FM_E_READING_FMI; }

  fmi = (fmi_t *) calloc(1, sizeof(fmi_t));
  if (fmi == NULL) { fclose(fp); return FM_E_ALLOCATING_FMI; }
  fmi->tag = head[0]; fmi->steps = head[1]; fmi->bwtsize = head[2];
  fmi->ncounters = head[3]; fmi->nentries = head[4]; fmi->chunk = head[5];
  fmi->nbitmaps = 2 * (fmi->chunk / 32);
  fmi->entry_words = fmi->nbitmaps * fmi->steps + fmi->ncounters;
  printf("Index Version: %u\n", fmi->tag);
  printf("Steps (k): %u \n", fmi->steps);
  printf("Reference Size: %u \n", fmi->bwtsize - 1);
  printf("rLF counters: %u \n", fmi->ncounters);
  printf("F entries: %u \n", fmi->nentries);
  printf("d Sampling: %u \n", fmi->chunk);

  fmi->h_dollarPositionBWT = (uint32_t *) malloc(fmi->steps * sizeof(uint32_t));
  fmi->h_dollarBaseBWT     = (uint32_t *) malloc(fmi->steps * sizeof(uint32_t));
  fmi->h_modposdollarBWT   = (uint32_t *) malloc(fmi->steps * sizeof(uint32_t));
  nwords = (size_t) fmi->nentries * fmi->entry_words;
  fmi->h_index = malloc(nwords * sizeof(uint32_t) + 16);
  if (!fmi->h_dollarPositionBWT || !fmi->h_dollarBaseBWT || !fmi->h_modposdollarBWT || !fmi->h_index) {
    fclose(fp); freeIndex((void **) &fmi); free(fmi); return FM_E_ALLOCATING_FMI;
  }
  if (fread(fmi->h_dollarPositionBWT, sizeof(uint32_t), fmi->steps, fp) != fmi->steps ||
      fread(fmi->h_dollarBaseBWT, sizeof(uint32_t), fmi->steps, fp) != fmi->steps ||
      fread(fmi->h_index, sizeof(uint32_t), nwords, fp) != nwords) {
    fclose(fp); freeIndex((void **) &fmi); free(fmi); return FM_E_READING_FMI;
  }
  for (i = 0; i < fmi->steps; i++) fmi->h_modposdollarBWT[i] = fmi->h_dollarPositionBWT[i] / fmi->chunk;
  fclose(fp);
  *index = fmi;
  return FM_SUCCESS;
}

/* like the reference, releases the inner host buffers and leaves the handle */
int32_t freeIndex(void **index)
{
  fmi_t *fmi;
  if (!index || !*index) return FM_SUCCESS;
  fmi = (fmi_t *) *index;
  free(fmi->h_index);              fmi->h_index = NULL;
  free(fmi->h_dollarPositionBWT);  fmi->h_dollarPositionBWT = NULL;
  free(fmi->h_dollarBaseBWT);      fmi->h_dollarBaseBWT = NULL;
  free(fmi->h_modposdollarBWT);    fmi->h_modposdollarBWT = NULL;
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ *
 * queries: multi-FASTA, one sequence line per read (common/common.c:167-173).
 * The buffer is pinned so transferCPUtoGPU / fmgpu_search_host copy at PCIe
 * speed; a line whose length is not sizeQuery is an error instead of the
 * reference's silent mis-stride.
 * ------------------------------------------------------------------------ */
int32_t loadQueries(char *fn, uint32_t sizeQuery, uint32_t numQueries, void **queries)
{
  FILE *fp;
  qrys_t *qrys;
  char line[2048];
  uint64_t got = 0;
  const uint64_t bytes = (uint64_t) sizeQuery * numQueries;

  if (!fn || !queries || sizeQuery == 0 || sizeQuery > 2000) return FM_E_READING_MFASTA_FILE;
  fp = fopen(fn, "rb");
  if (fp == NULL) return FM_E_OPENING_MFASTA_FILE;
  qrys = (qrys_t *) calloc(1, sizeof(qrys_t));
  if (!qrys) { fclose(fp); return FM_E_ALLOCATING_MFASTA; }
  qrys->num = numQueries; qrys->size = sizeQuery;
  qrys->h_queries = (char *) fmgpu_host_alloc(bytes ? bytes : 1);
  if (qrys->h_queries == NULL) { fclose(fp); free(qrys); return FM_E_ALLOCATING_MFASTA; }
  {
    /* fast path: mmap + all threads (fm_ingest.c); same contract as the stdio loop below, which stays as the
     * fallback for files that cannot be mapped */
    const int32_t fast = fm_parse_queries_mmap(fn, sizeQuery, numQueries, qrys->h_queries);
    if (fast != FM_E_NOT_IMPLEMENTED) {
      fclose(fp);
      if (fast != FM_SUCCESS) { fmgpu_host_free(qrys->h_queries); free(qrys); return fast; }
      *queries = qrys;
      return FM_SUCCESS;
    }
  }
  while (got < numQueries && fgets(line, sizeof line, fp) != NULL) {
    size_t m;
    if (line[0] == '>') continue;
    m = strlen(line);
    while (m && (line[m - 1] == '\n' || line[m - 1] == '\r')) m--;
    if (m != sizeQuery) { fclose(fp); fmgpu_host_free(qrys->h_queries); free(qrys); return FM_E_READING_MFASTA_FILE; }
    memcpy(qrys->h_queries + got * sizeQuery, line, sizeQuery);
    got++;
  }
  fclose(fp);
  if (got != numQueries) { fmgpu_host_free(qrys->h_queries); free(qrys); return FM_E_READING_MFASTA_FILE; }
  *queries = qrys;
  return FM_SUCCESS;
}

int32_t freeQueries(void **queries)
{
  qrys_t *qrys;
  if (!queries || !*queries) return FM_SUCCESS;
  qrys = (qrys_t *) *queries;
  if (qrys->h_queries) { fmgpu_host_free(qrys->h_queries); qrys->h_queries = NULL; }
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ *
 * results: 2*num u32, zeroed (common/common.c:248-260); text dump
 * "<num>\n" then "<L> <R>\n" (common/common.c:201-220)
 * ------------------------------------------------------------------------ */
int32_t initResults(uint32_t numresults, void **results)
{
  res_t *res;
  const size_t bytes = 2 * (size_t) numresults * sizeof(uint32_t);
  if (!results) return FM_E_ALLOCATING_RESULTS;
  res = (res_t *) calloc(1, sizeof(res_t));
  if (!res) return FM_E_ALLOCATING_RESULTS;
  res->num = numresults;
  res->h_results = (uint32_t *) fmgpu_host_alloc(bytes ? bytes : 1);
  if (res->h_results == NULL) { free(res); return FM_E_ALLOCATING_RESULTS; }
  memset(res->h_results, 0, bytes);
  *results = res;
  return FM_SUCCESS;
}

int32_t freeResults(void **results)
{
  res_t *res;
  if (!results || !*results) return FM_SUCCESS;
  res = (res_t *) *results;
  if (res->h_results) { fmgpu_host_free(res->h_results); res->h_results = NULL; }
  return FM_SUCCESS;
}

int32_t writeResults(char *fn, uint32_t *results, uint32_t numqueries)
{
  if (!fn || !results) return FM_E_OPENING_RESULTS_FILE;
  return fm_write_results_fast(fn, results, numqueries);     /* all threads format, bytes identical to the reference's */
}

int32_t loadResults(char *fn, void **results)
{
  FILE *fp;
  uint32_t num, i;
  int32_t err;
  res_t *res;
  if (!fn || !results) return FM_E_OPENING_RESULTS_FILE;
  fp = fopen(fn, "r");
  if (fp == NULL) return FM_E_OPENING_RESULTS_FILE;
  if (fscanf(fp, "%u\n", &num) != 1) { fclose(fp); return FM_E_READING_RESULTS_FILE; }
  if ((err = initResults(num, results)) != FM_SUCCESS) { fclose(fp); return err; }
  res = (res_t *) *results;
  for (i = 0; i < num; i++)
    if (fscanf(fp, "%u %u\n", &res->h_results[2 * (size_t) i], &res->h_results[2 * (size_t) i + 1]) != 2) {
      fclose(fp); freeResults(results); free(res); *results = NULL; return FM_E_READING_RESULTS_FILE;
    }
  fclose(fp);
  return FM_SUCCESS;
}

/* "<fn>.res.gpu": the name a reference CUDA build writes (common/common.c:331-332) */
int32_t saveResults(char *fn, void *results, void *index)
{
  res_t *res = (res_t *) results;
  char out[1024];
  (void) index;
  if (!fn || !res) return FM_E_OPENING_RESULTS_FILE;
  snprintf(out, sizeof out, "%s.res.gpu", fn);
  return writeResults(out, res->h_results, res->num);
}

/* ------------------------------------------------------------------------ *
 * host driver behind the reference's six *GPU symbols
 * ------------------------------------------------------------------------ */
int32_t fmgpu_set_devices(const int32_t *devices, int32_t ndevices)
{
  int32_t i;
  if (ndevices < 0 || ndevices > FM_MAX_GPUS || (ndevices && !devices)) return FM_E_BAD_ARGUMENT;
  for (i = 0; i < ndevices; i++) g_dev[i] = devices[i];
  g_ndev = ndevices;
  return FM_SUCCESS;
}

int32_t fmgpu_set_variant(const fmgpu_variant_t *v)
{
  if (v) { g_variant = *v; g_variant_set = 1; } else g_variant_set = 0;
  return FM_SUCCESS;
}

/* $FMGPU_MODE = task | coop | fused | sparse | wide | auto (default): kernel family used by searchIndexGPU */
#define FM_MODE_AUTO (-1)
static int fm_mode_from_env(void)
{
  const char *env = getenv("FMGPU_MODE");
  if (!env || !*env || !strcmp(env, "auto")) return FM_MODE_AUTO;
  if (!strcmp(env, "task")) return FMGPU_MODE_TASK;
  if (!strcmp(env, "coop")) return FMGPU_MODE_COOP;
  if (!strcmp(env, "fused")) return FMGPU_MODE_FUSED;
  if (!strcmp(env, "sparse")) return FMGPU_MODE_SPARSE;
  if (!strcmp(env, "wide")) return FMGPU_MODE_WIDE;
  return FM_MODE_AUTO;
}

/* configured devices, else $FMGPU_DEVICES ("0,1,2,3"), else device 0 */
static int32_t fm_resolve_devices(int32_t *dev)
{
  const char *env;
  int32_t n = 0;
  if (g_ndev > 0) { memcpy(dev, g_dev, g_ndev * sizeof(int32_t)); return g_ndev; }
  env = getenv("FMGPU_DEVICES");
  if (env && *env) {
    const char *p = env;
    while (*p && n < FM_MAX_GPUS) {
      char *end;
      long d = strtol(p, &end, 10);
      if (end == p) break;
      dev[n++] = (int32_t) d;
      p = (*end == ',') ? end + 1 : end;
    }
    if (n) return n;
  }
  dev[0] = 0;
  return 1;
}

/* detaches both owners (their d_* fields never dangle) and frees the shards */
static void fm_release_shards(fm_shard_set_t *ss)
{
  int32_t g;
  if (!ss) return;
  if (ss->owner_q && ss->owner_q->d_queries == (char *) ss) ss->owner_q->d_queries = NULL;
  if (ss->owner_r && ss->owner_r->d_results == (uint32_t *) ss) ss->owner_r->d_results = NULL;
  for (g = 0; g < ss->ndev; g++) fmgpu_batch_free(&ss->shard[g]);
  free(ss);
}

static double fm_wall(void)
{
  struct timespec tv;
  clock_gettime(CLOCK_MONOTONIC, &tv);
  return (double) tv.tv_sec + (double) tv.tv_nsec * 1e-9;
}

/* $FMGPU_STATS_FILE: the stats as one JSON line, written when the index leaves the GPUs (freeIndexGPU) -- the way to get
 * them out of a caller that cannot ask, like the reference's own main() */
static void fm_dump_stats(void)
{
  const char *fn = getenv("FMGPU_STATS_FILE");
  FILE *fp;
  int32_t g;
  if (!fn || !*fn || g_stats.ndev == 0) return;
  fp = fopen(fn, "a");
  if (!fp) return;
  fprintf(fp, "{\"ndev\": %d, \"searches\": %d, \"index_file_bytes\": %llu, \"table_bytes\": %llu, \"index_h2d_reblock_s\": %.6f, "
              "\"index_h2d_reblock_gbs\": %.3f, \"queries_h2d_pack_s\": %.6f, \"query_bytes\": %llu, \"results_d2h_s\": %.6f, \"result_bytes\": %llu",
          g_stats.ndev, g_stats.searches, (unsigned long long) g_stats.index_file_bytes, (unsigned long long) g_stats.table_bytes,
          g_stats.index_h2d_reblock_s, g_stats.index_h2d_reblock_s > 0 ? g_stats.index_file_bytes / g_stats.index_h2d_reblock_s / 1e9 : 0.0,
          g_stats.queries_h2d_pack_s, (unsigned long long) g_stats.query_bytes, g_stats.results_d2h_s, (unsigned long long) g_stats.result_bytes);
  fprintf(fp, ", \"context_init_s\": [");
  for (g = 0; g < g_stats.ndev; g++) fprintf(fp, "%s%.4f", g ? ", " : "", g_stats.context_init_s[g]);
  fprintf(fp, "], \"peer_copy_s\": [");
  for (g = 1; g < g_stats.ndev; g++) fprintf(fp, "%s%.6f", g > 1 ? ", " : "", g_stats.peer_copy_s[g]);
  fprintf(fp, "], \"replicate_s\": [");
  for (g = 1; g < g_stats.ndev; g++) fprintf(fp, "%s%.6f", g > 1 ? ", " : "", g_stats.replicate_s[g]);
  fprintf(fp, "], \"peer_copy_gbs\": [");
  for (g = 1; g < g_stats.ndev; g++) fprintf(fp, "%s%.2f", g > 1 ? ", " : "", g_stats.peer_copy_s[g] > 0 ? g_stats.table_bytes / g_stats.peer_copy_s[g] / 1e9 : 0.0);
  fprintf(fp, "], \"table_build_s\": [");
  for (g = 0; g < g_stats.ndev; g++) fprintf(fp, "%s%.4f", g ? ", " : "", g_stats.table_build_s[g]);
  fprintf(fp, "], \"search_ms_per_gpu\": [");
  for (g = 0; g < g_stats.ndev; g++) fprintf(fp, "%s%.4f", g ? ", " : "", g_stats.search_ms[g]);
  fprintf(fp, "]}\n");
  fclose(fp);
}

int32_t fmgpu_get_transfer_stats(fmgpu_transfer_stats_t *out)
{
  if (!out) return FM_E_BAD_ARGUMENT;
  *out = g_stats;
  return FM_SUCCESS;
}

int32_t transferCPUtoGPU(void *index, void *queries, void *results)
{
  fmi_t *fmi = (fmi_t *) index;
  qrys_t *qrys = (qrys_t *) queries;
  res_t *res = (res_t *) results;
  fm_replica_set_t *rs;
  fm_shard_set_t *ss;
  int32_t g, err;
  uint64_t per;

  if (!fmi || !qrys || !res || !fmi->h_index || !qrys->h_queries || !res->h_results) return FM_E_BAD_ARGUMENT;
  if (res->num < qrys->num) return FM_E_BAD_ARGUMENT;
  if (qrys->size % fmi->steps && fmi->steps != 2) return FM_E_QUERY_SHAPE;   /* k=2: the device layer serves odd lengths */

  /* index: one H2D + re-block on the first GPU, peer copies to the others */
  rs = (fm_replica_set_t *) fmi->d_index;
  if (rs == NULL) {
    double t0;
    fmgpu_index_meta_t meta0;
    memset(&g_stats, 0, sizeof g_stats);
    rs = (fm_replica_set_t *) calloc(1, sizeof(*rs));
    if (!rs) return FM_E_ALLOCATING_FMI;
    rs->ndev = fm_resolve_devices(rs->dev);
    g_stats.ndev = rs->ndev;
    /* CUDA contexts of all devices first (0.3-0.4 s each on a cold process), so that the copy timings below are copies */
    for (g = 0; g < rs->ndev; g++) {
      t0 = fm_wall();
      err = fmgpu_device_warmup(rs->dev[g]);
      g_stats.context_init_s[g] = fm_wall() - t0;
      if (err) { free(rs); return err; }
    }
    t0 = fm_wall();
    err = fmgpu_index_create(rs->dev[0], fmi->tag, fmi->steps, fmi->chunk, fmi->bwtsize, fmi->ncounters, fmi->nentries,
                             fmi->h_dollarPositionBWT, fmi->h_dollarBaseBWT, (const uint32_t *) fmi->h_index, &rs->replica[0]);
    g_stats.index_h2d_reblock_s = fm_wall() - t0;
    g_stats.index_file_bytes = (uint64_t) fmi->nentries * fmi->entry_words * 4ull;
    for (g = 1; g < rs->ndev && !err; g++) {
      t0 = fm_wall();
      err = fmgpu_index_replicate(rs->replica[0], rs->dev[g], &rs->replica[g]);
      g_stats.replicate_s[g] = fm_wall() - t0;                /* allocation + peer mapping + copy */
      g_stats.peer_copy_s[g] = fmgpu_last_peer_copy_seconds();   /* the cudaMemcpyPeer alone */
    }
    if (err) { for (g = 0; g < rs->ndev; g++) fmgpu_index_free(&rs->replica[g]); free(rs); return err; }
    if (fmgpu_index_get_meta(rs->replica[0], &meta0) == FM_SUCCESS) g_stats.table_bytes = meta0.nbytes;
    /* derived table on every replica unless $FMGPU_MODE asks for the plain kernels: the wide-step table when a step
     * width serves this read length (auto and "wide"), else the sparse-step table (auto and "sparse"; also when the wide
     * one cannot be built), or the fused-step table ("fused", and auto when neither can be built).  A table that cannot
     * be built -- no memory, over the table budget -- is never fatal: that replica keeps the plain 2-step kernels,
     * still on the GPU.  Only a broken context (FM_E_CUDA from anything but an allocation) aborts. */
    {
      /* auto: an index whose plain table stays L2-resident (config 1/2: 10.7 MB) is not worth a table build for one
       * batch: the plain kernels already run at 3-4.6 G reads/s there */
      const int mode = fm_mode_from_env();
      int want_sparse = (mode == FMGPU_MODE_SPARSE), want_fused = (mode == FMGPU_MODE_FUSED), want_wide = (mode == FMGPU_MODE_WIDE);
      if (mode == FM_MODE_AUTO) want_sparse = want_wide = g_stats.table_bytes > (96ull << 20);
      if (want_wide) want_sparse = 1;                            /* the fallback when no width serves the length or memory is short */
      /* every replica builds its own table from its own copy of the block table: one host thread per GPU */
      int32_t err_g[FM_MAX_GPUS] = { 0 };
      #pragma omp parallel for schedule(static, 1) num_threads(rs->ndev) private(err, t0, meta0) if (rs->ndev > 1)
      for (g = 0; g < rs->ndev; g++) {
        if (!(want_sparse || want_fused)) continue;
        t0 = fm_wall();
        err = FM_E_NOT_IMPLEMENTED;
        if (want_wide) {
          const uint32_t wb = fmgpu_wide_bases_for(rs->replica[g], qrys->size);
          if (wb) err = fmgpu_index_widen(rs->replica[g], wb, 0, 0);
          if (wb && err == FM_E_NOT_IMPLEMENTED) {               /* no room for the 96-bit table: the 64-bit width (a third less scratch) */
            const uint32_t wb2 = fmgpu_wide_bases_for_words(rs->replica[g], qrys->size, 2);
            if (wb2 && wb2 != wb) err = fmgpu_index_widen(rs->replica[g], wb2, 0, 0);
          }
        }
        if (err == FM_E_NOT_IMPLEMENTED && want_sparse) err = fmgpu_index_sparsify(rs->replica[g], 0, 0, 0);
        /* (round 1 switched repeat-rich texts to the fused-step table here; with search trees and per-read state machines
         * the sparse-step table wins on those too -- profiles/r02_skewed_text.md: 2 730 vs 1 689 and 4 034 vs 2 030 M reads/s
         * -- so the fused table is only the fallback when the sparse one cannot be built) */
        if (err == FM_E_NOT_IMPLEMENTED && (want_fused || mode == FM_MODE_AUTO)) err = fmgpu_index_fuse(rs->replica[g], 0, 0, 0);
        g_stats.table_build_s[g] = fm_wall() - t0;
        err_g[g] = (err && err != FM_E_NOT_IMPLEMENTED) ? err : 0;
      }
      for (g = 0; g < rs->ndev; g++)
        if (err_g[g]) { err = err_g[g]; for (g = 0; g < rs->ndev; g++) fmgpu_index_free(&rs->replica[g]); free(rs); return err; }
      if (getenv("FMGPU_VERBOSE") && fmgpu_index_get_meta(rs->replica[0], &meta0) == FM_SUCCESS)
        fprintf(stderr, "fmindex_b200: %d replica(s), SB96 %.1f MB, search table: %s\n", rs->ndev, meta0.nbytes / 1e6,
                meta0.wide_bases ? "wide-step" : meta0.sparse_bases ? "sparse-step" : meta0.fused_bases ? "fused-step" : "none (plain kernels)");
    }
    fmi->d_index = rs;
  }
  /* tables this read length uses (tail table for odd lengths, lead tables of the sparse plan): built here, never by a launch */
  for (g = 0; g < rs->ndev; g++) {
    err = fmgpu_index_prepare(rs->replica[g], qrys->size);
    if (err) return err;
    /* a resident wide-step table whose width does not serve this (new) read length: the sparse-step table joins it */
    if (!fmgpu_index_wide_serves(rs->replica[g], qrys->size)) {
      fmgpu_index_meta_t m;
      const int mode = fm_mode_from_env();
      if (fmgpu_index_get_meta(rs->replica[g], &m) == FM_SUCCESS && m.wide_bases && !m.sparse_bases &&
          (mode == FM_MODE_AUTO || mode == FMGPU_MODE_WIDE)) {
        err = fmgpu_index_sparsify(rs->replica[g], 0, 0, 0);
        if (err && err != FM_E_NOT_IMPLEMENTED) return err;
        err = fmgpu_index_prepare(rs->replica[g], qrys->size);
        if (err) return err;
      }
    }
  }

  /* queries + results: contiguous 32-aligned shards, one per GPU.  Whatever set either handle still holds is released
   * first, with BOTH of its owners detached -- a queries handle may come back with another results handle or vice versa */
  if (qrys->d_queries) fm_release_shards((fm_shard_set_t *) qrys->d_queries);
  if (res->d_results) fm_release_shards((fm_shard_set_t *) res->d_results);
  ss = (fm_shard_set_t *) calloc(1, sizeof(*ss));
  if (!ss) return FM_E_ALLOCATING_MFASTA;
  ss->ndev = rs->ndev; ss->index = rs;
  per = (((uint64_t) qrys->num + rs->ndev - 1) / rs->ndev + 31) & ~31ull;
  for (g = 0; g <= rs->ndev; g++) {
    uint64_t f = per * (uint64_t) g;
    ss->first[g] = f < qrys->num ? f : qrys->num;
  }
  {
    const double t0 = fm_wall();
    for (g = 0; g < rs->ndev; g++) {
      err = fmgpu_batch_create(rs->dev[g], ss->first[g + 1] - ss->first[g], qrys->size, fmi->steps, &ss->shard[g]);
      if (!err) err = fmgpu_batch_upload_ascii(ss->shard[g], qrys->h_queries + ss->first[g] * qrys->size);
      if (err) { fm_release_shards(ss); return err; }
    }
    g_stats.queries_h2d_pack_s = fm_wall() - t0;
    g_stats.query_bytes = (uint64_t) qrys->num * qrys->size;
  }
  ss->owner_q = qrys; ss->owner_r = res;
  qrys->d_queries = (char *) ss;
  res->d_results = (uint32_t *) ss;
  return FM_SUCCESS;
}

/* kernels only, all shards in flight at once.  Error-code form of searchIndexGPU for callers of the fmgpu_* layer. */
int32_t fmgpu_search_index(void *index, void *queries, void *resIntervals)
{
  fmi_t *fmi = (fmi_t *) index;
  qrys_t *qrys = (qrys_t *) queries;
  fm_replica_set_t *rs = fmi ? (fm_replica_set_t *) fmi->d_index : NULL;
  fm_shard_set_t *ss = qrys ? (fm_shard_set_t *) qrys->d_queries : NULL;
  int32_t g, err = FM_SUCCESS;
  (void) resIntervals;
  if (!rs || !ss || ss->index != rs) return FM_E_BAD_ARGUMENT;    /* transferCPUtoGPU has not been called for this pair */
  for (g = 0; g < ss->ndev && !err; g++) {
    fmgpu_variant_t v = g_variant;
    if (!g_variant_set) {                      /* $FMGPU_MODE, else the best table the replica has, else Coop */
      fmgpu_index_meta_t meta;
      const int mode = fm_mode_from_env();
      memset(&v, 0, sizeof v);
      err = fmgpu_index_get_meta(rs->replica[g], &meta);
      if (err) break;
      const int wide_ok = meta.wide_bases && fmgpu_index_wide_serves(rs->replica[g], qrys->size);
      v.mode = (mode == FM_MODE_AUTO) ? (wide_ok ? FMGPU_MODE_WIDE : meta.sparse_bases ? FMGPU_MODE_SPARSE : meta.fused_bases ? FMGPU_MODE_FUSED : FMGPU_MODE_COOP) : mode;
      if (v.mode == FMGPU_MODE_WIDE && !wide_ok) v.mode = FMGPU_MODE_SPARSE;
      if (v.mode == FMGPU_MODE_SPARSE && !meta.sparse_bases) v.mode = meta.fused_bases ? FMGPU_MODE_FUSED : FMGPU_MODE_COOP;
      if (v.mode == FMGPU_MODE_FUSED && !meta.fused_bases) v.mode = FMGPU_MODE_COOP;
      if (v.mode == FMGPU_MODE_SPARSE || v.mode == FMGPU_MODE_WIDE) v.queries_per_thread = 0;     /* the launcher's own default */
    }
    err = fmgpu_batch_search_timed_async(rs->replica[g], ss->shard[g], &v);
  }
  for (g = 0; g < ss->ndev && !err; g++) err = fmgpu_batch_sync(ss->shard[g]);
  for (g = 0; g < ss->ndev && !err; g++) err = fmgpu_batch_last_ms(ss->shard[g], &g_stats.search_ms[g]);
  if (!err) g_stats.searches += 1;
  return err;
}

/* void like the reference, so a failure prints and exits (reference HandleError, src/fmIndexGPU-Coop-2Step.cu:88-93) */
void searchIndexGPU(void *index, void *queries, void *resIntervals)
{
  const int32_t err = fmgpu_search_index(index, queries, resIntervals);
  if (err) {
    fprintf(stderr, "searchIndexGPU: %s (%s:%d)\n",
            err == FM_E_BAD_ARGUMENT ? "transferCPUtoGPU has not been called for this index/queries" : errorCommon(err), __FILE__, __LINE__);
    exit(EXIT_FAILURE);
  }
}

int32_t transferGPUtoCPU(void *results)
{
  res_t *res = (res_t *) results;
  fm_shard_set_t *ss = res ? (fm_shard_set_t *) res->d_results : NULL;
  int32_t g, err;
  if (!ss || !res->h_results) return FM_E_BAD_ARGUMENT;
  {
    const double t0 = fm_wall();
    for (g = 0; g < ss->ndev; g++) {
      if (ss->first[g + 1] == ss->first[g]) continue;
      err = fmgpu_batch_download(ss->shard[g], res->h_results + 2 * ss->first[g]);
      if (err) return err;
    }
    g_stats.results_d2h_s = fm_wall() - t0;
    g_stats.result_bytes = 8ull * ss->first[ss->ndev];
  }
  return FM_SUCCESS;
}

int32_t freeIndexGPU(void **index)
{
  fmi_t *fmi;
  fm_replica_set_t *rs;
  int32_t g;
  if (!index || !*index) return FM_SUCCESS;
  fmi = (fmi_t *) *index;
  rs = (fm_replica_set_t *) fmi->d_index;
  if (rs) {
    fm_dump_stats();
    for (g = 0; g < rs->ndev; g++) fmgpu_index_free(&rs->replica[g]);
    free(rs);
    fmi->d_index = NULL;
  }
  return FM_SUCCESS;
}

/* queries and results share one shard set: it is released when both sides have let go */
int32_t freeQueriesGPU(void **queries)
{
  qrys_t *qrys;
  fm_shard_set_t *ss;
  if (!queries || !*queries) return FM_SUCCESS;
  qrys = (qrys_t *) *queries;
  ss = (fm_shard_set_t *) qrys->d_queries;
  if (ss) {
    qrys->d_queries = NULL;
    ss->owner_q = NULL;
    if (!ss->owner_r) fm_release_shards(ss);
  }
  return FM_SUCCESS;
}

int32_t freeResultsGPU(void **results)
{
  res_t *res;
  fm_shard_set_t *ss;
  if (!results || !*results) return FM_SUCCESS;
  res = (res_t *) *results;
  ss = (fm_shard_set_t *) res->d_results;
  if (ss) {
    res->d_results = NULL;
    ss->owner_r = NULL;
    if (!ss->owner_q) fm_release_shards(ss);
  }
  return FM_SUCCESS;
}

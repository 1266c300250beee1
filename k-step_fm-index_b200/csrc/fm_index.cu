/*
 * fm_index.cu -- PART 2 of include/fmindex_b200.h: devices, index residency / re-blocking (SB96), replicas, tail table.
 * Replaces the cudaMalloc + cudaMemcpy of the reference's transferCPUtoGPU (src/fmIndexGPU-Coop-2Step.cu:250-285).
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 */
#include "fm_internal.h"
#include <mutex>
#include <time.h>
#include "fm_reblock.cuh"

const fmgpu_variant_t FM_DEFAULT_VARIANT = { FMGPU_MODE_TASK, 2, 256, 0 };

static thread_local char g_err[512] = "no error";

extern "C" const char *fmgpu_last_error(void) { return g_err; }

int32_t fm_fail(cudaError_t e, const char *what, const char *file, int line)
{
  snprintf(g_err, sizeof g_err, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  return FM_E_CUDA;
}
int32_t fm_fail_msg(int32_t code, const char *msg)
{
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
/* ------------------------------------------------------------------------ */
extern "C" int32_t fmgpu_device_count(void)
{
  int n = 0, usable = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  for (int i = 0; i < n; i++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) usable++;
  }
  return usable;
}

int32_t fm_use_device(int device)
{
  int major = 0;
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fm_fail_msg(FM_E_CUDA, "device is not sm_100 (this library carries sm_100a code only)");
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_device_warmup(int32_t device)
{
  const int32_t rc = fm_use_device(device);
  if (rc) return rc;
  CU_TRY(cudaFree(0));
  return FM_SUCCESS;
}

extern "C" uint32_t fmgpu_words_per_query(uint32_t len) { return (len + 15u) / 16u; }

/* ------------------------------------------------------------------------ *
 * index residency / layout stage
 * ------------------------------------------------------------------------ */
static uint32_t fm_nblocks_for(uint32_t bwtsize)
{
  uint64_t nb = (uint64_t) bwtsize / FM_SB_ROWS + 1;   /* block of X = bwtsize must exist */
  return (uint32_t)((nb + 7) & ~7ull);                 /* symbol stride = whole 128-byte lines */
}

/* AltCounters padding-entry quirk: in the last chunk, a
 * symbol whose counter lives in the padding entry comes out +1 per '$' row of
 * that chunk carrying the symbol.  Returns the per-symbol 2-bit table. */
static void fm_ac_quirk(uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize, uint32_t ncounters,
                        const uint32_t *dpos, const uint32_t *dbase, uint32_t *start, uint32_t *mask)
{
  *start = 0xFFFFFFFFu; *mask = 0;
  if (tag < 200) return;
  const uint32_t elast = (bwtsize - 1) / chunk;
  for (uint32_t s = 0; s < steps; s++) {
    if (dpos[s] / chunk != elast) continue;
    const uint32_t sigma = dbase[s];
    const bool next = ((elast & 1u) && sigma < ncounters) || (!(elast & 1u) && sigma >= ncounters);
    if (next) *mask += 1u << (2 * sigma);
  }
  if (*mask) *start = elast * chunk;
}

/* fm_build.cu */
cudaError_t fmb_counter_stage(uint32_t *entries, uint32_t k, uint32_t d, uint32_t entry_words, uint32_t nentries, uint32_t bwtsize,
                              const uint32_t *dpos, const uint32_t *dbase);

/* planes of BWT layers 0 and 1 of a k-step file entry (any tag) -> planes of a 2-step tag-100 entry */
__global__ void fm_project_planes_kernel(const FmRawIndex x, uint32_t *__restrict__ out, uint32_t out_entry_words)
{
  const uint32_t W = x.d / 32;
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t) x.nentries_std * W) return;
  const uint32_t e = (uint32_t)(t / W), n = (uint32_t)(t % W);
  uint32_t *o = out + (size_t) e * out_entry_words;
  for (uint32_t s = 0; s < 2; s++)
    for (uint32_t bit = 0; bit < 2; bit++) o[2 * W * s + W * bit + n] = fm_raw_plane(x, e, s, bit, n);
}

static int32_t fm_index_from_device_entries(int device, uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize,
                                            uint32_t ncounters, uint32_t nentries, const uint32_t *dpos,
                                            const uint32_t *dbase, const uint32_t *d_entries, fmgpu_index_t **out);

/* k = 3 or 4 files (CPU-only in the reference, makefile:226-230): the device layout does not depend on the file's k --
 * the first two BWT layers of the file ARE the 2-step index of the same text, so its planes are copied, the 2-step
 * counters are recomputed from them exactly as src/genFMindex.c:210-256 does (fmb_counter_stage, byte-identical to
 * gfmiBaseLine's 2-step file), and the search runs on that; the reference searchers of all k agree on (L,R) wherever
 * the read length is a multiple of k.  An AltCounters file whose padding-entry quirk is active cannot be reproduced
 * this way and is refused. */
static int32_t fm_index_from_wide_file(int device, uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize,
                                       uint32_t ncounters, uint32_t nentries, const uint32_t *dpos,
                                       const uint32_t *dbase, const uint32_t *d_entries, fmgpu_index_t **out)
{
  const bool ac = (tag == 200 || tag == 201);
  const uint32_t nsym = 1u << (2 * steps), W = chunk / 32;
  if (ncounters != (ac ? nsym / 2 : nsym)) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "counter count does not match k");
  const uint32_t nstd = (uint32_t)(((uint64_t) bwtsize + chunk - 1) / chunk);
  if (nentries != nstd + (ac ? 1u : 0u) || bwtsize < 2) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "entry count does not match bwtsize/d");
  uint32_t qmask = 0;
  {                                                          /* padding-entry quirk of a wide AltCounters file (any symbol) */
    const uint32_t elast = (bwtsize - 1) / chunk;
    for (uint32_t s = 0; ac && s < steps; s++)
      if (dpos[s] / chunk == elast) {
        const uint32_t sigma = dbase[s];
        if (((elast & 1u) && sigma < ncounters) || (!(elast & 1u) && sigma >= ncounters)) qmask = 1;
      }
  }
  if (qmask) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "k >= 3 AltCounters file with an active padding-entry quirk");
  const uint32_t ew2 = 4 * W + 16;
  uint32_t *proj = NULL;
  cudaError_t e = cudaMalloc((void **) &proj, (size_t) nstd * ew2 * 4);
  if (e != cudaSuccess) return fm_fail(e, "cudaMalloc(2-step projection)", __FILE__, __LINE__);
  FmRawIndex raw;
  raw.entries = d_entries; raw.tag = tag; raw.k = steps; raw.d = chunk; raw.ncounters = ncounters;
  raw.nentries = nentries; raw.entry_words = 2 * W * steps + ncounters; raw.bwtsize = bwtsize; raw.nentries_std = nstd;
  for (uint32_t s = 0; s < 2; s++) { raw.dpos[s] = dpos[s]; raw.dbase[s] = dbase[s]; }
  raw.quirk_start = 0xFFFFFFFFu; raw.quirk_mask = 0;
  const uint64_t nthreads = (uint64_t) nstd * W;
  fm_project_planes_kernel<<<(unsigned)((nthreads + 255) / 256), 256>>>(raw, proj, ew2);
  e = cudaGetLastError();
  const uint32_t dpos2[2] = { dpos[0], dpos[1] }, dbase2[2] = { dbase[0] & 15u, dbase[1] & 15u };
  if (e == cudaSuccess) e = fmb_counter_stage(proj, 2, chunk, ew2, nstd, bwtsize, dpos2, dbase2);
  if (e != cudaSuccess) { cudaFree(proj); return fm_fail(e, "2-step projection of a k >= 3 file", __FILE__, __LINE__); }
  int32_t rc = fm_index_from_device_entries(device, 100, 2, chunk, bwtsize, 16, nstd, dpos2, dbase2, proj, out);
  cudaFree(proj);
  if (rc == FM_SUCCESS) { (*out)->meta.source_tag = tag; (*out)->meta.source_steps = steps; }
  return rc;
}

static int32_t fm_index_from_device_entries(int device, uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize,
                                            uint32_t ncounters, uint32_t nentries, const uint32_t *dpos,
                                            const uint32_t *dbase, const uint32_t *d_entries, fmgpu_index_t **out)
{
  const bool ac = (tag == 200 || tag == 201);
  if (!(tag == 100 || tag == 101 || ac)) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "unknown index tag");
  if (steps < 1 || steps > 4) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "index files with k in {1,2,3,4} are supported (like the reference builders)");
  if (chunk == 0 || chunk % 32) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "d must be a multiple of 32");
  for (uint32_t s = 0; s < steps; s++)                          /* header fields of an untrusted file */
    if (dpos[s] >= bwtsize || dbase[s] >= (1u << (2 * steps)))
      return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "dollarPositionBWT / dollarBaseBWT out of range for this bwtsize and k");
  if (steps > 2) return fm_index_from_wide_file(device, tag, steps, chunk, bwtsize, ncounters, nentries, dpos, dbase, d_entries, out);
  const uint32_t nsym = 1u << (2 * steps);
  if (ncounters != (ac ? nsym / 2 : nsym)) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "counter count does not match k");
  const uint32_t need = (uint32_t)(((uint64_t) bwtsize + chunk - 1) / chunk) + (ac ? 1u : 0u);
  if (nentries != need || bwtsize < 2) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "entry count does not match bwtsize/d");

  fmgpu_index_t *idx = (fmgpu_index_t *) calloc(1, sizeof(*idx));
  if (!idx) return fm_fail_msg(FM_E_ALLOCATING_FMI, "host allocation failed");
  idx->device = device;
  idx->meta.steps = steps; idx->meta.bwtsize = bwtsize; idx->meta.nsymbols = nsym;
  idx->meta.nblocks = fm_nblocks_for(bwtsize); idx->meta.source_tag = tag; idx->meta.source_steps = steps;
  fm_ac_quirk(tag, steps, chunk, bwtsize, ncounters, dpos, dbase, &idx->meta.quirk_start, &idx->meta.quirk_mask);
  idx->meta.nbytes = (uint64_t) nsym * idx->meta.nblocks * sizeof(uint4);

  cudaError_t e = cudaMalloc((void **) &idx->blocks, idx->meta.nbytes);
  if (e != cudaSuccess) { free(idx); return fm_fail(e, "cudaMalloc(SB96 table)", __FILE__, __LINE__); }

  FmRawIndex raw;
  raw.entries = d_entries; raw.tag = tag; raw.k = steps; raw.d = chunk; raw.ncounters = ncounters;
  raw.nentries = nentries; raw.entry_words = 2 * (chunk / 32) * steps + ncounters; raw.bwtsize = bwtsize;
  raw.nentries_std = ac ? nentries - 1 : nentries;
  for (uint32_t s = 0; s < 2; s++) { raw.dpos[s] = s < steps ? dpos[s] : 0xFFFFFFFFu; raw.dbase[s] = s < steps ? dbase[s] : 0xFFFFFFFFu; }
  raw.quirk_start = idx->meta.quirk_start; raw.quirk_mask = idx->meta.quirk_mask;

  const uint32_t nb = idx->meta.nblocks;
  fm_reblock_kernel<<<(nb + 127) / 128, 128>>>(raw, idx->blocks, nb);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(idx->blocks); free(idx); return fm_fail(e, "fm_reblock_kernel", __FILE__, __LINE__); }

  /* constants of the derived 1-step rank that serves the last base of odd-length reads on a 2-step index */
  if (steps == 2) {
    uint4 first[16], last[16];
    const uint32_t bl = bwtsize / FM_SB_ROWS, rl = bwtsize - bl * FM_SB_ROWS;
    for (uint32_t s = 0; s < 16 && e == cudaSuccess; s++) {
      e = cudaMemcpy(&first[s], idx->blocks + (size_t) s * nb, sizeof(uint4), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(&last[s], idx->blocks + (size_t) s * nb + bl, sizeof(uint4), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) { cudaFree(idx->blocks); free(idx); return fm_fail(e, "tail constants", __FILE__, __LINE__); }
    uint32_t at0[16], total1[4] = { 0, 0, 0, 0 };
    for (uint32_t s = 0; s < 16; s++) {
      const uint32_t w[3] = { last[s].y, last[s].z, last[s].w };
      uint32_t end = last[s].x;
      for (uint32_t j = 0; j < 3; j++) {
        const int32_t nbits = (int32_t) rl - 32 * (int32_t) j;
        const uint32_t m = nbits <= 0 ? 0u : (nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u));
        end += (uint32_t) __builtin_popcount(w[j] & m);
      }
      at0[s] = first[s].x;                                  /* rank2(s, 0)        */
      total1[s & 3u] += end - at0[s];                       /* rows with 2-step symbol s, by layer-0 char */
    }
    const uint32_t t0 = dbase[1] & 3u;
    total1[t0] += 1;                                        /* the row whose layer-1 char is '$' still has a layer-0 char */
    uint32_t c1 = 1;                                        /* the '$' suffix precedes everything */
    for (uint32_t c = 0; c < 4; c++) {
      idx->meta.tail_const[c] = c1 - (at0[c] + at0[c | 4u] + at0[c | 8u] + at0[c | 12u]);
      c1 += total1[c];
    }
    idx->meta.tail_row = dpos[1]; idx->meta.tail_base = t0;
    idx->tail_consts_ok = 1;                                 /* the stored ranks are quirk-free: always the text's own 1-step index (locate) */
    idx->meta.tail_valid = idx->meta.quirk_mask == 0;         /* odd read lengths are served on quirk-free files only */
  }
  *out = idx;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_create(int32_t device, uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize,
                                      uint32_t ncounters, uint32_t nentries, const uint32_t *dpos, const uint32_t *dbase,
                                      const uint32_t *h_entries, fmgpu_index_t **out)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!h_entries || !out || !dpos || !dbase) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (steps < 1 || steps > 4 || chunk == 0 || chunk % 32) return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "index files need k in {1,2,3,4}, d multiple of 32");
  const uint64_t bytes = (uint64_t) nentries * (2 * (chunk / 32) * steps + ncounters) * 4ull;
  uint32_t *d_raw = NULL;
  CU_TRY(cudaMalloc((void **) &d_raw, bytes));
  cudaError_t e = cudaMemcpy(d_raw, h_entries, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d_raw); return fm_fail(e, "cudaMemcpy(index H2D)", __FILE__, __LINE__); }
  rc = fm_index_from_device_entries(device, tag, steps, chunk, bwtsize, ncounters, nentries, dpos, dbase, d_raw, out);
  cudaFree(d_raw);
  return rc;
}

extern "C" int32_t fmgpu_index_create_from_device(int32_t device, uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize,
                                                  uint32_t ncounters, uint32_t nentries, const uint32_t *dpos,
                                                  const uint32_t *dbase, const uint32_t *d_entries, fmgpu_index_t **out)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!d_entries || !out || !dpos || !dbase) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  return fm_index_from_device_entries(device, tag, steps, chunk, bwtsize, ncounters, nentries, dpos, dbase, d_entries, out);
}

extern "C" int32_t fmgpu_index_alloc_like(int32_t device, const fmgpu_index_meta_t *meta, fmgpu_index_t **out)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!meta || !out) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (meta->nbytes != (uint64_t) meta->nsymbols * meta->nblocks * sizeof(uint4)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "inconsistent index meta");
  fmgpu_index_t *idx = (fmgpu_index_t *) calloc(1, sizeof(*idx));
  if (!idx) return fm_fail_msg(FM_E_ALLOCATING_FMI, "host allocation failed");
  idx->device = device; idx->meta = *meta;
  /* derived tables are per replica: a fresh replica has none until fmgpu_index_fuse / fmgpu_index_sparsify run on it */
  idx->meta.fused_bases = 0; idx->meta.fused_lanes = 0; idx->meta.fused_bytes = 0; idx->meta.start_bases = 0;
  idx->meta.sparse_bases = 0; idx->meta.sparse_lambda = 0; idx->meta.sparse_bytes = 0; idx->meta.sparse_blocks = 0;
  idx->meta.sparse_overflow = 0; idx->meta.sparse_start_bases = 0; idx->meta.sparse_lanes = 0; idx->meta.tail_bytes = 0;
  idx->meta.sparse_uniform_nb = 0; idx->meta.sa_bytes = 0; idx->meta.sa_rate = 0; idx->meta.derived_bytes = 0;
  idx->meta.sparse_tree_nodes = 0; idx->meta.sparse_tree_rows = 0; idx->meta.sparse_tree_depth = 0;
  idx->meta.wide_bases = 0; idx->meta.wide_prefix_bits = 0; idx->meta.wide_row_bits = 0; idx->meta.wide_tree_depth = 0;
  idx->meta.wide_bytes = 0; idx->meta.wide_blocks = 0; idx->meta.wide_overflow = 0; idx->meta.wide_tree_nodes = 0;
  idx->meta.wide_tree_rows = 0; idx->meta.wide_exceptional = 0;
  cudaError_t e = cudaMalloc((void **) &idx->blocks, meta->nbytes);
  if (e != cudaSuccess) { free(idx); return fm_fail(e, "cudaMalloc(SB96 replica)", __FILE__, __LINE__); }
  *out = idx;
  return FM_SUCCESS;
}

static thread_local double g_last_peer_copy_s = 0.0;
extern "C" double fmgpu_last_peer_copy_seconds(void) { return g_last_peer_copy_s; }   /* the cudaMemcpyPeer of this thread's last fmgpu_index_replicate */

extern "C" int32_t fmgpu_index_replicate(const fmgpu_index_t *src, int32_t device, fmgpu_index_t **out)
{
  if (!src || !out) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  fmgpu_index_t *dst = NULL;
  int32_t rc = fmgpu_index_alloc_like(device, &src->meta, &dst);
  if (rc) return rc;
  int can = 0;
  cudaDeviceCanAccessPeer(&can, device, src->device);
  if (can) { cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0); if (pe != cudaSuccess) cudaGetLastError(); }
  cudaDeviceSynchronize();                                      /* allocation and peer mapping are done: what is timed below is the copy */
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  cudaError_t e = cudaMemcpyPeer(dst->blocks, device, src->blocks, src->device, src->meta.nbytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();            /* the replica is complete when this returns */
  clock_gettime(CLOCK_MONOTONIC, &t1);
  g_last_peer_copy_s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (e != cudaSuccess) { cudaFree(dst->blocks); free(dst); return fm_fail(e, "cudaMemcpyPeer(index replica)", __FILE__, __LINE__); }
  *out = dst;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_get_meta(const fmgpu_index_t *idx, fmgpu_index_meta_t *meta)
{
  if (!idx || !meta) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  *meta = idx->meta;
  return FM_SUCCESS;
}
extern "C" void *fmgpu_index_blocks(const fmgpu_index_t *idx) { return idx ? (void *) idx->blocks : NULL; }
extern "C" int32_t fmgpu_index_device(const fmgpu_index_t *idx) { return idx ? idx->device : -1; }

extern "C" int32_t fmgpu_index_free(fmgpu_index_t **pidx)
{
  if (!pidx || !*pidx) return FM_SUCCESS;
  fmgpu_index_t *idx = *pidx;
  if (idx->blocks || idx->fblocks || idx->start || idx->sblocks || idx->tail1 || idx->sa || idx->wblocks) {
    cudaSetDevice(idx->device);
    cudaFree(idx->blocks); cudaFree(idx->fblocks); cudaFree(idx->start);
    cudaFree(idx->sblocks); cudaFree(idx->sdir); cudaFree(idx->sstart); cudaFree(idx->tail1); cudaFree(idx->sa); cudaFree(idx->sa_marks);
    for (int b = 0; b < 16; b++) cudaFree(idx->slead[b]);
    cudaFree(idx->wblocks);
    for (int b = 0; b < 16; b++) cudaFree(idx->wlead[b]);
  }
  free(idx);
  *pidx = NULL;
  return FM_SUCCESS;
}

/* Tail table of a 2-step replica (fm_tail_table_kernel, a quarter of the SB96 table): built by fmgpu_index_prepare for
 * an odd read length (and by the suffix-array derivation), on that replica's device, from its own block table (so
 * replicas filled by a broadcast get theirs when they first need it).  Returns NULL -- the kernels then derive the rank
 * from four SB96 fetches -- when the index has no valid tail ($FMGPU_TAIL_TABLE=0 also forces that path), the table
 * budget is spent or the allocation fails.  Synchronous; never called from a launch path. */
cudaError_t fm_tail_table_into(const fmgpu_index_t *idx, uint4 *dst)
{
  const uint32_t nb = idx->meta.nblocks;
  const uint32_t *tc = idx->meta.tail_const;
  fm_tail_table_kernel<<<(nb + 255) / 256, 256>>>(idx->blocks, nb, tc[0], tc[1], tc[2], tc[3], idx->meta.tail_row, idx->meta.tail_base, dst);
  return cudaGetLastError();
}

static std::mutex g_tail_mutex;
const uint4 *fm_build_tail(fmgpu_index_t *idx)
{
  if (!idx->meta.tail_valid) return NULL;
  std::lock_guard<std::mutex> lock(g_tail_mutex);
  if (idx->tail1_tried) return idx->tail1;
  idx->tail1_tried = 1;
  const char *env = getenv("FMGPU_TAIL_TABLE");
  if (env && *env && atoi(env) == 0) return NULL;
  const uint32_t nb = idx->meta.nblocks;
  const uint64_t bytes = (uint64_t) 4 * nb * sizeof(uint4);
  if (!fm_budget_allows(idx, bytes)) return NULL;
  uint4 *t = NULL;
  if (cudaSetDevice(idx->device) != cudaSuccess || cudaMalloc((void **) &t, bytes) != cudaSuccess) { cudaGetLastError(); return NULL; }
  if (fm_tail_table_into(idx, t) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); cudaFree(t); return NULL; }
  idx->tail1 = t; idx->meta.tail_bytes = bytes;
  fm_budget_account(idx);
  return t;
}

/* ------------------------------------------------------------------------ *
 * derived-table memory budget: one knob for everything a replica derives from its SB96 table (sparse-step table +
 * directory + start / lead tables, fused-step table, tail table, suffix array).  $FMGPU_TABLE_BUDGET_GB or
 * fmgpu_set_table_budget(); 0 = no limit but the device's free memory.  meta.derived_bytes reports the sum.
 * ------------------------------------------------------------------------ */
static uint64_t g_table_budget = 0;
static bool g_table_budget_set = false;

extern "C" int32_t fmgpu_set_table_budget(uint64_t bytes) { g_table_budget = bytes; g_table_budget_set = true; return FM_SUCCESS; }

static uint64_t fm_table_budget(void)
{
  if (g_table_budget_set) return g_table_budget;
  const char *env = getenv("FMGPU_TABLE_BUDGET_GB");
  return env && *env ? (uint64_t)(atof(env) * 1e9) : 0;
}

static uint64_t fm_derived_bytes(const fmgpu_index_t *idx)
{
  return idx->meta.sparse_bytes + idx->meta.wide_bytes + idx->meta.fused_bytes + (idx->start ? ((uint64_t) 8 << 24) : 0) + idx->meta.tail_bytes + idx->meta.sa_bytes;
}

bool fm_budget_allows(const fmgpu_index_t *idx, uint64_t bytes)
{
  const uint64_t budget = fm_table_budget();
  return budget == 0 || fm_derived_bytes(idx) + bytes <= budget;
}

void fm_budget_account(fmgpu_index_t *idx)
{
  idx->meta.derived_bytes = fm_derived_bytes(idx);
  idx->meta.budget_bytes = fm_table_budget();
}

/* Everything a search of `len`-base reads on this replica may use is built NOW (synchronously): the tail table for odd
 * lengths on a 2-step index, the lead tables of the sparse-step plan.  The launch paths (fmgpu_batch_search,
 * fmgpu_search_device, searchIndexGPU) never build anything: they pick among the tables that exist and fall back to
 * SB96 steps otherwise -- same results, more fetches.  transferCPUtoGPU and fmgpu_search_host call this themselves. */
extern "C" int32_t fmgpu_index_prepare(fmgpu_index_t *idx, uint32_t len)
{
  if (!idx || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  CU_TRY(cudaSetDevice(idx->device));
  if (idx->meta.steps == 2 && (len & 1u)) fm_build_tail(idx);
  if (idx->sblocks) fm_sparse_prepare(idx, len);
  if (idx->wblocks) fm_wide_prepare(idx, len);
  fm_budget_account(idx);
  return FM_SUCCESS;
}

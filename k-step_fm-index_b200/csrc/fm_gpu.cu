/*
 * fm_gpu.cu -- PART 2 of include/fmindex_b200.h: the thin C ABI over the
 * hand-written CUDA of fm_kernels.cuh / fm_fused.cuh (index residency /
 * re-blocking, fused-step table, query packing, search launches, replicas,
 * pipelined end-to-end search with the hybrid host feed, roofline probes).
 *
 * Replaces the device-side support code every reference .cu carries
 * (transferCPUtoGPU / searchIndexGPU / transferGPUtoCPU / free*GPU, e.g.
 * src/fmIndexGPU-Coop-2Step.cu:231-338): one device, default stream,
 * pageable cudaMemcpy there; explicit devices, streams, pinned staging and
 * peer copies here.  No CPU fallback: without a usable sm_100 device every
 * entry point returns FM_E_CUDA.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <mutex>
#include <cuda_runtime.h>
#include "../../include/fmindex_b200.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "fm_kernels.cuh"
#include "fm_fused.cuh"
#include "fm_sparse.cuh"
#include "fm_locate.cuh"

struct fmgpu_index {
  int                device;
  fmgpu_index_meta_t meta;
  uint4             *blocks;
  uint4             *fblocks;      /* fused-step table (fmgpu_index_fuse), or NULL */
  uint32_t           nfblocks;     /* fused blocks per fused symbol */
  uint2             *start;        /* (L,R) of all 4^12 12-mers (start table of the fused kernel), or NULL */
  uint4             *sblocks;      /* sparse-step table (fmgpu_index_sparsify), or NULL */
  uint2             *sdir;         /* its directory: { first block, scale } per wide symbol */
  uint2             *sstart;       /* start table of the sparse kernel, or NULL */
  int                stables;      /* start / lead tables are in use for this replica's sparse table (large indexes, or $FMGPU_START_TABLE=1) */
  uint32_t           slead_tried;  /* bit b: building slead[b] was attempted */
  uint2             *slead[16];    /* lead tables: (L,R) of all b-mers, b = 6 .. sparse_bases - 1 (multiples of k), or NULL */
  uint4             *tail1;        /* tail table (fm_tail_table_kernel): built by the first odd-length search on this replica */
  uint32_t          *sa;           /* suffix array derived from the table (fmgpu_index_build_sa), or NULL */
  uint32_t           s_uni_nb, s_uni_scale;   /* sparse table is a uniform grid: blocks per symbol and the one scale (0 = directory) */
  int                tail1_tried;  /* 1 once that build was attempted (a failed allocation is not retried)               */
};

struct fmgpu_batch {
  int          device;
  uint64_t     nq;
  uint32_t     len, steps, wpq;
  char        *d_ascii;       /* staging for upload_ascii (allocated on first use) */
  uint32_t    *d_packed;
  uint32_t    *d_results;
  unsigned long long *d_counters;
  cudaStream_t stream;
  cudaEvent_t  ev0, ev1;
};

static thread_local char g_err[512] = "no error";

extern "C" const char *fmgpu_last_error(void) { return g_err; }

static int32_t fm_fail(cudaError_t e, const char *what, const char *file, int line)
{
  snprintf(g_err, sizeof g_err, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  return FM_E_CUDA;
}
static int32_t fm_fail_msg(int32_t code, const char *msg)
{
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
#define CU_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fm_fail(e_, #call, __FILE__, __LINE__); } while (0)

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

static int32_t fm_use_device(int device)
{
  int major = 0;
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fm_fail_msg(FM_E_CUDA, "device is not sm_100 (this library carries sm_100a code only)");
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
  uint32_t qstart, qmask;
  {                                                          /* padding-entry quirk of a wide AltCounters file (any symbol) */
    qstart = 0xFFFFFFFFu; qmask = 0;
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
  if (steps == 2 && idx->meta.quirk_mask == 0) {
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
    idx->meta.tail_row = dpos[1]; idx->meta.tail_base = t0; idx->meta.tail_valid = 1;
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
  idx->meta.sparse_uniform_nb = 0; idx->meta.sa_bytes = 0;
  cudaError_t e = cudaMalloc((void **) &idx->blocks, meta->nbytes);
  if (e != cudaSuccess) { free(idx); return fm_fail(e, "cudaMalloc(SB96 replica)", __FILE__, __LINE__); }
  *out = idx;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_replicate(const fmgpu_index_t *src, int32_t device, fmgpu_index_t **out)
{
  if (!src || !out) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  fmgpu_index_t *dst = NULL;
  int32_t rc = fmgpu_index_alloc_like(device, &src->meta, &dst);
  if (rc) return rc;
  int can = 0;
  cudaDeviceCanAccessPeer(&can, device, src->device);
  if (can) { cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0); if (pe != cudaSuccess) cudaGetLastError(); }
  cudaError_t e = cudaMemcpyPeer(dst->blocks, device, src->blocks, src->device, src->meta.nbytes);
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
  if (idx->blocks || idx->fblocks || idx->start || idx->sblocks || idx->tail1 || idx->sa) {
    cudaSetDevice(idx->device);
    cudaFree(idx->blocks); cudaFree(idx->fblocks); cudaFree(idx->start);
    cudaFree(idx->sblocks); cudaFree(idx->sdir); cudaFree(idx->sstart); cudaFree(idx->tail1); cudaFree(idx->sa);
    for (int b = 0; b < 16; b++) cudaFree(idx->slead[b]);
  }
  free(idx);
  *pidx = NULL;
  return FM_SUCCESS;
}

/* Tail table of a 2-step replica (fm_tail_table_kernel, a quarter of the SB96 table): built by the first search with an
 * odd read length, on that replica's device, from its own block table (so replicas filled by a broadcast get theirs
 * when they first need it).  Returns NULL -- the kernels then derive the rank from four SB96 fetches -- when the
 * index has no valid tail ($FMGPU_TAIL_TABLE=0 also forces that path) or the allocation fails. */
static std::mutex g_tail_mutex;
static const uint4 *fm_ensure_tail(const fmgpu_index_t *cidx, cudaStream_t stream)
{
  fmgpu_index_t *idx = const_cast<fmgpu_index_t *>(cidx);
  if (!idx->meta.tail_valid) return NULL;
  std::lock_guard<std::mutex> lock(g_tail_mutex);
  if (idx->tail1_tried) return idx->tail1;
  idx->tail1_tried = 1;
  const char *env = getenv("FMGPU_TAIL_TABLE");
  if (env && *env && atoi(env) == 0) return NULL;
  const uint32_t nb = idx->meta.nblocks;
  uint4 *t = NULL;
  if (cudaSetDevice(idx->device) != cudaSuccess || cudaMalloc((void **) &t, (size_t) 4 * nb * sizeof(uint4)) != cudaSuccess) { cudaGetLastError(); return NULL; }
  const uint32_t *tc = idx->meta.tail_const;
  fm_tail_table_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(idx->blocks, nb, tc[0], tc[1], tc[2], tc[3], idx->meta.tail_row, idx->meta.tail_base, t);
  /* waited for once: later searches may come on other streams */
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { cudaGetLastError(); cudaFree(t); return NULL; }
  idx->tail1 = t; idx->meta.tail_bytes = (uint64_t) 4 * nb * sizeof(uint4);
  return t;
}

/* ------------------------------------------------------------------------ *
 * fused-step table (fm_fused.cuh)
 * ------------------------------------------------------------------------ */
static uint32_t fm_fused_rows(uint32_t lanes) { return 32u * (8u * lanes - 1u); }
static const fmgpu_variant_t FM_DEFAULT_VARIANT = { FMGPU_MODE_TASK, 2, 256, 0 };
static int32_t fm_launch_fused(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                               uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters = NULL);
__global__ void fm_iota_kernel(uint32_t *out, uint32_t n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

template <int LANES>
static cudaError_t fm_fuse_build(const fmgpu_index_t *idx, const uint16_t *fsym, uint32_t nfsym, uint32_t nfb, uint32_t kbits,
                                 uint32_t hops, uint4 *fblocks)
{
  fm_fuse_write_kernel<LANES><<<nfb, 256, 256 * 8 * LANES * 4>>>(fsym, nfsym, nfb, fblocks);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fm_fuse_scan_kernel<LANES><<<nfsym, 1024>>>(idx->blocks, idx->meta.nblocks, kbits, hops, nfb, fblocks);
  return cudaGetLastError();
}

extern "C" int32_t fmgpu_index_fuse(fmgpu_index_t *idx, uint32_t fused_bases, uint32_t lanes, uint64_t budget_bytes)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->fblocks) return FM_SUCCESS;
  if (idx->meta.quirk_mask) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "fused steps are unavailable for an AltCounters index carrying the padding-entry quirk");
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t k = idx->meta.steps;
  if (lanes == 0) lanes = 2;
  if (!(lanes == 1 || lanes == 2 || lanes == 4)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "fused block lanes must be 1, 2 or 4");
  if (budget_bytes == 0) {
    const char *env = getenv("FMGPU_FUSE_BUDGET_GB");
    budget_bytes = (uint64_t)((env && *env ? atof(env) : 69.0) * 1e9);       /* flat part of the footprint curve ends near 68-70 GB */
  }
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  const uint32_t rows = fm_fused_rows(lanes);
  const uint32_t nfb = idx->meta.bwtsize / rows + 1;
  uint32_t kf = 0;
  for (uint32_t cand = 4; cand > k; cand--) {
    if (fused_bases && cand != fused_bases) continue;
    if (cand % k) continue;
    const uint64_t bytes = ((uint64_t) 1 << (2 * cand)) * nfb * 32ull * lanes;
    const uint64_t scratch = 3ull * idx->meta.bwtsize + (1ull << 30);       /* construction: 3 bytes per row, plus slack */
    if (bytes + scratch > free_b) continue;                                  /* does not fit in HBM right now */
    if (bytes <= budget_bytes || fused_bases) { kf = cand; break; }
  }
  if (!kf) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "no fused-step table fits the memory budget for this index (or k already is the requested width)");
  const uint32_t nfsym = 1u << (2 * kf), hops = kf / k, kbits = 2 * k;
  const uint64_t fbytes = (uint64_t) nfsym * nfb * 32ull * lanes;
  uint64_t nrows = (uint64_t) idx->meta.nblocks * FM_SB_ROWS;
  if ((uint64_t) nfb * rows > nrows) nrows = (uint64_t) nfb * rows;
  uint8_t *sym = NULL; uint16_t *fsym = NULL; uint4 *fblocks = NULL;
  cudaError_t e = cudaMalloc((void **) &fblocks, fbytes);
  if (e == cudaSuccess) e = cudaMalloc((void **) &sym, nrows);
  if (e == cudaSuccess) e = cudaMalloc((void **) &fsym, nrows * 2);
  if (e == cudaSuccess) {
    fm_fuse_symbols_kernel<<<(idx->meta.nblocks + 127) / 128, 128>>>(idx->blocks, idx->meta.nblocks, idx->meta.nsymbols, nrows, sym);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    fm_fuse_compose_kernel<<<(unsigned)((nrows + 255) / 256), 256>>>(idx->blocks, idx->meta.nblocks, sym, idx->meta.bwtsize, kbits, hops, nrows, fsym);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = lanes == 1 ? fm_fuse_build<1>(idx, fsym, nfsym, nfb, kbits, hops, fblocks)
                          : lanes == 2 ? fm_fuse_build<2>(idx, fsym, nfsym, nfb, kbits, hops, fblocks)
                                       : fm_fuse_build<4>(idx, fsym, nfsym, nfb, kbits, hops, fblocks);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(sym); cudaFree(fsym);
  if (e != cudaSuccess) { cudaFree(fblocks); return fm_fail(e, "fmgpu_index_fuse", __FILE__, __LINE__); }
  idx->fblocks = fblocks; idx->nfblocks = nfb;
  idx->meta.fused_bases = kf; idx->meta.fused_lanes = lanes; idx->meta.fused_bytes = fbytes;

  /* start table: the fused kernel itself searches all 4^12 12-mers once (a packed 12-mer IS its 24-bit key);
   * only worth it when the table (134 MB) is small next to the index, and FM_START_BASES must be whole fused steps */
  {
    const char *env = getenv("FMGPU_START_TABLE");
    const bool want = env && *env ? atoi(env) != 0 : idx->meta.nbytes >= (1ull << 30);
    if (want && FM_START_BASES % kf == 0 && idx->meta.bwtsize > (1u << 24)) {
      const uint32_t nkeys = 1u << (2 * FM_START_BASES);
      uint32_t *keys = NULL; uint2 *table = NULL;
      e = cudaMalloc((void **) &keys, (size_t) nkeys * 4);
      if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nkeys * 8);
      if (e == cudaSuccess) { fm_iota_kernel<<<(nkeys + 255) / 256, 256>>>(keys, nkeys); e = cudaGetLastError(); }
      int32_t rc = FM_SUCCESS;
      if (e == cudaSuccess) rc = fm_launch_fused(idx, keys, nkeys, FM_START_BASES, (uint32_t *) table, FM_DEFAULT_VARIANT, 0);
      if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
      cudaFree(keys);
      if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); }   /* the table is optional */
      else { idx->start = table; idx->meta.start_bases = FM_START_BASES; }
    }
  }
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_unfuse(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->fblocks || idx->start) { CU_TRY(cudaSetDevice(idx->device)); cudaFree(idx->fblocks); cudaFree(idx->start); idx->fblocks = NULL; idx->start = NULL; }
  idx->nfblocks = 0; idx->meta.fused_bases = 0; idx->meta.fused_lanes = 0; idx->meta.fused_bytes = 0; idx->meta.start_bases = 0;
  return FM_SUCCESS;
}

typedef void (*fm_fused_fn)(const FmFusedParams);

template <int KF, int K, int LANES>
static fm_fused_fn fm_pick_fused_q(int qpt)
{
  if (qpt == 0) return fm_search_fused_kernel<KF, K, LANES, 1, 256, 6, true>;      /* instrumented */
  if (qpt == 1) return fm_search_fused_kernel<KF, K, LANES, 1, 256, 6, false>;
  if (qpt == 2) return fm_search_fused_kernel<KF, K, LANES, 2, 256, 4, false>;
  return NULL;
}
template <int KF, int K>
static fm_fused_fn fm_pick_fused_l(int lanes, int qpt)
{
  if (lanes == 1) return fm_pick_fused_q<KF, K, 1>(qpt);
  if (lanes == 2) return fm_pick_fused_q<KF, K, 2>(qpt);
  if (lanes == 4) return fm_pick_fused_q<KF, K, 4>(qpt);
  return NULL;
}
static fm_fused_fn fm_pick_fused(uint32_t kf, uint32_t k, int lanes, int qpt)
{
  if (kf == 4 && k == 2) return fm_pick_fused_l<4, 2>(lanes, qpt);
  if (kf == 4 && k == 1) return fm_pick_fused_l<4, 1>(lanes, qpt);
  if (kf == 3 && k == 1) return fm_pick_fused_l<3, 1>(lanes, qpt);
  if (kf == 2 && k == 1) return fm_pick_fused_l<2, 1>(lanes, qpt);
  return NULL;
}

static int32_t fm_launch_fused(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                               uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters)
{
  if (!idx->fblocks) return fm_fail_msg(FM_E_BAD_ARGUMENT, "FMGPU_MODE_FUSED needs fmgpu_index_fuse() on this replica first");
  const uint32_t k = idx->meta.steps, kf = idx->meta.fused_bases, lanes = idx->meta.fused_lanes, hops = kf / k;
  if (v.queries_per_thread != 1 && v.queries_per_thread != 2) v.queries_per_thread = 2;
  FmFusedParams p;
  p.fblocks = idx->fblocks; p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results;
  p.nfblocks = idx->nfblocks; p.nblocks = idx->meta.nblocks; p.nq = (uint32_t) nq;
  p.nlead = (len / k) % hops; p.nfused = (len / k) / hops;
  p.wpq = fmgpu_words_per_query(len); p.wpq_pad = (p.wpq + 1) | 1u; p.bwtsize = idx->meta.bwtsize;
  p.fetch_counters = d_counters;
  p.start = idx->start; p.start_steps = idx->start ? FM_START_BASES / kf : 0u;
  p.has_tail = len % k; p.tail_row = idx->meta.tail_row; p.tail_base = idx->meta.tail_base;
  for (int c = 0; c < 4; c++) p.tail_const[c] = idx->meta.tail_const[c];
  p.tail1 = p.has_tail ? fm_ensure_tail(idx, stream) : NULL;
  if (d_counters) v.queries_per_thread = 1;
  uint32_t qper; size_t smem;
  for (;;) {
    qper = (256 / lanes) * v.queries_per_thread;
    smem = 16 + ((size_t) qper * p.wpq + 4) * 4;             /* mbarrier + reads + one readable spare word */
    if (smem <= 200 * 1024) break;
    if (v.queries_per_thread > 1) v.queries_per_thread = 1;
    else return fm_fail_msg(FM_E_QUERY_SHAPE, "reads too long to stage in shared memory");
  }
  fm_fused_fn fn = fm_pick_fused(kf, k, (int) lanes, d_counters ? 0 : v.queries_per_thread);
  if (!fn) return fm_fail_msg(FM_E_BAD_ARGUMENT, "no fused kernel for this (fused bases, k, lanes)");
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const uint32_t grid = (uint32_t)((nq + qper - 1) / qper);
  void *args[] = { (void *) &p };
  CU_TRY(cudaLaunchKernel((const void *) fn, dim3(grid), dim3(256), args, smem, stream));
  return FM_SUCCESS;
}


/* ------------------------------------------------------------------------ *
 * sparse-step table (fm_sparse.cuh)
 * ------------------------------------------------------------------------ */
static int32_t fm_launch_sparse(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters = NULL);
static const uint2 *fm_ensure_lead(const fmgpu_index_t *cidx, uint32_t b);

extern "C" int32_t fmgpu_index_unsparsify(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sblocks || idx->sdir || idx->sstart) {
    CU_TRY(cudaSetDevice(idx->device));
    cudaFree(idx->sblocks); cudaFree(idx->sdir); cudaFree(idx->sstart);
    idx->sblocks = NULL; idx->sdir = NULL; idx->sstart = NULL;
    for (int b = 0; b < 16; b++) { cudaFree(idx->slead[b]); idx->slead[b] = NULL; }
    idx->slead_tried = 0;
  }
  idx->stables = 0;
  idx->s_uni_nb = 0; idx->s_uni_scale = 0; idx->meta.sparse_uniform_nb = 0;
  idx->meta.sparse_bases = 0; idx->meta.sparse_lambda = 0; idx->meta.sparse_bytes = 0; idx->meta.sparse_blocks = 0;
  idx->meta.sparse_overflow = 0; idx->meta.sparse_start_bases = 0; idx->meta.sparse_lanes = 0;
  return FM_SUCCESS;
}

static thread_local bool g_require_uniform = false;          /* set while the automatic choice tries 12 bases per step */
extern "C" int32_t fmgpu_index_sparsify(fmgpu_index_t *idx, uint32_t sparse_bases, uint32_t lambda, uint32_t lanes)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sblocks) return FM_SUCCESS;
  if (idx->meta.quirk_mask) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "sparse steps are unavailable for an AltCounters index carrying the padding-entry quirk");
  if (idx->meta.bwtsize >= FM_SP_OVF) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "text too long for the sparse-step table");
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t k = idx->meta.steps, n = idx->meta.bwtsize;
  if (lanes == 0) lanes = 2;
  if (lanes != 2 && lanes != 4) return fm_fail_msg(FM_E_BAD_ARGUMENT, "sparse block lanes must be 2 (64-byte blocks) or 4 (128-byte blocks)");
  const uint32_t slots = 8 * lanes - 1, bbytes = 32 * lanes;
  if (lambda == 0) lambda = lanes == 4 ? 12 : 5;
  if (lambda > slots) return fm_fail_msg(FM_E_BAD_ARGUMENT, "lambda must not exceed the slots of a block (15 or 31)");
  uint32_t ks = sparse_bases;
  if (ks == 0) {
    /* 12 bases per step when that table can be a uniform grid (a directory of 4^12 entries would not stay in L2): tried first,
     * given up as soon as the symbol counts turn out uneven.  Else the widest multiple of k up to 10 with >= 64 rows per symbol. */
    const char *env = getenv("FMGPU_SPARSE_UNIFORM");
    if (!g_require_uniform && !(env && *env && atoi(env) == 0)) {
      /* 14 bases need >= lambda rows per 14-mer on average, 12 bases >= 64 rows per 12-mer */
      const uint32_t wide[2] = { 14, 12 };
      const uint64_t least[2] = { (uint64_t) lambda << 28, (uint64_t) 64 << 24 };
      for (int c = 0; c < 2; c++) {
        if (wide[c] % k || least[c] > n) continue;
        g_require_uniform = true;
        const int32_t rcw = fmgpu_index_sparsify(idx, wide[c], lambda, lanes);
        g_require_uniform = false;
        if (rcw == FM_SUCCESS) return FM_SUCCESS;
      }
    }
    for (uint32_t cand = 10; cand >= 2 * k; cand--)
      if (cand % k == 0 && (((uint64_t) 64) << (2 * cand)) <= n) { ks = cand; break; }
    if (ks == 0) ks = 2 * k;
  }
  if (ks % k || ks <= k || ks > 14) return fm_fail_msg(FM_E_BAD_ARGUMENT, "sparse bases must be a multiple of k, larger than k and at most 14");
  const uint32_t nsym = 1u << (2 * ks), hops = ks / k, kbits = 2 * k;
  const uint64_t nrows = (uint64_t) idx->meta.nblocks * FM_SB_ROWS;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  const uint64_t est_blocks = (uint64_t) n / lambda + nsym;
  const uint64_t need = 16ull * n + nrows + est_blocks * bbytes + 32ull * nsym + (1ull << 30);
  if (need > free_b) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough free device memory to build the sparse-step table");

  uint8_t *sym = NULL; uint32_t *keys = NULL, *rows = NULL, *keys2 = NULL, *rows2 = NULL, *symstart = NULL, *nb = NULL, *first = NULL, *rank0 = NULL;
  uint2 *dir = NULL; uint4 *sblocks = NULL; void *tmp = NULL; unsigned long long *d_novf = NULL;
  size_t tmp_bytes = 0, tmp2 = 0;
  uint64_t total_blocks = 0; unsigned long long novf = 0;
  cudaError_t e = cudaMalloc((void **) &sym, nrows);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rows, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys2, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rows2, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &symstart, 4ull * (nsym + 1));
  if (e == cudaSuccess) e = cudaMalloc((void **) &nb, 4ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &first, 4ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &rank0, 4ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &dir, 8ull * nsym);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_novf, 8);
  if (e == cudaSuccess) e = cudaMemset(d_novf, 0, 8);
  if (e == cudaSuccess) {
    fm_fuse_symbols_kernel<<<(idx->meta.nblocks + 127) / 128, 128>>>(idx->blocks, idx->meta.nblocks, idx->meta.nsymbols, nrows, sym);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    fm_sparse_compose_kernel<<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(idx->blocks, idx->meta.nblocks, sym, n, kbits, hops, nsym, keys, rows);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, keys, keys2, rows, rows2, (int64_t) n, 0, (int)(2 * ks + 1));
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(NULL, tmp2, nb, first, (int) nsym);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, rows, rows2, (int64_t) n, 0, (int)(2 * ks + 1));
  if (e == cudaSuccess) {
    fm_sparse_symstart_kernel<<<(nsym + 1 + 255) / 256, 256>>>(keys2, n, nsym, symstart);
    e = cudaGetLastError();
  }
  /* uniform grid or per-symbol block counts?  $FMGPU_SPARSE_UNIFORM = 0 / 1 forces; default: uniform when no symbol occurs
   * more than 1.6 x as often as the mean (its blocks then expect <= 8 rows at lambda 5: < 1 % of THOSE symbols' blocks
   * overflow 15 slots, far cheaper than a directory lookup in every step) nor less than 0.4 x as often (wasted blocks) */
  uint32_t uni_nb = 0, uni_scale = 0;
  if (e == cudaSuccess) {
    uint32_t range[2] = { 0xFFFFFFFFu, 0u }, carrying = 0;
    uint32_t *d_range = first;                                 /* scratch: `first` is written by the scan below */
    e = cudaMemcpy(d_range, range, 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { fm_sparse_count_range_kernel<<<(nsym + 255) / 256, 256>>>(symstart, nsym, d_range); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(range, d_range, 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&carrying, symstart + nsym, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
      const double mean = (double) carrying / nsym;
      const char *env = getenv("FMGPU_SPARSE_UNIFORM");
      const uint64_t per = ((uint64_t) carrying + (uint64_t) nsym * lambda - 1) / ((uint64_t) nsym * lambda);
      /* rows living in symbols that cannot fit their share of the grid even if spread perfectly (count > slots x blocks) */
      unsigned long long heavy = 0, *d_heavy = d_novf;            /* (d_novf is zero here: the fill kernel runs later) */
      fm_sparse_heavy_rows_kernel<<<(nsym + 255) / 256, 256>>>(symstart, nsym, (uint32_t)(slots * (per ? per : 1)), d_heavy);
      e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaMemcpy(&heavy, d_heavy, 8, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemset(d_heavy, 0, 8);
      const bool even = mean >= lambda && range[1] <= 1.6 * mean && range[0] >= 0.4 * mean;          /* many rows per symbol: tight counts */
      const bool sparse_even = mean >= lambda && mean < 64 && heavy * 1000ull <= carrying;           /* few rows per symbol (Poisson scatter): no heavy tail */
      const bool want = env && *env ? atoi(env) != 0 : (even || sparse_even);
      if (want && per >= 1 && per * nsym < (1ull << 32)) {
        uni_nb = (uint32_t) per;
        unsigned long long sc = ((((unsigned long long) uni_nb) << 32) - 1ull) / n;    /* as fm_sparse_dir_kernel */
        uni_scale = sc > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) sc;
      }
    }
  }
  if (e == cudaSuccess && g_require_uniform && !uni_nb) {      /* the 12-base attempt of the automatic choice: counts are uneven */
    cudaFree(sym); cudaFree(keys); cudaFree(rows); cudaFree(keys2); cudaFree(rows2); cudaFree(symstart); cudaFree(nb); cudaFree(first);
    cudaFree(rank0); cudaFree(tmp); cudaFree(d_novf); cudaFree(dir);
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "symbol counts too uneven for a uniform grid");
  }
  if (e == cudaSuccess) {
    fm_sparse_nblocks_kernel<<<(nsym + 255) / 256, 256>>>(symstart, nsym, lambda, uni_nb, nb);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, nb, first, (int) nsym);
  if (e == cudaSuccess) {
    uint32_t last_first = 0, last_nb = 0;
    e = cudaMemcpy(&last_first, first + (nsym - 1), 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&last_nb, nb + (nsym - 1), 4, cudaMemcpyDeviceToHost);
    total_blocks = (uint64_t) last_first + last_nb;
  }
  /* the sort's input buffers are dead now: release them before the table is allocated */
  cudaFree(keys); keys = NULL; cudaFree(rows); rows = NULL; cudaFree(sym); sym = NULL;
  if (e == cudaSuccess && total_blocks >= (1ull << 32)) e = cudaErrorInvalidValue;
  if (e == cudaSuccess) e = cudaMalloc((void **) &sblocks, total_blocks * bbytes);
  if (e == cudaSuccess) {
    fm_sparse_dir_kernel<<<(nsym + 255) / 256, 256>>>(idx->blocks, idx->meta.nblocks, kbits, hops, nsym, n, nb, first, dir, rank0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    if (uni_nb && uni_nb < 32) {                               /* many symbols, few blocks each: one thread per block */
      const unsigned grid = (unsigned)((total_blocks + 255) / 256);
      if (lanes == 4) fm_sparse_fill_uniform_kernel<4><<<grid, 256>>>(rows2, symstart, nsym, uni_nb, uni_scale, rank0, sblocks, d_novf);
      else            fm_sparse_fill_uniform_kernel<2><<<grid, 256>>>(rows2, symstart, nsym, uni_nb, uni_scale, rank0, sblocks, d_novf);
    } else if (lanes == 4) fm_sparse_fill_kernel<4><<<nsym, 128>>>(rows2, symstart, dir, nb, rank0, sblocks, d_novf);
    else                   fm_sparse_fill_kernel<2><<<nsym, 128>>>(rows2, symstart, dir, nb, rank0, sblocks, d_novf);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(&novf, d_novf, 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(sym); cudaFree(keys); cudaFree(rows); cudaFree(keys2); cudaFree(rows2); cudaFree(symstart); cudaFree(nb); cudaFree(first);
  cudaFree(rank0); cudaFree(tmp); cudaFree(d_novf);
  if (e != cudaSuccess) { cudaFree(sblocks); cudaFree(dir); return fm_fail(e, "fmgpu_index_sparsify", __FILE__, __LINE__); }
  idx->sblocks = sblocks; idx->sdir = dir; idx->s_uni_nb = uni_nb; idx->s_uni_scale = uni_scale; idx->meta.sparse_uniform_nb = uni_nb;
  idx->meta.sparse_bases = ks; idx->meta.sparse_lambda = lambda; idx->meta.sparse_blocks = total_blocks;
  idx->meta.sparse_overflow = novf; idx->meta.sparse_bytes = total_blocks * bbytes + 8ull * nsym; idx->meta.sparse_lanes = lanes;

  /* start table: the sparse kernel itself searches every SB-mer once (a packed SB-mer IS its key); SB = the
   * largest whole number of sparse steps within 12 bases */
  {
    const char *env = getenv("FMGPU_START_TABLE");
    const bool want = env && *env ? atoi(env) != 0 : idx->meta.nbytes >= (1ull << 30);
    const uint32_t ssteps = 12 / ks, sb = ssteps * ks;
    if (want && ssteps && ((uint64_t) 1 << (2 * sb)) < n) {
      const uint32_t nkeys = 1u << (2 * sb);
      uint32_t *skeys = NULL; uint2 *table = NULL;
      e = cudaMalloc((void **) &skeys, (size_t) nkeys * 4);
      if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nkeys * 8);
      if (e == cudaSuccess) { fm_iota_kernel<<<(nkeys + 255) / 256, 256>>>(skeys, nkeys); e = cudaGetLastError(); }
      int32_t rc = FM_SUCCESS;
      if (e == cudaSuccess) rc = fm_launch_sparse(idx, skeys, nkeys, sb, (uint32_t *) table, FM_DEFAULT_VARIANT, 0);
      if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
      cudaFree(skeys);
      if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); }   /* the table is optional */
      else { idx->sstart = table; idx->meta.sparse_start_bases = sb; idx->meta.sparse_bytes += (uint64_t) nkeys * 8; }
    }
    /* lead tables (fm_ensure_lead): the small ones now, the wide ones (12 .. 15 bases, up to 8.6 GB) when a read length asks */
    idx->stables = want ? 1 : 0;
    if (want)
      for (uint32_t b = 1; b <= ks + 1 && b < 12; b++) fm_ensure_lead(idx, b);
  }
  return FM_SUCCESS;
}

/* Lead table of width b: (L,R) of every b-mer, computed by the search kernel itself (a packed b-mer is its own key).
 * A read whose length leaves b bases over -- or b - KS, giving up one sparse step -- starts from it and then runs only
 * whole sparse steps: no SB96 fetches behind them and no tail fetch.  Any parity: on a 2-step index an odd width ends
 * with the derived 1-step rank while the TABLE is computed (bases may be grouped into steps in any way; every grouping
 * composes the same LF steps).  Widths 6 .. 11 are built with the sparse table, 12 .. 15 (134 MB .. 8.6 GB) by the first
 * search that needs them.  NULL when tables are off for this index, the width is not representable or memory is short. */
static std::mutex g_lead_mutex;
static thread_local bool g_building_lead = false;
static const uint2 *fm_ensure_lead(const fmgpu_index_t *cidx, uint32_t b)
{
  fmgpu_index_t *idx = const_cast<fmgpu_index_t *>(cidx);
  if (b < 1 || b >= 16 || !idx->stables) return NULL;
  std::lock_guard<std::mutex> lock(g_lead_mutex);
  if (idx->slead[b]) return idx->slead[b];
  if (idx->slead_tried & (1u << b)) return NULL;
  idx->slead_tried |= 1u << b;
  const uint32_t k = idx->meta.steps, n = idx->meta.bwtsize;
  if ((idx->sstart && b == idx->meta.sparse_start_bases) || (b % k && !(k == 2 && idx->meta.tail_valid)) || ((uint64_t) 1 << (2 * b)) >= n) return NULL;
  if (cudaSetDevice(idx->device) != cudaSuccess) { cudaGetLastError(); return NULL; }
  const uint32_t nkeys = 1u << (2 * b);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return NULL; }
  if ((uint64_t) nkeys * 12 + (2ull << 30) > free_b) return NULL;
  uint32_t *skeys = NULL; uint2 *table = NULL;
  cudaError_t e = cudaMalloc((void **) &skeys, (size_t) nkeys * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nkeys * 8);
  if (e == cudaSuccess) { fm_iota_kernel<<<(nkeys + 255) / 256, 256>>>(skeys, nkeys); e = cudaGetLastError(); }
  int32_t rc = FM_SUCCESS;
  g_building_lead = true;
  if (e == cudaSuccess) rc = fm_launch_sparse(idx, skeys, nkeys, b, (uint32_t *) table, FM_DEFAULT_VARIANT, 0);
  g_building_lead = false;
  if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
  cudaFree(skeys);
  if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); return NULL; }
  idx->slead[b] = table; idx->meta.sparse_bytes += (uint64_t) nkeys * 8;
  return table;
}

typedef void (*fm_sparse_fn)(const FmSparseParams);
template <int K, int LANES>
static fm_sparse_fn fm_pick_sparse(int qpt)
{
  if (qpt == 0) return fm_search_sparse_kernel<K, LANES, 1, 256, 6, true>;        /* instrumented */
  if (qpt == 1) return fm_search_sparse_kernel<K, LANES, 1, 256, 6, false>;
  if (qpt == 2) return fm_search_sparse_kernel<K, LANES, 2, 256, 4, false>;
  if (qpt == 3) return fm_search_sparse_kernel<K, LANES, 3, 256, 4, false>;
  if (qpt == 4) return fm_search_sparse_kernel<K, LANES, 4, 256, 3, false>;
  return NULL;
}

static int32_t fm_launch_sparse(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters)
{
  if (!idx->sblocks) return fm_fail_msg(FM_E_BAD_ARGUMENT, "FMGPU_MODE_SPARSE needs fmgpu_index_sparsify() on this replica first");
  const uint32_t k = idx->meta.steps, ks = idx->meta.sparse_bases, hops = ks / k, lanes = idx->meta.sparse_lanes;
  if (v.queries_per_thread < 1 || v.queries_per_thread > 4) v.queries_per_thread = 4;
  FmSparseParams p;
  p.sblocks = idx->sblocks; p.dir = idx->sdir; p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results;
  p.nblocks = idx->meta.nblocks; p.nq = (uint32_t) nq;
  /* plan: S whole sparse steps + lb leftover bases (rem base-k steps + an odd tail base).  (B) the leftover bases -- or
   * leftover + one sparse step's bases -- are taken first from their lead table and only whole sparse steps follow:
   * S (or S - 1) block fetches plus one lookup, no tail fetch; (A) no such table: the start table replaces the first
   * sparse step(s) and the rem steps follow the sparse ones, S - m + rem block fetches from DRAM (+ tail); (C) no
   * tables at all (small indexes): the rem steps run first on the upper (L2-resident) levels of SB96. */
  const uint32_t S = (len / k) / hops, rem = (len / k) % hops;
  const uint32_t m = idx->sstart ? idx->meta.sparse_start_bases / ks : 0u;
  const uint32_t lb = len - S * ks;                            /* leftover bases, the odd one included */
  uint32_t lead = 0;
  if (idx->stables && !g_building_lead && (lb >= 1 || !m)) {   /* (a lead table is computed without lead tables) */
    /* the leftover bases themselves when the interval they leave is much narrower than a bucket (4^lb >= 8 x blocks per
     * symbol: the first sparse step then rarely needs two fetches), else leftover + one sparse step's bases (12 .. 15;
     * with no leftover and no start table -- 14 bases per step -- the table of all 14-mers IS the start table) */
    const uint64_t nb_mean = idx->meta.sparse_blocks >> (2 * ks);
    const bool narrow = lb >= 1 && lb < 16 && ((uint64_t) 1 << (2 * lb)) >= 8 * (nb_mean ? nb_mean : 1);
    if (narrow && S >= 1 && fm_ensure_lead(idx, lb)) lead = lb;
    else if (!narrow && S >= 2 && lb + ks < 16 && lb + ks > idx->meta.sparse_start_bases && fm_ensure_lead(idx, lb + ks)) lead = lb + ks;
    else if (!narrow && lb >= 1 && S >= 1 && !m && fm_ensure_lead(idx, lb)) lead = lb;      /* better two fetches in the first step than SB96 steps */
    else if (S == 0 && lb >= 1 && lb < 16 && fm_ensure_lead(idx, lb)) lead = lb;            /* a read shorter than one sparse step: one lookup */
  }
  p.nfront = 0; p.nback = 0; p.nsteps = S; p.start = NULL; p.start_bits = 0;
  if (lead) { p.start = idx->slead[lead]; p.start_bits = 2 * lead; p.nsteps = S - (lead > lb ? 1u : 0u); }
  else if (m && S >= m) { p.start = idx->sstart; p.start_bits = 2 * ks * m; p.nsteps = S - m; p.nback = rem; }
  else p.nfront = rem;
  p.wpq = fmgpu_words_per_query(len); p.bwtsize = idx->meta.bwtsize;
  p.sbits = 2 * ks; p.hops = hops;
  p.uni_nb = idx->s_uni_nb; p.uni_scale = idx->s_uni_scale;
  p.fetch_counters = d_counters;
  p.has_tail = lead ? 0u : len % k;                            /* a lead table already holds the odd base */
  p.tail_row = idx->meta.tail_row; p.tail_base = idx->meta.tail_base;
  for (int c = 0; c < 4; c++) p.tail_const[c] = idx->meta.tail_const[c];
  p.tail1 = p.has_tail ? fm_ensure_tail(idx, stream) : NULL;
  if (d_counters) v.queries_per_thread = 1;
  uint32_t qper; size_t smem;
  for (;;) {
    qper = (256 / lanes) * v.queries_per_thread;
    smem = 16 + ((size_t) qper * p.wpq + 4) * 4;
    if (smem <= 200 * 1024) break;
    if (v.queries_per_thread > 1) v.queries_per_thread -= 1;
    else return fm_fail_msg(FM_E_QUERY_SHAPE, "reads too long to stage in shared memory");
  }
  const int qsel = d_counters ? 0 : v.queries_per_thread;
  fm_sparse_fn fn = k == 2 ? (lanes == 4 ? fm_pick_sparse<2, 4>(qsel) : fm_pick_sparse<2, 2>(qsel))
                           : (lanes == 4 ? fm_pick_sparse<1, 4>(qsel) : fm_pick_sparse<1, 2>(qsel));
  if (!fn) return fm_fail_msg(FM_E_BAD_ARGUMENT, "no sparse kernel for this variant");
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const uint32_t grid = (uint32_t)((nq + qper - 1) / qper);
  void *args[] = { (void *) &p };
  CU_TRY(cudaLaunchKernel((const void *) fn, dim3(grid), dim3(256), args, smem, stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_count_fetches_sparse_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                                     uint32_t *d_results, void *stream, uint64_t *nsparse_blocks, uint64_t *nsb96_blocks,
                                                     uint64_t *noverflows)
{
  if (!idx || !d_packed || !d_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (len == 0 || (len % idx->meta.steps && !idx->meta.tail_valid)) return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a positive multiple of k");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long *d_c = NULL, h[3] = { 0, 0, 0 };
  CU_TRY(cudaMalloc((void **) &d_c, 24));
  CU_TRY(cudaMemsetAsync(d_c, 0, 24, (cudaStream_t) stream));
  int32_t rc = nq ? fm_launch_sparse(idx, d_packed, nq, len, d_results, FM_DEFAULT_VARIANT, (cudaStream_t) stream, d_c) : FM_SUCCESS;
  if (rc == FM_SUCCESS) {
    cudaError_t e = cudaMemcpyAsync(h, d_c, 24, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) rc = fm_fail(e, "fetch counters D2H", __FILE__, __LINE__);
  }
  cudaFree(d_c);
  if (nsparse_blocks) *nsparse_blocks = h[0];
  if (nsb96_blocks) *nsb96_blocks = h[1];
  if (noverflows) *noverflows = h[2];
  return rc;
}

/* ------------------------------------------------------------------------ *
 * locate (fm_locate.cuh): suffix array derived from the replica's own table, SA[L..R) gathers
 * ------------------------------------------------------------------------ */
extern "C" int32_t fmgpu_index_build_sa(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sa) return FM_SUCCESS;
  if (idx->meta.quirk_mask) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "locate is unavailable for an AltCounters index carrying the padding-entry quirk");
  if (idx->meta.steps == 2 && !idx->meta.tail_valid) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "this 2-step index has no derived 1-step rank");
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t n = idx->meta.bwtsize, nb = idx->meta.nblocks;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  if (20ull * n + 64ull * nb + (256ull << 20) > free_b) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough free device memory to derive the suffix array");
  /* the 4-symbol table that holds rank1 at block starts + per-row char bits: SB96 itself (k = 1) or the tail table (k = 2) */
  const uint4 *t1 = idx->blocks;
  uint4 *t1_tmp = NULL;
  if (idx->meta.steps == 2) {
    t1 = fm_ensure_tail(idx, 0);
    if (!t1) {                                                  /* $FMGPU_TAIL_TABLE=0: a private copy for this build */
      CU_TRY(cudaMalloc((void **) &t1_tmp, (size_t) 4 * nb * sizeof(uint4)));
      const uint32_t *tc = idx->meta.tail_const;
      fm_tail_table_kernel<<<(nb + 255) / 256, 256>>>(idx->blocks, nb, tc[0], tc[1], tc[2], tc[3], idx->meta.tail_row, idx->meta.tail_base, t1_tmp);
      t1 = t1_tmp;
    }
  }
  uint2 *na = NULL, *nbuf = NULL; uint32_t *sa = NULL, *d_term = NULL; unsigned long long *d_status = NULL;
  uint32_t term[2] = { 0xFFFFFFFFu, 0u }; unsigned long long bad = 0;
  cudaError_t e = cudaMalloc((void **) &na, 8ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &nbuf, 8ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &sa, 4ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_term, 8);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_status, 8);
  if (e == cudaSuccess) e = cudaMemcpy(d_term, term, 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(d_status, 0, 8);
  if (e == cudaSuccess) e = cudaMemset(na, 0xFF, 8ull * n);     /* a row the table does not cover would point nowhere valid */
  if (e == cudaSuccess) { fm_locate_lf_kernel<<<(nb + 127) / 128, 128>>>(t1, nb, n, na, d_term); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(term, d_term, 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && term[1] != 1u) {
    cudaFree(na); cudaFree(nbuf); cudaFree(sa); cudaFree(d_term); cudaFree(d_status); cudaFree(t1_tmp);
    return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "index table is inconsistent: expected exactly one row without a BWT character");
  }
  if (e == cudaSuccess) {
    uint32_t rounds = 0;
    while (rounds < 32 && (1ull << rounds) < n) rounds++;
    for (uint32_t r = 0; r < rounds && e == cudaSuccess; r++) {
      fm_locate_jump_kernel<<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(na, nbuf, n);
      e = cudaGetLastError();
      uint2 *t = na; na = nbuf; nbuf = t;
    }
  }
  if (e == cudaSuccess) { fm_locate_extract_kernel<<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(na, n, term[0], sa, d_status); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpy(&bad, d_status, 8, cudaMemcpyDeviceToHost);
  cudaFree(na); cudaFree(nbuf); cudaFree(d_term); cudaFree(d_status); cudaFree(t1_tmp);
  if (e != cudaSuccess) { cudaFree(sa); return fm_fail(e, "fmgpu_index_build_sa", __FILE__, __LINE__); }
  if (bad) { cudaFree(sa); return fm_fail_msg(FM_E_UNSUPPORTED_INDEX, "index table is inconsistent: its LF mapping is not a single cycle"); }
  idx->sa = sa; idx->meta.sa_bytes = 4ull * n;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_index_drop_sa(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->sa) { CU_TRY(cudaSetDevice(idx->device)); cudaFree(idx->sa); idx->sa = NULL; }
  idx->meta.sa_bytes = 0;
  return FM_SUCCESS;
}

extern "C" void *fmgpu_index_sa(const fmgpu_index_t *idx) { return idx ? (void *) idx->sa : NULL; }

extern "C" int32_t fmgpu_locate_device(const fmgpu_index_t *idx, const uint32_t *d_results, uint64_t nq, uint32_t max_hits,
                                       uint32_t *d_positions, uint32_t *d_nhits, void *stream)
{
  if (!idx || !d_results || !d_positions || max_hits == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  if (!idx->sa) return fm_fail_msg(FM_E_BAD_ARGUMENT, "locate needs fmgpu_index_build_sa() on this replica first");
  if (nq == 0) return FM_SUCCESS;
  const uint64_t total = nq * max_hits;
  if (total >= (1ull << 39)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "too many (read, hit) slots in one launch; shard the batch");
  CU_TRY(cudaSetDevice(idx->device));
  fm_locate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t) stream>>>(idx->sa, (const uint2 *) d_results, nq, max_hits, d_positions, d_nhits);
  CU_TRY(cudaGetLastError());
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ *
 * kernel dispatch
 * ------------------------------------------------------------------------ */
typedef void (*fm_kernel_fn)(const FmSearchParams);

template <int K, int QPT, int THREADS, int MINB>
static fm_kernel_fn fm_pick_task(bool quirk, bool count)
{
  if (count) return quirk ? fm_search_task_kernel<K, QPT, THREADS, MINB, true, true> : fm_search_task_kernel<K, QPT, THREADS, MINB, false, true>;
  return quirk ? fm_search_task_kernel<K, QPT, THREADS, MINB, true, false> : fm_search_task_kernel<K, QPT, THREADS, MINB, false, false>;
}
template <int K, int QPT, int THREADS, int MINB>
static fm_kernel_fn fm_pick_coop(bool quirk)
{
  return quirk ? fm_search_coop_kernel<K, QPT, THREADS, MINB, true> : fm_search_coop_kernel<K, QPT, THREADS, MINB, false>;
}

/* register budgets: QPT=1 -> 32 regs (2048 thr/SM), QPT=2 -> 40, QPT=4 -> 64 */
template <int K>
static fm_kernel_fn fm_pick(int mode, int qpt, int tpb, bool quirk, bool count)
{
  if (mode == FMGPU_MODE_TASK) {
    if (qpt == 1 && tpb == 128) return fm_pick_task<K, 1, 128, 16>(quirk, count);
    if (qpt == 1 && tpb == 256) return fm_pick_task<K, 1, 256, 8>(quirk, count);
    if (qpt == 1 && tpb == 512) return fm_pick_task<K, 1, 512, 4>(quirk, count);
    if (qpt == 2 && tpb == 128) return fm_pick_task<K, 2, 128, 12>(quirk, count);
    if (qpt == 2 && tpb == 256) return fm_pick_task<K, 2, 256, 6>(quirk, count);
    if (qpt == 2 && tpb == 512) return fm_pick_task<K, 2, 512, 3>(quirk, count);
    if (qpt == 4 && tpb == 128) return fm_pick_task<K, 4, 128, 8>(quirk, count);
    if (qpt == 4 && tpb == 256) return fm_pick_task<K, 4, 256, 4>(quirk, count);
    if (qpt == 4 && tpb == 512) return fm_pick_task<K, 4, 512, 2>(quirk, count);
  } else if (mode == FMGPU_MODE_COOP && !count) {
    if (qpt == 1 && tpb == 128) return fm_pick_coop<K, 1, 128, 16>(quirk);
    if (qpt == 1 && tpb == 256) return fm_pick_coop<K, 1, 256, 8>(quirk);
    if (qpt == 1 && tpb == 512) return fm_pick_coop<K, 1, 512, 4>(quirk);
    if (qpt == 2 && tpb == 128) return fm_pick_coop<K, 2, 128, 16>(quirk);
    if (qpt == 2 && tpb == 256) return fm_pick_coop<K, 2, 256, 8>(quirk);
    if (qpt == 2 && tpb == 512) return fm_pick_coop<K, 2, 512, 4>(quirk);
    if (qpt == 4 && tpb == 128) return fm_pick_coop<K, 4, 128, 12>(quirk);
    if (qpt == 4 && tpb == 256) return fm_pick_coop<K, 4, 256, 6>(quirk);
    if (qpt == 4 && tpb == 512) return fm_pick_coop<K, 4, 512, 3>(quirk);
  }
  return NULL;
}

static int32_t fm_launch_search(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                uint32_t *d_results, const fmgpu_variant_t *vin, cudaStream_t stream,
                                unsigned long long *d_counters)
{
  if (!idx || !d_packed || !d_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  const uint32_t k = idx->meta.steps;
  if (len == 0 || (len % k && !idx->meta.tail_valid))
    return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a positive multiple of k (undefined in the reference, SURVEY.md App. C-5; "
                                         "odd lengths are served on 2-step indexes without the AltCounters quirk only)");
  if (nq == 0) return FM_SUCCESS;
  if (nq >= (1ull << 31)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "more than 2^31 reads in one launch; shard the batch");
  fmgpu_variant_t v = vin ? *vin : FM_DEFAULT_VARIANT;
  if (v.queries_per_thread == 0) v.queries_per_thread = FM_DEFAULT_VARIANT.queries_per_thread;
  if (v.threads_per_block == 0) v.threads_per_block = FM_DEFAULT_VARIANT.threads_per_block;
  const bool count = d_counters != NULL;
  if (count) { v.mode = FMGPU_MODE_TASK; v.queries_per_thread = 1; v.threads_per_block = 256; }
  if (v.mode == FMGPU_MODE_FUSED) return fm_launch_fused(idx, d_packed, nq, len, d_results, vin ? *vin : FM_DEFAULT_VARIANT, stream);
  if (v.mode == FMGPU_MODE_SPARSE) return fm_launch_sparse(idx, d_packed, nq, len, d_results, vin ? *vin : FM_DEFAULT_VARIANT, stream);

  FmSearchParams p;
  p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results; p.fetch_counters = d_counters;
  p.nblocks = idx->meta.nblocks; p.nq = (uint32_t) nq; p.nsteps = len / k;
  p.wpq = fmgpu_words_per_query(len); p.wpq_pad = p.wpq | 1u;
  p.bwtsize = idx->meta.bwtsize; p.quirk_start = idx->meta.quirk_start; p.quirk_mask = idx->meta.quirk_mask;
  p.has_tail = len % k; p.tail_row = idx->meta.tail_row; p.tail_base = idx->meta.tail_base;
  for (int c = 0; c < 4; c++) p.tail_const[c] = idx->meta.tail_const[c];
  p.tail1 = p.has_tail ? fm_ensure_tail(idx, stream) : NULL;
  const bool quirk = idx->meta.quirk_mask != 0;

  /* shrink the CTA's read count until the staged reads fit in shared memory */
  uint32_t qper; size_t smem;
  for (;;) {
    qper = (v.mode == FMGPU_MODE_COOP ? v.threads_per_block / 2 : v.threads_per_block) * v.queries_per_thread;
    smem = (size_t) qper * p.wpq_pad * 4;
    if (smem <= 200 * 1024) break;
    if (v.queries_per_thread > 1) v.queries_per_thread /= 2;
    else if (v.threads_per_block > 128) v.threads_per_block /= 2;
    else return fm_fail_msg(FM_E_QUERY_SHAPE, "reads too long to stage in shared memory");
  }
  fm_kernel_fn fn = (k == 1) ? fm_pick<1>(v.mode, v.queries_per_thread, v.threads_per_block, quirk, count)
                             : fm_pick<2>(v.mode, v.queries_per_thread, v.threads_per_block, quirk, count);
  if (!fn) return fm_fail_msg(FM_E_BAD_ARGUMENT, "unsupported kernel variant (mode 0/1, queries_per_thread 1/2/4, threads_per_block 128/256/512)");
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const uint32_t grid = (uint32_t)((nq + qper - 1) / qper);
  void *args[] = { (void *) &p };
  CU_TRY(cudaLaunchKernel((const void *) fn, dim3(grid), dim3(v.threads_per_block), args, smem, stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_search_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                       uint32_t *d_results, const fmgpu_variant_t *v, void *stream)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  CU_TRY(cudaSetDevice(idx->device));
  return fm_launch_search(idx, d_packed, nq, len, d_results, v, (cudaStream_t) stream, NULL);
}

extern "C" int32_t fmgpu_pack_queries_device(int32_t device, const char *d_ascii, uint64_t nq, uint32_t len,
                                             uint32_t *d_packed, void *stream)
{
  if (!d_ascii || !d_packed || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(device));
  if (nq == 0) return FM_SUCCESS;
  const uint32_t wpq = fmgpu_words_per_query(len);
  const uint64_t total = nq * wpq;
  if ((total + 255) / 256 >= (1ull << 31)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "batch too large for one pack launch");
  fm_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t) stream>>>(d_ascii, nq, len, wpq, d_packed);
  CU_TRY(cudaGetLastError());
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_unstream_device(int32_t device, const uint32_t *d_stream, uint64_t nq, uint32_t len,
                                         uint32_t *d_packed, void *stream)
{
  if (!d_stream || !d_packed || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(device));
  if (nq == 0) return FM_SUCCESS;
  const uint32_t wpq = fmgpu_words_per_query(len);
  if ((nq * wpq + 255) / 256 >= (1ull << 31)) return fm_fail_msg(FM_E_BAD_ARGUMENT, "batch too large for one launch");
  fm_unstream_kernel<<<(unsigned)((nq * wpq + 255) / 256), 256, 0, (cudaStream_t) stream>>>(d_stream, nq, len, wpq, d_packed);
  CU_TRY(cudaGetLastError());
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ *
 * query shards
 * ------------------------------------------------------------------------ */
extern "C" int32_t fmgpu_batch_create(int32_t device, uint64_t nq, uint32_t len, uint32_t steps, fmgpu_batch_t **out)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!out || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  fmgpu_batch_t *b = (fmgpu_batch_t *) calloc(1, sizeof(*b));
  if (!b) return fm_fail_msg(FM_E_ALLOCATING_MFASTA, "host allocation failed");
  b->device = device; b->nq = nq; b->len = len; b->steps = steps; b->wpq = fmgpu_words_per_query(len);
  const size_t pw = (size_t)(nq ? nq : 1) * b->wpq * 4, rw = (size_t)(nq ? nq : 1) * 8;
  cudaError_t e = cudaMalloc((void **) &b->d_packed, pw);
  if (e == cudaSuccess) e = cudaMalloc((void **) &b->d_results, rw);
  if (e == cudaSuccess) e = cudaMemset(b->d_results, 0, rw);                 /* reference: cudaMemset of results */
  if (e == cudaSuccess) e = cudaMalloc((void **) &b->d_counters, 16);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&b->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
  if (e != cudaSuccess) { fmgpu_batch_free(&b); return fm_fail(e, "fmgpu_batch_create", __FILE__, __LINE__); }
  *out = b;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_upload_ascii(fmgpu_batch_t *b, const char *h_ascii)
{
  if (!b || !h_ascii) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(b->device));
  if (b->nq == 0) return FM_SUCCESS;
  /* staged in slices so the ASCII staging buffer stays small next to the packed shard */
  const uint64_t slice = 4ull << 20;                                         /* reads per slice */
  const uint64_t cap = b->nq < slice ? b->nq : slice;
  if (!b->d_ascii) CU_TRY(cudaMalloc((void **) &b->d_ascii, cap * b->len));
  for (uint64_t q0 = 0; q0 < b->nq; q0 += slice) {
    const uint64_t n = (b->nq - q0 < slice) ? b->nq - q0 : slice;
    CU_TRY(cudaMemcpyAsync(b->d_ascii, h_ascii + q0 * b->len, n * b->len, cudaMemcpyHostToDevice, b->stream));
    int32_t rc = fmgpu_pack_queries_device(b->device, b->d_ascii, n, b->len, b->d_packed + q0 * b->wpq, b->stream);
    if (rc) return rc;
  }
  CU_TRY(cudaStreamSynchronize(b->stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_search(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v)
{
  if (!idx || !b) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (idx->device != b->device) return fm_fail_msg(FM_E_BAD_ARGUMENT, "index replica and shard live on different devices");
  CU_TRY(cudaSetDevice(b->device));
  return fm_launch_search(idx, b->d_packed, b->nq, b->len, b->d_results, v, b->stream, NULL);
}

extern "C" int32_t fmgpu_batch_sync(fmgpu_batch_t *b)
{
  if (!b) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaStreamSynchronize(b->stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_download(fmgpu_batch_t *b, uint32_t *h_results)
{
  if (!b || !h_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(b->device));
  if (b->nq) CU_TRY(cudaMemcpyAsync(h_results, b->d_results, b->nq * 8, cudaMemcpyDeviceToHost, b->stream));
  CU_TRY(cudaStreamSynchronize(b->stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_search_timed(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v,
                                            int32_t iters, float *ms_per_iter)
{
  if (!idx || !b || !ms_per_iter || iters < 1) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaEventRecord(b->ev0, b->stream));
  for (int i = 0; i < iters; i++) {
    int32_t rc = fmgpu_batch_search(idx, b, v);
    if (rc) return rc;
  }
  CU_TRY(cudaEventRecord(b->ev1, b->stream));
  CU_TRY(cudaEventSynchronize(b->ev1));
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  *ms_per_iter = ms / (float) iters;
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_batch_count_fetches(const fmgpu_index_t *idx, fmgpu_batch_t *b, uint64_t *nblocks, uint64_t *nsectors)
{
  if (!idx || !b) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (idx->device != b->device) return fm_fail_msg(FM_E_BAD_ARGUMENT, "index replica and shard live on different devices");
  CU_TRY(cudaSetDevice(b->device));
  CU_TRY(cudaMemsetAsync(b->d_counters, 0, 16, b->stream));
  int32_t rc = fm_launch_search(idx, b->d_packed, b->nq, b->len, b->d_results, NULL, b->stream, b->d_counters);
  if (rc) return rc;
  unsigned long long h[2] = { 0, 0 };
  CU_TRY(cudaMemcpyAsync(h, b->d_counters, 16, cudaMemcpyDeviceToHost, b->stream));
  CU_TRY(cudaStreamSynchronize(b->stream));
  if (nblocks) *nblocks = h[0];
  if (nsectors) *nsectors = h[1];
  return FM_SUCCESS;
}

/* locate for a shard: its (L,R) -> positions and hit counts in host memory */
extern "C" int32_t fmgpu_batch_locate(const fmgpu_index_t *idx, fmgpu_batch_t *b, uint32_t max_hits, uint32_t *h_positions, uint32_t *h_nhits)
{
  if (!idx || !b || !h_positions || max_hits == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  if (b->nq == 0) return FM_SUCCESS;
  CU_TRY(cudaSetDevice(b->device));
  uint32_t *d_pos = NULL, *d_n = NULL;
  CU_TRY(cudaMalloc((void **) &d_pos, b->nq * max_hits * 4ull));
  cudaError_t e = cudaMalloc((void **) &d_n, b->nq * 4ull);
  int32_t rc = FM_SUCCESS;
  if (e != cudaSuccess) rc = fm_fail(e, "cudaMalloc(hit counts)", __FILE__, __LINE__);
  if (rc == FM_SUCCESS) rc = fmgpu_locate_device(idx, b->d_results, b->nq, max_hits, d_pos, d_n, b->stream);
  if (rc == FM_SUCCESS) {
    e = cudaMemcpyAsync(h_positions, d_pos, b->nq * max_hits * 4ull, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess && h_nhits) e = cudaMemcpyAsync(h_nhits, d_n, b->nq * 4ull, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    if (e != cudaSuccess) rc = fm_fail(e, "locate D2H", __FILE__, __LINE__);
  }
  cudaFree(d_pos); cudaFree(d_n);
  return rc;
}

extern "C" int32_t fmgpu_index_download_sa(const fmgpu_index_t *idx, uint32_t *h_sa)
{
  if (!idx || !h_sa) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (!idx->sa) return fm_fail_msg(FM_E_BAD_ARGUMENT, "no suffix array on this replica (fmgpu_index_build_sa)");
  CU_TRY(cudaSetDevice(idx->device));
  CU_TRY(cudaMemcpy(h_sa, idx->sa, 4ull * idx->meta.bwtsize, cudaMemcpyDeviceToHost));
  return FM_SUCCESS;
}

extern "C" void *fmgpu_batch_packed(const fmgpu_batch_t *b)  { return b ? (void *) b->d_packed : NULL; }
extern "C" void *fmgpu_batch_results(const fmgpu_batch_t *b) { return b ? (void *) b->d_results : NULL; }
extern "C" void *fmgpu_batch_stream(const fmgpu_batch_t *b)  { return b ? (void *) b->stream : NULL; }

extern "C" int32_t fmgpu_batch_free(fmgpu_batch_t **pb)
{
  if (!pb || !*pb) return FM_SUCCESS;
  fmgpu_batch_t *b = *pb;
  cudaSetDevice(b->device);
  if (b->d_ascii) cudaFree(b->d_ascii);
  if (b->d_packed) cudaFree(b->d_packed);
  if (b->d_results) cudaFree(b->d_results);
  if (b->d_counters) cudaFree(b->d_counters);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream) cudaStreamDestroy(b->stream);
  free(b);
  *pb = NULL;
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ *
 * end to end, host buffers in and out: chunks of the batch flow through
 * H2D -> pack -> search -> D2H on FM_PIPE_STREAMS streams per GPU so the
 * PCIe copies of one chunk overlap the kernels of another.
 * ------------------------------------------------------------------------ */
#define FM_PIPE_STREAMS 4
#define FM_MAX_DEVICES  16

struct fm_pipe_lane {
  cudaStream_t stream;
  cudaEvent_t  h2d_done;      /* the lane's pinned staging buffer may be overwritten after this */
  char     *d_ascii;
  uint32_t *d_packed;
  uint32_t *d_results;
  uint32_t *h_packed;         /* pinned staging for host-packed reads */
  uint64_t  h2d_pending;      /* H2D bytes queued on this lane since h2d_done was last seen complete */
  size_t    cap_ascii, cap_packed, cap_results, cap_hpacked;
};
static fm_pipe_lane g_pipe[FM_MAX_DEVICES][FM_PIPE_STREAMS];

extern "C" int  fm_hostpack_has_simd(void);
extern "C" int  fm_hostpack_threads(void);
extern "C" void fm_hostpack_reads(const char *ascii, uint64_t nq, uint32_t len, uint32_t *packed, int nthreads);
extern "C" void fm_hostpack_stream(const char *ascii, uint64_t nbases, unsigned char *out, int nthreads);

static bool g_pipe_allocated = false;          /* set when a call had to (re)allocate lane buffers: its timing is not representative */

static int32_t fm_pipe_reserve(int device, fm_pipe_lane *ln, size_t ascii, size_t packed, size_t results, size_t hpacked)
{
  if (!ln->stream || ln->cap_ascii < ascii || ln->cap_packed < packed || ln->cap_results < results || ln->cap_hpacked < hpacked) g_pipe_allocated = true;
  CU_TRY(cudaSetDevice(device));
  if (!ln->stream) CU_TRY(cudaStreamCreateWithFlags(&ln->stream, cudaStreamNonBlocking));
  if (!ln->h2d_done) CU_TRY(cudaEventCreateWithFlags(&ln->h2d_done, cudaEventDisableTiming));
  if (ln->cap_ascii < ascii)     { if (ln->d_ascii) cudaFree(ln->d_ascii);     ln->cap_ascii = 0;   CU_TRY(cudaMalloc((void **) &ln->d_ascii, ascii));     ln->cap_ascii = ascii; }
  if (ln->cap_packed < packed)   { if (ln->d_packed) cudaFree(ln->d_packed);   ln->cap_packed = 0;  CU_TRY(cudaMalloc((void **) &ln->d_packed, packed));   ln->cap_packed = packed; }
  if (ln->cap_results < results) { if (ln->d_results) cudaFree(ln->d_results); ln->cap_results = 0; CU_TRY(cudaMalloc((void **) &ln->d_results, results)); ln->cap_results = results; }
  if (ln->cap_hpacked < hpacked) { if (ln->h_packed) cudaFreeHost(ln->h_packed); ln->cap_hpacked = 0; CU_TRY(cudaMallocHost((void **) &ln->h_packed, hpacked)); ln->cap_hpacked = hpacked; }
  return FM_SUCCESS;
}

/* Feeding the GPU from host ASCII reads (variant.reserved, or $FMGPU_FEED):
 *   1 = ASCII over PCIe, 2-bit packing on the device        (PCIe-bound: 100 B/read)
 *   2 = 2-bit packing on the host, 25 B/read over PCIe      (CPU-bound)
 *   3 = hybrid: the copy engine pulls ASCII chunks while the CPU threads pack other chunks; each chunk goes
 *       to whichever resource would otherwise idle (greedy on a PCIe-busy-until estimate)
 *   0 = auto: self-tuning.  Which of 1 and 3 wins depends on the host (with one rank on a 16-core box the hybrid
 *       is 2x faster; with 8 ranks sharing a 32-core box's memory the plain ASCII copy is 8 % faster, because a
 *       host-packed read costs 150 B of host memory traffic against 100 B): the first large calls time mode 3
 *       and mode 1 once each and later calls use the faster one, re-probing the other every 64 calls. */
enum { FM_FEED_AUTO = 0, FM_FEED_ASCII = 1, FM_FEED_HOSTPACK = 2, FM_FEED_HYBRID = 3 };

static double g_feed_rate[4] = { 0, 0, 0, 0 };   /* reads/s last measured per mode on large calls */
static unsigned g_feed_calls = 0;

static int fm_feed_mode(const fmgpu_variant_t *v, uint64_t nq, bool *probe)
{
  int mode = v ? v->reserved : 0;
  const char *env = getenv("FMGPU_FEED");
  *probe = false;
  if (mode == FM_FEED_AUTO && env && *env) mode = atoi(env);
  if (mode >= FM_FEED_ASCII && mode <= FM_FEED_HYBRID) return mode;
  if (!(fm_hostpack_has_simd() && fm_hostpack_threads() >= 2)) return FM_FEED_ASCII;
  if (nq < (1ull << 20)) return g_feed_rate[FM_FEED_ASCII] > g_feed_rate[FM_FEED_HYBRID] ? FM_FEED_ASCII : FM_FEED_HYBRID;
  *probe = true;                                  /* large call: its rate is recorded */
  const unsigned c = g_feed_calls++;
  if (g_feed_rate[FM_FEED_HYBRID] == 0) return FM_FEED_HYBRID;
  if (g_feed_rate[FM_FEED_ASCII] == 0) return FM_FEED_ASCII;
  const int best = g_feed_rate[FM_FEED_ASCII] > g_feed_rate[FM_FEED_HYBRID] ? FM_FEED_ASCII : FM_FEED_HYBRID;
  if (c % 64 == 63) return best == FM_FEED_ASCII ? FM_FEED_HYBRID : FM_FEED_ASCII;     /* re-probe the loser now and then */
  return best;
}

static double fm_now(void)
{
  struct timespec tv;
  clock_gettime(CLOCK_MONOTONIC, &tv);
  return (double) tv.tv_sec + (double) tv.tv_nsec * 1e-9;
}

static double g_pack_s_per_read = 1.6e-9;     /* running estimate of the host packer (all threads), seconds per read */
static const double FM_H2D_BYTES_PER_S = 50e9; /* PCIe gen5 x16 pinned H2D as measured on this pool (47-52 GB/s) */

extern "C" int32_t fmgpu_search_host(fmgpu_index_t *const *replicas, int32_t nrep, const char *h_ascii, uint64_t nq,
                                     uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v)
{
  if (!replicas || nrep < 1 || nrep > FM_MAX_DEVICES || !h_ascii || !h_results || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  for (int g = 0; g < nrep; g++)
    if (!replicas[g] || replicas[g]->device < 0 || replicas[g]->device >= FM_MAX_DEVICES) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad replica");
  if (len % replicas[0]->meta.steps && !replicas[0]->meta.tail_valid) return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a multiple of k");
  if (nq == 0) return FM_SUCCESS;
  const uint32_t wpq = fmgpu_words_per_query(len);
  bool probe = false;
  const int feed = fm_feed_mode(v, nq, &probe);
  const double t_call = fm_now();
  g_pipe_allocated = false;
  /* chunk: 512 K reads, 32-aligned; small batches still get one chunk per lane */
  uint64_t chunk = 1ull << 19;
  const uint64_t lanes = (uint64_t) nrep * FM_PIPE_STREAMS;
  if (nq < chunk * lanes) chunk = ((nq + lanes - 1) / lanes + 31) & ~31ull;
  if (chunk == 0) chunk = 32;
  for (int g = 0; g < nrep; g++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) {
      /* d_ascii doubles as the landing buffer of host-packed streams (len/4 bytes per read) */
      int32_t rc = fm_pipe_reserve(replicas[g]->device, &g_pipe[replicas[g]->device][s],
                                   feed != FM_FEED_HOSTPACK ? chunk * len + 64 : chunk * len / 4 + 64,
                                   chunk * wpq * 4, chunk * 8, feed != FM_FEED_ASCII ? chunk * wpq * 4 + 64 : 0);
      if (rc) return rc;
    }
  /* hybrid feed: keep each GPU's PCIe link fed with just enough ASCII chunks that it does not idle while the CPU
   * threads pack the next chunk; everything else is packed on the host.  The bytes still queued on a link are
   * tracked with the lanes' H2D events, so a slower link (shared host memory, several ranks) shifts work to the
   * CPU by itself and a slower CPU shifts it to the link. */
  for (int g = 0; g < nrep; g++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) g_pipe[replicas[g]->device][s].h2d_pending = 0;
  uint64_t c = 0;
  for (uint64_t q0 = 0; q0 < nq; q0 += chunk, c++) {
    const uint64_t n = (nq - q0 < chunk) ? nq - q0 : chunk;
    const int g = (int)(c % nrep), s = (int)((c / nrep) % FM_PIPE_STREAMS);
    const fmgpu_index_t *idx = replicas[g];
    fm_pipe_lane *ln = &g_pipe[idx->device][s];
    CU_TRY(cudaSetDevice(idx->device));
    int32_t rc;
    bool on_host = (feed == FM_FEED_HOSTPACK);
    if (feed == FM_FEED_HYBRID) {
      uint64_t pending = 0;
      for (int t = 0; t < FM_PIPE_STREAMS; t++) {
        fm_pipe_lane *o = &g_pipe[idx->device][t];
        if (o->h2d_pending && cudaEventQuery(o->h2d_done) == cudaSuccess) o->h2d_pending = 0;
        pending += o->h2d_pending;
      }
      cudaGetLastError();                                             /* cudaErrorNotReady from the queries is not an error */
      double need = FM_H2D_BYTES_PER_S * g_pack_s_per_read * (double) n;   /* bytes the link moves while one chunk is packed */
      const double lo = 0.5 * (double)(n * len), hi = 2.0 * (double)(n * len);
      need = need < lo ? lo : (need > hi ? hi : need);
      on_host = (double) pending >= need;
    }
    if (on_host) {
      CU_TRY(cudaEventSynchronize(ln->h2d_done));                     /* staging buffer free again? */
      ln->h2d_pending = 0;
      const double t0 = fm_now();
      /* the host only streams ASCII -> 2 bit (64 bases per AVX-512 iteration, no per-read work);
       * cutting into reads, reversal and word alignment happen on the GPU (fm_unstream_kernel) */
      const uint64_t sbytes = (((n * len + 3) / 4) + 19) & ~15ull;
      fm_hostpack_stream(h_ascii + q0 * len, n * len, (unsigned char *) ln->h_packed, 0);
      if (n >= 4096) g_pack_s_per_read = 0.75 * g_pack_s_per_read + 0.25 * (fm_now() - t0) / (double) n;
      CU_TRY(cudaMemcpyAsync(ln->d_ascii, ln->h_packed, sbytes, cudaMemcpyHostToDevice, ln->stream));
      CU_TRY(cudaEventRecord(ln->h2d_done, ln->stream));
      ln->h2d_pending += sbytes;
      fm_unstream_kernel<<<(unsigned)((n * wpq + 255) / 256), 256, 0, ln->stream>>>((const uint32_t *) ln->d_ascii, n, len, wpq, ln->d_packed);
      CU_TRY(cudaGetLastError());
    } else {
      CU_TRY(cudaMemcpyAsync(ln->d_ascii, h_ascii + q0 * len, n * len, cudaMemcpyHostToDevice, ln->stream));
      CU_TRY(cudaEventRecord(ln->h2d_done, ln->stream));
      ln->h2d_pending += n * len;
      rc = fmgpu_pack_queries_device(idx->device, ln->d_ascii, n, len, ln->d_packed, ln->stream);
      if (rc) return rc;
    }
    rc = fm_launch_search(idx, ln->d_packed, n, len, ln->d_results, v, ln->stream, NULL);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(h_results + 2 * q0, ln->d_results, n * 8, cudaMemcpyDeviceToHost, ln->stream));
  }
  for (int g = 0; g < nrep; g++) {
    CU_TRY(cudaSetDevice(replicas[g]->device));
    for (int s = 0; s < FM_PIPE_STREAMS; s++) CU_TRY(cudaStreamSynchronize(g_pipe[replicas[g]->device][s].stream));
  }
  if (probe && !g_pipe_allocated) g_feed_rate[feed] = (double) nq / (fm_now() - t_call);   /* self-tuning of the auto feed */
  return FM_SUCCESS;
}

/* Same pipeline for reads that already are 2-bit packed on the host (binary read format: per-read reversed
 * words as produced by fm_hostpack_reads / the device pack kernel): 28 instead of 100 bytes per 100-bp read over
 * PCIe and no conversion anywhere. */
extern "C" int32_t fmgpu_search_host_packed(fmgpu_index_t *const *replicas, int32_t nrep, const uint32_t *h_packed, uint64_t nq,
                                            uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v)
{
  if (!replicas || nrep < 1 || nrep > FM_MAX_DEVICES || !h_packed || !h_results || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  for (int g = 0; g < nrep; g++)
    if (!replicas[g] || replicas[g]->device < 0 || replicas[g]->device >= FM_MAX_DEVICES) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad replica");
  if (len % replicas[0]->meta.steps && !replicas[0]->meta.tail_valid) return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a multiple of k");
  if (nq == 0) return FM_SUCCESS;
  const uint32_t wpq = fmgpu_words_per_query(len);
  uint64_t chunk = 1ull << 19;
  const uint64_t lanes = (uint64_t) nrep * FM_PIPE_STREAMS;
  if (nq < chunk * lanes) chunk = ((nq + lanes - 1) / lanes + 31) & ~31ull;
  if (chunk == 0) chunk = 32;
  for (int g = 0; g < nrep; g++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) {
      int32_t rc = fm_pipe_reserve(replicas[g]->device, &g_pipe[replicas[g]->device][s], 0, chunk * wpq * 4, chunk * 8, 0);
      if (rc) return rc;
    }
  uint64_t c = 0;
  for (uint64_t q0 = 0; q0 < nq; q0 += chunk, c++) {
    const uint64_t n = (nq - q0 < chunk) ? nq - q0 : chunk;
    const fmgpu_index_t *idx = replicas[c % nrep];
    fm_pipe_lane *ln = &g_pipe[idx->device][(c / nrep) % FM_PIPE_STREAMS];
    CU_TRY(cudaSetDevice(idx->device));
    CU_TRY(cudaMemcpyAsync(ln->d_packed, h_packed + q0 * wpq, n * wpq * 4, cudaMemcpyHostToDevice, ln->stream));
    int32_t rc = fm_launch_search(idx, ln->d_packed, n, len, ln->d_results, v, ln->stream, NULL);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(h_results + 2 * q0, ln->d_results, n * 8, cudaMemcpyDeviceToHost, ln->stream));
  }
  for (int g = 0; g < nrep; g++) {
    CU_TRY(cudaSetDevice(replicas[g]->device));
    for (int s = 0; s < FM_PIPE_STREAMS; s++) CU_TRY(cudaStreamSynchronize(g_pipe[replicas[g]->device][s].stream));
  }
  return FM_SUCCESS;
}

/* releases the streams and staging buffers fmgpu_search_host keeps between calls (they are re-created on demand).
 * fmgpu_search_host and this function are not re-entrant: one caller thread at a time, like the reference driver. */
extern "C" int32_t fmgpu_release_pipeline(void)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return FM_SUCCESS; }
  for (int d = 0; d < ndev && d < FM_MAX_DEVICES; d++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) {
      fm_pipe_lane *ln = &g_pipe[d][s];
      if (!ln->stream && !ln->d_ascii && !ln->d_packed && !ln->d_results && !ln->h_packed) continue;
      CU_TRY(cudaSetDevice(d));
      if (ln->stream) { cudaStreamSynchronize(ln->stream); cudaStreamDestroy(ln->stream); }
      if (ln->h2d_done) cudaEventDestroy(ln->h2d_done);
      cudaFree(ln->d_ascii); cudaFree(ln->d_packed); cudaFree(ln->d_results);
      if (ln->h_packed) cudaFreeHost(ln->h_packed);
      memset(ln, 0, sizeof(*ln));
    }
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* page-aligned host memory, pinned when a CUDA device is there to pin it for
 * (on a box without a GPU the loaders still work; nothing can be searched) */
extern "C" void *fmgpu_host_alloc(size_t bytes)
{
  void *p = NULL;
  if (posix_memalign(&p, 4096, bytes ? bytes : 1) != 0) return NULL;
  if (cudaHostRegister(p, bytes ? bytes : 1, cudaHostRegisterPortable) != cudaSuccess) cudaGetLastError();
  return p;
}
extern "C" void fmgpu_host_free(void *p)
{
  if (!p) return;
  if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
  free(p);
}
extern "C" int32_t fmgpu_host_register(void *p, size_t bytes)
{
  CU_TRY(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return FM_SUCCESS;
}
extern "C" int32_t fmgpu_host_unregister(void *p)
{
  CU_TRY(cudaHostUnregister(p));
  return FM_SUCCESS;
}

/* ------------------------------------------------------------------------ */
template <int WIDTH>
static cudaError_t fm_probe_launch(uint32_t grid, const uint4 *table, uint64_t naccess, uint32_t lpt, uint32_t *sink)
{
  fm_gather_probe_kernel<(WIDTH >= 4 ? 2 : 4), WIDTH><<<grid, 256>>>(table, naccess, lpt, sink);
  return cudaGetLastError();
}

extern "C" int32_t fmgpu_gather_probe_ex(int32_t device, uint64_t table_bytes, uint32_t access_bytes,
                                         uint64_t loads_per_thread, int32_t iters, double *accesses_per_second)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!accesses_per_second || table_bytes < 4096 || iters < 1) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  if (!(access_bytes == 16 || access_bytes == 32 || access_bytes == 64 || access_bytes == 128))
    return fm_fail_msg(FM_E_BAD_ARGUMENT, "access_bytes must be 16, 32, 64 or 128");
  const uint64_t naccess = table_bytes / access_bytes;
  uint4 *table = NULL; uint32_t *sink = NULL;
  CU_TRY(cudaMalloc((void **) &table, naccess * access_bytes));
  CU_TRY(cudaMalloc((void **) &sink, 4));
  CU_TRY(cudaMemset(table, 0x5A, naccess * access_bytes));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const uint32_t lpt = (uint32_t)((loads_per_thread + 3) & ~3ull);
  const uint32_t grid = (uint32_t) sms * 8 * 4;                    /* 4 waves of 8 CTAs per SM */
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int i = 0; i <= iters; i++) {                               /* i == 0 is the warm-up */
    CU_TRY(cudaEventRecord(e0));
    cudaError_t e = access_bytes == 16 ? fm_probe_launch<1>(grid, table, naccess, lpt, sink)
                  : access_bytes == 32 ? fm_probe_launch<2>(grid, table, naccess, lpt, sink)
                  : access_bytes == 64 ? fm_probe_launch<4>(grid, table, naccess, lpt, sink)
                                       : fm_probe_launch<8>(grid, table, naccess, lpt, sink);
    if (e != cudaSuccess) return fm_fail(e, "fm_gather_probe_kernel", __FILE__, __LINE__);
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (i > 0 && ms < best) best = ms;
  }
  *accesses_per_second = (double) grid * 256.0 * lpt / (best * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(table); cudaFree(sink);
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_gather_probe(int32_t device, uint64_t table_bytes, uint64_t loads_per_thread, int32_t iters,
                                      double *loads_per_second)
{
  return fmgpu_gather_probe_ex(device, table_bytes, 16, loads_per_thread, iters, loads_per_second);
}

/* locality probe: the 32 lanes of each warp-level load fall in one random window of `window_bytes` */
extern "C" int32_t fmgpu_gather_probe_local(int32_t device, uint64_t table_bytes, uint64_t window_bytes,
                                            uint64_t loads_per_thread, int32_t iters, double *loads_per_second)
{
  int32_t rc = fm_use_device(device);
  if (rc) return rc;
  if (!loads_per_second || window_bytes < 512 || table_bytes < window_bytes || iters < 1 || window_bytes > (1ull << 34))
    return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  const uint32_t window16 = (uint32_t)(window_bytes / 16);
  const uint64_t nwindows = table_bytes / window_bytes;
  uint4 *table = NULL; uint32_t *sink = NULL;
  CU_TRY(cudaMalloc((void **) &table, nwindows * window_bytes));
  CU_TRY(cudaMalloc((void **) &sink, 4));
  CU_TRY(cudaMemset(table, 0x5A, nwindows * window_bytes));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const uint32_t lpt = (uint32_t)((loads_per_thread + 3) & ~3ull);
  const uint32_t grid = (uint32_t) sms * 8 * 4;
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0)); CU_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int i = 0; i <= iters; i++) {
    CU_TRY(cudaEventRecord(e0));
    fm_gather_probe_local_kernel<4><<<grid, 256>>>(table, nwindows, window16, lpt, sink);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (i > 0 && ms < best) best = ms;
  }
  *loads_per_second = (double) grid * 256.0 * lpt / (best * 1e-3);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(table); cudaFree(sink);
  return FM_SUCCESS;
}

/* fused-step fetch counter: blocks of the fused table and SB96 blocks of the leading steps that one search must fetch */
extern "C" int32_t fmgpu_count_fetches_fused_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                                    uint32_t *d_results, void *stream, uint64_t *nfused_blocks, uint64_t *nlead_blocks)
{
  if (!idx || !d_packed || !d_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  if (len == 0 || (len % idx->meta.steps && !idx->meta.tail_valid)) return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a positive multiple of k");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long *d_c = NULL, h[2] = { 0, 0 };
  CU_TRY(cudaMalloc((void **) &d_c, 16));
  CU_TRY(cudaMemsetAsync(d_c, 0, 16, (cudaStream_t) stream));
  int32_t rc = nq ? fm_launch_fused(idx, d_packed, nq, len, d_results, FM_DEFAULT_VARIANT, (cudaStream_t) stream, d_c) : FM_SUCCESS;
  if (rc == FM_SUCCESS) {
    cudaError_t e = cudaMemcpyAsync(h, d_c, 16, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) rc = fm_fail(e, "fetch counters D2H", __FILE__, __LINE__);
  }
  cudaFree(d_c);
  if (nfused_blocks) *nfused_blocks = h[0];
  if (nlead_blocks) *nlead_blocks = h[1];
  return rc;
}

/* fetch counter on caller-owned device memory: one instrumented search (results are written too) */
extern "C" int32_t fmgpu_count_fetches_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                              uint32_t *d_results, void *stream, uint64_t *nblocks, uint64_t *nsectors)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long *d_c = NULL, h[2] = { 0, 0 };
  CU_TRY(cudaMalloc((void **) &d_c, 16));
  CU_TRY(cudaMemsetAsync(d_c, 0, 16, (cudaStream_t) stream));
  int32_t rc = fm_launch_search(idx, d_packed, nq, len, d_results, NULL, (cudaStream_t) stream, d_c);
  if (rc == FM_SUCCESS) {
    cudaError_t e = cudaMemcpyAsync(h, d_c, 16, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) rc = fm_fail(e, "fetch counters D2H", __FILE__, __LINE__);
  }
  cudaFree(d_c);
  if (nblocks) *nblocks = h[0];
  if (nsectors) *nsectors = h[1];
  return rc;
}

/*
 * fm_pipeline.cu -- end to end, host buffers in and out: chunk pipeline with the hybrid host feed; pinned host memory.
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 *
 * Chunks of the batch flow through H2D -> pack -> search -> D2H on FM_PIPE_STREAMS streams ("lanes") per GPU so the
 * PCIe copies of one chunk (H2D and D2H run on separate copy engines) overlap the kernels of another.  All state lives
 * in a handle (fmgpu_pipeline_t): lanes with their staging buffers, and the self-tuning state of the host feed.  One
 * caller thread per handle at a time; fmgpu_search_host / fmgpu_search_host_packed use a process-default handle under
 * a mutex, so they are safe (serialised) from any thread.
 */
#include "fm_internal.h"
#include <time.h>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#define FM_PIPE_STREAMS 4

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

struct fmgpu_pipeline {
  fm_pipe_lane lane[FM_MAX_DEVICES][FM_PIPE_STREAMS];
  bool     allocated;         /* the current call had to (re)allocate lane buffers: its timing is not representative */
  double   feed_rate[4];      /* reads/s last measured per feed mode on large calls */
  unsigned feed_calls;
  double   pack_s_per_read;   /* running estimate of the host packer (all threads), seconds per read */
  fmgpu_pipeline_stats_t stats;
};

extern "C" int  fm_hostpack_has_simd(void);
extern "C" int  fm_hostpack_threads(void);
extern "C" void fm_hostpack_stream(const char *ascii, uint64_t nbases, unsigned char *out, int nthreads);

extern "C" int32_t fmgpu_pipeline_create(fmgpu_pipeline_t **out)
{
  if (!out) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  fmgpu_pipeline_t *p = (fmgpu_pipeline_t *) calloc(1, sizeof(*p));
  if (!p) return fm_fail_msg(FM_E_ALLOCATING_MFASTA, "host allocation failed");
  p->pack_s_per_read = 1.6e-9;
  *out = p;
  return FM_SUCCESS;
}

static int32_t fm_pipe_reserve(fmgpu_pipeline_t *pp, int device, fm_pipe_lane *ln, size_t ascii, size_t packed, size_t results, size_t hpacked)
{
  if (!ln->stream || ln->cap_ascii < ascii || ln->cap_packed < packed || ln->cap_results < results || ln->cap_hpacked < hpacked) pp->allocated = true;
  CU_TRY(cudaSetDevice(device));
  if (!ln->stream) CU_TRY(cudaStreamCreateWithFlags(&ln->stream, cudaStreamNonBlocking));
  if (!ln->h2d_done) CU_TRY(cudaEventCreateWithFlags(&ln->h2d_done, cudaEventDisableTiming));
  if (ln->cap_ascii < ascii)     { if (ln->d_ascii) cudaFree(ln->d_ascii);     ln->cap_ascii = 0;   CU_TRY(cudaMalloc((void **) &ln->d_ascii, ascii));     ln->cap_ascii = ascii; }
  if (ln->cap_packed < packed)   { if (ln->d_packed) cudaFree(ln->d_packed);   ln->cap_packed = 0;  CU_TRY(cudaMalloc((void **) &ln->d_packed, packed));   ln->cap_packed = packed; }
  if (ln->cap_results < results) { if (ln->d_results) cudaFree(ln->d_results); ln->cap_results = 0; CU_TRY(cudaMalloc((void **) &ln->d_results, results)); ln->cap_results = results; }
  if (ln->cap_hpacked < hpacked) { if (ln->h_packed) cudaFreeHost(ln->h_packed); ln->cap_hpacked = 0; CU_TRY(cudaMallocHost((void **) &ln->h_packed, hpacked)); ln->cap_hpacked = hpacked; }
  return FM_SUCCESS;
}

/* Feeding the GPU from host ASCII reads (variant.feed, or $FMGPU_FEED):
 *   1 = ASCII over PCIe, 2-bit packing on the device        (PCIe-bound: 100 B/read)
 *   2 = 2-bit packing on the host, 25 B/read over PCIe      (CPU-bound)
 *   3 = hybrid: the copy engine pulls ASCII chunks while the CPU threads pack other chunks; each chunk goes
 *       to whichever resource would otherwise idle (greedy on a PCIe-busy-until estimate)
 *   0 = auto: self-tuning.  Which of 1 and 3 wins depends on the host (with one rank on a 16-core box the hybrid
 *       is 2x faster; with 8 ranks sharing a 32-core box's memory the plain ASCII copy is 8 % faster, because a
 *       host-packed read costs 150 B of host memory traffic against 100 B): the first large calls time mode 3
 *       and mode 1 once each and later calls use the faster one, re-probing the other every 64 calls. */
static int fm_feed_mode(fmgpu_pipeline_t *pp, const fmgpu_variant_t *v, uint64_t nq, bool *probe)
{
  int mode = v ? v->feed : 0;
  const char *env = getenv("FMGPU_FEED");
  *probe = false;
  if (mode == FMGPU_FEED_AUTO && env && *env) mode = atoi(env);
  if (mode >= FMGPU_FEED_ASCII && mode <= FMGPU_FEED_HYBRID) return mode;
  if (!(fm_hostpack_has_simd() && fm_hostpack_threads() >= 2)) return FMGPU_FEED_ASCII;
  /* a priori: the hybrid feed pays when this process has cores to pack with (16 threads: 2 x the ASCII feed; 4 threads on a
   * host shared by 8 ranks: 10 % slower, profiles/r02_e2e_n8_torchrun_feed*.json) */
  const int first = fm_hostpack_threads() >= 6 ? FMGPU_FEED_HYBRID : FMGPU_FEED_ASCII;
  const int second = first == FMGPU_FEED_HYBRID ? FMGPU_FEED_ASCII : FMGPU_FEED_HYBRID;
  if (nq < (1ull << 20)) {
    if (pp->feed_rate[first] == 0 || pp->feed_rate[second] == 0) return first;
    return pp->feed_rate[first] >= pp->feed_rate[second] ? first : second;
  }
  *probe = true;                                  /* large call: its rate is recorded (the best of a mode's measurements counts) */
  const unsigned c = pp->feed_calls++;
  /* the first large calls measure both modes twice, alternating (a mode's first call also pays first-use costs, and the
   * ranks of a shared host disturb each other's measurements); then the faster one, re-probing the other every 64 calls */
  if (c < 4) return (c & 1u) ? second : first;
  const int best = pp->feed_rate[first] >= pp->feed_rate[second] ? first : second;
  if (c % 64 == 63) return best == first ? second : first;
  return best;
}

static double fm_now(void)
{
  struct timespec tv;
  clock_gettime(CLOCK_MONOTONIC, &tv);
  return (double) tv.tv_sec + (double) tv.tv_nsec * 1e-9;
}

static const double FM_H2D_BYTES_PER_S = 50e9; /* PCIe gen5 x16 pinned H2D as measured on this pool (47-52 GB/s) */

/* waits for everything queued on the lanes of these replicas: called on every exit path after the first enqueue, so
 * that no copy still reads the caller's input or writes its output once the call has returned */
static int32_t fm_pipe_drain(fmgpu_pipeline_t *pp, fmgpu_index_t *const *replicas, int32_t nrep)
{
  int32_t rc = FM_SUCCESS;
  for (int g = 0; g < nrep; g++) {
    if (cudaSetDevice(replicas[g]->device) != cudaSuccess) { cudaGetLastError(); continue; }
    for (int s = 0; s < FM_PIPE_STREAMS; s++) {
      fm_pipe_lane *ln = &pp->lane[replicas[g]->device][s];
      if (!ln->stream) continue;
      const cudaError_t e = cudaStreamSynchronize(ln->stream);
      if (e != cudaSuccess && rc == FM_SUCCESS) rc = fm_fail(e, "cudaStreamSynchronize(pipeline lane)", __FILE__, __LINE__);
    }
  }
  return rc;
}

static int32_t fm_pipe_check_args(fmgpu_pipeline_t *pp, fmgpu_index_t *const *replicas, int32_t nrep, const void *in, const void *out, uint32_t len)
{
  if (!pp || !replicas || nrep < 1 || nrep > FM_MAX_DEVICES || !in || !out || len == 0) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad argument");
  for (int g = 0; g < nrep; g++)
    if (!replicas[g] || replicas[g]->device < 0 || replicas[g]->device >= FM_MAX_DEVICES) return fm_fail_msg(FM_E_BAD_ARGUMENT, "bad replica");
  if (len % replicas[0]->meta.steps && !replicas[0]->meta.tail_valid) return fm_fail_msg(FM_E_QUERY_SHAPE, "read length must be a multiple of k");
  return FM_SUCCESS;
}

static uint64_t fm_pipe_chunk(uint64_t nq, int32_t nrep)
{
  /* chunk: 512 K reads, 32-aligned; small batches still get one chunk per lane */
  uint64_t chunk = 1ull << 19;
  const uint64_t lanes = (uint64_t) nrep * FM_PIPE_STREAMS;
  if (nq < chunk * lanes) chunk = ((nq + lanes - 1) / lanes + 31) & ~31ull;
  return chunk ? chunk : 32;
}

#define PIPE_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fm_fail(e_, #call, __FILE__, __LINE__); goto fail; } } while (0)

extern "C" int32_t fmgpu_pipeline_search_host(fmgpu_pipeline_t *pp, fmgpu_index_t *const *replicas, int32_t nrep, const char *h_ascii,
                                              uint64_t nq, uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v)
{
  int32_t rc = fm_pipe_check_args(pp, replicas, nrep, h_ascii, h_results, len);
  if (rc) return rc;
  if (nq == 0) return FM_SUCCESS;
  for (int g = 0; g < nrep; g++) { rc = fmgpu_index_prepare(replicas[g], len); if (rc) return rc; }   /* tables this length uses: built before anything is queued */
  const uint32_t wpq = fmgpu_words_per_query(len);
  bool probe = false;
  const int feed = fm_feed_mode(pp, v, nq, &probe);
  const double t_call = fm_now();
  pp->allocated = false;
  const uint64_t chunk = fm_pipe_chunk(nq, nrep);
  for (int g = 0; g < nrep; g++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) {
      /* d_ascii doubles as the landing buffer of host-packed streams (len/4 bytes per read) */
      rc = fm_pipe_reserve(pp, replicas[g]->device, &pp->lane[replicas[g]->device][s],
                           feed != FMGPU_FEED_HOSTPACK ? chunk * len + 64 : chunk * len / 4 + 64,
                           chunk * wpq * 4, chunk * 8, feed != FMGPU_FEED_ASCII ? chunk * wpq * 4 + 64 : 0);
      if (rc) return rc;                                              /* nothing queued yet */
    }
  /* hybrid feed: keep each GPU's PCIe link fed with just enough ASCII chunks that it does not idle while the CPU
   * threads pack the next chunk; everything else is packed on the host.  The bytes still queued on a link are
   * tracked with the lanes' H2D events, so a slower link (shared host memory, several ranks) shifts work to the
   * CPU by itself and a slower CPU shifts it to the link. */
  for (int g = 0; g < nrep; g++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) pp->lane[replicas[g]->device][s].h2d_pending = 0;
  /* one feeder per GPU: with several replicas every GPU's chunks (c = g, g + nrep, ...) are issued by their own host thread,
   * so that no link waits for the thread that is busy packing or queueing for another GPU (8 GPUs, one process: one issuing
   * thread was the limit).  The packer threads of the box are split between the feeders. */
  {
    const uint64_t nchunks = (nq + chunk - 1) / chunk;
    const int pack_threads = nrep > 1 ? (fm_hostpack_threads() / nrep > 0 ? fm_hostpack_threads() / nrep : 1) : 0;
    std::atomic<uint64_t> n_host(0), n_link(0);
    std::atomic<int32_t> first_rc(FM_SUCCESS);
    char first_err[512] = "";
    std::mutex err_mutex;
    auto feeder = [&](int g) {
      const fmgpu_index_t *idx = replicas[g];
      int32_t frc = FM_SUCCESS;
      cudaError_t fe = cudaSetDevice(idx->device);
      if (fe != cudaSuccess) frc = fm_fail(fe, "cudaSetDevice(pipeline feeder)", __FILE__, __LINE__);
      for (uint64_t c = (uint64_t) g; c < nchunks && frc == FM_SUCCESS && first_rc.load() == FM_SUCCESS; c += (uint64_t) nrep) {
        const uint64_t q0 = c * chunk, n = (nq - q0 < chunk) ? nq - q0 : chunk;
        fm_pipe_lane *ln = &pp->lane[idx->device][(c / nrep) % FM_PIPE_STREAMS];
        bool on_host = (feed == FMGPU_FEED_HOSTPACK);
        if (feed == FMGPU_FEED_HYBRID) {
          uint64_t pending = 0;
          for (int t = 0; t < FM_PIPE_STREAMS; t++) {
            fm_pipe_lane *o = &pp->lane[idx->device][t];
            if (o->h2d_pending && cudaEventQuery(o->h2d_done) == cudaSuccess) o->h2d_pending = 0;
            pending += o->h2d_pending;
          }
          cudaGetLastError();                                         /* cudaErrorNotReady from the queries is not an error */
          double need = FM_H2D_BYTES_PER_S * pp->pack_s_per_read * (double) n;   /* bytes the link moves while one chunk is packed */
          const double lo = 0.5 * (double)(n * len), hi = 2.0 * (double)(n * len);
          need = need < lo ? lo : (need > hi ? hi : need);
          on_host = (double) pending >= need;
        }
#define FEED_TRY(call) { fe = (call); if (fe != cudaSuccess) { frc = fm_fail(fe, #call, __FILE__, __LINE__); break; } }
        if (on_host) {
          FEED_TRY(cudaEventSynchronize(ln->h2d_done));               /* staging buffer free again? */
          ln->h2d_pending = 0;
          const double t0 = fm_now();
          /* the host only streams ASCII -> 2 bit (64 bases per AVX-512 iteration, no per-read work);
           * cutting into reads, reversal and word alignment happen on the GPU (fm_unstream_kernel) */
          const uint64_t sbytes = (((n * len + 3) / 4) + 19) & ~15ull;
          fm_hostpack_stream(h_ascii + q0 * len, n * len, (unsigned char *) ln->h_packed, pack_threads);
          if (n >= 4096) pp->pack_s_per_read = 0.75 * pp->pack_s_per_read + 0.25 * (fm_now() - t0) / (double) n;   /* per feeder: the link it competes with is its own GPU's */
          FEED_TRY(cudaMemcpyAsync(ln->d_ascii, ln->h_packed, sbytes, cudaMemcpyHostToDevice, ln->stream));
          FEED_TRY(cudaEventRecord(ln->h2d_done, ln->stream));
          ln->h2d_pending += sbytes;
          frc = fmgpu_unstream_device(idx->device, (const uint32_t *) ln->d_ascii, n, len, ln->d_packed, ln->stream);
          if (frc) break;
          n_host += n;
        } else {
          FEED_TRY(cudaMemcpyAsync(ln->d_ascii, h_ascii + q0 * len, n * len, cudaMemcpyHostToDevice, ln->stream));
          FEED_TRY(cudaEventRecord(ln->h2d_done, ln->stream));
          ln->h2d_pending += n * len;
          frc = fmgpu_pack_queries_device(idx->device, ln->d_ascii, n, len, ln->d_packed, ln->stream);
          if (frc) break;
          n_link += n;
        }
        frc = fm_launch_search(idx, ln->d_packed, n, len, ln->d_results, v, ln->stream, NULL);
        if (frc) break;
        FEED_TRY(cudaMemcpyAsync(h_results + 2 * q0, ln->d_results, n * 8, cudaMemcpyDeviceToHost, ln->stream));
#undef FEED_TRY
      }
      if (frc != FM_SUCCESS) {                                        /* the message lives in this thread: hand it to the caller's */
        std::lock_guard<std::mutex> lock(err_mutex);
        if (first_rc.load() == FM_SUCCESS) { first_rc.store(frc); snprintf(first_err, sizeof first_err, "%s", fmgpu_last_error()); }
      }
    };
    /* replicas that share a device share its lanes (a test arrangement: "logical shards"): one thread then issues all chunks,
     * in order -- two feeders would interleave their copies into the same staging buffers */
    bool shared_device = false;
    for (int a = 0; a < nrep; a++)
      for (int b = a + 1; b < nrep; b++) shared_device |= replicas[a]->device == replicas[b]->device;
    if (nrep == 1) feeder(0);
    else if (shared_device) { for (int g = 0; g < nrep; g++) feeder(g); }
    else {
      std::vector<std::thread> th;
      for (int g = 0; g < nrep; g++) th.emplace_back(feeder, g);
      for (auto &t : th) t.join();
    }
    rc = first_rc.load();
    if (rc != FM_SUCCESS) {
      fm_pipe_drain(pp, replicas, nrep);                              /* queued copies still touch the caller's buffers */
      cudaGetLastError();
      return fm_fail_msg(rc, first_err);
    }
    rc = fm_pipe_drain(pp, replicas, nrep);
    if (rc) return rc;
    const double dt = fm_now() - t_call;
    if (probe && !pp->allocated && (double) nq / dt > pp->feed_rate[feed]) pp->feed_rate[feed] = (double) nq / dt;   /* self-tuning of the auto feed */
    if (probe && pp->feed_calls % 64 == 0) { pp->feed_rate[FMGPU_FEED_ASCII] *= 0.97; pp->feed_rate[FMGPU_FEED_HYBRID] *= 0.97; }   /* old maxima fade, so a changed host is noticed */
    pp->stats.calls += 1; pp->stats.last_feed = feed; pp->stats.last_seconds = dt;
    pp->stats.last_reads_host_packed = n_host.load(); pp->stats.last_reads_ascii_over_link = n_link.load();
    pp->stats.host_pack_seconds_per_read = pp->pack_s_per_read;
  }
  return FM_SUCCESS;
}

/* Same pipeline for reads that already are 2-bit packed on the host (binary read format: per-read reversed
 * words as produced by fm_hostpack_reads / the device pack kernel): 28 instead of 100 bytes per 100-bp read over
 * PCIe and no conversion anywhere. */
extern "C" int32_t fmgpu_pipeline_search_host_packed(fmgpu_pipeline_t *pp, fmgpu_index_t *const *replicas, int32_t nrep, const uint32_t *h_packed,
                                                     uint64_t nq, uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v)
{
  int32_t rc = fm_pipe_check_args(pp, replicas, nrep, h_packed, h_results, len);
  if (rc) return rc;
  if (nq == 0) return FM_SUCCESS;
  for (int g = 0; g < nrep; g++) { rc = fmgpu_index_prepare(replicas[g], len); if (rc) return rc; }
  const uint32_t wpq = fmgpu_words_per_query(len);
  const uint64_t chunk = fm_pipe_chunk(nq, nrep);
  for (int g = 0; g < nrep; g++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) {
      rc = fm_pipe_reserve(pp, replicas[g]->device, &pp->lane[replicas[g]->device][s], 0, chunk * wpq * 4, chunk * 8, 0);
      if (rc) return rc;
    }
  uint64_t c = 0;
  for (uint64_t q0 = 0; q0 < nq; q0 += chunk, c++) {
    const uint64_t n = (nq - q0 < chunk) ? nq - q0 : chunk;
    const fmgpu_index_t *idx = replicas[c % nrep];
    fm_pipe_lane *ln = &pp->lane[idx->device][(c / nrep) % FM_PIPE_STREAMS];
    PIPE_TRY(cudaSetDevice(idx->device));
    PIPE_TRY(cudaMemcpyAsync(ln->d_packed, h_packed + q0 * wpq, n * wpq * 4, cudaMemcpyHostToDevice, ln->stream));
    rc = fm_launch_search(idx, ln->d_packed, n, len, ln->d_results, v, ln->stream, NULL);
    if (rc) goto fail;
    PIPE_TRY(cudaMemcpyAsync(h_results + 2 * q0, ln->d_results, n * 8, cudaMemcpyDeviceToHost, ln->stream));
  }
  return fm_pipe_drain(pp, replicas, nrep);
fail:
  fm_pipe_drain(pp, replicas, nrep);
  cudaGetLastError();
  return rc;
}

extern "C" int32_t fmgpu_pipeline_get_stats(const fmgpu_pipeline_t *pp, fmgpu_pipeline_stats_t *out)
{
  if (!pp || !out) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  *out = pp->stats;
  return FM_SUCCESS;
}

/* releases the streams and staging buffers of a pipeline and the handle itself */
extern "C" int32_t fmgpu_pipeline_free(fmgpu_pipeline_t **ppp)
{
  if (!ppp || !*ppp) return FM_SUCCESS;
  fmgpu_pipeline_t *pp = *ppp;
  for (int d = 0; d < FM_MAX_DEVICES; d++)
    for (int s = 0; s < FM_PIPE_STREAMS; s++) {
      fm_pipe_lane *ln = &pp->lane[d][s];
      if (!ln->stream && !ln->d_ascii && !ln->d_packed && !ln->d_results && !ln->h_packed) continue;
      if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
      if (ln->stream) { cudaStreamSynchronize(ln->stream); cudaStreamDestroy(ln->stream); }
      if (ln->h2d_done) cudaEventDestroy(ln->h2d_done);
      cudaFree(ln->d_ascii); cudaFree(ln->d_packed); cudaFree(ln->d_results);
      if (ln->h_packed) cudaFreeHost(ln->h_packed);
    }
  free(pp);
  *ppp = NULL;
  return FM_SUCCESS;
}

/* process-default pipeline behind the handle-less entry points (created on first use, serialised by a mutex) */
static std::mutex g_default_pipe_mutex;
static fmgpu_pipeline_t *g_default_pipe = NULL;

extern "C" int32_t fmgpu_search_host(fmgpu_index_t *const *replicas, int32_t nrep, const char *h_ascii, uint64_t nq,
                                     uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v)
{
  std::lock_guard<std::mutex> lock(g_default_pipe_mutex);
  if (!g_default_pipe) { const int32_t rc = fmgpu_pipeline_create(&g_default_pipe); if (rc) return rc; }
  return fmgpu_pipeline_search_host(g_default_pipe, replicas, nrep, h_ascii, nq, len, h_results, v);
}

extern "C" int32_t fmgpu_search_host_packed(fmgpu_index_t *const *replicas, int32_t nrep, const uint32_t *h_packed, uint64_t nq,
                                            uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v)
{
  std::lock_guard<std::mutex> lock(g_default_pipe_mutex);
  if (!g_default_pipe) { const int32_t rc = fmgpu_pipeline_create(&g_default_pipe); if (rc) return rc; }
  return fmgpu_pipeline_search_host_packed(g_default_pipe, replicas, nrep, h_packed, nq, len, h_results, v);
}

extern "C" int32_t fmgpu_release_pipeline(void)
{
  std::lock_guard<std::mutex> lock(g_default_pipe_mutex);
  return fmgpu_pipeline_free(&g_default_pipe);
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

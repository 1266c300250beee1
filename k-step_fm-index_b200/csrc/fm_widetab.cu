/*
 * fm_widetab.cu -- wide-step table (fm_wide.cuh): construction on the GPU, lead tables, launch plan, fetch counter.
 * (one translation unit of libfmindex_b200.so; shared declarations in fm_internal.h)
 */
#include "fm_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <mutex>
#include <math.h>
#include "fm_wide.cuh"

extern "C" int32_t fmgpu_index_unwiden(fmgpu_index_t *idx)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->wblocks) {
    CU_TRY(cudaSetDevice(idx->device));
    cudaFree(idx->wblocks); idx->wblocks = NULL;
    for (int b = 0; b < 16; b++) { cudaFree(idx->wlead[b]); idx->wlead[b] = NULL; }
    idx->wlead_tried = 0;
  }
  idx->meta.wide_bases = 0; idx->meta.wide_prefix_bits = 0; idx->meta.wide_row_bits = 0; idx->meta.wide_bytes = 0; idx->meta.wide_blocks = 0;
  idx->meta.wide_overflow = 0; idx->meta.wide_tree_nodes = 0; idx->meta.wide_tree_rows = 0; idx->meta.wide_tree_depth = 0;
  idx->meta.wide_exceptional = 0; idx->meta.wide_lanes = 0; idx->meta.wide_entry_words = 0; idx->meta.wide_block_entries = 0;
  fm_budget_account(idx);
  return FM_SUCCESS;
}

static uint32_t fm_bits_for(uint32_t v) { uint32_t b = 0; while (b < 32 && (v >> b)) b++; return b ? b : 1; }

/* prefix bits the automatic choice takes: the most with at least a quarter of a block's slots filled on average (3.75 .. 7.5
 * rows per 15-slot block, 1.875 .. 3.75 per 7-slot block: ~17 .. 34 bytes per text base either way) */
static uint32_t fm_wide_auto_prefix(uint32_t n, uint32_t lanes)
{
  uint32_t pb = 1;
  while (pb < 29 && ((uint64_t) 15 << (pb + 1)) <= (uint64_t) n * 4) pb++;
  return pb + (lanes == 2 ? 1u : 0u);
}
static uint32_t fm_wide_default_lanes(void)
{
  const char *env = getenv("FMGPU_WIDE_LANES");
  return env && *env && atoi(env) == 4 ? 4u : 2u;
}

#define FM_WIDE_MAX_BASES 46
/* Widest step a table over this text can take with 64-bit entries: W bases, a multiple of k, at most 30, with
 * sub_bits + row_bits <= 64; with 96-bit entries: at most 46 bases (a 92-bit key), sub_bits + row_bits <= 96 */
static uint32_t fm_wide_max_bases_ew(uint32_t k, uint32_t pb, uint32_t rb, uint32_t ew)
{
  uint32_t w = (32 * ew + pb - rb) / 2;
  const uint32_t cap = ew == 3 ? FM_WIDE_MAX_BASES : 30u;
  if (w > cap) w = cap;
  return w - w % k;
}
static uint32_t fm_wide_max_bases(uint32_t k, uint32_t pb, uint32_t rb) { return fm_wide_max_bases_ew(k, pb, rb, 2); }

/* Step width for reads of `len` bases: the fewest wide steps S with len = b + S * W for a lead table of b <= 12 bases
 * (134 MB at most), W <= wmax a multiple of k; the smallest such b (the lead table then stays in L2).  On a 2-step index
 * an odd b ends with the derived 1-step rank (tail_ok).  0 when no such width exists (reads shorter than 16 bases are
 * not worth a table). */
static uint32_t fm_wide_bases_for_len(uint32_t k, uint32_t len, uint32_t wmax, bool tail_ok, uint32_t *steps)
{
  *steps = 0;
  if (len < 16 || wmax < 8) return 0;
  for (uint32_t S = 1; S <= len / 8; S++)
    for (uint32_t b = 0; b <= 12 && b < len; b++) {
      if ((len - b) % S) continue;
      const uint32_t w = (len - b) / S;
      if (w > wmax || w < 8 || w % k) continue;
      if (b % k && !(k == 2 && tail_ok)) continue;
      *steps = S;
      return w;
    }
  return 0;
}

extern "C" uint32_t fmgpu_wide_bases_for_words(const fmgpu_index_t *idx, uint32_t len, uint32_t max_entry_words)
{
  if (!idx) return 0;
  const uint32_t n = idx->meta.bwtsize, rb = fm_bits_for(n), pb = fm_wide_auto_prefix(n, fm_wide_default_lanes());
  if (idx->meta.quirk_mask) return 0;
  const uint32_t k = idx->meta.steps;
  const bool tail_ok = idx->meta.tail_valid != 0;
  uint32_t s2 = 0, s3 = 0;
  const uint32_t w2 = fm_wide_bases_for_len(k, len, fm_wide_max_bases_ew(k, pb, rb, 2), tail_ok, &s2);
  /* 96-bit entries (5 instead of 7 per 64-byte block: 4 % instead of 0.3 % of the steps meet a search tree on a random text)
   * when they save a whole step: 100 bp = 8 + 2 x 46 instead of 10 + 3 x 30.  $FMGPU_WIDE_ENTRY_WORDS=2 keeps 64-bit entries. */
  const char *env = getenv("FMGPU_WIDE_ENTRY_WORDS");
  if (max_entry_words < 3 || (env && *env && atoi(env) == 2)) return w2;
  const uint32_t w3 = fm_wide_bases_for_len(k, len, fm_wide_max_bases_ew(k, pb, rb, 3), tail_ok, &s3);
  if (!w2) return w3;
  if (!w3) return w2;
  if (s3 >= s2) return w2;
  /* fewer steps, but fewer entries per block as well: expected fetches = steps x (1 + share of the steps that continue into a
   * search tree), and a tree step costs about twice its fetch (it is a dependent one, and its warp waits for it).  Rows per
   * bucket on the roomy grid, Poisson tail at the block's capacity: 5 packed / 4 unpacked 96-bit entries against 7 64-bit ones
   * (2 Gbp: 2 x 1.08 against 3 x 1.01 -> 46 bases per step; 3.1 Gbp, unpacked: 2 x 1.66 against 3 x 1.06 -> 30). */
  const uint32_t pbr = pb < 30 ? pb + 1 : pb;
  const double mean = (double) n / (double) (1ull << pbr);
  const uint32_t cap3 = (fm_wide_default_lanes() == 2 && n < 0x7FFFFFF0u) ? 5u : 2u * fm_wide_default_lanes(), cap2 = 4u * fm_wide_default_lanes() - 1u;
  auto tail = [&](uint32_t cap) { double p = exp(-mean), acc = 0; for (uint32_t j = 0; j < cap; j++) { acc += p; p *= mean / (double)(j + 1); } return 1.0 - acc; };
  return (double) s3 * (1.0 + 2.0 * tail(cap3)) < (double) s2 * (1.0 + 2.0 * tail(cap2)) ? w3 : w2;
}

extern "C" uint32_t fmgpu_wide_bases_for(const fmgpu_index_t *idx, uint32_t len) { return fmgpu_wide_bases_for_words(idx, len, 3); }

struct fm_wide_shape { uint32_t W, pb, rb, lanes, ew, force_every; };
struct fm_wide_built { uint4 *wblocks; uint64_t total_blocks; uint32_t total_ext; unsigned long long stats[4]; };

/* the whole construction for one block shape; *over_budget is set when the finished size (grid + trees) exceeds the table budget */
template <int LANES, int EW>
static cudaError_t fm_wide_build(fmgpu_index_t *idx, const fm_wide_shape &sh, fm_wide_built *out, bool *over_budget)
{
  typedef typename FmWideKey<EW>::type KeyT;
  const uint32_t k = idx->meta.steps, n = idx->meta.bwtsize, kbits = 2 * k, wbits = 2 * sh.W, sub_bits = wbits - sh.pb;
  const uint32_t hops = sh.W / k, nroots = 1u << sh.pb, bbytes = 32 * LANES;
  const uint64_t nrows = (uint64_t) idx->meta.nblocks * FM_SB_ROWS;
  uint8_t *sym = NULL; KeyT *keys = NULL, *keys2 = NULL; uint64_t *vals = NULL, *vals2 = NULL;
  uint32_t *bstart = NULL, *g0 = NULL, *exc = NULL, *ext = NULL, *extoff = NULL;
  uint4 *wblocks = NULL; void *tmp = NULL; unsigned long long *d_stats = NULL;
  size_t tmp_bytes = 0, tmp2 = 0;
  uint32_t total_ext = 0, nvalid = 0;
  memset(out, 0, sizeof *out);
  *over_budget = false;
  cudaError_t e = cudaMalloc((void **) &sym, nrows);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys, sizeof(KeyT) * (size_t) n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &vals, 8ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &keys2, sizeof(KeyT) * (size_t) n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &vals2, 8ull * n);
  if (e == cudaSuccess) e = cudaMalloc((void **) &d_stats, 32);
  if (e == cudaSuccess) e = cudaMemset(d_stats, 0, 32);
  if (e == cudaSuccess) e = fm_row_symbols(idx, nrows, sym);
  if (e == cudaSuccess) {
    fm_wide_compose_kernel<KeyT><<<(unsigned)(((uint64_t) n + 255) / 256), 256>>>(idx->blocks, idx->meta.nblocks, sym, n, kbits, hops, wbits, keys, vals);
    e = cudaGetLastError();
  }
  /* order of (F(i), i): the sort is stable and the input rows ascend.  Double-buffer form: the two key / value arrays are the
   * sort's only big buffers (the plain form would add a third pair inside its temporary storage: 48 GB at 2 Gbp) */
  cub::DoubleBuffer<KeyT> dk(keys, keys2);
  cub::DoubleBuffer<uint64_t> dv(vals, vals2);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, dk, dv, (int64_t) n, 0, (int)(wbits + 1));
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(NULL, tmp2, (uint32_t *) NULL, (uint32_t *) NULL, (int64_t) nroots);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int64_t) n, 0, (int)(wbits + 1));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess && dk.Current() != keys2) { KeyT *t = keys; keys = keys2; keys2 = t; }       /* sorted pairs end up in keys2 / vals2 */
  if (e == cudaSuccess && dv.Current() != vals2) { uint64_t *t = vals; vals = vals2; vals2 = t; }
  cudaFree(keys); keys = NULL; cudaFree(vals); vals = NULL; cudaFree(sym); sym = NULL;
  if (e == cudaSuccess) e = cudaMalloc((void **) &bstart, 4ull * ((uint64_t) nroots + 1));
  if (e == cudaSuccess) e = cudaMalloc((void **) &g0, 4ull * nroots);
  if (e == cudaSuccess) e = cudaMalloc((void **) &exc, 4ull * ((nroots + 31) / 32));
  if (e == cudaSuccess) e = cudaMemset(exc, 0, 4ull * ((nroots + 31) / 32));
  if (e == cudaSuccess) e = cudaMalloc((void **) &ext, 4ull * nroots);
  if (e == cudaSuccess) e = cudaMalloc((void **) &extoff, 4ull * nroots);
  if (e == cudaSuccess) {
    fm_wide_bstart_kernel<KeyT><<<(nroots + 1 + 255) / 256, 256>>>(keys2, n, sub_bits, nroots, bstart);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    fm_wide_g0_kernel<KeyT><<<(nroots + 255) / 256, 256>>>(idx->blocks, idx->meta.nblocks, kbits, hops, sub_bits, nroots, g0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(&nvalid, bstart + nroots, 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && nvalid) {
    fm_wide_verify_entries_kernel<KeyT><<<(unsigned)(((uint64_t) nvalid + 255) / 256), 256>>>(keys2, vals2, nvalid, sub_bits, bstart, g0, exc);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    fm_wide_verify_buckets_kernel<<<(nroots + 255) / 256, 256>>>(bstart, g0, nroots, n, sh.force_every, exc);
    e = cudaGetLastError();
  }
  FmWideBuild<KeyT> x;
  x.keys = keys2; x.vals = vals2; x.bstart = bstart; x.g0 = g0; x.exc = exc; x.extoff = extoff;
  x.nroots = nroots; x.sub_bits = sub_bits; x.row_bits = sh.rb;
  if (e == cudaSuccess) {
    fm_wide_count_kernel<LANES, EW><<<(nroots + 255) / 256, 256>>>(x, ext, d_stats);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, ext, extoff, (int64_t) nroots);
  if (e == cudaSuccess) {
    uint32_t last_off = 0, last_ext = 0;
    e = cudaMemcpy(&last_off, extoff + (nroots - 1), 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&last_ext, ext + (nroots - 1), 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out->stats, d_stats, 32, cudaMemcpyDeviceToHost);
    total_ext = last_off + last_ext;                             /* <= n / (slots - 1) + ...: far below 2^32 */
  }
  cudaFree(ext); ext = NULL;
  const uint64_t total_blocks = (uint64_t) nroots + total_ext;
  if (e == cudaSuccess && total_blocks >= 0xFFFFFFF0ull) e = cudaErrorInvalidValue;
  if (e == cudaSuccess && !fm_budget_allows(idx, total_blocks * bbytes)) *over_budget = true;
  if (e == cudaSuccess && !*over_budget) e = cudaMalloc((void **) &wblocks, total_blocks * bbytes);
  if (e == cudaSuccess && !*over_budget) {
    fm_wide_fill_roots_kernel<LANES, EW><<<(nroots + 255) / 256, 256>>>(x, wblocks);
    e = cudaGetLastError();
    if (e == cudaSuccess && total_ext) {
      fm_wide_fill_ext_kernel<LANES, EW><<<(total_ext + 255) / 256, 256>>>(x, total_ext, wblocks);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(sym); cudaFree(keys); cudaFree(vals); cudaFree(keys2); cudaFree(vals2); cudaFree(bstart); cudaFree(g0); cudaFree(exc);
  cudaFree(ext); cudaFree(extoff); cudaFree(tmp); cudaFree(d_stats);
  if (e != cudaSuccess || *over_budget) { cudaFree(wblocks); wblocks = NULL; }
  out->wblocks = wblocks; out->total_blocks = total_blocks; out->total_ext = total_ext;
  return e;
}

extern "C" int32_t fmgpu_index_widen(fmgpu_index_t *idx, uint32_t wide_bases, uint32_t prefix_bits, uint32_t lanes)
{
  if (!idx) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null index");
  if (idx->wblocks) return FM_SUCCESS;
  if (idx->meta.quirk_mask) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "AltCounters file with an active padding quirk: the sparse-step table serves it");
  CU_TRY(cudaSetDevice(idx->device));
  const uint32_t k = idx->meta.steps, n = idx->meta.bwtsize;
  const uint32_t rb = fm_bits_for(n);
  if (lanes == 0) lanes = fm_wide_default_lanes();
  if (lanes != 2 && lanes != 4) return fm_fail_msg(FM_E_BAD_ARGUMENT, "wide block lanes must be 2 (64-byte blocks) or 4 (128-byte blocks)");
  const uint32_t bbytes = 32 * lanes;
  uint32_t pb = prefix_bits ? prefix_bits : fm_wide_auto_prefix(n, lanes);
  if (pb > 30) return fm_fail_msg(FM_E_BAD_ARGUMENT, "at most 30 prefix bits");
  if (!prefix_bits && pb < 30) {
    /* roomy grid: one more prefix bit (half the rows per bucket: 0.3 % instead of 8 % of the steps meet a search tree on a
     * random text, +9 % reads/s at 2 Gbp, profiles/r02w_wide_sweep.jsonl) when the doubled grid still is a modest share of
     * this device's memory (40 %: 68.7 GB of a B200 for 2 Gbp) and fits the table budget; $FMGPU_WIDE_ROOMY=0/1 forces */
    size_t fb = 0, tb = 0;
    const char *renv = getenv("FMGPU_WIDE_ROOMY");
    bool roomy = false;
    if (renv && *renv) roomy = atoi(renv) != 0;
    else if (cudaMemGetInfo(&fb, &tb) == cudaSuccess)
      roomy = ((uint64_t) bbytes << (pb + 1)) <= (uint64_t) tb * 2 / 5 && ((uint64_t) bbytes << (pb + 1)) + 24ull * n + (2ull << 30) <= fb &&
              fm_budget_allows(idx, ((uint64_t) bbytes << (pb + 1)) + (pb >= 3 ? (uint64_t) bbytes << (pb - 3) : 0));
    else cudaGetLastError();
    if (roomy) pb += 1;
  }
  uint32_t W = wide_bases ? wide_bases : fm_wide_max_bases(k, pb, rb);
  if (W % k || W < 2 * k || W > FM_WIDE_MAX_BASES) return fm_fail_msg(FM_E_BAD_ARGUMENT, "wide bases must be a multiple of k, at least 2k and at most 46");
  const uint32_t wbits = 2 * W;
  if (pb > wbits) pb = wbits;
  const uint32_t sub_bits = wbits - pb;
  /* entry = rest of the symbol + row number: 64 bits up to 30 bases per step (at 2 Gbp), 96 bits beyond */
  const uint32_t ew = (sub_bits + rb <= 64 && wbits <= 62) ? 2u : 3u;   /* (a 64-bit key also holds the "no symbol" value 2^wbits) */
  if (sub_bits + rb > 96) return fm_fail_msg(FM_E_BAD_ARGUMENT, "wide step too wide for this text: (2 * bases - prefix_bits) + bits of a row number must fit 96");
  const uint32_t nroots = 1u << pb;

  const uint64_t nrows = (uint64_t) idx->meta.nblocks * FM_SB_ROWS;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t) 0; }
  const uint64_t kv = ew == 3 ? 24 : 16;                              /* sorted (key, value) bytes per row */
  const uint64_t build_peak = 2 * kv * n + nrows + 16ull * nroots + (1ull << 30);
  const uint64_t final_peak = kv * n + 16ull * nroots + (uint64_t) nroots * bbytes + (uint64_t) n / 4 * bbytes / 4 + (1ull << 30);
  if ((build_peak > final_peak ? build_peak : final_peak) > free_b)
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough free device memory to build the wide-step table");
  if (!fm_budget_allows(idx, (uint64_t) nroots * bbytes)) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "the wide-step table would exceed the derived-table budget");

  const char *fenv = getenv("FMGPU_WIDE_FORCE_EXC");           /* tests: every N-th bucket is made exceptional */
  fm_wide_shape sh;
  sh.W = W; sh.pb = pb; sh.rb = rb; sh.lanes = lanes; sh.ew = ew; sh.force_every = fenv && *fenv ? (uint32_t) atoi(fenv) : 0u;
  fm_wide_built bt;
  bool over_budget = false;
  /* 96-bit entries on 64-byte blocks: packed five to a block (one-word header) when row numbers and block numbers fit 31 bits */
  const char *penv = getenv("FMGPU_WIDE_PACK");
  const bool packed = ew == 3 && lanes == 2 && n < 0x7FFFFFF0u && (uint64_t) nroots + n / 4 + n / 16 < 0x7FFFFF00ull && !(penv && *penv && atoi(penv) == 0);
  cudaError_t e = lanes == 4 ? (ew == 3 ? fm_wide_build<4, 3>(idx, sh, &bt, &over_budget) : fm_wide_build<4, 2>(idx, sh, &bt, &over_budget))
                             : (packed ? fm_wide_build<2, 5>(idx, sh, &bt, &over_budget) :
                                ew == 3 ? fm_wide_build<2, 3>(idx, sh, &bt, &over_budget) : fm_wide_build<2, 2>(idx, sh, &bt, &over_budget));
  if (over_budget) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "the wide-step table would exceed the derived-table budget");
  if (e != cudaSuccess) {
    cudaGetLastError();                                          /* a failed cudaMalloc stays "last error" otherwise and fails the next attempt's first check */
    if (e == cudaErrorMemoryAllocation) return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "not enough device memory for the wide-step table (the other kernels still serve this index)");
    return fm_fail(e, "fmgpu_index_widen", __FILE__, __LINE__);
  }
  idx->wblocks = bt.wblocks;
  idx->meta.wide_bases = W; idx->meta.wide_prefix_bits = pb; idx->meta.wide_row_bits = rb;
  idx->meta.wide_blocks = bt.total_blocks; idx->meta.wide_bytes = bt.total_blocks * bbytes; idx->meta.wide_lanes = lanes; idx->meta.wide_entry_words = ew;
  idx->meta.wide_overflow = bt.stats[0]; idx->meta.wide_tree_rows = bt.stats[1]; idx->meta.wide_tree_depth = (uint32_t) bt.stats[2];
  idx->meta.wide_tree_nodes = bt.total_ext; idx->meta.wide_exceptional = bt.stats[3];
  idx->meta.wide_block_entries = packed ? 5u : ew == 3 ? 2u * lanes : 4u * lanes - 1u;
  fm_budget_account(idx);
  return FM_SUCCESS;
}

/* Lead table of width b: (L,R) of every b-mer, computed by the plain Coop kernel (a packed b-mer is its own key).  On a
 * 2-step index an odd width ends with the derived 1-step rank.  NULL when the width is not representable or memory is short. */
static std::mutex g_wlead_mutex;
static const uint2 *fm_wide_lead(fmgpu_index_t *idx, uint32_t b)
{
  if (b < 1 || b >= 16) return NULL;
  std::lock_guard<std::mutex> lock(g_wlead_mutex);
  if (idx->wlead[b]) return idx->wlead[b];
  if (idx->wlead_tried & (1u << b)) return NULL;
  idx->wlead_tried |= 1u << b;
  const uint32_t k = idx->meta.steps;
  if (b % k && !(k == 2 && idx->meta.tail_valid)) return NULL;
  if (cudaSetDevice(idx->device) != cudaSuccess) { cudaGetLastError(); return NULL; }
  const uint32_t nkeys = 1u << (2 * b);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return NULL; }
  if ((uint64_t) nkeys * 12 + (1ull << 30) > free_b || !fm_budget_allows(idx, (uint64_t) nkeys * 8)) return NULL;
  if (b % k) fm_build_tail(idx);
  uint32_t *skeys = NULL; uint2 *table = NULL;
  cudaError_t e = cudaMalloc((void **) &skeys, (size_t) nkeys * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **) &table, (size_t) nkeys * 8);
  if (e == cudaSuccess) { fm_iota_kernel<<<(nkeys + 255) / 256, 256>>>(skeys, nkeys); e = cudaGetLastError(); }
  int32_t rc = FM_SUCCESS;
  fmgpu_variant_t v = FM_DEFAULT_VARIANT;
  v.mode = FMGPU_MODE_COOP; v.queries_per_thread = 1; v.threads_per_block = 256;
  if (e == cudaSuccess) rc = fm_launch_search(idx, skeys, nkeys, b, (uint32_t *) table, &v, 0, NULL);
  if (e == cudaSuccess && rc == FM_SUCCESS) e = cudaDeviceSynchronize();
  cudaFree(skeys);
  if (e != cudaSuccess || rc != FM_SUCCESS) { cudaFree(table); cudaGetLastError(); return NULL; }
  idx->wlead[b] = table; idx->meta.wide_bytes += (uint64_t) nkeys * 8;
  return table;
}

/* The plan of a wide search of `len`-base reads on a table of W-base steps: len = b + S * W with a lead table of b bases
 * (0 = none: the search starts from the whole BWT), the fewest steps first.  false = this table does not serve the length. */
struct fm_wide_plan { uint32_t S, b; };
static bool fm_wide_plan_for(const fmgpu_index_t *idx, uint32_t len, fm_wide_plan *pl)
{
  const uint32_t W = idx->meta.wide_bases, k = idx->meta.steps;
  if (!idx->wblocks || !W || len == 0) return false;
  const char *env = getenv("FMGPU_WIDE_LEAD_MAX");             /* widest lead table a plan may use (default 15 bases = 8.6 GB; 12 = 134 MB) */
  const uint32_t maxlead = env && *env ? (uint32_t) atoi(env) : 15u;
  for (uint32_t S = 0; S <= len / W; S++) {
    if (len < S * W) break;
    const uint32_t b = len - S * W;
    if (b >= 16 || b > maxlead) continue;
    if (b % k && !(k == 2 && idx->meta.tail_valid)) continue;
    if (S == 0 && b == 0) continue;
    pl->S = S; pl->b = b;
    return true;
  }
  return false;
}

extern "C" int32_t fmgpu_index_wide_serves(const fmgpu_index_t *idx, uint32_t len)
{
  fm_wide_plan pl;
  if (!idx || !fm_wide_plan_for(idx, len, &pl)) return 0;
  return (pl.b == 0 || idx->wlead[pl.b]) ? 1 : 0;
}

void fm_wide_prepare(fmgpu_index_t *idx, uint32_t len)
{
  fm_wide_plan pl;
  if (!fm_wide_plan_for(idx, len, &pl)) return;
  if (pl.b) fm_wide_lead(idx, pl.b);
  fm_budget_account(idx);
}

typedef void (*fm_wide_fn)(const FmWideParams);
/* one read per lane group in CTAs of 128 or 64 threads: a CTA's registers and warp slots are free again when ITS slowest
 * read is done, so smaller CTAs waste less of the SM on the reads that walk a tree */
template <int LANES, int EW>
static fm_wide_fn fm_pick_wide_small(int tpb)
{
  if (tpb == 128) return fm_search_wide_kernel<LANES, EW, 1, 128, EW != 2 ? 16 : 12, false>;
  if (tpb == 64)  return fm_search_wide_kernel<LANES, EW, 1, 64, 32, false>;
  return NULL;
}

template <int LANES, int EW>
static fm_wide_fn fm_pick_wide(int qpt)
{
  if (qpt == 0) return fm_search_wide_kernel<LANES, EW, 1, 256, 4, true>;         /* instrumented */
  if (qpt == 1) return fm_search_wide_kernel<LANES, EW, 1, 256, EW != 2 ? 8 : 6, false>;   /* 96-bit entries at 32 registers, 8 CTAs per SM: 0.615 vs 0.640 ms at 40 registers */   /* (96-bit entries: 32 registers, 8 CTAs per SM: 0.640 vs 0.649 ms at 38) */
  if (qpt == 2) return fm_search_wide_kernel<LANES, EW, 2, 256, EW != 2 ? 3 : 4, false>;
  if (qpt == 3) return fm_search_wide_kernel<LANES, EW, 3, 256, EW != 2 ? 2 : 3, false>;
  if (qpt == 4) return fm_search_wide_kernel<LANES, EW, 4, 256, 2, false>;
  return NULL;
}

typedef void (*fm_wide_dyn_fn)(const FmWideParams, uint32_t);
template <int LANES, int EW>
static fm_wide_dyn_fn fm_pick_wide_dyn(int qpt)
{
  if (qpt == 1) return fm_search_wide_dyn_kernel<LANES, EW, 1, 256, EW != 2 ? 5 : 6>;
  if (qpt == 2) return fm_search_wide_dyn_kernel<LANES, EW, 2, 256, 3>;
  if (qpt == 3) return fm_search_wide_dyn_kernel<LANES, EW, 3, 256, 2>;
  if (qpt == 4) return fm_search_wide_dyn_kernel<LANES, EW, 4, 256, 2>;
  return NULL;
}

/* dynamic read assignment pays when reads differ much in the length of their walk: trees two or more levels deep holding
 * more than 1 % of the rows (repeat-rich texts).  The one-level trees of a random text's Poisson tail -- 12 % of the steps
 * with 96-bit entries -- do not: the static kernel's prologue overlaps the lead-table lookup with the first block fetch,
 * the dynamic one spends an iteration on it (0.65 vs 0.88 ms per 10 M reads at 46 bases per step).  $FMGPU_WIDE_DYNAMIC=0/1 forces. */
static bool fm_wide_dynamic_enabled(const fmgpu_index_t *idx)
{
  const char *env = getenv("FMGPU_WIDE_DYNAMIC");
  if (env && *env) return atoi(env) != 0;
  return idx->meta.wide_tree_depth >= 2 && idx->meta.wide_tree_rows * 100ull > idx->meta.bwtsize;
}

/* burst kernels (all grid blocks of a read in flight at once): QPT reads per lane group x up to PF blocks per read and chunk */
template <int LANES>
static fm_wide_fn fm_pick_wide_burst(int qpt, int pf)
{
  if (qpt == 0) return fm_search_wide_burst_kernel<LANES, 1, 3, 256, 4, true>;   /* instrumented */
  if (pf <= 3) {
    if (qpt == 1) return fm_search_wide_burst_kernel<LANES, 1, 3, 256, 4, false>;
    if (qpt == 2) return fm_search_wide_burst_kernel<LANES, 2, 3, 256, 2, false>;
    if (qpt == 3) return fm_search_wide_burst_kernel<LANES, 3, 3, 256, 2, false>;
    if (qpt == 4) return fm_search_wide_burst_kernel<LANES, 4, 3, 256, 1, false>;
  } else {
    if (qpt == 1) return fm_search_wide_burst_kernel<LANES, 1, 4, 256, 4, false>;
    if (qpt == 2) return fm_search_wide_burst_kernel<LANES, 2, 4, 256, 2, false>;
    if (qpt == 3) return fm_search_wide_burst_kernel<LANES, 3, 4, 256, 2, false>;
    if (qpt == 4) return fm_search_wide_burst_kernel<LANES, 4, 4, 256, 1, false>;
  }
  return NULL;
}

int32_t fm_launch_wide(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                       uint32_t *d_results, fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters)
{
  if (!idx->wblocks) return fm_fail_msg(FM_E_BAD_ARGUMENT, "FMGPU_MODE_WIDE needs fmgpu_index_widen() on this replica first");
  fm_wide_plan pl;
  if (!fm_wide_plan_for(idx, len, &pl))
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "this wide-step table does not serve the read length (length = lead bases (< 16) + whole steps); see fmgpu_index_wide_serves");
  if (pl.b && !idx->wlead[pl.b])
    return fm_fail_msg(FM_E_NOT_IMPLEMENTED, "the lead table of this read length has not been built: call fmgpu_index_prepare(idx, len) first");
  const uint32_t k = idx->meta.steps, W = idx->meta.wide_bases, lanes = idx->meta.wide_lanes;
  if (v.queries_per_thread < 1 || v.queries_per_thread > 4) v.queries_per_thread = 1;
  FmWideParams p;
  p.wblocks = idx->wblocks; p.blocks = idx->blocks; p.packed = d_packed; p.results = d_results;
  p.nblocks = idx->meta.nblocks; p.nq = (uint32_t) nq; p.nsteps = pl.S;
  p.wpq = fmgpu_words_per_query(len); p.bwtsize = idx->meta.bwtsize;
  p.wbits = 2 * W; p.sub_bits = 2 * W - idx->meta.wide_prefix_bits; p.row_bits = idx->meta.wide_row_bits;
  p.hops = W / k; p.kbits = 2 * k;
  p.nroots = 1u << idx->meta.wide_prefix_bits; p.total_blocks = (uint32_t) idx->meta.wide_blocks;
  p.start = pl.b ? idx->wlead[pl.b] : NULL; p.start_bits = 2 * pl.b;
  p.fetch_counters = d_counters;
  if (d_counters) v.queries_per_thread = 1;
  const char *tenv = getenv("FMGPU_WIDE_TPB");
  uint32_t tpb = 256;
  if (!d_counters && v.queries_per_thread == 1) {
    tpb = tenv && *tenv ? (uint32_t) atoi(tenv) : (v.threads_per_block == 64 || v.threads_per_block == 128 ? (uint32_t) v.threads_per_block : 256u);
    if (tpb != 64 && tpb != 128) tpb = 256;
  }
  uint32_t qper; size_t smem;
  for (;;) {
    qper = (tpb / lanes) * v.queries_per_thread;
    smem = 16 + ((size_t) qper * p.wpq + 4) * 4;
    if (smem <= 200 * 1024) break;
    if (v.queries_per_thread > 1) v.queries_per_thread -= 1;
    else return fm_fail_msg(FM_E_QUERY_SHAPE, "reads too long to stage in shared memory");
  }
  /* $FMGPU_WIDE_BURST=1: the burst kernel instead of the chained state-machine kernel (one dependent fetch per iteration) --
   * measured slower so far (profiles/r02_wide_sweep.jsonl): fewer reads fit an SM with all their blocks in registers;
   * $FMGPU_WIDE_PF: blocks per read and chunk the burst kernel keeps in flight (3 or 4; default: 3 up to three steps, else 4) */
  const char *benv = getenv("FMGPU_WIDE_BURST"), *penv = getenv("FMGPU_WIDE_PF");
  const uint32_t ew = idx->meta.wide_entry_words;
  const bool packed = ew == 3 && idx->meta.wide_block_entries == 5;
  const bool burst = benv && *benv && atoi(benv) != 0 && ew == 2;
  const int pf = penv && *penv ? atoi(penv) : (pl.S <= 3 ? 3 : 4);
  if (!d_counters && !burst && p.nsteps >= 1 && fm_wide_dynamic_enabled(idx)) {
    const int q = v.queries_per_thread;
    fm_wide_dyn_fn dfn = packed ? fm_pick_wide_dyn<2, 5>(q) : ew == 3 ? (lanes == 4 ? fm_pick_wide_dyn<4, 3>(q) : fm_pick_wide_dyn<2, 3>(q))
                                 : (lanes == 4 ? fm_pick_wide_dyn<4, 2>(q) : fm_pick_wide_dyn<2, 2>(q));
    const char *renv = getenv("FMGPU_WIDE_ROUNDS");
    uint32_t rounds = renv && *renv && atoi(renv) >= 1 ? (uint32_t) atoi(renv) : 8u, rpc; size_t dsmem;
    for (;;) {
      rpc = (256 / lanes) * q * rounds;
      dsmem = 16 + ((size_t) rpc * p.wpq + 4) * 4;
      if (dsmem <= 48 * 1024 || rounds == 1) break;
      rounds--;
    }
    if (dfn && dsmem <= 200 * 1024) {
      if (dsmem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) dfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dsmem));
      const uint32_t dgrid = (uint32_t)((nq + rpc - 1) / rpc);
      void *dargs[] = { (void *) &p, (void *) &rpc };
      CU_TRY(cudaLaunchKernel((const void *) dfn, dim3(dgrid), dim3(256), dargs, dsmem, stream));
      return FM_SUCCESS;
    }
  }
  const int qsel = d_counters ? 0 : v.queries_per_thread;
  fm_wide_fn fn = burst ? (lanes == 4 ? fm_pick_wide_burst<4>(qsel, pf) : fm_pick_wide_burst<2>(qsel, pf))
                        : packed ? fm_pick_wide<2, 5>(qsel)
                        : ew == 3 ? (lanes == 4 ? fm_pick_wide<4, 3>(qsel) : fm_pick_wide<2, 3>(qsel))
                                  : (lanes == 4 ? fm_pick_wide<4, 2>(qsel) : fm_pick_wide<2, 2>(qsel));
  if (tpb != 256 && !burst)
    fn = packed ? fm_pick_wide_small<2, 5>(tpb) : ew == 3 ? (lanes == 4 ? fm_pick_wide_small<4, 3>(tpb) : fm_pick_wide_small<2, 3>(tpb))
                                                          : (lanes == 4 ? fm_pick_wide_small<4, 2>(tpb) : fm_pick_wide_small<2, 2>(tpb));
  else tpb = 256;
  if (!fn) return fm_fail_msg(FM_E_BAD_ARGUMENT, "no wide kernel for this variant");
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const uint32_t grid = (uint32_t)((nq + qper - 1) / qper);
  void *args[] = { (void *) &p };
  CU_TRY(cudaLaunchKernel((const void *) fn, dim3(grid), dim3(tpb), args, smem, stream));
  return FM_SUCCESS;
}

extern "C" int32_t fmgpu_count_fetches_wide_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len,
                                                   uint32_t *d_results, void *stream, uint64_t *ngrid_blocks, uint64_t *nsb96_blocks,
                                                   uint64_t *ntree_blocks)
{
  if (!idx || !d_packed || !d_results) return fm_fail_msg(FM_E_BAD_ARGUMENT, "null argument");
  CU_TRY(cudaSetDevice(idx->device));
  unsigned long long *d_c = NULL, h[3] = { 0, 0, 0 };
  CU_TRY(cudaMalloc((void **) &d_c, 24));
  CU_TRY(cudaMemsetAsync(d_c, 0, 24, (cudaStream_t) stream));
  int32_t rc = nq ? fm_launch_wide(idx, d_packed, nq, len, d_results, FM_DEFAULT_VARIANT, (cudaStream_t) stream, d_c) : FM_SUCCESS;
  if (rc == FM_SUCCESS) {
    cudaError_t e = cudaMemcpyAsync(h, d_c, 24, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) rc = fm_fail(e, "fetch counters D2H", __FILE__, __LINE__);
  }
  cudaFree(d_c);
  if (ngrid_blocks) *ngrid_blocks = h[0];
  if (nsb96_blocks) *nsb96_blocks = h[1];
  if (ntree_blocks) *ntree_blocks = h[2];
  return rc;
}

/*
 * fmindex_b200.h -- C ABI of libfmindex_b200.so
 *
 * B200-native (sm_100a) batched k-step FM-index backward search, built from
 * scratch behind the API of achacond/k-step_FM-index so that it drops in for
 * that one path.  Plain C: pointers and sizes only, no C++/torch types.
 *
 * The header has two parts.
 *
 *   PART 1  is the reference's own plugin surface for this path -- the
 *           symbols a search binary links today (common/interface.h:27-41 and
 *           common/common.h:64-96 of the reference).  Names, argument meaning,
 *           ownership and error codes are kept; what changes is that k, d and
 *           the index flavour are read from the file header at run time
 *           instead of -DK_STEPS/-DNUM_CHUNK/-DNUM_COUNTERS, and that sizes
 *           are 64-bit clean internally.
 *
 *   PART 2  is the thin layer the host C code uses to reach the hand-written
 *           CUDA (index residency / re-blocking, query packing, the search
 *           kernels, multi-GPU replicas, the gather roofline probe).  It is
 *           also what a host written in another language would bind.
 *
 * There is NO CPU search path in this library: searchIndexCPU is not
 * exported, and every entry point fails with FM_E_CUDA when no sm_100 device
 * is usable.
 */
#ifndef FMINDEX_B200_H_
#define FMINDEX_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ *
 * Error codes: values of the reference's error_t (common/common.h:36-62).   *
 * 100/101/200/201 double as "this index type is required" codes there; here *
 * loadIndex accepts all four tags, so they are only reported for a file     *
 * whose tag is none of them.                                                *
 * ------------------------------------------------------------------------ */
typedef enum {
  FM_SUCCESS = 0,
  FM_E_OPENING_INDEX_FILE = 1,
  FM_E_ALLOCATING_BWT = 2,
  FM_E_ALLOCATING_FMI = 3,
  FM_E_READING_BWT = 4,
  FM_E_READING_FMI = 5,
  FM_E_SAVING_INDEX_FILE = 6,
  FM_E_SAVING_BWT_FILE = 7,
  FM_E_BUILDING_BWT = 8,
  FM_E_BUILDING_FMI = 9,
  FM_E_OPENING_REFERENCE_FILE = 10,
  FM_E_ALLOCATING_REFERENCE = 11,
  FM_E_READING_MFASTA_FILE = 12,
  FM_E_READING_REFERENCE_FILE = 13,
  FM_E_OPENING_MFASTA_FILE = 14,
  FM_E_ALLOCATING_MFASTA = 15,
  FM_E_ALLOCATING_RESULTS = 16,
  FM_E_OPENING_RESULTS_FILE = 17,
  FM_E_READING_RESULTS_FILE = 18,
  FM_E_NOT_IMPLEMENTED = 19,
  /* new codes (outside the reference's range) */
  FM_E_CUDA = 50,              /* CUDA runtime error or no usable device        */
  FM_E_BAD_ARGUMENT = 51,
  FM_E_UNSUPPORTED_INDEX = 52, /* k not in {1..4}, d not a multiple of 32, ...     */
  FM_E_QUERY_SHAPE = 53,       /* read length not a multiple of k               */
  FM_E_INDEX_VER_BASELINE = 100,
  FM_E_INDEX_VER_INTERLEAVE = 101,
  FM_E_INDEX_VER_BASELINE_AC = 200,
  FM_E_INDEX_VER_INTERLEAVE_AC = 201
} fm_error_t;

/* ======================================================================== *
 * PART 1 -- the reference's interface for this path                         *
 * ======================================================================== */

/* Containers: field order of the reference structs is kept so that a caller
 * compiled against the reference headers sees the same leading layout. */

/* common/common.h:64-69 */
typedef struct {
  uint32_t num;        /* reads in the batch                                    */
  uint32_t size;       /* bases per read                                        */
  char    *h_queries;  /* num*size ASCII bases, no terminators (plain order)    */
  char    *d_queries;  /* opaque: device-side batch state (fmgpu_batch_t)       */
} qrys_t;

/* common/common.h:77-81 */
typedef struct {
  uint32_t  num;
  uint32_t *h_results; /* 2*num: [2q]=L, [2q+1]=R, half-open [L,R) in BWT rows  */
  uint32_t *d_results; /* opaque: device-side state                             */
} res_t;

/* src/fmIndexCPUBaseline.c:54-69 (first 14 fields), then extensions */
typedef struct {
  uint32_t  steps;      /* k                                                    */
  uint32_t  bwtsize;    /* n + 1                                                */
  uint32_t  ncounters;  /* counters per FILE entry: 4^k, or 4^k/2 (AltCounters) */
  uint32_t  nentries;   /* FILE entries (AltCounters: includes padding entry)   */
  uint32_t  chunk;      /* d                                                    */
  uint32_t  nbitmaps;   /* 2*(d/32) (never read by the reference)               */
  uint32_t *h_dollarPositionBWT;
  uint32_t *h_dollarBaseBWT;
  uint32_t *h_modposdollarBWT;
  void     *h_index;    /* raw file entries (bitcnt_t[nentries])                */
  uint32_t *d_dollarPositionBWT; /* unused: '$' rows are folded into the device layout */
  uint32_t *d_dollarBaseBWT;     /* unused                                       */
  uint32_t *d_modposdollarBWT;   /* unused                                       */
  void     *d_index;    /* opaque: fmgpu replica set created by transferCPUtoGPU */
  /* --- extensions --- */
  uint32_t  tag;        /* 100 | 101 | 200 | 201                                */
  uint32_t  entry_words;
} fmi_t;

/* common/interface.h:27 ; loaders src/fmIndexCPUBaseline.c:71-143 and
 * src/fmIndexCPUBaseline-AltCounters.c:71-143.  Accepts tags 100/101/200/201. */
int32_t loadIndex(const char *fn, void **index);
/* common/interface.h:33 ; src/fmIndexCPUBaseline.c:145-155 (frees h_index only) */
int32_t freeIndex(void **index);
/* common/common.h:86 ; common/common.c:132-199.  Plain (non-interleaved) order:
 * the warp interleave of INTERLEAVING_QUERIES is replaced by device-side packing. */
int32_t loadQueries(char *fn, uint32_t sizeQuery, uint32_t numQueries, void **queries);
/* common/interface.h:29 ; common/common.c:248-260 */
int32_t initResults(uint32_t numresults, void **results);
/* common/common.h:89-90,93 ; common/common.c:201-246,324-341.  saveResults
 * writes "<fn>.res.gpu" in the reference's text format. */
int32_t writeResults(char *fn, uint32_t *results, uint32_t numqueries);
int32_t loadResults(char *fn, void **results);
int32_t saveResults(char *fn, void *results, void *index);
/* common/common.h:91-92 ; common/common.c:262-280 */
int32_t freeQueries(void **queries);
int32_t freeResults(void **results);
/* common/common.h:94 ; common/common.c:282-310 */
char   *errorCommon(int32_t e);
/* common/common.h:83 ; common/common.c:28-33 */
double  sampleTime(void);

/* The six symbols a reference .cu object exports (e.g.
 * src/fmIndexGPU-Coop-2Step.cu:231,250,287,295,319,330).
 *
 * transferCPUtoGPU: re-blocks the file entries into the device layout on the
 *   first configured GPU, replicates it to the others over NVLink peer copies,
 *   builds the sparse-step table on every replica of an index larger than L2
 *   ($FMGPU_MODE=sparse|fused|task|coop forces a kernel family),
 *   shards the batch contiguously (32-aligned) over the GPUs, uploads the
 *   ASCII reads and packs them to 2 bit on the device, allocates results.
 * searchIndexGPU: launches the search on every shard and waits (kernels only,
 *   like the reference's timed region); kernel family = $FMGPU_MODE, else the
 *   sparse-step (or fused-step) kernel when the replica has that table, else Coop;
 *   or whatever fmgpu_set_variant selected.  Returns void like the reference;
 *   a CUDA failure prints file:line and exits (reference HandleError, :88-93).
 * transferGPUtoCPU: per-GPU D2H of its (L,R) shard straight into h_results.
 */
int32_t transferCPUtoGPU(void *index, void *queries, void *results);
void    searchIndexGPU(void *index, void *queries, void *resIntervals);
int32_t transferGPUtoCPU(void *results);
int32_t freeIndexGPU(void **index);
int32_t freeQueriesGPU(void **queries);
int32_t freeResultsGPU(void **results);

/* ======================================================================== *
 * PART 2 -- thin C ABI to the CUDA side                                     *
 * ======================================================================== */

typedef struct fmgpu_index fmgpu_index_t;   /* device-resident re-blocked index (one GPU) */
typedef struct fmgpu_batch fmgpu_batch_t;   /* device-resident query shard + its results  */

enum { FMGPU_MODE_TASK = 0, FMGPU_MODE_COOP = 1, FMGPU_MODE_FUSED = 2, FMGPU_MODE_SPARSE = 3, FMGPU_MODE_WIDE = 4 };

/* Kernel variant.  Zero-initialised = library defaults. */
typedef struct {
  int32_t mode;               /* FMGPU_MODE_TASK: one thread per query (both endpoints);
                                 FMGPU_MODE_COOP: lane pair per query, L lane + R lane,
                                 block fetch shared through warp shuffles;
                                 FMGPU_MODE_FUSED: fused-step table (fmgpu_index_fuse): a lane group fetches one
                                 32/64/128-byte block with 256-bit loads and consumes up to 4 bases per step;
                                 FMGPU_MODE_SPARSE: sparse-step table (fmgpu_index_sparsify): one 64-byte block (grid
                                 root or search-tree node) per fetch, up to 14 bases per step, one state machine per read;
                                 FMGPU_MODE_WIDE: wide-step table (fmgpu_index_widen): one 64-byte block per fetch, up to 46
                                 bases per step, the block computed from the read alone (both interval ends share it) */
  int32_t queries_per_thread; /* independent queries interleaved per thread / lane pair / lane group: 1, 2 or 4 (sparse: 1..4);
                                 0 = the kernel family's default (sparse: 3 with static, 1 with dynamic read assignment) */
  int32_t threads_per_block;  /* 128, 256 or 512                                      */
  int32_t feed;               /* fmgpu_search_host only (also $FMGPU_FEED): FMGPU_FEED_AUTO, _ASCII (upload ASCII, pack
                                 on the GPU), _HOSTPACK (pack to 2 bit on the host with OpenMP + AVX-512, upload
                                 25 B/read), _HYBRID (copy engine pulls ASCII chunks while the CPU packs others) */
} fmgpu_variant_t;
enum { FMGPU_FEED_AUTO = 0, FMGPU_FEED_ASCII = 1, FMGPU_FEED_HOSTPACK = 2, FMGPU_FEED_HYBRID = 3 };

/* Shape of the device layout ("SB96": per-symbol blocks of 96 BWT rows,
 * 16 bytes = {u32 rank at block start, 96 indicator bits}); see DESIGN.md. */
typedef struct {
  uint32_t steps;        /* k                                                   */
  uint32_t bwtsize;
  uint32_t nsymbols;     /* 4^k                                                 */
  uint32_t nblocks;      /* blocks per symbol (stride)                          */
  uint32_t source_tag;   /* tag of the file it was derived from                 */
  uint32_t quirk_start;  /* first BWT row of the AltCounters padding quirk, or 0xFFFFFFFF */
  uint32_t quirk_mask;   /* 2 bits per symbol: value the reference AC searcher adds there  */
  uint32_t source_steps; /* k of the file it was derived from (files with k = 3, 4 are searched through the
                            2-step index their first two BWT layers define; steps is 2 then)              */
  uint64_t nbytes;       /* size of the block table                             */
  uint32_t fused_bases;  /* bases per fused step (0 = no fused table), see fmgpu_index_fuse */
  uint32_t fused_lanes;  /* lanes per fused block (block = 32 bytes per lane)   */
  uint64_t fused_bytes;  /* size of the fused table                             */
  /* odd read lengths on a 2-step index: the last base is consumed by a 1-step rank derived from the 2-step
   * table (result = what the 1-step index of the same text gives; the reference itself is undefined there,
   * SURVEY.md App. C-5).  Valid for k = 2 indexes without the AltCounters padding quirk. */
  uint32_t tail_valid;
  uint32_t tail_row;      /* row whose layer-1 char is '$' (dollarPositionBWT[1])             */
  uint32_t tail_base;     /* its layer-0 char (dollarBaseBWT[1] & 3)                           */
  uint32_t tail_const[4]; /* C1[c] - sum over c1 of rank2(c | c1<<2, 0)                         */
  uint32_t start_bases;   /* bases covered by the fused kernel's start table (12), 0 = none    */
  /* sparse-step table (fmgpu_index_sparsify), 0 = none */
  uint32_t sparse_bases;       /* bases per sparse step                                         */
  uint32_t sparse_lambda;      /* target occurrences per block                                  */
  uint64_t sparse_bytes;       /* blocks (+ start / lead tables)                                */
  uint64_t sparse_blocks;      /* number of blocks: grid + tree nodes                           */
  uint64_t sparse_overflow;    /* buckets holding more occurrences than slots (roots of search trees) */
  uint32_t sparse_start_bases; /* bases covered by the sparse kernel's start table, 0 = none    */
  uint32_t sparse_lanes;       /* lanes per block: 2 = 64-byte blocks (15 slots), 4 = 128-byte blocks (31 slots) */
  /* tail table: the derived 1-step rank above re-blocked so that the last base of an odd-length read costs one
   * block fetch instead of four; built on this replica by its first odd-length search (a quarter of nbytes), 0 = not
   * built (yet, or $FMGPU_TAIL_TABLE=0, or no memory: the four-fetch derivation is used) */
  uint64_t tail_bytes;
  /* grid of the sparse-step table: every wide symbol owns this many blocks (root of (symbol, row) is computed, not looked up) */
  uint32_t sparse_uniform_nb;
  uint32_t sa_rate;            /* suffix array kept for locate: 1 = every row (fmgpu_index_build_sa), s > 1 = the rows whose
                                  text position is a multiple of s (fmgpu_index_build_sa_sampled), 0 = none */
  uint64_t sa_bytes;           /* its size: 4 bytes per BWT row when full, ~4/s + 1/6 bytes per row when sampled */
  /* search trees of the sparse-step table: buckets with more occurrences than a block has slots (repeats, skewed
   * texts) are the root of a tree of blocks over their occurrence list; sparse_overflow counts those buckets */
  uint64_t sparse_tree_nodes;  /* blocks below the grid                                          */
  uint64_t sparse_tree_rows;   /* occurrences living in trees (of bwtsize)                       */
  uint32_t sparse_tree_depth;  /* levels below the grid of the deepest tree                      */
  uint32_t reserved1;
  uint64_t derived_bytes;      /* everything this replica derived from its SB96 table: sparse + fused + tail + SA tables */
  uint64_t budget_bytes;       /* the limit derived_bytes is kept under ($FMGPU_TABLE_BUDGET_GB / fmgpu_set_table_budget), 0 = none */
  /* wide-step table (fmgpu_index_widen), 0 = none */
  uint32_t wide_bases;         /* bases per wide step (<= 30)                                    */
  uint32_t wide_prefix_bits;   /* top bits of a wide symbol that select its 128-byte block (2^bits grid blocks) */
  uint32_t wide_row_bits;      /* bits of a row number inside a 64-bit entry                     */
  uint32_t wide_tree_depth;    /* levels below the grid of the deepest search tree               */
  uint64_t wide_bytes;         /* blocks (+ lead tables)                                         */
  uint64_t wide_blocks;        /* grid + tree nodes                                              */
  uint64_t wide_overflow;      /* buckets with more than 15 rows (roots of search trees)         */
  uint64_t wide_tree_nodes;    /* blocks below the grid                                          */
  uint64_t wide_tree_rows;     /* rows living in trees (of bwtsize)                              */
  uint64_t wide_exceptional;   /* buckets whose steps run on the block table (a suffix shorter than the step sorts into them) */
  uint32_t wide_lanes;         /* lanes per block: 2 = 64-byte blocks (7 entries), 4 = 128-byte blocks (15 entries) */
  uint32_t wide_entry_words;   /* 2 = 64-bit entries (steps up to 30 bases), 3 = 96-bit entries (steps up to 46 bases) */
  uint32_t wide_block_entries; /* entries per block: 7 / 15 (64-bit entries, 64- / 128-byte blocks), 5 (96-bit entries packed into 64 bytes), 4 / 8 (96-bit, unpacked) */
  uint32_t reserved3;
} fmgpu_index_meta_t;

/* what the last transferCPUtoGPU / searchIndexGPU / transferGPUtoCPU sequence of this process did (wall-clock seconds of
 * each stage, CUDA-event milliseconds of each GPU's search kernels): the numbers the reference main() cannot print */
typedef struct {
  int32_t  ndev;
  int32_t  searches;              /* searchIndexGPU calls since the index was transferred                          */
  double   index_h2d_reblock_s;   /* file entries H2D + re-block on the first GPU                                  */
  double   peer_copy_s[16];       /* [g]: the cudaMemcpyPeer of the block table to GPU g (g >= 1), alone             */
  double   table_build_s[16];     /* [g]: sparse-step / fused-step table built on GPU g                             */
  double   queries_h2d_pack_s;    /* all shards: ASCII H2D + 2-bit pack                                            */
  double   results_d2h_s;         /* all shards: (L,R) D2H                                                         */
  float    search_ms[16];         /* [g]: kernels of the last searchIndexGPU on GPU g, CUDA events on its stream   */
  uint64_t index_file_bytes, table_bytes, query_bytes, result_bytes;
  double   context_init_s[16];    /* [g]: first use of GPU g by this process (CUDA context), before anything is copied */
  double   replicate_s[16];       /* [g]: whole fmgpu_index_replicate for GPU g (allocation + peer mapping + the copy in peer_copy_s) */
} fmgpu_transfer_stats_t;
int32_t fmgpu_get_transfer_stats(fmgpu_transfer_stats_t *out);
/* searchIndexGPU with an error code instead of exit(): FM_E_BAD_ARGUMENT when transferCPUtoGPU was not called for this pair */
int32_t fmgpu_search_index(void *index, void *queries, void *resIntervals);

/* devices ---------------------------------------------------------------- */
int32_t fmgpu_device_count(void);                 /* usable sm_100 devices; 0 if none */
int32_t fmgpu_device_warmup(int32_t device);      /* creates the device's CUDA context (first use costs 0.3-0.4 s) */
/* devices used by transferCPUtoGPU / searchIndexGPU.  Default: the list in
 * $FMGPU_DEVICES ("0,1,2"), else device 0. */
int32_t fmgpu_set_devices(const int32_t *devices, int32_t ndevices);
int32_t fmgpu_set_variant(const fmgpu_variant_t *v);   /* variant used by searchIndexGPU */
const char *fmgpu_last_error(void);

/* index residency / layout stage ------------------------------------------ */
/* Uploads raw file entries (any of the four tags) to `device`, re-blocks them
 * there into SB96 and frees the raw copy.  Replaces the cudaMalloc+cudaMemcpy
 * of the reference's transferCPUtoGPU (src/fmIndexGPU-Coop-2Step.cu:250-285). */
int32_t fmgpu_index_create(int32_t device, uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize,
                           uint32_t ncounters, uint32_t nentries,
                           const uint32_t *dollarPositionBWT, const uint32_t *dollarBaseBWT,
                           const uint32_t *h_entries, fmgpu_index_t **out);
/* same, entries already on `device` (e.g. written by the GPU index builder) */
int32_t fmgpu_index_create_from_device(int32_t device, uint32_t tag, uint32_t steps, uint32_t chunk, uint32_t bwtsize,
                                       uint32_t ncounters, uint32_t nentries,
                                       const uint32_t *dollarPositionBWT, const uint32_t *dollarBaseBWT,
                                       const uint32_t *d_entries, fmgpu_index_t **out);
/* replica on another GPU of this process: cudaMemcpyPeer over NVLink */
int32_t fmgpu_index_replicate(const fmgpu_index_t *src, int32_t device, fmgpu_index_t **out);
double  fmgpu_last_peer_copy_seconds(void);       /* duration of the copy inside this thread's last fmgpu_index_replicate */
/* replica in another PROCESS: allocate an empty table of the same shape, then
 * fill fmgpu_index_blocks() with a broadcast (NCCL) from the owner */
int32_t fmgpu_index_alloc_like(int32_t device, const fmgpu_index_meta_t *meta, fmgpu_index_t **out);
/* Fused-step table, composed on the GPU from this replica's own block table: one fused step = fused_bases/k
 * reference LF steps (exactly), one 256-bit load per lane of a `lanes`-lane group per rank.  fused_bases 0 =
 * the widest of 4,3,2 that is a multiple of k and fits `budget_bytes` (0 = ~69 GB or free memory); lanes 0 = 2.
 * AltCounters files with an active padding quirk are served (a few phantom occurrences ride beside the bitmaps).
 * FM_E_NOT_IMPLEMENTED when nothing fits. */
int32_t fmgpu_index_fuse(fmgpu_index_t *idx, uint32_t fused_bases, uint32_t lanes, uint64_t budget_bytes);
int32_t fmgpu_index_unfuse(fmgpu_index_t *idx);
/* Sparse-step table, built on the GPU from this replica's own block table: one sparse step = sparse_bases/k
 * reference LF steps (exactly).  Per wide symbol the occurrence rows are cut into a uniform grid of buckets, the same
 * number for every symbol (meta.sparse_uniform_nb, ~lambda rows per bucket on average), so a bucket's block is
 * computed, never looked up: one block fetch per rank (csrc/fm_sparse.cuh).  A bucket with more occurrences than a block
 * has slots (repeats, skewed texts) becomes the root of a search tree of blocks over its occurrence list
 * (meta.sparse_overflow / sparse_tree_*): depth + 1 fetches for those, on any text -- nothing falls back to the block
 * table.  sparse_bases 0 = the widest multiple of k up to 14 that leaves lambda rows per symbol on average (14 bases from
 * 1.34 Gbp, 12 from 84 Mbp ...; explicit widths: multiples of k up to 14); lanes 0 = 2 (64-byte blocks, 15 slots; 4 =
 * 128-byte blocks, 31 slots); lambda 0 = 5 / 12.  The grid takes ~32*lanes/lambda bytes per text base whatever the
 * width, the trees at most 32*lanes/(slots-1) bytes per row living in them.  AltCounters files with an active
 * padding-entry quirk are served too (a few phantom occurrences).  FM_E_NOT_IMPLEMENTED when memory or the table budget
 * does not suffice. */
int32_t fmgpu_index_sparsify(fmgpu_index_t *idx, uint32_t sparse_bases, uint32_t lambda, uint32_t lanes);
int32_t fmgpu_index_unsparsify(fmgpu_index_t *idx);
/* Wide-step table, built on the GPU from this replica's own block table: one wide step = wide_bases/k reference LF steps
 * (exactly), ONE block fetch for both interval ends.  The rows are sorted by the wide symbol in front of them (wide_bases <= 46
 * bases = a 92-bit key); the top prefix_bits of the symbol select the block -- computed from the read, never looked up, the same
 * for both ends -- and the block's entries carry the rest of the symbol with the row (csrc/fm_wide.cuh): 64-bit entries (7 per
 * 64-byte block) for steps up to 30 bases, 96-bit entries (5 per block) up to 46.  Buckets with more rows than a block holds
 * become search trees as in the sparse-step table; the few buckets a suffix shorter than the step sorts into are detected by
 * the builder (every entry and every bucket is verified against the composed LF walk) and answered with plain steps.
 * A read of len = b + S * wide_bases bases (b < 16 from a lead table of all b-mers) costs S block fetches: 2 for 100 bp at 46
 * bases per step, 3 at 30.  wide_bases 0 = the widest 64-bit entries allow (30 up to 4 G rows); prefix_bits 0 = 1.9 .. 3.75 rows
 * per 64-byte bucket on average (17 .. 34 bytes per text base), one bit more ("roomy", 0.9 .. 1.9 rows) when that grid is at most
 * 40 % of the device's memory; lanes 0 = 2 (64-byte blocks; 4 = 128-byte blocks).  A table serves the read lengths its width
 * divides (after the lead bases): fmgpu_wide_bases_for(idx, len) names the width to build for a length,
 * fmgpu_index_wide_serves(idx, len) tells whether an existing table (with its lead table, see fmgpu_index_prepare) does.
 * AltCounters files with an active padding quirk are refused (FM_E_NOT_IMPLEMENTED, like memory or budget shortage): the
 * sparse-step table serves them. */
int32_t  fmgpu_index_widen(fmgpu_index_t *idx, uint32_t wide_bases, uint32_t prefix_bits, uint32_t lanes);
int32_t  fmgpu_index_unwiden(fmgpu_index_t *idx);
uint32_t fmgpu_wide_bases_for(const fmgpu_index_t *idx, uint32_t len);      /* 0 = no width serves this length */
/* the same restricted to entries of at most max_entry_words 32-bit words (2: steps up to 30 bases, whose table needs a third
 * less scratch memory to build): the width to retry with when the 96-bit table does not fit */
uint32_t fmgpu_wide_bases_for_words(const fmgpu_index_t *idx, uint32_t len, uint32_t max_entry_words);
int32_t  fmgpu_index_wide_serves(const fmgpu_index_t *idx, uint32_t len);   /* 1 / 0 */
/* Builds NOW (synchronously) whatever a search of `len`-base reads on this replica may use: the tail table for odd
 * lengths on a 2-step index, the lead tables of the sparse-step plan.  The search entry points themselves never build
 * or allocate: they pick among the tables that exist (same results, more fetches when one is missing).
 * transferCPUtoGPU and fmgpu_search_host call this; callers of fmgpu_batch_search / fmgpu_search_device may. */
int32_t fmgpu_index_prepare(fmgpu_index_t *idx, uint32_t len);
/* one budget for all derived tables of a replica (sparse, fused, tail, lead tables, suffix array); a table that would
 * exceed it is not built (FM_E_NOT_IMPLEMENTED from the explicit builders).  0 = no limit ($FMGPU_TABLE_BUDGET_GB). */
int32_t fmgpu_set_table_budget(uint64_t bytes);
int32_t fmgpu_index_get_meta(const fmgpu_index_t *idx, fmgpu_index_meta_t *meta);
void   *fmgpu_index_blocks(const fmgpu_index_t *idx);     /* device pointer */
int32_t fmgpu_index_device(const fmgpu_index_t *idx);
int32_t fmgpu_index_free(fmgpu_index_t **idx);

/* query shards ------------------------------------------------------------ */
int32_t fmgpu_batch_create(int32_t device, uint64_t nqueries, uint32_t len, uint32_t steps, fmgpu_batch_t **out);
/* H2D of ASCII reads (nqueries*len bytes, plain order) + 2-bit pack kernel */
int32_t fmgpu_batch_upload_ascii(fmgpu_batch_t *b, const char *h_ascii);
/* async launch of the search kernels on the shard's stream */
int32_t fmgpu_batch_search(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v);
int32_t fmgpu_batch_sync(fmgpu_batch_t *b);
/* the same launch bracketed by the shard's two CUDA events (no host wait); after fmgpu_batch_sync,
 * fmgpu_batch_last_ms gives the milliseconds between them */
int32_t fmgpu_batch_search_timed_async(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v);
int32_t fmgpu_batch_last_ms(fmgpu_batch_t *b, float *ms);
/* D2H of 2*nqueries u32 */
int32_t fmgpu_batch_download(fmgpu_batch_t *b, uint32_t *h_results);
/* `iters` searches timed with CUDA events on the shard's stream; ms per search */
int32_t fmgpu_batch_search_timed(const fmgpu_index_t *idx, fmgpu_batch_t *b, const fmgpu_variant_t *v,
                                 int32_t iters, float *ms_per_iter);
/* sum over LF steps of |{block(L), block(R)}| (necessary 16-byte block fetches)
 * and of |{sector(L), sector(R)}| (necessary 32-byte sectors) for this shard */
int32_t fmgpu_batch_count_fetches(const fmgpu_index_t *idx, fmgpu_batch_t *b, uint64_t *nblocks, uint64_t *nsectors);
void   *fmgpu_batch_packed(const fmgpu_batch_t *b);       /* device pointers */
void   *fmgpu_batch_results(const fmgpu_batch_t *b);
void   *fmgpu_batch_stream(const fmgpu_batch_t *b);       /* cudaStream_t */
int32_t fmgpu_batch_free(fmgpu_batch_t **b);

/* caller-owned device memory (e.g. torch tensors) ------------------------- */
uint32_t fmgpu_words_per_query(uint32_t len);
int32_t fmgpu_pack_queries_device(int32_t device, const char *d_ascii, uint64_t nqueries, uint32_t len,
                                  uint32_t *d_packed, void *stream);
int32_t fmgpu_search_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nqueries, uint32_t len,
                            uint32_t *d_results, const fmgpu_variant_t *v, void *stream);

/* end to end: host ASCII reads in, host (L,R) out; the batch is cut into
 * chunks and upload / pack / search / download overlap on several streams per
 * GPU, shards spread over `nreplicas` GPUs */
int32_t fmgpu_search_host(fmgpu_index_t *const *replicas, int32_t nreplicas, const char *h_ascii,
                          uint64_t nqueries, uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v);

/* same for reads that already are 2-bit packed on the host (binary read format = nqueries *
 * fmgpu_words_per_query(len) words as written by fm_hostpack_reads): no conversion, 28 B per 100-bp read over PCIe */
int32_t fmgpu_search_host_packed(fmgpu_index_t *const *replicas, int32_t nreplicas, const uint32_t *h_packed,
                                 uint64_t nqueries, uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v);
/* frees the process-default pipeline the two calls above keep between calls (they serialise on one mutex) */
int32_t fmgpu_release_pipeline(void);

/* the same two calls on an explicit pipeline handle (streams, staging buffers and the self-tuning state of the host
 * feed live in it): one caller thread per handle at a time, any number of handles */
typedef struct fmgpu_pipeline fmgpu_pipeline_t;
typedef struct {
  uint64_t calls;
  int32_t  last_feed;                    /* FMGPU_FEED_* the last call used */
  int32_t  pad;
  double   last_seconds;
  uint64_t last_reads_host_packed;       /* reads the CPU packed / reads that crossed the link as ASCII in the last call */
  uint64_t last_reads_ascii_over_link;
  double   host_pack_seconds_per_read;
} fmgpu_pipeline_stats_t;
int32_t fmgpu_pipeline_create(fmgpu_pipeline_t **out);
int32_t fmgpu_pipeline_search_host(fmgpu_pipeline_t *p, fmgpu_index_t *const *replicas, int32_t nreplicas, const char *h_ascii,
                                   uint64_t nqueries, uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v);
int32_t fmgpu_pipeline_search_host_packed(fmgpu_pipeline_t *p, fmgpu_index_t *const *replicas, int32_t nreplicas, const uint32_t *h_packed,
                                          uint64_t nqueries, uint32_t len, uint32_t *h_results, const fmgpu_variant_t *v);
int32_t fmgpu_pipeline_get_stats(const fmgpu_pipeline_t *p, fmgpu_pipeline_stats_t *out);
int32_t fmgpu_pipeline_free(fmgpu_pipeline_t **p);

/* host-side ASCII -> reversed 2-bit packing (same words as the device pack kernel); OpenMP over reads,
 * AVX-512 VBMI when the CPU has it.  packed holds nqueries * fmgpu_words_per_query(len) words. */
void    fm_hostpack_reads(const char *ascii, uint64_t nqueries, uint32_t len, uint32_t *packed, int nthreads);
void    fm_hostpack_reads_scalar(const char *ascii, uint64_t nqueries, uint32_t len, uint32_t *packed);
/* stream variant used by fmgpu_search_host: the whole batch as one 2-bit sequence (base g at bits 2(g%4) of byte
 * g/4), no per-read work on the host; fmgpu_unstream_device cuts / reverses / aligns it into packed reads */
void    fm_hostpack_stream(const char *ascii, uint64_t nbases, unsigned char *out, int nthreads);
int32_t fmgpu_unstream_device(int32_t device, const uint32_t *d_stream, uint64_t nqueries, uint32_t len,
                              uint32_t *d_packed, void *stream);
void    fm_hostpack_set_streams(int streams);  /* interleaved sub-streams per packer thread (default 4, 1 = one; $FM_HOSTPACK_STREAMS) */
void    fm_hostpack_set_prefetch(int bytes);   /* software prefetch distance of the stream packer (default 8192, 0 = off) */
/* host DRAM read bandwidth over a caller's buffer, GB/s (best of iters, all threads): the ceiling of any ASCII feed */
double  fm_host_read_bandwidth(const void *buf, uint64_t bytes, int nthreads, int iters);
int     fm_hostpack_has_simd(void);
int     fm_hostpack_threads(void);

/* pinned host memory for query / result buffers */
void   *fmgpu_host_alloc(size_t bytes);
void    fmgpu_host_free(void *p);
int32_t fmgpu_host_register(void *p, size_t bytes);
int32_t fmgpu_host_unregister(void *p);

/* index construction on the GPU -------------------
 * Builds, on `device`, the image of the reference's tag-100 ".fmi" FILE
 * (header + entries, byte-identical to what genFMindex writes,
 * src/genFMindex.c:155-181,457-543) for an ASCII text or for the synthetic
 * text of fm_synth.h.  k in {1,2,3,4} (the reference builds k = 3, 4 for its CPU searchers only, makefile:226-230); any text
 * (repetitive texts take a prefix-doubling
 * pass over their tied suffixes); 2k <= n < 2^32 - 2. */
typedef struct fmgpu_build fmgpu_build_t;
int32_t  fmgpu_build_from_text(int32_t device, const char *h_ascii, uint64_t n, uint32_t steps, uint32_t chunk, fmgpu_build_t **out);
int32_t  fmgpu_build_from_synth(int32_t device, uint64_t n, uint64_t seed, uint32_t steps, uint32_t chunk, fmgpu_build_t **out);
uint64_t fmgpu_build_image_words(const fmgpu_build_t *b);
void    *fmgpu_build_image_device(const fmgpu_build_t *b);          /* device pointer */
int32_t  fmgpu_build_download(const fmgpu_build_t *b, uint32_t *h_image);
int32_t  fmgpu_build_to_index(const fmgpu_build_t *b, fmgpu_index_t **out); /* re-block, no host round trip */
/* the reference's layout transformers on the GPU, byte-identical file images: tag 101
 * (src/transformIndexBitmaps.c:269-295), 200 and 201 (src/transformIndexAlternateCounters.c:387-479) */
int32_t  fmgpu_build_transform(const fmgpu_build_t *tag100, uint32_t tag, fmgpu_build_t **out);
/* writes the image as an index file readable by the reference tools and by loadIndex (saveIndex of
 * src/genFMindex.c:155-181); bin/gfmi_b200 is the reference's generateIndex main rebuilt on these calls */
int32_t  fmgpu_build_save(const fmgpu_build_t *b, const char *path);
int32_t  fmgpu_build_free(fmgpu_build_t **b);
const char *fmgpu_build_last_error(void);
/* reads of fm_synth.h (exact substrings, uniform start) as ASCII into device memory */
int32_t  fmgpu_synth_reads_device(int32_t device, uint64_t n, uint64_t seed_ref, uint64_t nqueries, uint32_t len,
                                  uint64_t seed_reads, uint64_t first, char *d_ascii, void *stream);

/* locate (the reference stops at (L,R)) ----------------------
 * fmgpu_index_build_sa derives the full suffix array SA[0 .. bwtsize) of the indexed text from the replica's own
 * table -- so it works for index FILES, which carry no SA: SA[r] = number of 1-step LF steps from row r to the '$'
 * row, by list ranking over the LF permutation on the GPU (csrc/fm_locate.cuh).  4 bytes per row stay resident
 * (8 GB for 2 Gbp), 16 bytes per row of scratch while it is built.  AltCounters files with an active padding quirk are
 * served too (the table holds quirk-free ranks: the text's own index).  FM_E_NOT_IMPLEMENTED when memory does not suffice.
 * fmgpu_locate_device turns intervals into text positions: positions[q * max_hits + j] = start (0-based) of the
 * j-th occurrence of read q in suffix-array order, 0xFFFFFFFF beyond min(R - L, max_hits); nhits[q] = R - L (may
 * be NULL).  d_results is what the search wrote ([2q] = L, [2q+1] = R). */
int32_t fmgpu_index_build_sa(fmgpu_index_t *idx);
/* Sampled suffix array: only the rows whose text position is a multiple of `rate` (0 = 32) keep their SA value; locate
 * walks the 1-step LF mapping from an occurrence's row to the next marked row, one 64-byte fetch per step, on a table
 * made for it (0.5 bytes per row).  1.3 GB instead of 8 GB for 2 Gbp at rate 32; a located occurrence costs
 * (rate - 1) / 2 + 2 fetches on average instead of one.  Same positions as the full array.  Replaces whichever array
 * the replica holds.  fmgpu_index_sa / fmgpu_index_download_sa serve the full array only. */
int32_t fmgpu_index_build_sa_sampled(fmgpu_index_t *idx, uint32_t rate);
int32_t fmgpu_index_drop_sa(fmgpu_index_t *idx);
void   *fmgpu_index_sa(const fmgpu_index_t *idx);          /* device pointer to SA, or NULL */
int32_t fmgpu_locate_device(const fmgpu_index_t *idx, const uint32_t *d_results, uint64_t nqueries, uint32_t max_hits,
                            uint32_t *d_positions, uint32_t *d_nhits, void *stream);
/* the same for a shard (its own (L,R), its stream), positions [nqueries][max_hits] and hit counts to host memory */
int32_t fmgpu_batch_locate(const fmgpu_index_t *idx, fmgpu_batch_t *b, uint32_t max_hits, uint32_t *h_positions, uint32_t *h_nhits);
int32_t fmgpu_index_download_sa(const fmgpu_index_t *idx, uint32_t *h_sa);   /* bwtsize words */

/* one-mismatch search (the reference matches exactly only) -------------------------------
 * Every read is searched as it is and in all its 3 * len single-substitution variants, by the ordinary exact-match kernels
 * of variant `v` (variants are generated and reduced on the device, in chunks).  d_out[q] = the exact interval, how many of
 * the variants occur in the text, and their occurrences in total (saturating); d_variant_lr, when not NULL, receives all
 * variant intervals: [q][j][2] with j = 3 * t + s for packed position t (base len-1-t of the read) changed to
 * (code + 1 + s) & 3.  Synchronous (scratch is allocated and released inside).  Any read length the index serves. */
typedef struct { uint32_t L, R, variants_found, occurrences_1mm; } fmgpu_mm1_t;
int32_t fmgpu_search_device_mm1(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nqueries, uint32_t len, fmgpu_mm1_t *d_out,
                                uint32_t *d_variant_lr, const fmgpu_variant_t *v, void *stream);

/* HBM random-access roofline probe: independent uniformly random 16-byte
 * loads over a table of `table_bytes`, full occupancy.  Returns loads/s. */
int32_t fmgpu_gather_probe(int32_t device, uint64_t table_bytes, uint64_t loads_per_thread,
                           int32_t iters, double *loads_per_second);
/* same with `access_bytes` (16/32/64/128) consecutive bytes per random access */
int32_t fmgpu_gather_probe_ex(int32_t device, uint64_t table_bytes, uint32_t access_bytes, uint64_t loads_per_thread,
                              int32_t iters, double *accesses_per_second);
/* same for FMGPU_MODE_FUSED: fused-table blocks and SB96 blocks (leading steps) one search must fetch */
int32_t fmgpu_count_fetches_fused_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nqueries, uint32_t len,
                                         uint32_t *d_results, void *stream, uint64_t *nfused_blocks, uint64_t *nlead_blocks);
/* same for FMGPU_MODE_SPARSE: grid blocks (roots), SB96 blocks (leftover base steps), tree blocks below the roots */
int32_t fmgpu_count_fetches_sparse_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nqueries, uint32_t len,
                                          uint32_t *d_results, void *stream, uint64_t *nroot_blocks, uint64_t *nsb96_blocks,
                                          uint64_t *ntree_blocks);
/* same for FMGPU_MODE_WIDE: grid blocks, SB96 blocks (steps of exceptional buckets), tree blocks below the grid */
int32_t fmgpu_count_fetches_wide_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nqueries, uint32_t len,
                                        uint32_t *d_results, void *stream, uint64_t *ngrid_blocks, uint64_t *nsb96_blocks,
                                        uint64_t *ntree_blocks);
/* locality variant: the 32 lanes of every warp-level load fall inside ONE random window of
 * `window_bytes` (e.g. one 2 MB page), random blocks inside it: isolates address-translation cost */
int32_t fmgpu_gather_probe_local(int32_t device, uint64_t table_bytes, uint64_t window_bytes, uint64_t loads_per_thread,
                                 int32_t iters, double *loads_per_second);
/* fetch counter (see fmgpu_batch_count_fetches) on caller-owned device memory */
int32_t fmgpu_count_fetches_device(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nqueries, uint32_t len,
                                   uint32_t *d_results, void *stream, uint64_t *nblocks, uint64_t *nsectors);

#ifdef __cplusplus
}
#endif
#endif /* FMINDEX_B200_H_ */

/*
 * fm_internal.h -- declarations shared by the translation units behind PART 2 of include/fmindex_b200.h
 * (fm_index.cu, fm_fusedtab.cu, fm_sparsetab.cu, fm_widetab.cu, fm_locate.cu, fm_search.cu, fm_pipeline.cu, fm_probe.cu).
 * Not installed; C++ (nvcc) only.
 */
#ifndef FM_INTERNAL_H_
#define FM_INTERNAL_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/fmindex_b200.h"

#define FM_MAX_DEVICES  16
#define FM_MAX_PHANTOMS 16

/* Extra occurrences of wide symbols that an AltCounters padding-entry quirk adds to the composed rank function (see
 * fm_sparsetab.cu "quirk"): rank_F(sym, X) gains one for every entry with this symbol and row < X. */
struct fm_phantoms {
  uint32_t n;
  uint32_t sym[FM_MAX_PHANTOMS];
  uint32_t row[FM_MAX_PHANTOMS];
};

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
  uint2             *slead[16];    /* lead tables: (L,R) of all b-mers, or NULL */
  uint4             *tail1;        /* tail table (fm_tail_table_kernel), built by fmgpu_index_prepare for odd read lengths */
  uint32_t          *sa;           /* suffix array, full (fmgpu_index_build_sa) or sampled (fmgpu_index_build_sa_sampled), or NULL */
  uint32_t           sa_rate;      /* 1 = full: sa[row]; s > 1: sa[] = the SA values of the marked rows, in row order (fm_locate.cuh) */
  uint32_t          *sa_marks;     /* sampled SA: the walk table (64 bytes per 128 rows: ranks, character planes, marks) */
  uint32_t          *sa_markrank;  /* sampled SA: marked rows before each 128-row block */
  uint32_t           sa_norow, sa_nlb;   /* the row without a BWT character; number of 128-row blocks */
  int                tail_consts_ok;     /* meta.tail_* describe the text's 1-step index (k = 2), whether or not odd lengths are served */
  uint32_t           s_uni_nb, s_uni_scale;   /* sparse table is a uniform grid: blocks per symbol and the one scale (0 = directory) */
  int                tail1_tried;  /* 1 once that build was attempted (a failed allocation is not retried) */
  fm_phantoms        fphantoms;    /* quirk: extra occurrences of fused symbols */
  uint4             *wblocks;      /* wide-step table (fmgpu_index_widen), or NULL */
  uint32_t           wlead_tried;  /* bit b: building wlead[b] was attempted */
  uint2             *wlead[16];    /* its lead tables: (L,R) of all b-mers, or NULL */
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

/* error reporting (fm_index.cu): message kept per thread, returned by fmgpu_last_error() */
int32_t fm_fail(cudaError_t e, const char *what, const char *file, int line);
int32_t fm_fail_msg(int32_t code, const char *msg);
#define CU_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fm_fail(e_, #call, __FILE__, __LINE__); } while (0)

int32_t fm_use_device(int device);

extern const fmgpu_variant_t FM_DEFAULT_VARIANT;

/* derived-table memory budget (fm_index.cu): FM_SUCCESS when `bytes` more of derived tables fit the budget of this replica */
bool     fm_budget_allows(const fmgpu_index_t *idx, uint64_t bytes);
void     fm_budget_account(fmgpu_index_t *idx);

/* tail table of a 2-step replica, built if missing (synchronous; fm_index.cu).  NULL when the index has no valid tail,
 * $FMGPU_TAIL_TABLE=0, or memory is short: kernels then derive the rank from four SB96 fetches. */
const uint4 *fm_build_tail(fmgpu_index_t *idx);
cudaError_t fm_tail_table_into(const fmgpu_index_t *idx, uint4 *dst);   /* the same table into caller-owned memory (async, legacy stream) */

/* the same for any SB96-shaped table of nsym symbols (e.g. the 4-symbol tail table) */
cudaError_t fm_table_symbols(const uint4 *table, uint32_t nblocks, uint32_t nsym, uint64_t nrows_alloc, uint8_t *d_sym);
/* k-step symbol of every row of the SB96 table into sym[nrows_alloc] (FM_SYM_NONE for '$' rows and rows past the end) (fm_fusedtab.cu) */
cudaError_t fm_row_symbols(const fmgpu_index_t *idx, uint64_t nrows_alloc, uint8_t *d_sym);

/* launches (asynchronous on `stream`; they only pick among tables that exist, nothing is built here) */
int32_t fm_launch_search(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len, uint32_t *d_results,
                         const fmgpu_variant_t *v, cudaStream_t stream, unsigned long long *d_counters);
int32_t fm_launch_fused(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len, uint32_t *d_results,
                        fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters);
int32_t fm_launch_sparse(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len, uint32_t *d_results,
                         fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters, bool use_lead_tables);
/* lead tables a sparse search of `len`-base reads would use, built if missing (synchronous; fm_sparsetab.cu) */
void    fm_sparse_prepare(fmgpu_index_t *idx, uint32_t len);
/* wide-step table (fm_widetab.cu): launch (FM_E_NOT_IMPLEMENTED when the table's width does not serve `len`), lead table of `len` */
int32_t fm_launch_wide(const fmgpu_index_t *idx, const uint32_t *d_packed, uint64_t nq, uint32_t len, uint32_t *d_results,
                       fmgpu_variant_t v, cudaStream_t stream, unsigned long long *d_counters);
void    fm_wide_prepare(fmgpu_index_t *idx, uint32_t len);

#endif /* FM_INTERNAL_H_ */

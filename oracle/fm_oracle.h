/*
 * oracle/fm_oracle.h -- TEST INFRASTRUCTURE ONLY.  See fm_oracle.c.
 */
#ifndef FM_ORACLE_H_
#define FM_ORACLE_H_

#include <stdint.h>

#define FMO_MAX_STEPS 4

typedef struct {
  uint32_t tag;         /* 100 | 101 | 200 | 201                               */
  uint32_t steps;       /* k                                                   */
  uint32_t bwtsize;     /* n + 1                                               */
  uint32_t ncounters;   /* 4^k (100/101) or 4^k/2 (200/201)                    */
  uint32_t nentries;
  uint32_t chunk;       /* d                                                   */
  uint32_t dollarPositionBWT[FMO_MAX_STEPS];
  uint32_t dollarBaseBWT[FMO_MAX_STEPS];
  uint32_t entry_words; /* 2*(d/32)*k + ncounters                              */
  uint32_t *entries;    /* nentries * entry_words, exactly as in the file      */
  int      owns_entries;
} fmo_index_t;

int32_t  fmo_load_index(const char *fn, fmo_index_t **out);
int32_t  fmo_wrap_image(const uint32_t *image, uint64_t nwords, fmo_index_t **out);
void     fmo_free_index(fmo_index_t *idx);

uint32_t fmo_base_code(uint32_t ascii);
uint32_t fmo_plane_word(const fmo_index_t *idx, uint32_t entry, uint32_t step, uint32_t bit, uint32_t n);
uint32_t fmo_counter(const fmo_index_t *idx, uint32_t entry, uint32_t slot);
uint32_t fmo_lf(const fmo_index_t *idx, uint32_t sigma, uint32_t X);
void     fmo_search(const fmo_index_t *idx, const char *queries, uint64_t num, uint32_t len, uint32_t *results);
uint64_t fmo_count_sectors(const fmo_index_t *idx, const char *queries, uint64_t num, uint32_t len,
                           uint32_t block_positions, uint32_t blocks_per_sector);
int32_t  fmo_load_queries(const char *fn, uint32_t len, uint64_t num, char *out);
int32_t  fmo_write_results(const char *fn, const uint32_t *results, uint32_t num);

#endif

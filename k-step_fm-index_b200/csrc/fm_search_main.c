/*
 * fmIndexSearchGPU_b200 <index> <queries.fa> <read length> <number of reads>
 *
 * The benchmark driver of the reference (common/searchQueries.c:34-132, CUDA
 * flow) rebuilt on libfmindex_b200: same argv, same call sequence
 *   loadIndex -> loadQueries -> initResults -> transferCPUtoGPU ->
 *   5 x searchIndexGPU (timed) -> transferGPUtoCPU -> saveResults -> free*,
 * same "TIME:" line (seconds per iteration) and the same "<index>.res.gpu"
 * output.  Any of the four index layouts is accepted (the flavour is read
 * from the file header); GPUs are chosen with $FMGPU_DEVICES ("0,1,2,3").
 */
#include <stdio.h>
#include <stdlib.h>
#include "../../include/fmindex_b200.h"

#define HOST_HANDLE_ERROR(error) { if (error) { fprintf(stderr, "%s\n", errorCommon(error)); exit(EXIT_FAILURE); } }

int main(int argc, char *argv[])
{
  void *index = NULL, *queries = NULL, *results = NULL;
  uint32_t qrysize, numqueries, iter = 5, n;
  double ts, ts1;
  int32_t error;

  if (argc < 5) {
    fprintf(stderr, "usage: %s <index> <queries.fa> <read length> <number of reads>\n", argv[0]);
    return EXIT_FAILURE;
  }
  qrysize = (uint32_t) atoll(argv[3]);
  numqueries = (uint32_t) atoll(argv[4]);

  error = loadIndex(argv[1], &index);                       HOST_HANDLE_ERROR(error);
  error = loadQueries(argv[2], qrysize, numqueries, &queries); HOST_HANDLE_ERROR(error);
  error = initResults(numqueries, &results);                HOST_HANDLE_ERROR(error);
  error = transferCPUtoGPU(index, queries, results);        HOST_HANDLE_ERROR(error);

  ts = sampleTime();
  for (n = 0; n < iter; n++) searchIndexGPU(index, queries, results);
  ts1 = sampleTime();

  error = transferGPUtoCPU(results);                        HOST_HANDLE_ERROR(error);
  error = saveResults(argv[1], results, index);             HOST_HANDLE_ERROR(error);
  error = freeIndexGPU(&index);                             HOST_HANDLE_ERROR(error);
  error = freeQueriesGPU(&queries);                         HOST_HANDLE_ERROR(error);
  error = freeResultsGPU(&results);                         HOST_HANDLE_ERROR(error);

  printf("TIME: \t %f \n", (ts1 - ts) / iter);

  error = freeIndex(&index);                                HOST_HANDLE_ERROR(error);
  error = freeQueries(&queries);                            HOST_HANDLE_ERROR(error);
  error = freeResults(&results);                            HOST_HANDLE_ERROR(error);
  return FM_SUCCESS;
}

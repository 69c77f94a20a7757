/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's D-SOFT seeding; see dsoft_oracle.c. */
#ifndef DSOFT_ORACLE_H
#define DSOFT_ORACLE_H
#include <stdint.h>
#include "../include/darwin_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct DsoftIndex {
    int k, w, max_stride;
    uint32_t bin_size, kmer_max_occurence;
    uint64_t n_buckets, n_positions;
    uint32_t* buckets;        /* n_buckets + 1 prefix sums (seedBuckets) */
    uint32_t* positions;      /* seedPositions */
} DsoftIndex;

uint32_t dsoft_hash32(uint32_t key, int k);
uint64_t dsoft_minimizers(const char* seq, uint32_t len, int k, int w, uint64_t* out);
int  dsoft_index_build(DsoftIndex* ix, const char* dram, const uint32_t* chr_start, const uint32_t* chr_len_unpadded, int n_chr,
                       uint32_t ref_length, int k, int w, uint32_t seed_occurence_multiple, uint32_t bin_size, int max_stride);
void dsoft_index_free(DsoftIndex* ix);
int  dsoft_query(const DsoftIndex* ix, const char* query, uint32_t query_length, int N, int threshold, int overlap,
                 DarwinSeedAnchor* anchors, int anchors_cap, uint64_t* pool, uint64_t pool_cap, uint64_t* pool_used);

#ifdef __cplusplus
}
#endif
#endif

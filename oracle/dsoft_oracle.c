/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * CPU restatement of the reference's D-SOFT seeding: minimizer iteration
 * (software/seed_pos_table.h:280-372 iterate_minimizers_qw, ntcoding.h:35-67), the seed position table
 * (software/seed_pos_table.cpp:41-160) and SeedPosTable::DSOFT (software/seed_pos_table.cpp:252-553), in plain C.
 * Pinned against the compiled reference's own seeder_body in tests/test_oracle_dsoft.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "dsoft_oracle.h"

/* _q_to_2_bit (seed_pos_table.h:63-84): a 16-entry table indexed by the LOW NIBBLE of the character:
 * 'A'/'a' (1) -> 0, 'C'/'c' (3) -> 1, 'G'/'g' (7) -> 2, 'T'/'t' (4) -> 3, everything else ('N' = 14) -> 0. */
static inline uint32_t nt2(char c) {
    switch (c & 0x0f) { case 3: return 1; case 7: return 2; case 4: return 3; default: return 0; }
}

uint32_t dsoft_hash32(uint32_t key, int k) {                     /* ntcoding.h:56-67 */
    uint32_t m = (1u << 2 * k) - 1;
    key = (~key + (key << 21)) & m;
    key = key ^ (key >> 24);
    key = ((key + (key << 3)) + (key << 8)) & m;
    key = key ^ (key >> 14);
    key = ((key + (key << 2)) + (key << 4)) & m;
    key = key ^ (key >> 28);
    key = (key + (key << 31)) & m;
    return key;
}

/* iterate_minimizers_qw (seed_pos_table.h:280-372; also what the w = 5 / 9 specialisations compute): positions
 * p in [0, round16(len) - k); the seed at p packs characters p .. p+k-1 (character p in the lowest two bits; characters at
 * or beyond `len` are whatever follows in memory -- the reader pads with 'N', which codes as 0); window minimum over the
 * last w hashes (Min_Window starts from 2^31 - 1); a minimizer is emitted at p >= w-1 when it differs from the last
 * emitted one or w positions have passed (last_m = last_p = 0 initially).  out[i] = (p << 32) | m. */
uint64_t dsoft_minimizers(const char* seq, uint32_t len, int k, int w, uint64_t* out) {
    uint32_t centinel = (~0x0fu & (len + 15u)) - (uint32_t)k;
    uint32_t mask = (1u << 2 * k) - 1;
    uint32_t* win = (uint32_t*)calloc((size_t)w, sizeof(uint32_t));
    uint64_t n = 0, last_m = 0, last_p = 0;
    /* the first batch always covers p = 0..15 (seed_pos_table.h:300-318), the vector loop then stops at the centinel */
    uint32_t end = centinel < 16 ? 16 : centinel;
    for (uint32_t p = 0; p < end; p++) {
        uint32_t seed = 0;
        for (int c = 0; c < k; c++) seed |= nt2(seq[p + c]) << (2 * c);
        win[p % w] = dsoft_hash32(seed & mask, k);
        if (p < (uint32_t)w - 1) continue;
        uint32_t m = (1u << 31) - 1;
        for (int c = 0; c < w; c++) if (win[c] < m) m = win[c];
        if (m != last_m || p - last_p >= (uint64_t)w) { out[n++] = ((uint64_t)p << 32) | m; last_m = m; last_p = p; }
    }
    free(win);
    return n;
}

/* Seed position table (seed_pos_table.cpp:41-160): buckets = prefix sums of the minimizer histogram, positions grouped by
 * hashed minimizer; a bucket is sorted ascending when it is non-empty and no larger than kmer_max_occurence. */
static int cmp_u32(const void* a, const void* b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

int dsoft_index_build(DsoftIndex* ix, const char* dram, const uint32_t* chr_start, const uint32_t* chr_len_unpadded, int n_chr,
                      uint32_t ref_length, int k, int w, uint32_t seed_occurence_multiple, uint32_t bin_size, int max_stride) {
    memset(ix, 0, sizeof(*ix));
    ix->k = k; ix->w = w; ix->bin_size = bin_size; ix->max_stride = max_stride;
    ix->kmer_max_occurence = seed_occurence_multiple * (1 + (ref_length >> (2 * k)));            /* :57 */
    ix->n_buckets = 1ull << (2 * k);
    ix->buckets = (uint32_t*)calloc(ix->n_buckets + 1, sizeof(uint32_t));
    if (!ix->buckets) return -1;
    uint64_t** lists = (uint64_t**)calloc((size_t)n_chr, sizeof(uint64_t*));
    uint64_t* counts = (uint64_t*)calloc((size_t)n_chr, sizeof(uint64_t));
    uint64_t total = 0;
    for (int c = 0; c < n_chr; c++) {                                                             /* main.cpp:323-341 */
        lists[c] = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)chr_len_unpadded[c] + 32));
        counts[c] = dsoft_minimizers(dram + chr_start[c], chr_len_unpadded[c], k, w, lists[c]);
        for (uint64_t i = 0; i < counts[c]; i++) ix->buckets[(uint32_t)lists[c][i] + 1]++;
        total += counts[c];
    }
    for (uint64_t b = 0; b < ix->n_buckets; b++) ix->buckets[b + 1] += ix->buckets[b];
    ix->n_positions = total;
    ix->positions = (uint32_t*)malloc(sizeof(uint32_t) * (total ? total : 1));
    uint32_t* cursor = (uint32_t*)malloc(sizeof(uint32_t) * (ix->n_buckets));
    memcpy(cursor, ix->buckets, sizeof(uint32_t) * ix->n_buckets);
    for (int c = 0; c < n_chr; c++) {
        for (uint64_t i = 0; i < counts[c]; i++) {
            uint32_t m = (uint32_t)lists[c][i], p = (uint32_t)(lists[c][i] >> 32);
            ix->positions[cursor[m]++] = p + chr_start[c];                                        /* main.cpp:336-338 */
        }
        free(lists[c]);
    }
    for (uint64_t b = 0; b < ix->n_buckets; b++) {
        uint32_t lo = ix->buckets[b], hi = ix->buckets[b + 1];
        if (lo < hi && hi - lo <= ix->kmer_max_occurence) qsort(ix->positions + lo, hi - lo, sizeof(uint32_t), cmp_u32);   /* :146-150 */
    }
    free(cursor); free(lists); free(counts);
    return 0;
}

void dsoft_index_free(DsoftIndex* ix) { free(ix->buckets); free(ix->positions); memset(ix, 0, sizeof(*ix)); }

typedef struct { uint64_t bin_offset; uint32_t hit; uint32_t order; } Hit;
static int cmp_hit(const void* a_, const void* b_) {            /* std::stable_sort by bin_offset (seed_pos_table.h:42-45) */
    const Hit* a = (const Hit*)a_; const Hit* b = (const Hit*)b_;
    if (a->bin_offset != b->bin_offset) return a->bin_offset < b->bin_offset ? -1 : 1;
    return a->order < b->order ? -1 : (a->order > b->order);
}
static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
typedef struct { uint64_t hit_offset; int num; uint64_t loff, ln, roff, rn; } Anc;
static int cmp_anc(const void* a_, const void* b_) {            /* seed_pos_table.cpp:506-510 */
    const Anc* a = (const Anc*)a_; const Anc* b = (const Anc*)b_;
    if (a->num != b->num) return a->num > b->num ? -1 : 1;
    return a->hit_offset < b->hit_offset ? -1 : (a->hit_offset > b->hit_offset);
}

/* SeedPosTable::DSOFT (seed_pos_table.cpp:252-553) for one query.  Anchors come back in the reference's output order;
 * chained hits go to pool (left ascending, right descending).  Returns the number of anchors or < 0 (capacity). */
int dsoft_query(const DsoftIndex* ix, const char* query, uint32_t query_length, int N, int threshold, int overlap,
                DarwinSeedAnchor* anchors, int anchors_cap, uint64_t* pool, uint64_t pool_cap, uint64_t* pool_used) {
    uint64_t* mins = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)query_length + 32));
    uint64_t min_count = dsoft_minimizers(query, query_length, ix->k, ix->w, mins);
    size_t hcap = 1024, nh = 0;
    Hit* hits = (Hit*)malloc(sizeof(Hit) * hcap);
    int stride = 1;
    for (int64_t i = 0; i < (int64_t)min_count; i += stride) {                                    /* :307-336 */
        uint32_t offset = (uint32_t)(mins[i] >> 32), index = (uint32_t)mins[i];
        uint32_t s = ix->buckets[index], e = ix->buckets[index + 1];
        if (e - s <= ix->kmer_max_occurence) {
            for (uint32_t j = s; j < e; j++) {
                uint32_t hit = ix->positions[j];
                if (hit >= offset) {
                    if (nh == hcap) { hcap *= 2; hits = (Hit*)realloc(hits, sizeof(Hit) * hcap); }
                    uint32_t bin = (hit - offset) / ix->bin_size;
                    hits[nh].bin_offset = ((uint64_t)bin << 32) + offset; hits[nh].hit = hit; hits[nh].order = (uint32_t)nh; nh++;
                }
            }
        }
        if (i > N) { if (overlap == 0) stride = ix->max_stride; else break; }
    }
    qsort(hits, nh, sizeof(Hit), cmp_hit);
    /* candidate bins (:352-392); num_candidates is never incremented in the reference, so max_candidates never bites */
    size_t acap = 64, na = 0;
    Anc* anc = (Anc*)malloc(sizeof(Anc) * acap);
    uint32_t* cbin = (uint32_t*)malloc(sizeof(uint32_t) * acap);
    uint32_t last_bin = 1u << 31, last_offset = 0, curr_count = 0, ks = (uint32_t)ix->k;
    for (size_t i = 0; i < nh; i++) {
        uint32_t offset = (uint32_t)hits[i].bin_offset, bin = (uint32_t)(hits[i].bin_offset >> 32), hit = hits[i].hit;
        int push = 0;
        if (bin == last_bin) {
            if (curr_count < (uint32_t)threshold) {
                curr_count = ((offset - last_offset > ks) || (curr_count == 0)) ? curr_count + ks : curr_count + (offset - last_offset);
                if (curr_count >= (uint32_t)threshold) push = 1;
            }
        } else {
            last_bin = bin; curr_count = ks;
            if (curr_count >= (uint32_t)threshold) push = 1;
        }
        if (push) {
            if (na == acap) { acap *= 2; anc = (Anc*)realloc(anc, sizeof(Anc) * acap); cbin = (uint32_t*)realloc(cbin, sizeof(uint32_t) * acap); }
            memset(&anc[na], 0, sizeof(Anc)); anc[na].hit_offset = ((uint64_t)hit << 32) + offset; cbin[na] = bin; na++;
        }
        last_offset = offset;
    }
    /* chained hits inside the SV window of every candidate bin (:394-497) */
    uint32_t sv = (overlap == 0) ? (1u << 12) / ix->bin_size : 1;
    size_t start_idx = 0;
    uint64_t used = 0;
    uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (2 * nh + 2));
    int rc = 0;
    for (size_t k = 0; k < na && rc == 0; k++) {
        uint32_t cb = cbin[k];
        uint64_t* L = tmp; uint64_t* R = tmp + nh + 1; size_t nl = 0, nr = 0;
        int start_assigned = 0;
        for (size_t i = start_idx; i < nh; i++) {
            uint32_t bin = (uint32_t)(hits[i].bin_offset >> 32);
            if ((bin + sv >= cb) && (bin < cb + sv)) {
                if (!start_assigned) { start_assigned = 1; start_idx = i; }
                uint64_t ho = ((uint64_t)hits[i].hit << 32) + (uint32_t)hits[i].bin_offset;
                if (ho <= anc[k].hit_offset) L[nl++] = ho;
                if (ho >= anc[k].hit_offset) R[nr++] = ho;
            } else if (bin >= cb + sv) break;
        }
        qsort(L, nl, sizeof(uint64_t), cmp_u64); qsort(R, nr, sizeof(uint64_t), cmp_u64);
        /* collinear left chain, walked from the anchor outwards (:437-463); kept ascending */
        if (used + nl + nr > pool_cap) { rc = -2; break; }
        uint64_t* outL = pool + used; size_t kl = 0;
        uint64_t cur = L[nl - 1]; outL[kl++] = cur;
        for (size_t h = nl - 1; h-- > 0;) {
            uint32_t h1 = (uint32_t)(cur >> 32), o1 = (uint32_t)cur, h2 = (uint32_t)(L[h] >> 32), o2 = (uint32_t)L[h];
            if (h1 >= h2 && o1 >= o2) { outL[kl++] = L[h]; cur = L[h]; }
        }
        qsort(outL, kl, sizeof(uint64_t), cmp_u64);
        anc[k].loff = used; anc[k].ln = kl; used += kl;
        /* collinear right chain (:465-489); reversed -> descending */
        uint64_t* outR = pool + used; size_t kr = 0;
        cur = R[0]; outR[kr++] = cur;
        for (size_t h = 1; h < nr; h++) {
            uint32_t h1 = (uint32_t)(cur >> 32), o1 = (uint32_t)cur, h2 = (uint32_t)(R[h] >> 32), o2 = (uint32_t)R[h];
            if (h1 <= h2 && o1 <= o2) { outR[kr++] = R[h]; cur = R[h]; }
        }
        for (size_t a = 0, b = kr; a + 1 < b; ) { b--; uint64_t t = outR[a]; outR[a] = outR[b]; outR[b] = t; a++; }
        anc[k].roff = used; anc[k].rn = kr; used += kr;
        anc[k].num = (int)(kl + kr);
    }
    if (rc == 0) {
        qsort(anc, na, sizeof(Anc), cmp_anc);                                                     /* :506-510 */
        if ((int)na > anchors_cap) rc = -3;
        else for (size_t k = 0; k < na; k++) {
            anchors[k].hit_offset = anc[k].hit_offset;
            anchors[k].left_off = anc[k].loff; anchors[k].left_n = (uint32_t)anc[k].ln;
            anchors[k].right_off = anc[k].roff; anchors[k].right_n = (uint32_t)anc[k].rn;
        }
    }
    if (pool_used) *pool_used = used;
    free(tmp); free(cbin); free(anc); free(hits); free(mins);
    return rc ? rc : (int)na;
}

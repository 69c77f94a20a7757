/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * CPU restatement ("port") of the GACT alignment-extension path of
 * yatisht/darwin, in plain C.  It exists to check the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_ref.py)
 * against the reference's own translation units compiled unmodified into
 * oracle/_ref/libdarwin_ref{,_patched}.so, and against the committed fixtures
 * under tests/golden/ generated from that library (tests/golden/make_golden.py).
 * The reference has no golden vectors of its own for this path; the
 * RTL known-answer scores (RTL/GACT/test_data/test_align.txt) are reproduced at
 * score level in tests/test_golden.py.
 */
#ifndef GACT_ORACLE_H
#define GACT_ORACLE_H

#include <stdint.h>
#include "../include/darwin_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct GactScoring {
    int sub[25];               /* row = reference nt, col = query nt, 4 = N (Processor.cpp:50-74) */
    int go, ge, lgo, lge;      /* Processor.cpp:75-78 */
    int tri[11];               /* cfg.gact_sub_mat (main.cpp:183-197) for AlignmentScore */
} GactScoring;

/* tile rules */
#define GACT_RULE_STRIPED 0    /* literal scalar emulation of the AVX2 striped kernel + lazy-F (== patched reference) */
#define GACT_RULE_STREAM  1    /* streaming closed form -- what the exact CUDA kernel computes */
#define GACT_RULE_CLEAN   2    /* textbook rule; sets GACT_TILE_LFLAG when it may differ */

/* flags returned per tile */
#define GACT_TILE_LFLAG     1u  /* traceback met a long-gap candidate while in DIAG state (clean rule only) */
#define GACT_TILE_LONG_INS  2u  /* traceback entered the long-insertion state (reference UB bits) */

void gact_scoring_init(GactScoring* sc, const DarwinScoring* s);
int  gact_nt2int(char nt, int complement);

/* One tile == one request of BatchAlignmentSIMD (Processor.cpp:718-762).
 * tb_words: 32 ops per word (Processor.cpp:568-582); ops (optional): 1 byte per op in emission order. */
int gact_tile(const GactScoring* sc, const char* dram, const DarwinTileReq* req, int do_traceback, int rule,
              DarwinTileRes* res, uint64_t* tb_words, int tb_words_cap, uint8_t* ops, int ops_cap, uint32_t* flags);

int gact_tiles(const GactScoring* sc, const char* dram, int do_traceback, int rule, const DarwinTileReq* req, int n,
               DarwinTileRes* res, uint64_t* tb_words, int tb_words_per_req, uint32_t* flags);

/* extender_body for one anchor (extender.cpp:9-1065 + makeForward/BackwardAlignment :1067-1159). */
int gact_extend(const GactScoring* sc, const char* dram, const DarwinExtendParams* p, int rule,
                const DarwinAnchor* anchors, int n, const uint64_t* hit_pool,
                DarwinAlnRes* res, uint8_t* ops_pool, uint64_t ops_pool_bytes);

/* debugging aid: log every tile request gact_extend issues (cap entries kept, count keeps running) */
void gact_set_tile_log(DarwinTileReq* buf, int cap);
int  gact_tile_log_count(void);

/* AlignmentScore (extender.cpp:1161-1200) evaluated on an op string + the sequences it was cut from. */
int gact_alignment_score(const GactScoring* sc, const char* ref_str, const char* query_str, uint64_t n);

/* Rebuild the reference's gapped strings from an alignment result (what printer.cpp consumes). */
int gact_build_strings(const char* dram, const DarwinAnchor* a, const DarwinAlnRes* r, const uint8_t* ops,
                       char* ref_str, char* query_str);

/* The tile part of filter_body::operator() (filter.cpp:28-122, :131-223) for n candidates. */
int gact_filter(const GactScoring* sc, const char* dram, const DarwinFilterParams* p, const DarwinFilterCand* cands, int n,
                DarwinFilterRes* res);

/* filter_body::slopeFilter (filter.cpp:227-289) for the locations of one strand; returns the number kept and their
 * indices in output order. */
int gact_slope_filter(const int* read_num, const int* score, const uint32_t* reference_pos, const uint32_t* query_pos, int n,
                      float slope_threshold, int* order_out);

#ifdef __cplusplus
}
#endif
#endif

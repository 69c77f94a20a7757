/*
 * darwin_gpu.h -- C-ABI of the B200-native GACT alignment-extension library
 * (libdarwin_gact.so).  Plain pointers and sizes only; no C++/torch types.
 *
 * This is the boundary a maintainer of yatisht/darwin binds instead of the
 * software "Processor" (the reference's FPGA/DLL seam):
 *
 *   reference seam (software/Processor.h:50-63)        replacement entry point
 *   -------------------------------------------        -----------------------
 *   InitializeProcessor_ptr           (:50)            darwin_gpu_create / _destroy
 *   InitializeScoringParameters_ptr   (:51)            darwin_gpu_set_scoring
 *   InitializeReferenceMemory_ptr     (:52)            darwin_gpu_upload
 *   InitializeReadMemory_ptr          (:53)            darwin_gpu_upload
 *   BatchAlignmentSIMD_ptr            (:55)            darwin_gpu_tiles
 *   filter_body::operator()  (software/filter.cpp:8-225, graph.h:205-217)
 *                                                      darwin_gpu_filter
 *   extender_body::operator()  (software/extender.cpp:9, graph.h:219-229)
 *                                                      darwin_gpu_extend
 *
 * All functions return DARWIN_OK (0) or a negative DarwinStatus; they never
 * fall back to a CPU implementation -- without a usable CUDA device they fail
 * with DARWIN_ERR_NO_DEVICE.  A handle is single-threaded; use one handle per
 * host thread / GPU (the reference's `token`, main.cpp:615-624).
 */
#ifndef DARWIN_GPU_H
#define DARWIN_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum DarwinStatus {
    DARWIN_OK              =  0,
    DARWIN_ERR_NO_DEVICE   = -1,  /* no CUDA device / driver */
    DARWIN_ERR_INVALID     = -2,  /* bad argument (== Darwin::Status::InvalidData, Darwin.bond:36-40) */
    DARWIN_ERR_CUDA        = -3,  /* a CUDA call failed; see darwin_gpu_last_error */
    DARWIN_ERR_CAPACITY    = -4,  /* caller-provided output buffer too small */
    DARWIN_ERR_NOT_READY   = -5,  /* scoring / arena not initialised */
    DARWIN_ERR_NOMEM       = -6   /* host allocation failed inside the library (no C++ exception crosses this boundary) */
} DarwinStatus;

/* align_fields bits (software/graph.h:22-26, Darwin.bond:97) */
#define DARWIN_REVERSE_REF      (1u << 4)
#define DARWIN_COMPLEMENT_REF   (1u << 3)
#define DARWIN_REVERSE_QUERY    (1u << 2)
#define DARWIN_COMPLEMENT_QUERY (1u << 1)
#define DARWIN_START_END        (1u << 0)

/* traceback op codes inside TB words / op strings (software/Processor.h:14:
 * enum states {Z,I,D,M,...}; the long-gap states are stored `% 4`,
 * Processor.cpp:570, so only I, D, M ever appear). */
#define DARWIN_OP_I 1u   /* consumes query only  ('-' in the reference string) */
#define DARWIN_OP_D 2u   /* consumes reference only ('-' in the query string)  */
#define DARWIN_OP_M 3u   /* consumes both */

/* Largest tile edge the reference ever requests (extender.cpp:70-75). */
#define DARWIN_MAX_TILE 1984

/* Scoring, in Darwin.bond:45-65 field order (AlignmentScoringParams). */
typedef struct DarwinScoring {
    int32_t sub_AA, sub_AC, sub_AG, sub_AT;
    int32_t sub_CC, sub_CG, sub_CT;
    int32_t sub_GG, sub_GT;
    int32_t sub_TT;
    int32_t sub_N;
    int32_t gap_open, gap_extend;
    int32_t long_gap_open, long_gap_extend;
} DarwinScoring;

/* One tile request == AlignmentInputFieldsDRAM (Darwin.bond:95-112). */
typedef struct DarwinTileReq {
    uint64_t ref_bases_start_addr;    /* arena offset of the first reference base of the tile */
    uint64_t query_bases_start_addr;  /* arena offset of the first query base of the tile */
    uint32_t score_threshold;         /* carried, unused (as in the reference) */
    uint16_t index;
    uint16_t ref_size;
    uint16_t query_size;
    uint16_t max_tb_steps;
    uint8_t  align_fields;
    uint8_t  reserved[3];
} DarwinTileReq;

/* One tile result == AlignmentResult (Darwin.bond:114-129) without the vector;
 * the TB_pointers words go to tb_words[i * tb_words_per_req ...]. */
typedef struct DarwinTileRes {
    int32_t  score;
    uint16_t ref_offset;
    uint16_t query_offset;
    uint16_t ref_max_pos;
    uint16_t query_max_pos;
    uint16_t total_TB_pointers;
    uint8_t  index;                   /* uint8 in the reference too (BatchSize) */
    uint8_t  status;                  /* low nibble: 0 = OK, 1 = tile larger than DARWIN_MAX_TILE, 2 = tb_words_per_req too small;
                                       * bit 4: DARWIN_TILE_LONG_INS_PATH */
} DarwinTileRes;

/* DarwinTileRes.status bit: the traceback of this tile entered the long-insertion state.  The reference reads
 * uninitialised vectors there (software/Processor.cpp:259-260, used :405-408, :444), so its answer for such a tile
 * depends on the build; tiles WITHOUT this bit are identical under every build of the reference. */
#define DARWIN_TILE_LONG_INS_PATH 0x10u

/* One anchor == ExtendLocations (software/graph.h:83-91) plus what
 * makeForwardAlignment / makeBackwardAlignment look up (extender.cpp:1067-1159). */
typedef struct DarwinAnchor {
    uint64_t read_addr;        /* arena offset of the FORWARD read (query_start_addr, both strands) */
    uint32_t reference_pos;    /* absolute arena offset (ExtendLocations::reference_pos) */
    uint32_t query_pos;        /* strand-local read offset (ExtendLocations::query_pos) */
    uint32_t chr_start;        /* Index::chr_coord[chr_id] */
    uint32_t ref_len;          /* Index::chr_len[chr_id] -- the PADDED length (main.cpp:453) */
    uint32_t read_len;
    int32_t  read_num;
    int32_t  chr_id;
    int32_t  score;            /* first-tile score, carried through */
    uint32_t left_hits_off;    /* left_hit_offsets  = hit_pool[left_hits_off  .. +left_hits_n)  ascending  */
    uint32_t left_hits_n;
    uint32_t right_hits_off;   /* right_hit_offsets = hit_pool[right_hits_off .. +right_hits_n) descending */
    uint32_t right_hits_n;
    uint8_t  strand;           /* 0 = '+', 1 = '-' */
    uint8_t  reserved[7];
} DarwinAnchor;

/* DarwinAlnRes.flags */
#define DARWIN_ALN_EMITTED       (1u << 0)  /* the reference would push this ExtendAlignments */
#define DARWIN_ALN_OPS_OVERFLOW  (1u << 1)  /* op string did not fit ops_cap; coordinates still valid */
#define DARWIN_ALN_EXACT_RERUN   (1u << 2)  /* >=1 tile was recomputed with the exact (lazy-F faithful) rule */
#define DARWIN_ALN_LONG_INS_PATH (1u << 3)  /* a traceback entered the long-insertion state (reference UB bits) */

/* One extended alignment == the fields of ExtendAlignments (graph.h:97-121)
 * that survive to the printer; gapped strings are rebuilt from the op string. */
typedef struct DarwinAlnRes {
    uint64_t ops_offset;        /* first op of this alignment inside ops_pool (1 byte per op, left-to-right) */
    uint64_t cells;             /* sum of ref_size*query_size over the tile requests issued */
    uint32_t n_ops;
    uint32_t reference_start_offset;
    uint32_t reference_end_offset;
    uint32_t query_start_offset;
    uint32_t query_end_offset;
    uint32_t n_left_ops;        /* ops [0, n_left_ops) came from the left extension */
    uint32_t n_tiles;           /* tile requests issued (num_active_tiles share, extender.cpp:233) */
    uint32_t n_large_tiles;     /* 1984x960 / 960x1984 requests (num_large_tiles, extender.cpp:77) */
    int32_t  score;             /* AlignmentScore (extender.cpp:1161-1200) of the final strings */
    uint32_t flags;
} DarwinAlnRes;

typedef struct DarwinExtendParams {
    int32_t tile_size;      /* params.cfg [GACT_extend] tile_size    */
    int32_t tile_overlap;   /* params.cfg [GACT_extend] tile_overlap */
    int32_t do_overlap;     /* argv[3] of the reference (0 = reference-guided, 1 = de novo overlap) */
    int32_t reserved;
} DarwinExtendParams;

/* One first-tile candidate == one element of seeder_data::fwAnchors / rcAnchors (software/graph.h:131-138,
 * Anchors::hit_offset, software/seed_pos_table.h:30-40) plus what filter_body looks up for it (filter.cpp:44-56). */
typedef struct DarwinFilterCand {
    uint64_t read_addr;        /* arena offset of the FORWARD read (read.seq.data() - g_DRAM->buffer) */
    uint32_t hit;              /* hit_offset >> 32: absolute arena offset of the seed hit on the reference */
    uint32_t offset;           /* hit_offset & 0xffffffff: strand-local read offset of the seed hit */
    uint32_t chr_start;        /* Index::chr_coord[chr_id] */
    uint32_t chr_len;          /* Index::chr_len[chr_id] (padded) */
    uint32_t read_len;
    uint8_t  strand;           /* 0 = fwAnchors, 1 = rcAnchors (reverse_query + complement_query, filter.cpp:181) */
    uint8_t  reserved[3];
} DarwinFilterCand;

#define DARWIN_FILTER_SCORE_OK   (1u << 0)  /* score >= first_tile_score_threshold (filter.cpp:87) */
#define DARWIN_FILTER_OVERLAP_OK (1u << 1)  /* offset + (chr_end - hit) > min_overlap / 2 (filter.cpp:102-104) */

/* What filter_body derives from one candidate's tile (filter.cpp:87-116): the ExtendLocations fields. */
typedef struct DarwinFilterRes {
    int32_t  score;            /* AlignmentResult.score of the first tile */
    uint32_t reference_pos;    /* ref_tile_start + ref_max_pos   (absolute arena offset) */
    uint32_t query_pos;        /* query_tile_start + query_max_pos (strand-local) */
    uint32_t flags;            /* DARWIN_FILTER_* */
} DarwinFilterRes;

typedef struct DarwinFilterParams {
    int32_t first_tile_size;              /* params.cfg [GACT_first_tile] first_tile_size */
    int32_t first_tile_score_threshold;   /* first_tile_score_threshold */
    int32_t min_overlap;                  /* min_overlap */
    int32_t reserved;
} DarwinFilterParams;

/* D-SOFT parameters: params.cfg [DSOFT_params] (software/params.cfg:18-27) as SeedPosTable's constructor and
 * seeder_body pass them on (seed_pos_table.cpp:41-57, seeder.cpp:27-33). */
typedef struct DarwinSeedParams {
    int32_t seed_size;                /* k, 4..15 */
    int32_t minimizer_window;         /* w */
    int32_t bin_size;
    int32_t threshold;                /* dsoft_threshold: bases of one bin covered by seed hits */
    int32_t num_seeds;                /* N: all minimizers up to index N+1, every max_stride-th afterwards */
    int32_t seed_occurence_multiple;
    int32_t max_stride;
    int32_t do_overlap;               /* argv[3]: 1 = stop after N+1 seeds, SV window of one bin */
} DarwinSeedParams;

/* One chromosome for the index build (main.cpp:418-466): arena offset and UNPADDED length. */
typedef struct DarwinChrom { uint32_t start; uint32_t len_unpadded; } DarwinChrom;

/* One read to seed (both strands are seeded, seeder.cpp:35-49). */
typedef struct DarwinSeedRead { uint64_t read_addr; uint32_t read_len; uint32_t reserved; } DarwinSeedRead;

/* One D-SOFT candidate == Anchors (software/seed_pos_table.h:30-40): the hit that pushed its bin over the threshold and
 * the collinear chained hits around it (left ascending incl. the anchor, right descending incl. the anchor). */
typedef struct DarwinSeedAnchor {
    uint64_t hit_offset;              /* (reference arena offset << 32) | strand-local read offset */
    uint64_t left_off;                /* left_chained_hits  = pool[left_off  .. left_off  + left_n)  */
    uint64_t right_off;               /* right_chained_hits = pool[right_off .. right_off + right_n) */
    uint32_t left_n;
    uint32_t right_n;
} DarwinSeedAnchor;

/* Everything darwin_gpu_align_reads needs beyond the seeding parameters fixed at darwin_gpu_seed_index time. */
typedef struct DarwinAlignParams {
    DarwinFilterParams filter;        /* params.cfg [GACT_first_tile] */
    DarwinExtendParams extend;        /* params.cfg [GACT_extend] + do_overlap */
    float   slope_threshold;          /* params.cfg slope_threshold (filter.cpp:271) */
    int32_t reserved;
} DarwinAlignParams;

typedef struct DarwinGpuStats {
    uint64_t kernel_launches;   /* kernels of this library launched since create */
    uint64_t tiles_fast;        /* tiles finished by the packed fast path */
    uint64_t tiles_exact;       /* tiles (re)computed by the unpacked exact path */
    uint64_t tiles_rerun;       /* fast tiles whose traceback asked for the exact path (subset of tiles_exact) */
    uint64_t cells;             /* DP cells requested (algorithmic) */
    uint64_t cells_exact;       /* DP cells computed by the unpacked exact path (incl. reruns) */
    uint64_t tiles_xfast;       /* tiles computed by the packed exact path (large tiles, reruns) */
    float    last_kernel_ms;    /* CUDA-event time of the last tiles/extend/filter kernel(s) */
    float    reserved;
    uint64_t tiles_filter;      /* score-only tiles finished by the packed filter path (two tiles per warp) */
    float    last_seed_ms;      /* darwin_gpu_align_reads: CUDA-event time of the D-SOFT kernels of the last call */
    float    last_filter_ms;    /*   ... of the first-tile filter kernels */
    float    last_extend_ms;    /*   ... of the extension kernels (anchor walking + score / compaction) */
    float    reserved2;
    uint64_t tiles_scoreonly;   /* large tiles settled by the score-only pre-pass (corner ZERO: no traceback needed) */
} DarwinGpuStats;

typedef struct DarwinGpu DarwinGpu;   /* opaque */

/* replaces InitializeProcessor (Processor.h:50): one handle per device/host thread.
 * arena_bytes = size of the byte-addressed sequence arena this handle mirrors
 * (the reference's g_DRAM, DRAM.cpp:8); the device keeps it 4-bit packed. */
int darwin_gpu_create(DarwinGpu** h, int device, uint64_t arena_bytes);   /* on failure *h = NULL, nothing leaks, and darwin_gpu_last_error(NULL) has the reason */
/* A further handle ("lane") on the parent's device that SHARES the parent's arena replica: own stream, own scratch,
 * own result buffers.  Lets several host threads keep kernels of independent batches in flight on one GPU (the
 * reference's tokens, main.cpp:615-624) without one arena copy per thread.  Uploads through any lane are visible to
 * all of them once darwin_gpu_upload returns.  Ordering rules (enforced): darwin_gpu_destroy(parent) fails with
 * DARWIN_ERR_INVALID while lanes are alive; darwin_gpu_set_scoring on the parent also reaches its lanes; the first
 * darwin_gpu_seed_index on the parent is adopted by existing lanes, a REbuild is refused while lanes share the table. */
int darwin_gpu_create_shared(DarwinGpu** h, DarwinGpu* parent);
int darwin_gpu_destroy(DarwinGpu* h);

/* replaces g_InitializeScoringParameters (Processor.cpp:48-80). */
int darwin_gpu_set_scoring(DarwinGpu* h, const DarwinScoring* s);

/* replaces g_InitializeReferenceMemory / g_InitializeReadMemory (Processor.cpp:82-85,
 * sender.cpp:4-97): copy n ASCII bases to arena offset arena_addr (any alignment). */
int darwin_gpu_upload(DarwinGpu* h, uint64_t arena_addr, const char* ascii, uint64_t n);

/* Several uploads in one call (the reads of many host batches ahead of one merged device call): same effect as calling
 * darwin_gpu_upload once per span, but the spans share the staging buffers, the copies are queued back to back and the
 * call synchronises once. */
typedef struct DarwinSpan { uint64_t arena_addr; const char* ascii; uint64_t n; } DarwinSpan;
int darwin_gpu_upload_spans(DarwinGpu* h, const DarwinSpan* spans, int n_spans);

/* replaces g_BatchAlignmentSIMD (Processor.cpp:718-762): n independent tiles.
 * tb_words may be NULL when do_traceback == 0. */
int darwin_gpu_tiles(DarwinGpu* h, int do_traceback, const DarwinTileReq* req, int n,
                     DarwinTileRes* res, uint64_t* tb_words, int tb_words_per_req);

/* replaces extender_body::operator() (extender.cpp:9-1065) for n anchors of any
 * number of reads.  ops_pool receives the op strings; per-anchor capacity is
 * derived from read_len, overflow is flagged per alignment. */
int darwin_gpu_extend(DarwinGpu* h, const DarwinExtendParams* p,
                      const DarwinAnchor* anchors, int n,
                      const uint64_t* hit_pool, uint64_t n_hits,
                      DarwinAlnRes* res, uint8_t* ops_pool, uint64_t ops_pool_bytes);

/* replaces the tile part of filter_body::operator() (filter.cpp:28-122 forward, :131-223 reverse complement) for n
 * candidates of any number of reads: builds the first-tile requests (first_tile_size x first_tile_size, score-only,
 * max-cell mode), runs them and applies the score and overlap tests.  The slope filter (filter.cpp:227-289) and the
 * copying of the chained hits stay on the host (darwin_b200/host: gpu_filter_body). */
int darwin_gpu_filter(DarwinGpu* h, const DarwinFilterParams* p, const DarwinFilterCand* cands, int n,
                      DarwinFilterRes* res);

/* replaces the SeedPosTable constructor + the minimizer pass over the reference (main.cpp:323-341, :508,
 * seed_pos_table.cpp:41-160): builds the seed position table in HBM from the chromosomes already uploaded. */
int darwin_gpu_seed_index(DarwinGpu* h, const DarwinSeedParams* p, const DarwinChrom* chroms, int n_chroms, uint64_t reference_size);

/* lanes created BEFORE the table was built adopt the parent's table with this call (lanes created afterwards share it
 * automatically) */
int darwin_gpu_seed_index_share(DarwinGpu* h, DarwinGpu* parent);

/* diagnostics / tests: copy the table back (buckets: n_buckets + 1 prefix sums; positions of buckets larger than
 * max_occ are unordered, D-SOFT never reads them) */
int darwin_gpu_seed_index_read(DarwinGpu* h, uint32_t* buckets, uint64_t buckets_cap, uint32_t* positions, uint64_t positions_cap,
                               uint64_t* n_buckets, uint64_t* n_positions, uint32_t* max_occ);

/* replaces seeder_body::operator() (seeder.cpp:6-55) -> SeedPosTable::DSOFT (seed_pos_table.cpp:252-553) for n reads
 * already resident in the arena: anchors of read r, strand s (0 = forward, 1 = reverse complement) are
 * anchors[anchor_begin[2r+s] .. anchor_begin[2r+s+1]) in the reference's output order.  Capacities are in elements;
 * DARWIN_ERR_CAPACITY reports the needed sizes in *n_anchors / *n_pool. */
int darwin_gpu_seed(DarwinGpu* h, const DarwinSeedRead* reads, int n, uint32_t* anchor_begin /* 2n+1 */,
                    DarwinSeedAnchor* anchors, uint64_t anchors_cap, uint64_t* n_anchors,
                    uint64_t* pool, uint64_t pool_cap, uint64_t* n_pool);

/* The reference-guided pipeline of one read batch in ONE call -- seeder_body, filter_body (incl. slopeFilter) and
 * extender_body (main.cpp:590-624's seeder -> filter -> extender chain) -- with the chained hits kept in HBM between the
 * stages.  Reads must be resident; darwin_gpu_seed_index must have been called.  anchors_out[i] / res[i] describe the
 * i-th location handed to the extension (read_num = index into `reads`), forward-strand locations first, in the
 * reference's order; *n_out = their number (also set on DARWIN_ERR_CAPACITY). */
int darwin_gpu_align_reads(DarwinGpu* h, const DarwinAlignParams* p, const DarwinSeedRead* reads, int n,
                           DarwinAnchor* anchors_out, DarwinAlnRes* res, uint64_t cap, uint64_t* n_out,
                           uint8_t* ops_pool, uint64_t ops_pool_bytes);

/* Output stage (host functions, no device work): what printer_body needs from an alignment, straight from the op string.
 *
 * replaces the CIGAR construction of printer_body::AlignmentToSam (software/printer.cpp:236-301): leading soft clip
 * (query_start_offset), one run per maximal run of equal ops (I: '-' in the reference string, D: '-' in the query string,
 * M otherwise), trailing soft clip (query_length - query_end_offset - 1); "*" when empty.  Writes *len characters (no
 * terminator); DARWIN_ERR_CAPACITY reports the needed size in *len. */
int darwin_gpu_cigar(const DarwinAlnRes* r, const uint8_t* ops_pool, uint32_t query_length, char* out, uint64_t cap, uint64_t* len);

/* replaces the ordering and overlap suppression of printer_body::sam_printer (printer.cpp:15-47) for the alignments of any
 * number of reads: order[k] = index (into anchors / res) of the k-th EMITTED alignment in print order -- stable sort by
 * (read_num, score descending) -- and keep[k] = 1 when it is printed, 0 when more than half of its query span is covered by
 * a better alignment of the same read.  order and keep need room for n entries; *n_order = number of emitted alignments. */
int darwin_gpu_sam_select(const DarwinAnchor* anchors, const DarwinAlnRes* res, uint64_t n, uint32_t* order, uint8_t* keep,
                          uint64_t* n_order);

/* device-resident variants used by bench.py's `value` leg: same work, inputs and
 * outputs stay in HBM (pointers are device pointers of this handle's device). */
int darwin_gpu_tiles_device(DarwinGpu* h, int do_traceback, const void* d_req, int n,
                            void* d_res, void* d_tb_words, int tb_words_per_req,
                            int max_ref_size, int max_query_size);

/* Page-locked host memory for request / result buffers: DMA goes straight to it (no staging copy, full PCIe rate).
 * Any buffer handed to the calls above may come from here; plain malloc'ed memory works too, slower. */
void* darwin_gpu_host_alloc(uint64_t bytes);
void  darwin_gpu_host_free(void* p);

int darwin_gpu_stats(DarwinGpu* h, DarwinGpuStats* out);

/* Batch sizing hint for darwin_gpu_extend / darwin_gpu_align_reads (no counterpart in the reference, whose batch is one read,
 * main.cpp:299,:688): *slots = number of anchors one launch keeps in flight at this tile_size (one persistent warp per
 * anchor).  With many short anchors it does not matter; a batch of FEW LONG reads (50 kbp: ~270 tiles per anchor) runs in
 * whole waves of that size, so a host that has the choice hands over a little under a multiple of it.  Needs the scoring. */
int darwin_gpu_extend_slots(DarwinGpu* h, int tile_size, int* slots);

/* Roofline denominator: measured issue rate, in 1e9 32-bit lane-ops per second, of packed-int16 /
 * integer instructions on this GPU (8 independent chains per thread, operands all loop-variant):
 * out[0] VIMNMX.U16x2 (2-input min/max), out[1] VIADDMNMX.U16x2, out[2] VIMNMX3.U16x2, out[3] IADD3, out[4] LOP3,
 * out[5] IMAD (fma pipe), out[6] 2-input LOP3, out[7] LOP3 with immediate, out[8] PRMT, out[9] SHFL.IDX.
 * Each packed lane-op updates two int16 cells. */
int darwin_gpu_int_peak(DarwinGpu* h, double out_glaneops[10]);
const char* darwin_gpu_last_error(DarwinGpu* h);
const char* darwin_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DARWIN_GPU_H */

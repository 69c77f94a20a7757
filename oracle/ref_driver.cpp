// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// C-callable driver around the UNMODIFIED reference translation units of
// yatisht/darwin (compiled where they lie under /root/reference/software by
// oracle/Makefile, against the shim headers in oracle/ref_shim).  It replaces
// the reference's main.cpp (which needs kseq.h and a real TBB flow graph):
// the arena/index/read bookkeeping below follows main.cpp:294-296, :418-466,
// :508, :631-698 step by step, then the reference's own stage functors are
// called directly (seeder_body, filter_body, extender_body).
//
// Built into oracle/_ref/libdarwin_ref{,_patched}.so.  Only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference leg may load it.
#include "../include/darwin_gpu.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <mutex>
#include <chrono>

#include "graph.h"        // reference: software/graph.h (pulls Processor.h, seed_pos_table.h, Index.h, DRAM.h)
#include "ConfigFile.h"   // reference: software/ConfigFile.h

// defined (external linkage) in the reference's Processor.cpp:718 but not declared in Processor.h
void BatchAlignmentSIMD(size_t token, char* dram, Darwin::BatchAlignmentInputFieldsDRAM& request,
                        Darwin::BatchAlignmentResultDRAM& result);

// ---- globals main.cpp would have defined (main.cpp:45, :55) -------------------
Configuration cfg;
SeedPosTable* sa = nullptr;

using namespace Darwin;

static std::vector<Read> g_reads;                 // every read added so far
static tbb::concurrent_vector<mini_list>* g_minimizers = nullptr;
static uint32_t* g_seed_hist = nullptr;
static size_t g_hist_size = 0;
static filter_data g_last_filter;                 // output of the last dref_seed_filter
static std::atomic<uint64_t> g_cells(0);          // ref_size*query_size over non-dummy requests
static std::atomic<uint64_t> g_tiles(0);
static int g_count_cells = 0;

// Counting wrapper installed into the reference's own seam (Processor.h:61).
static void CountingBatchAlignmentSIMD(size_t token, char* dram, BatchAlignmentInputFieldsDRAM& request,
                                       BatchAlignmentResultDRAM& result) {
    if (g_count_cells) {
        uint64_t c = 0, t = 0;
        for (auto& r : request.requests) {
            // idle-slot filler of the extender: 8x8 at address 0 with no flags (extender.cpp:209-220)
            bool dummy = (r.ref_size == 8 && r.query_size == 8 && r.ref_bases_start_addr == 0 &&
                          r.query_bases_start_addr == 0 && r.align_fields == 0);
            if (!dummy) { c += (uint64_t)r.ref_size * r.query_size; t++; }
        }
        g_cells += c; g_tiles += t;
    }
    BatchAlignmentSIMD(token, dram, request, result);
}

static void to_params(const DarwinScoring* s, AlignmentScoringParams& p) {
    p.sub_AA = s->sub_AA; p.sub_AC = s->sub_AC; p.sub_AG = s->sub_AG; p.sub_AT = s->sub_AT;
    p.sub_CC = s->sub_CC; p.sub_CG = s->sub_CG; p.sub_CT = s->sub_CT;
    p.sub_GG = s->sub_GG; p.sub_GT = s->sub_GT; p.sub_TT = s->sub_TT; p.sub_N = s->sub_N;
    p.gap_open = s->gap_open; p.gap_extend = s->gap_extend;
    p.long_gap_open = s->long_gap_open; p.long_gap_extend = s->long_gap_extend;
}

// main.cpp:59-121 equivalent (reverse complement, case preserved, 'N'-padded to WORD_SIZE)
static char* rev_comp(const char* seq, size_t n) {
    size_t padded = n + ((n % WORD_SIZE) ? WORD_SIZE - (n % WORD_SIZE) : 0);
    char* rc = (char*)scalable_aligned_malloc(padded ? padded : WORD_SIZE, 64);
    size_t r = 0;
    for (size_t i = n; i-- > 0;) {
        char c = seq[i], o;
        switch (c) {
            case 'a': o = 't'; break; case 'A': o = 'T'; break;
            case 'c': o = 'g'; break; case 'C': o = 'G'; break;
            case 'g': o = 'c'; break; case 'G': o = 'C'; break;
            case 't': o = 'a'; break; case 'T': o = 'A'; break;
            case 'n': o = 'n'; break; default: o = 'N'; break;
        }
        rc[r++] = o;
    }
    for (; r < padded; r++) rc[r] = 'N';
    return rc;
}

extern "C" {

const char* dref_flavour(void) {
#ifdef DREF_PATCHED
    return "patched";   // Processor.cpp:384 also initialises vF_La / vF_La_ext
#else
    return "as-is";
#endif
}

// main.cpp:178-230 (+ :262-288): read params.cfg, push scoring into the Processor.
int dref_load_cfg(const char* path, int do_overlap) {
    try {
        ConfigFile f(path);
        const char* S = "GACT_scoring";
        const char* names[11] = {"sub_AA","sub_AC","sub_AG","sub_AT","sub_CC","sub_CG","sub_CT","sub_GG","sub_GT","sub_TT","sub_N"};
        for (int i = 0; i < 11; i++) cfg.gact_sub_mat[i] = f.Value(S, names[i]);
        cfg.gap_open = f.Value(S, "gap_open"); cfg.gap_extend = f.Value(S, "gap_extend");
        cfg.long_gap_open = f.Value(S, "long_gap_open"); cfg.long_gap_extend = f.Value(S, "long_gap_extend");
        const char* D = "DSOFT_params";
        cfg.seed_size = f.Value(D, "seed_size"); cfg.minimizer_window = f.Value(D, "minimizer_window");
        cfg.bin_size = f.Value(D, "bin_size"); cfg.dsoft_threshold = f.Value(D, "threshold");
        cfg.num_seeds = f.Value(D, "num_seeds"); cfg.seed_occurence_multiple = f.Value(D, "seed_occurence_multiple");
        cfg.max_candidates = f.Value(D, "max_candidates"); cfg.max_stride = f.Value(D, "max_stride");
        cfg.do_overlap = do_overlap;
        const char* F = "GACT_first_tile";
        cfg.first_tile_size = f.Value(F, "first_tile_size");
        cfg.first_tile_score_threshold = f.Value(F, "first_tile_score_threshold");
        cfg.first_tile_batch_size = f.Value(F, "first_tile_batch_size");
        cfg.min_overlap = f.Value(F, "min_overlap");
        cfg.slope_threshold = (float)f.Value(F, "slope_threshold");
        const char* E = "GACT_extend";
        cfg.tile_size = f.Value(E, "tile_size"); cfg.tile_overlap = f.Value(E, "tile_overlap");
        cfg.batch_size = f.Value(E, "batch_size");
        cfg.num_threads = f.Value("Multithreading", "num_threads");
    } catch (...) { return -1; }
    AlignmentScoringParams p; AlignmentScoringParamsResponse resp;
    p.sub_AA = cfg.gact_sub_mat[0]; p.sub_AC = cfg.gact_sub_mat[1]; p.sub_AG = cfg.gact_sub_mat[2]; p.sub_AT = cfg.gact_sub_mat[3];
    p.sub_CC = cfg.gact_sub_mat[4]; p.sub_CG = cfg.gact_sub_mat[5]; p.sub_CT = cfg.gact_sub_mat[6];
    p.sub_GG = cfg.gact_sub_mat[7]; p.sub_GT = cfg.gact_sub_mat[8]; p.sub_TT = cfg.gact_sub_mat[9];
    p.sub_N = cfg.gact_sub_mat[10];
    p.gap_open = cfg.gap_open; p.gap_extend = cfg.gap_extend;
    p.long_gap_open = cfg.long_gap_open; p.long_gap_extend = cfg.long_gap_extend;
    g_InitializeScoringParameters(0, p, resp);
    g_BatchAlignmentSIMD = CountingBatchAlignmentSIMD;
    return 0;
}

// Set scoring both in the Processor statics and in cfg (AlignmentScore reads cfg).
int dref_set_scoring(const DarwinScoring* s) {
    AlignmentScoringParams p; AlignmentScoringParamsResponse resp;
    to_params(s, p);
    g_InitializeScoringParameters(0, p, resp);
    const int32_t* v = (const int32_t*)s;
    for (int i = 0; i < 11; i++) cfg.gact_sub_mat[i] = v[i];
    cfg.gap_open = s->gap_open; cfg.gap_extend = s->gap_extend;
    cfg.long_gap_open = s->long_gap_open; cfg.long_gap_extend = s->long_gap_extend;
    g_BatchAlignmentSIMD = CountingBatchAlignmentSIMD;
    return 0;
}

int dref_set_extend(int tile_size, int tile_overlap, int batch_size, int do_overlap) {
    cfg.tile_size = tile_size; cfg.tile_overlap = tile_overlap; cfg.batch_size = batch_size;
    cfg.do_overlap = do_overlap;
    return 0;
}

// DSOFT / first-tile parameters for callers that do not load a params.cfg.
int dref_set_dsoft(int seed_size, int minimizer_window, int bin_size, int threshold, int num_seeds,
                   int seed_occurence_multiple, int max_candidates, int max_stride,
                   int first_tile_size, int first_tile_score_threshold, int first_tile_batch_size,
                   int min_overlap, float slope_threshold) {
    cfg.seed_size = seed_size; cfg.minimizer_window = minimizer_window; cfg.bin_size = bin_size;
    cfg.dsoft_threshold = threshold; cfg.num_seeds = num_seeds;
    cfg.seed_occurence_multiple = seed_occurence_multiple; cfg.max_candidates = max_candidates;
    cfg.max_stride = max_stride; cfg.first_tile_size = first_tile_size;
    cfg.first_tile_score_threshold = first_tile_score_threshold;
    cfg.first_tile_batch_size = first_tile_batch_size; cfg.min_overlap = min_overlap;
    cfg.slope_threshold = slope_threshold;
    return 0;
}

// main.cpp:294-296: one arena per process (4 GiB of lazily committed virtual memory).
int dref_reset_arena(void) {
    if (!g_DRAM) g_DRAM = new DRAM;
    if (!g_DRAM->buffer) return -1;
    g_DRAM->referenceSize = 0;
    g_DRAM->bufferPosition = 0;
    Index::chr_id.clear(); Index::chr_coord.clear(); Index::chr_len.clear(); Index::chr_len_unpadded.clear();
    Index::init();
    for (auto& r : g_reads) scalable_aligned_free((void*)r.rc_seq.data());
    g_reads.clear();
    if (g_minimizers) { delete g_minimizers; g_minimizers = nullptr; }
    if (g_seed_hist) { scalable_free(g_seed_hist); g_seed_hist = nullptr; }
    if (sa) { delete sa; sa = nullptr; }
    return 0;
}

char* dref_arena(void) { return g_DRAM ? g_DRAM->buffer : nullptr; }
uint64_t dref_arena_reference_size(void) { return g_DRAM ? g_DRAM->referenceSize : 0; }
uint64_t dref_arena_position(void) { return g_DRAM ? g_DRAM->bufferPosition : 0; }

// main.cpp:418-466 (reference reader) + :323-341 (minimizer node) for one sequence.
// Returns the arena offset of the sequence, or 0 when the reference would stop reading (len <= 64).
uint64_t dref_add_chr(const char* name, const char* seq, uint64_t len, int collect_minimizers) {
    const size_t readBufferLimit = 1 << 6;                       // main.cpp:299
    if (len <= readBufferLimit) return 0;
    size_t seq_len = len, seq_len_unpadded = len;
    uint64_t at = g_DRAM->referenceSize;
    memcpy(g_DRAM->buffer + at, seq, seq_len);
    size_t extra = seq_len % WORD_SIZE;
    if (extra != 0) { extra = WORD_SIZE - extra; memset(g_DRAM->buffer + at + seq_len, 'N', extra); seq_len += extra; }
    g_DRAM->referenceSize += seq_len;
    Index::chr_id.push_back(std::string(name));
    Index::chr_len.push_back(seq_len);
    Index::chr_len_unpadded.push_back(seq_len_unpadded);
    Index::chr_coord.push_back(g_DRAM->referenceSize);
    g_DRAM->bufferPosition = g_DRAM->referenceSize;              // main.cpp:483
    if (collect_minimizers) {
        int kmer_size = cfg.seed_size;
        if (!g_minimizers) {
            g_minimizers = new tbb::concurrent_vector<mini_list>();
            g_hist_size = 1ull << (kmer_size << 1);
            g_seed_hist = (uint32_t*)scalable_calloc(g_hist_size, sizeof(uint32_t));
        }
        auto miniList = g_minimizers->grow_by(1);
        miniList->reserve(seq_len_unpadded);
        uint32_t seq_start = (uint32_t)at;
        iterate_minimizers(g_DRAM->buffer + at, (uint32_t)seq_len_unpadded, kmer_size, cfg.minimizer_window,
            [&](uint64_t p, uint32_t m) {
                g_seed_hist[m]++;
                miniList->push_back(((uint64_t)m << 32) + p + seq_start);
            });
    }
    return at;
}

// main.cpp:508
int dref_build_index(void) {
    if (!g_minimizers) return -1;
    sa = new SeedPosTable((uint32_t)g_DRAM->referenceSize, cfg.seed_size, cfg.minimizer_window, cfg.max_stride,
                          cfg.seed_occurence_multiple, cfg.bin_size, *g_minimizers, g_seed_hist, g_hist_size);
    return 0;
}

// main.cpp:631-698 for one read.  Returns the read number (index into the driver's read list) or -1 if skipped.
int dref_add_read(const char* name, const char* seq, uint64_t len, uint64_t* arena_addr) {
    const size_t readBufferLimit = 1 << 6;
    size_t extra = g_DRAM->bufferPosition % WORD_SIZE;
    if (extra != 0) g_DRAM->bufferPosition += WORD_SIZE - extra;
    size_t seq_len = len;
    if (seq_len <= readBufferLimit) return -1;
    if (g_DRAM->bufferPosition + WORD_SIZE + seq_len > g_DRAM->size) g_DRAM->bufferPosition = g_DRAM->referenceSize;
    memcpy(g_DRAM->buffer + g_DRAM->bufferPosition, seq, seq_len);
    Read read;
    read.description = std::string(name);
    read.seq = bond::blob(g_DRAM->buffer + g_DRAM->bufferPosition, seq_len);
    read.rc_seq = bond::blob(rev_comp(read.seq.data(), seq_len), seq_len);
    if (arena_addr) *arena_addr = g_DRAM->bufferPosition;
    extra = seq_len % WORD_SIZE;
    if (extra != 0) { extra = WORD_SIZE - extra; memset(g_DRAM->buffer + g_DRAM->bufferPosition + seq_len, 'N', extra); seq_len += extra; }
    g_DRAM->bufferPosition += seq_len;
    g_reads.push_back(read);
    return (int)g_reads.size() - 1;
}

// ---- tile level: the reference's BatchAlignmentSIMD on an arbitrary byte buffer -------------
int dref_tiles(const char* dram, int do_traceback, const DarwinTileReq* req, int n,
               DarwinTileRes* res, uint64_t* tb_words, int tb_words_per_req) {
    if (!dram) dram = g_DRAM->buffer;
    BatchAlignmentInputFieldsDRAM in; BatchAlignmentResultDRAM out;
    in.do_traceback = (uint8_t)do_traceback;
    in.requests.resize(1);
    for (int i = 0; i < n; i++) {
        AlignmentInputFieldsDRAM& r = in.requests[0];
        r.align_fields = req[i].align_fields; r.index = req[i].index;
        r.ref_bases_start_addr = req[i].ref_bases_start_addr; r.query_bases_start_addr = req[i].query_bases_start_addr;
        r.ref_size = req[i].ref_size; r.query_size = req[i].query_size;
        r.max_tb_steps = req[i].max_tb_steps; r.score_threshold = req[i].score_threshold;
        BatchAlignmentSIMD(0, (char*)dram, in, out);
        const AlignmentResult& a = out.results[0];
        res[i].score = (int32_t)a.score; res[i].ref_offset = a.ref_offset; res[i].query_offset = a.query_offset;
        res[i].ref_max_pos = a.ref_max_pos; res[i].query_max_pos = a.query_max_pos;
        res[i].total_TB_pointers = a.total_TB_pointers; res[i].index = a.index; res[i].status = 0;
        if (tb_words && do_traceback) {
            if ((int)a.TB_pointers.size() > tb_words_per_req) return DARWIN_ERR_CAPACITY;
            for (size_t w = 0; w < a.TB_pointers.size(); w++) tb_words[(size_t)i * tb_words_per_req + w] = a.TB_pointers[w];
        }
    }
    return 0;
}

// Multi-threaded timing leg for the tile-only workload: std::thread x nthreads,
// one stream of BatchAlignmentSIMD calls per thread.  Returns wall seconds.
double dref_tiles_mt(const char* dram, int do_traceback, const DarwinTileReq* req, int n,
                     DarwinTileRes* res, uint64_t* tb_words, int tb_words_per_req, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([=]() {
            int lo = (int)((int64_t)n * t / nthreads), hi = (int)((int64_t)n * (t + 1) / nthreads);
            if (hi > lo)
                dref_tiles(dram, do_traceback, req + lo, hi - lo, res + lo,
                           tb_words ? tb_words + (size_t)lo * tb_words_per_req : nullptr, tb_words_per_req);
        });
    }
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- anchors: seeder_body + filter_body on reads [first, first+count) ------------------------
// Results are kept in the driver and fetched with dref_get_anchors.
int dref_seed_filter(int first, int count) {
    if (!sa) return -1;
    reader_output reads(g_reads.begin() + first, g_reads.begin() + first + count);
    seeder_input sin(reads, 0);
    filter_input fin = seeder_body()(sin);
    extender_input ein = filter_body()(fin);
    g_last_filter = std::get<1>(std::get<0>(ein));
    // read_num inside the locations is relative to the batch: make it absolute
    for (auto& l : g_last_filter.fwLocations) l.read_num += first;
    for (auto& l : g_last_filter.rcLocations) l.read_num += first;
    return (int)(g_last_filter.fwLocations.size() + g_last_filter.rcLocations.size());
}

// ---- seeder alone, then the reference's filter_body on the kept seeder output (first-tile filter parity) ----------
static filter_input* g_last_seed = nullptr;
static int g_last_seed_first = 0;

int dref_seed(int first, int count) {
    if (!sa) return -1;
    reader_output reads(g_reads.begin() + first, g_reads.begin() + first + count);
    seeder_input sin(reads, 0);
    delete g_last_seed;
    g_last_seed = new filter_input(seeder_body()(sin));
    g_last_seed_first = first;
    auto& d = std::get<1>(std::get<0>(*g_last_seed));
    return (int)(d.fwAnchors.size() + d.rcAnchors.size());
}

// Hand-made seeder output for reads [first, first+count): candidate k = (hit_offset[k], read read_num[k] (absolute),
// strand[k]); within each strand the candidates must be sorted by read (seeder.cpp:38-50 appends read by read).
// Lets the tests drive the reference's filter_body with candidates D-SOFT would rarely propose (tiles clamped at
// chromosome / read ends, low scores).
int dref_seed_custom(int first, int count, const uint64_t* hit_offset, const int* read_num, const uint8_t* strand, int n) {
    reader_output reads(g_reads.begin() + first, g_reads.begin() + first + count);
    seeder_data d;
    for (int s = 0; s < 2; s++) {
        auto& anchors = s ? d.rcAnchors : d.fwAnchors;
        auto& buckets = s ? d.rcAnchorBuckets : d.fwAnchorBuckets;
        buckets.push_back(0ull);
        int k = 0;
        for (int r = 0; r < count; r++) {
            for (; k < n; k++) {
                if (strand[k] != s) continue;
                if (read_num[k] - first < r) return -3;                      // not sorted by read
                if (read_num[k] - first > r) break;
                Anchors a(hit_offset[k]);
                a.num_chained_hits = 1; a.anchor_score = 0;
                a.left_chained_hits.push_back(hit_offset[k]); a.right_chained_hits.push_back(hit_offset[k]);
                anchors.push_back(a);
            }
            buckets.push_back(anchors.size());
        }
    }
    delete g_last_seed;
    g_last_seed = new filter_input(filter_payload(reads, d), 0);
    g_last_seed_first = first;
    return (int)(d.fwAnchors.size() + d.rcAnchors.size());
}

// the candidates of the last dref_seed as filter_body sees them (filter.cpp:44-56): forward strand first
int dref_get_candidates(DarwinFilterCand* out, int* read_num_out, int cap) {
    if (!g_last_seed) return -1;
    auto& reads = std::get<0>(std::get<0>(*g_last_seed));
    auto& d = std::get<1>(std::get<0>(*g_last_seed));
    int n = 0;
    for (int strand = 0; strand < 2; strand++) {
        auto& anchors = strand ? d.rcAnchors : d.fwAnchors;
        auto& buckets = strand ? d.rcAnchorBuckets : d.fwAnchorBuckets;
        for (size_t c = 0; c < anchors.size(); c++) {
            if (n >= cap) return DARWIN_ERR_CAPACITY;
            uint32_t hit = (uint32_t)(anchors[c].hit_offset >> 32), offset = (uint32_t)((anchors[c].hit_offset << 32) >> 32);
            size_t chr_id = std::upper_bound(Index::chr_coord.cbegin(), Index::chr_coord.cend(), hit) - Index::chr_coord.cbegin() - 1;
            size_t read_num = std::upper_bound(buckets.cbegin(), buckets.cend(), c) - buckets.cbegin() - 1;
            const Read& rd = reads[read_num];
            DarwinFilterCand& k = out[n];
            memset(&k, 0, sizeof(k));
            k.read_addr = (uint64_t)(rd.seq.data() - g_DRAM->buffer);
            k.hit = hit; k.offset = offset; k.chr_start = Index::chr_coord[chr_id]; k.chr_len = Index::chr_len[chr_id];
            k.read_len = (uint32_t)rd.seq.size(); k.strand = (uint8_t)strand;
            if (read_num_out) read_num_out[n] = (int)read_num + g_last_seed_first;
            n++;
        }
    }
    return n;
}

// the full output of the last dref_seed (SeedPosTable::DSOFT per read and strand): anchors with their chained hits in
// the layout of darwin_gpu_seed -- anchors of read r (relative to `first`), strand s at [begin[2r+s], begin[2r+s+1])
int64_t dref_get_seed_anchors(DarwinSeedAnchor* out, uint64_t cap, uint32_t* begin, uint64_t* pool, uint64_t pool_cap, uint64_t* pool_used) {
    if (!g_last_seed) return -1;
    auto& reads = std::get<0>(std::get<0>(*g_last_seed));
    auto& d = std::get<1>(std::get<0>(*g_last_seed));
    uint64_t n = 0, used = 0;
    for (size_t r = 0; r < reads.size(); r++) {
        for (int strand = 0; strand < 2; strand++) {
            auto& anchors = strand ? d.rcAnchors : d.fwAnchors;
            auto& buckets = strand ? d.rcAnchorBuckets : d.fwAnchorBuckets;
            begin[2 * r + strand] = (uint32_t)n;
            for (size_t c = buckets[r]; c < buckets[r + 1]; c++) {
                const Anchors& a = anchors[c];
                if (n >= cap || used + a.left_chained_hits.size() + a.right_chained_hits.size() > pool_cap) return DARWIN_ERR_CAPACITY;
                DarwinSeedAnchor& o = out[n++];
                o.hit_offset = a.hit_offset;
                o.left_off = used; o.left_n = (uint32_t)a.left_chained_hits.size();
                for (auto x : a.left_chained_hits) pool[used++] = x;
                o.right_off = used; o.right_n = (uint32_t)a.right_chained_hits.size();
                for (auto x : a.right_chained_hits) pool[used++] = x;
            }
        }
    }
    begin[2 * reads.size()] = (uint32_t)n;
    if (pool_used) *pool_used = used;
    return (int64_t)n;
}

// D-SOFT state of the driver: what SeedPosTable's constructor was given (dref_build_index) and the chromosomes
int dref_get_chroms(DarwinChrom* out, int cap) {
    int n = (int)Index::chr_id.size();
    if (n > cap) return DARWIN_ERR_CAPACITY;
    for (int k = 0; k < n; k++) { out[k].start = Index::chr_coord[k]; out[k].len_unpadded = (uint32_t)Index::chr_len_unpadded[k]; }
    return n;
}
void dref_get_seed_params(DarwinSeedParams* p) {
    p->seed_size = cfg.seed_size; p->minimizer_window = cfg.minimizer_window; p->bin_size = (int32_t)cfg.bin_size;
    p->threshold = cfg.dsoft_threshold; p->num_seeds = cfg.num_seeds; p->seed_occurence_multiple = cfg.seed_occurence_multiple;
    p->max_stride = cfg.max_stride; p->do_overlap = cfg.do_overlap;
}

// the reference's filter_body (first tiles through g_BatchAlignmentSIMD + slopeFilter) on the last dref_seed output;
// fetch the locations with dref_get_anchors
int dref_filter_last(void) {
    if (!g_last_seed) return -1;
    extender_input ein = filter_body()(*g_last_seed);
    g_last_filter = std::get<1>(std::get<0>(ein));
    for (auto& l : g_last_filter.fwLocations) l.read_num += g_last_seed_first;
    for (auto& l : g_last_filter.rcLocations) l.read_num += g_last_seed_first;
    return (int)(g_last_filter.fwLocations.size() + g_last_filter.rcLocations.size());
}

uint64_t dref_anchor_hits_total(void) {
    uint64_t n = 0;
    for (auto& l : g_last_filter.fwLocations) n += l.left_hit_offsets.size() + l.right_hit_offsets.size();
    for (auto& l : g_last_filter.rcLocations) n += l.left_hit_offsets.size() + l.right_hit_offsets.size();
    return n;
}

// Serialise the last filter output: forward-strand anchors first, then reverse (extender order).
int dref_get_anchors(DarwinAnchor* out, int cap, uint64_t* hit_pool, uint64_t hit_cap, uint64_t hit_base) {
    int n = 0; uint64_t h = 0;
    for (int strand = 0; strand < 2; strand++) {
        auto& v = strand ? g_last_filter.rcLocations : g_last_filter.fwLocations;
        for (auto& l : v) {
            if (n >= cap) return DARWIN_ERR_CAPACITY;
            if (h + l.left_hit_offsets.size() + l.right_hit_offsets.size() > hit_cap) return DARWIN_ERR_CAPACITY;
            DarwinAnchor& a = out[n++];
            memset(&a, 0, sizeof(a));
            const Read& rd = g_reads[l.read_num];
            a.read_addr = (uint64_t)(rd.seq.data() - g_DRAM->buffer);
            a.reference_pos = l.reference_pos; a.query_pos = l.query_pos;
            a.chr_start = Index::chr_coord[l.chr_id]; a.ref_len = Index::chr_len[l.chr_id];
            a.read_len = (uint32_t)rd.seq.size(); a.read_num = l.read_num; a.chr_id = l.chr_id; a.score = l.score;
            a.left_hits_off = (uint32_t)(hit_base + h); a.left_hits_n = (uint32_t)l.left_hit_offsets.size();
            for (auto x : l.left_hit_offsets) hit_pool[h++] = x;
            a.right_hits_off = (uint32_t)(hit_base + h); a.right_hits_n = (uint32_t)l.right_hit_offsets.size();
            for (auto x : l.right_hit_offsets) hit_pool[h++] = x;
            a.strand = (uint8_t)strand;
        }
    }
    return n;
}

static ExtendLocations to_location(const DarwinAnchor& a, const uint64_t* hit_pool, int read_num) {
    ExtendLocations l;
    l.read_num = read_num; l.chr_id = a.chr_id; l.score = a.score;
    l.reference_pos = a.reference_pos; l.query_pos = a.query_pos;
    l.left_hit_offsets.assign(hit_pool + a.left_hits_off, hit_pool + a.left_hits_off + a.left_hits_n);
    l.right_hit_offsets.assign(hit_pool + a.right_hits_off, hit_pool + a.right_hits_off + a.right_hits_n);
    return l;
}

static void fill_result(const ExtendAlignments& e, DarwinAlnRes& r, uint8_t* ops_pool, uint64_t& used, uint64_t cap) {
    r.flags |= DARWIN_ALN_EMITTED;
    r.reference_start_offset = e.reference_start_offset; r.reference_end_offset = e.reference_end_offset;
    r.query_start_offset = e.query_start_offset; r.query_end_offset = e.query_end_offset;
    r.score = e.score;
    size_t n = e.aligned_reference_str.size();
    r.n_ops = (uint32_t)n; r.ops_offset = used;
    if (used + n > cap) { r.flags |= DARWIN_ALN_OPS_OVERFLOW; return; }
    for (size_t k = 0; k < n; k++) {
        char rc = e.aligned_reference_str[k], qc = e.aligned_query_str[k];
        ops_pool[used + k] = (rc == '-') ? DARWIN_OP_I : (qc == '-') ? DARWIN_OP_D : DARWIN_OP_M;
    }
    used += n;
}

// ---- anchor level: extender_body, ONE anchor per call so results map 1:1 onto anchors --------
// (cfg.batch_size lock-step batching does not change any anchor's result, only the output order;
//  SURVEY Appendix B.)  The anchor's read must have been added with dref_add_read (read_num).
int dref_extend(const DarwinAnchor* anchors, int n, const uint64_t* hit_pool,
                DarwinAlnRes* res, uint8_t* ops_pool, uint64_t ops_pool_bytes) {
    uint64_t used = 0;
    g_count_cells = 1;
    for (int i = 0; i < n; i++) {
        const DarwinAnchor& a = anchors[i];
        DarwinAlnRes& r = res[i];
        memset(&r, 0, sizeof(r));
        if (a.read_num < 0 || a.read_num >= (int)g_reads.size()) return DARWIN_ERR_INVALID;
        reader_output reads(1, g_reads[a.read_num]);
        filter_data fd;
        (a.strand ? fd.rcLocations : fd.fwLocations).push_back(to_location(a, hit_pool, 0));
        extender_input in(extender_payload(reads, fd), 0);
        extender_node::output_ports_type ports;
        int large0 = extender_body::num_large_tiles;
        g_cells = 0; g_tiles = 0;
        extender_body()(in, ports);
        r.n_tiles = (uint32_t)g_tiles; r.cells = g_cells;
        r.n_large_tiles = (uint32_t)(extender_body::num_large_tiles - large0);
        auto& outv = std::get<1>(std::get<0>(std::get<0>(ports).items[0])).extend_alignments;
        if (!outv.empty()) fill_result(outv[0], r, ops_pool, used, ops_pool_bytes);
    }
    g_count_cells = 0;
    return 0;
}

// Timing leg: the reference's own call pattern (all anchors of one read per extender_body call,
// cfg.batch_size slots in lock-step), std::thread x nthreads over reads.  anchors must be grouped by
// read_num (as dref_get_anchors emits per dref_seed_filter call).  Returns wall seconds; *cells gets
// the algorithmic cell count, *n_alignments the number of emitted alignments.
double dref_extend_mt(const DarwinAnchor* anchors, int n, const uint64_t* hit_pool, int nthreads,
                      uint64_t* cells, uint64_t* n_alignments) {
    struct Group { int read_num; std::vector<int> idx; };
    std::vector<Group> groups;
    for (int i = 0; i < n; i++) {
        if (groups.empty() || groups.back().read_num != anchors[i].read_num) groups.push_back({anchors[i].read_num, {}});
        groups.back().idx.push_back(i);
    }
    std::atomic<size_t> next(0); std::atomic<uint64_t> alns(0);
    g_cells = 0; g_tiles = 0; g_count_cells = 1;
    if (nthreads < 1) nthreads = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([&, t]() {
            for (;;) {
                size_t g = next++;
                if (g >= groups.size()) break;
                reader_output reads(1, g_reads[groups[g].read_num]);
                filter_data fd;
                for (int i : groups[g].idx)
                    (anchors[i].strand ? fd.rcLocations : fd.fwLocations).push_back(to_location(anchors[i], hit_pool, 0));
                extender_input in(extender_payload(reads, fd), (size_t)t);
                extender_node::output_ports_type ports;
                extender_body()(in, ports);
                alns += std::get<1>(std::get<0>(std::get<0>(ports).items[0])).extend_alignments.size();
            }
        });
    }
    for (auto& x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    g_count_cells = 0;
    if (cells) *cells = g_cells;
    if (n_alignments) *n_alignments = alns;
    return s;
}

// Timing leg for reads/s: the reference's whole per-read chain -- seeder_body -> filter_body -> extender_body, one read per
// batch exactly as main.cpp:590-702 feeds them (readBufferLimit = 64 bytes) -- on `nthreads` std::threads playing the
// reference's tokens.  stats[0] = wall seconds, [1] = alignments, [2..4] = seconds inside the three stages (summed over
// threads), [5] = DP cells of the extension's tile requests.  Returns the number of alignments.
int dref_pipeline_cpu_mt(int first, int count, int nthreads, double* stats) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0); std::atomic<uint64_t> alns(0);
    std::vector<double> ts(nthreads, 0.0), tf(nthreads, 0.0), te(nthreads, 0.0);
    g_cells = 0; g_tiles = 0; g_count_cells = 1;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([&, t]() {
            for (;;) {
                const int k = next.fetch_add(1);
                if (k >= count) break;
                reader_output reads(1, g_reads[first + k]);
                const auto a0 = now();
                filter_input fin = seeder_body()(seeder_input(reads, (size_t)t));
                const auto a1 = now();
                extender_input ein = filter_body()(fin);
                const auto a2 = now();
                extender_node::output_ports_type ports;
                extender_body()(ein, ports);
                const auto a3 = now();
                ts[t] += secs(a0, a1); tf[t] += secs(a1, a2); te[t] += secs(a2, a3);
                alns += std::get<1>(std::get<0>(std::get<0>(ports).items[0])).extend_alignments.size();
            }
        });
    }
    for (auto& x : th) x.join();
    if (stats) {
        stats[0] = secs(t0, now()); stats[1] = (double)alns.load(); stats[2] = stats[3] = stats[4] = 0;
        for (int t = 0; t < nthreads; t++) { stats[2] += ts[t]; stats[3] += tf[t]; stats[4] += te[t]; }
        stats[5] = (double)g_cells.load();
    }
    g_count_cells = 0;
    return (int)alns.load();
}

int dref_num_reads(void) { return (int)g_reads.size(); }
int dref_read_len(int k) { return (k >= 0 && k < (int)g_reads.size()) ? (int)g_reads[k].seq.size() : -1; }
uint64_t dref_read_addr(int k) { return (k >= 0 && k < (int)g_reads.size()) ? (uint64_t)(g_reads[k].seq.data() - g_DRAM->buffer) : 0; }
unsigned dref_chr_start(int k) { return Index::chr_coord[k]; }
unsigned dref_chr_len(int k) { return Index::chr_len[k]; }
int dref_num_chr(void) { return (int)Index::chr_id.size(); }
int dref_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

} // extern "C"

// ---- drop-in check: the reference pipeline with the GPU library swapped in through the host adapter ----------
// (darwin_b200/host/darwin_gpu_processor.cpp).  Built only into libdarwin_ref_gpu.so.
#ifdef DREF_WITH_GPU
#include "../darwin_b200/host/darwin_gpu_processor.h"
#include <algorithm>
#include <iostream>
#include <unistd.h>

// software defaults of the reference (external linkage in Processor.cpp:48,:82; not declared in Processor.h)
void InitializeScoringParams(size_t token, Darwin::AlignmentScoringParams& request, Darwin::AlignmentScoringParamsResponse& response);
void InitializeMemory(size_t token, char* dram, Darwin::InitializeDRAMMessage& request, Darwin::InitializeDRAMMessageResponse& response);

static Darwin::AlignmentScoringParams cfg_params() {
    Darwin::AlignmentScoringParams p;
    p.sub_AA = cfg.gact_sub_mat[0]; p.sub_AC = cfg.gact_sub_mat[1]; p.sub_AG = cfg.gact_sub_mat[2]; p.sub_AT = cfg.gact_sub_mat[3];
    p.sub_CC = cfg.gact_sub_mat[4]; p.sub_CG = cfg.gact_sub_mat[5]; p.sub_CT = cfg.gact_sub_mat[6];
    p.sub_GG = cfg.gact_sub_mat[7]; p.sub_GT = cfg.gact_sub_mat[8]; p.sub_TT = cfg.gact_sub_mat[9]; p.sub_N = cfg.gact_sub_mat[10];
    p.gap_open = cfg.gap_open; p.gap_extend = cfg.gap_extend; p.long_gap_open = cfg.long_gap_open; p.long_gap_extend = cfg.long_gap_extend;
    return p;
}

extern "C" {

// what main() would do once: create the GPU processors, install the g_* table, push the
// scoring and the arena (reference + reads) through the reference's own upload messages (sender.cpp chunking).
int dref_gpu_init(int gpus) {
    try {
        if (darwin_gpu_host::InitializeProcessor(cfg.num_threads, gpus, "") == 0) return -1;
        darwin_gpu_host::InstallProcessorTable();
        Darwin::AlignmentScoringParams p = cfg_params();
        Darwin::AlignmentScoringParamsResponse resp;
        g_InitializeScoringParameters(0, p, resp);
        if (resp.status != Darwin::Status::OK) return -2;
        // sender.cpp:26-44: 2048-byte, 128-aligned messages, 8 bases per u64
        const uint64_t total = g_DRAM->bufferPosition;
        for (uint64_t at = 0; at < total; at += MAX_CHAR_TO_SEND) {
            uint64_t nb = std::min<uint64_t>(MAX_CHAR_TO_SEND, total - at);
            if (nb % 8) nb += 8 - nb % 8;
            Darwin::InitializeDRAMMessage m; Darwin::InitializeDRAMMessageResponse r;
            m.start_addr = at; m.num_bytes = (uint16_t)nb; m.data.resize(nb / 8);
            memcpy(m.data.data(), g_DRAM->buffer + at, nb);
            g_InitializeReferenceMemory(0, g_DRAM->buffer, m, r);
            if (r.status != Darwin::Status::OK) return -3;
        }
    } catch (const std::exception& e) { fprintf(stderr, "dref_gpu_init: %s\n", e.what()); return -4; }
    return 0;
}

void dref_gpu_shutdown(void) {
    darwin_gpu_host::ShutdownProcessor();
}

// GPU filter stage on the last dref_seed output (darwin_gpu_host::gpu_filter_body); fetch with dref_get_anchors
int dref_filter_last_gpu(void) {
    if (!g_last_seed) return -1;
    try {
        extender_input ein = darwin_gpu_host::gpu_filter_body()(*g_last_seed);
        g_last_filter = std::get<1>(std::get<0>(ein));
    } catch (const std::exception& e) { fprintf(stderr, "dref_filter_last_gpu: %s\n", e.what()); return -2; }
    for (auto& l : g_last_filter.fwLocations) l.read_num += g_last_seed_first;
    for (auto& l : g_last_filter.rcLocations) l.read_num += g_last_seed_first;
    return (int)(g_last_filter.fwLocations.size() + g_last_filter.rcLocations.size());
}

// seeder -> filter -> extender for reads [first, first+count); use_gpu = 1 selects gpu_extender_body, use_gpu = 2 also
// gpu_filter_body (with use_gpu <= 1 the filter still goes through g_BatchAlignmentSIMD, i.e. through the GPU once
// dref_gpu_init has installed the table).
// Writes one canonical text line per alignment, sorted; returns the number of alignments or <0.
int dref_pipeline(int first, int count, int use_gpu, char* out, uint64_t cap) {
    try {
        const bool sam = (use_gpu & 8) != 0;                   // bit 3: emit what the reference's printer_body prints
        use_gpu &= 7;
        reader_output reads(g_reads.begin() + first, g_reads.begin() + first + count);
        seeder_input sin(reads, 0);
        if (sam && use_gpu == 5) {                             // native output stage: reads in, SAM text out (gpu_sam_body)
            printer_body::done_header = 0;
            const std::string text = darwin_gpu_host::gpu_sam_body()(sin);
            if (text.size() + 1 > cap) return -2;
            memcpy(out, text.data(), text.size()); out[text.size()] = 0;
            int lines = 0;
            for (char c : text) lines += c == '\n';
            return lines;
        }
        filter_input fin = (use_gpu >= 3) ? darwin_gpu_host::gpu_seeder_body()(sin) : seeder_body()(sin);
        extender_input ein = (use_gpu >= 2) ? darwin_gpu_host::gpu_filter_body()(fin) : filter_body()(fin);
        extender_node::output_ports_type ports;
        if (use_gpu) darwin_gpu_host::gpu_extender_body()(ein, ports);
        else extender_body()(ein, ports);
        if (sam) {
            // printer.cpp:7-98 (SAM) / :100-180 (MHAP, cfg.do_overlap): the reference's own output stage, unmodified;
            // its std::cout / printf output is captured through a pipe-less redirect of both streams into a file
            printer_input pin = std::get<0>(ports).items[0];
            fflush(stdout); std::cout.flush();
            char path[] = "/tmp/dref_sam_XXXXXX";
            int fd = mkstemp(path);
            if (fd < 0) return -3;
            int saved = dup(1);
            dup2(fd, 1);
            printer_body::done_header = 0;
            printer_body()(pin);
            std::cout.flush(); fflush(stdout);
            dup2(saved, 1); close(saved);
            off_t len = lseek(fd, 0, SEEK_END);
            if ((uint64_t)len + 1 > cap) { close(fd); unlink(path); return -2; }
            lseek(fd, 0, SEEK_SET);
            ssize_t got = read(fd, out, (size_t)len);
            close(fd); unlink(path);
            if (got != len) return -3;
            out[len] = 0;
            int lines = 0;
            for (off_t k = 0; k < len; k++) lines += out[k] == '\n';
            return lines;
        }
        auto& al = std::get<1>(std::get<0>(std::get<0>(ports).items[0])).extend_alignments;
        std::vector<std::string> lines;
        for (auto& e : al) {
            std::string o = std::to_string(e.read_num) + " " + std::to_string(e.chr_id) + " " + std::string(1, e.strand) + " " +
                            std::to_string(e.reference_start_offset) + " " + std::to_string(e.reference_end_offset) + " " +
                            std::to_string(e.query_start_offset) + " " + std::to_string(e.query_end_offset) + " " +
                            std::to_string(e.score) + " " + e.aligned_reference_str + " " + e.aligned_query_str;
            lines.push_back(o);
        }
        std::sort(lines.begin(), lines.end());
        uint64_t pos = 0;
        for (auto& l : lines) {
            if (pos + l.size() + 2 > cap) return -2;
            memcpy(out + pos, l.data(), l.size()); pos += l.size(); out[pos++] = '\n';
        }
        out[pos] = 0;
        return (int)lines.size();
    } catch (const std::exception& e) { fprintf(stderr, "dref_pipeline: %s\n", e.what()); return -1; }
}

// ---- multi-threaded end-to-end run: `threads` host threads play the reference's tokens
// (main.cpp:615-624); each pulls batches of `reads_per_batch` reads and runs seeder_body -> filter -> extender on them.
// mode 0: the reference's CPU stages; mode 1: GPU extender only (filter tiles through g_BatchAlignmentSIMD);
// mode 2: gpu_filter_body + gpu_extender_body; mode 3: gpu_seeder_body as well (dref_gpu_seed_index first);
// mode 4: gpu_align_body (one resident device call per merged batch).  With the GPU stages all threads share the per-GPU combiner
// (darwin_b200/host/darwin_gpu_combiner.h).  Output: canonical sorted lines like dref_pipeline (out may be NULL);
// stats[0] = wall seconds, [1] = alignments, [2] = seconds inside seeder_body (summed over threads), [3] = filter stage,
// [4] = extender stage, [5] = DP cells of the alignments' tile requests when dref_count_cells(1) was set (CPU modes).
int dref_pipeline_mt(int first, int count, int threads, int reads_per_batch, int mode, char* out, uint64_t cap, double* stats) {
    if (threads < 1) threads = 1;
    if (reads_per_batch < 1) reads_per_batch = 1;
    const int nbatches = (count + reads_per_batch - 1) / reads_per_batch;
    std::atomic<int> next(0);
    std::atomic<int> failed(0);
    std::mutex out_mutex;
    std::vector<std::string> lines;
    std::atomic<uint64_t> n_aln(0);
    std::vector<double> t_seed(threads, 0.0), t_filter(threads, 0.0), t_extend(threads, 0.0);
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        th.emplace_back([&, t] {
            try {
                for (;;) {
                    const int b = next.fetch_add(1);
                    if (b >= nbatches || failed.load()) break;
                    const int lo = first + b * reads_per_batch, hi = std::min(first + count, lo + reads_per_batch);
                    reader_output reads(g_reads.begin() + lo, g_reads.begin() + hi);
                    if (mode == 5) {                                   // all stages + the native SAM output stage
                        const auto b0 = now();
                        const std::string text = darwin_gpu_host::gpu_sam_body()(seeder_input(reads, (size_t)t));
                        t_extend[t] += secs(b0, now());
                        uint64_t nl = 0;
                        for (char c : text) nl += c == '\n';
                        n_aln += nl;
                        if (out) { std::lock_guard<std::mutex> g(out_mutex); lines.push_back(text); }
                        continue;
                    }
                    if (mode >= 4) {                                   // all stages in one device call
                        const auto b0 = now();
                        extender_node::output_ports_type ports;
                        darwin_gpu_host::gpu_align_body()(seeder_input(reads, (size_t)t), ports);
                        t_extend[t] += secs(b0, now());
                        auto& al = std::get<1>(std::get<0>(std::get<0>(ports).items[0])).extend_alignments;
                        n_aln += al.size();
                        if (out) {
                            std::vector<std::string> mine;
                            for (auto& e : al)
                                mine.push_back(std::to_string(e.read_num + lo) + " " + std::to_string(e.chr_id) + " " + std::string(1, e.strand) + " " +
                                               std::to_string(e.reference_start_offset) + " " + std::to_string(e.reference_end_offset) + " " +
                                               std::to_string(e.query_start_offset) + " " + std::to_string(e.query_end_offset) + " " +
                                               std::to_string(e.score) + " " + e.aligned_reference_str + " " + e.aligned_query_str);
                            std::lock_guard<std::mutex> g(out_mutex);
                            for (auto& l : mine) lines.push_back(std::move(l));
                        }
                        continue;
                    }
                    const auto a0 = now();
                    filter_input fin = (mode >= 3) ? darwin_gpu_host::gpu_seeder_body()(seeder_input(reads, (size_t)t))
                                                   : seeder_body()(seeder_input(reads, (size_t)t));
                    const auto a1 = now();
                    extender_input ein = (mode >= 2) ? darwin_gpu_host::gpu_filter_body()(fin) : filter_body()(fin);
                    const auto a2 = now();
                    extender_node::output_ports_type ports;
                    if (mode >= 1) darwin_gpu_host::gpu_extender_body()(ein, ports);
                    else extender_body()(ein, ports);
                    const auto a3 = now();
                    t_seed[t] += secs(a0, a1); t_filter[t] += secs(a1, a2); t_extend[t] += secs(a2, a3);
                    auto& al = std::get<1>(std::get<0>(std::get<0>(ports).items[0])).extend_alignments;
                    n_aln += al.size();
                    if (out) {
                        std::vector<std::string> mine;
                        for (auto& e : al)
                            mine.push_back(std::to_string(e.read_num + lo) + " " + std::to_string(e.chr_id) + " " + std::string(1, e.strand) + " " +
                                           std::to_string(e.reference_start_offset) + " " + std::to_string(e.reference_end_offset) + " " +
                                           std::to_string(e.query_start_offset) + " " + std::to_string(e.query_end_offset) + " " +
                                           std::to_string(e.score) + " " + e.aligned_reference_str + " " + e.aligned_query_str);
                        std::lock_guard<std::mutex> g(out_mutex);
                        for (auto& l : mine) lines.push_back(std::move(l));
                    }
                }
            } catch (const std::exception& e) { fprintf(stderr, "dref_pipeline_mt: %s\n", e.what()); failed = 1; }
        });
    }
    for (auto& x : th) x.join();
    const double wall = secs(t0, now());
    if (stats) {
        stats[0] = wall; stats[1] = (double)n_aln.load(); stats[2] = stats[3] = stats[4] = 0;
        for (int t = 0; t < threads; t++) { stats[2] += t_seed[t]; stats[3] += t_filter[t]; stats[4] += t_extend[t]; }
        stats[5] = (double)g_cells.load();
    }
    if (failed.load()) return -1;
    if (out) {
        std::sort(lines.begin(), lines.end());
        uint64_t pos = 0;
        for (auto& l : lines) {
            if (pos + l.size() + 2 > cap) return -2;
            memcpy(out + pos, l.data(), l.size()); pos += l.size();
            if (mode != 5) out[pos++] = '\n';                   // mode 5 chunks are whole SAM blocks (sorted by their first line)
        }
        out[pos] = 0;
    }
    return (int)n_aln.load();
}

// merged-call statistics summed over all combiners (GPUs x lanes): out[0..2] device calls (tiles, filter, extend), [3..5] requests, [6..8] items,
// [9..11] largest number of requests merged into one call
void dref_combiner_stats(uint64_t* out) {
    darwin_gpu_host::CombinerStats s = darwin_gpu_host::combiner_stats_total();
    for (int k = 0; k < 3; k++) { out[k] = s.device_calls[k]; out[3 + k] = s.requests[k]; out[6 + k] = s.items[k]; out[9 + k] = s.max_merged[k]; }
}

void dref_host_profile(double* out3) { darwin_gpu_host::host_profile(out3); }

// where the combining threads' time went, summed over all combiners since they were created (seconds): uploads, inside
// the device calls, merging + scattering; then device calls / requests / reads of the ALIGN kind
void dref_combiner_phases(double* out6) {
    darwin_gpu_host::CombinerStats s = darwin_gpu_host::combiner_stats_total();
    for (int k = 0; k < 3; k++) out6[k] = (double)s.phase_ns[k] * 1e-9;
    out6[3] = (double)s.device_calls[4]; out6[4] = (double)s.requests[4]; out6[5] = (double)s.items[4];
}

// seed position table on the GPUs (after dref_gpu_init)
int dref_gpu_seed_index(void) {
    try { darwin_gpu_host::BuildSeedIndex(); } catch (const std::exception& e) { fprintf(stderr, "dref_gpu_seed_index: %s\n", e.what()); return -1; }
    return 0;
}

// back to the software Processor (the reference's defaults, Processor.cpp:1063-1069)
void dref_use_cpu_table(void) {
    g_InitializeScoringParameters = InitializeScoringParams;
    g_InitializeReferenceMemory = InitializeMemory; g_InitializeReadMemory = InitializeMemory;
    g_BatchAlignmentSIMD = BatchAlignmentSIMD;
    Darwin::AlignmentScoringParams p = cfg_params(); Darwin::AlignmentScoringParamsResponse resp;
    g_InitializeScoringParameters(0, p, resp);
}

} // extern "C"
#endif

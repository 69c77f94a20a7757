// See darwin_gpu_processor.h.  Host-side only; all arithmetic of the path runs in libdarwin_gact.so.
#include "darwin_gpu_processor.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace darwin_gpu_host {

static std::vector<DarwinGpu*> g_handles;
static std::vector<GpuCombiner*> g_combiners;                            // one per GPU: merges the requests of all host threads
static std::string g_error;
static std::mutex g_error_mutex;
static uint64_t g_arena_bytes = 4ull * 1024ull * 1024ull * 1024ull;       // DRAM.cpp:8

static void fail(DarwinGpu* h, int rc, const char* what) {
    std::string msg = std::string(what) + ": " + darwin_gpu_last_error(h) + " (" + std::to_string(rc) + ")";
    { std::lock_guard<std::mutex> g(g_error_mutex); g_error = msg; }
    throw std::runtime_error(msg);                                       // no silent CPU fallback
}
static void fail_msg(int rc, const char* what, const std::string& detail) {
    std::string msg = std::string(what) + ": " + detail + " (" + std::to_string(rc) + ")";
    { std::lock_guard<std::mutex> g(g_error_mutex); g_error = msg; }
    throw std::runtime_error(msg);
}

// a thread-local copy: the pointer stays valid while other host threads record their own errors
const char* last_error() {
    static thread_local std::string copy;
    { std::lock_guard<std::mutex> g(g_error_mutex); copy = g_error; }
    return copy.c_str();
}

DarwinGpu* handle_for_token(size_t token) {
    if (g_handles.empty()) fail_msg(DARWIN_ERR_NOT_READY, "handle_for_token", "InitializeProcessor was not called");
    return g_handles[token % g_handles.size()];
}

GpuCombiner& combiner_for_token(size_t token) {
    if (g_combiners.empty()) fail_msg(DARWIN_ERR_NOT_READY, "combiner_for_token", "InitializeProcessor was not called");
    return *g_combiners[token % g_combiners.size()];
}

// The reads of one batch as arena spans.  The reader lays the reads of a batch out back to back, WORD_SIZE-aligned with
// 'N' padding in between (main.cpp:645-686), so they normally merge into ONE span; a wrap of the arena (main.cpp:652-655)
// simply starts a second one.
static std::vector<UploadSpan> read_spans(const std::vector<Read>& reads) {
    std::vector<UploadSpan> spans;
    for (const auto& rd : reads) {
        const uint64_t at = (uint64_t)(rd.seq.data() - g_DRAM->buffer), n = rd.seq.size();
        if (!spans.empty()) {
            UploadSpan& s = spans.back();
            const uint64_t end = s.arena_addr + s.n;
            if (at >= end && at - end < 2 * WORD_SIZE) { s.n = at + n - s.arena_addr; continue; }
        }
        spans.push_back(UploadSpan{at, rd.seq.data(), n});
    }
    return spans;
}

// Layout of g_handles / g_combiners: lane-major, g_handles[lane * g_gpus + gpu]; lane 0 of every GPU owns the arena replica.
static int g_gpus = 0, g_lanes = 0;

// where the host threads' time goes in gpu_align_body (nanoseconds, summed over threads): [0] building the request,
// [1] blocked in the combiner (device call + waiting for it), [2] rebuilding ExtendAlignments (gapped strings)
static std::atomic<uint64_t> g_prof_ns[3];
static inline uint64_t now_ns() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
void host_profile(double out_seconds[3]) { for (int k = 0; k < 3; k++) out_seconds[k] = (double)g_prof_ns[k].exchange(0) * 1e-9; }

size_t InitializeProcessor(int threads, int gpus, std::string /*chip_ids*/) {
    if (gpus < 1) gpus = 1;
    if (g_DRAM) g_arena_bytes = g_DRAM->size;
    // lanes per GPU: independent handles (stream + scratch) sharing the GPU's arena replica, so that the latency-bound
    // kernels of small batches and the host-side pre/post work of different host threads overlap
    int lanes = 4;
    if (const char* e = getenv("DARWIN_GPU_LANES")) lanes = atoi(e);
    if (threads > 0 && lanes > (threads + gpus - 1) / gpus) lanes = (threads + gpus - 1) / gpus;
    if (lanes < 1) lanes = 1;
    for (int d = 0; d < gpus; d++) {
        DarwinGpu* h = nullptr;
        int rc = darwin_gpu_create(&h, d, g_arena_bytes);
        if (rc != DARWIN_OK) { if (d == 0) fail(h, rc, "darwin_gpu_create"); break; }
        g_handles.push_back(h);
    }
    g_gpus = (int)g_handles.size(); g_lanes = 1;
    for (int l = 1; l < lanes; l++) {
        for (int d = 0; d < g_gpus; d++) {
            DarwinGpu* h = nullptr;
            int rc = darwin_gpu_create_shared(&h, g_handles[d]);
            if (rc != DARWIN_OK) fail(h, rc, "darwin_gpu_create_shared");
            g_handles.push_back(h);
        }
        g_lanes = l + 1;
    }
    for (auto h : g_handles) g_combiners.push_back(new GpuCombiner(h, GpuCalls::library()));
    return g_handles.size();
}

void ShutdownProcessor() {
    for (auto c : g_combiners) delete c;
    g_combiners.clear();
    for (size_t k = g_handles.size(); k-- > 0;) darwin_gpu_destroy(g_handles[k]);       // arena owners (lane 0) last
    g_handles.clear();
    g_gpus = g_lanes = 0;
}

CombinerStats combiner_stats(size_t token) { return combiner_for_token(token).stats(); }

CombinerStats combiner_stats_total() {
    CombinerStats t; memset(&t, 0, sizeof(t));
    for (auto c : g_combiners) {
        const CombinerStats s = c->stats();
        for (int k = 0; k < 5; k++) {
            t.device_calls[k] += s.device_calls[k]; t.requests[k] += s.requests[k]; t.items[k] += s.items[k];
            if (s.max_merged[k] > t.max_merged[k]) t.max_merged[k] = s.max_merged[k];
        }
        for (int k = 0; k < 3; k++) t.phase_ns[k] += s.phase_ns[k];
    }
    return t;
}

void InitializeScoringParameters(size_t /*token*/, Darwin::AlignmentScoringParams& r,
                                 Darwin::AlignmentScoringParamsResponse& response) {
    DarwinScoring s{r.sub_AA, r.sub_AC, r.sub_AG, r.sub_AT, r.sub_CC, r.sub_CG, r.sub_CT, r.sub_GG, r.sub_GT, r.sub_TT,
                    r.sub_N, r.gap_open, r.gap_extend, r.long_gap_open, r.long_gap_extend};
    response.status = Darwin::Status::OK;
    for (auto h : g_handles)                                             // scoring is global state in the reference (Processor.cpp:15-19)
        if (darwin_gpu_set_scoring(h, &s) != DARWIN_OK) response.status = Darwin::Status::InvalidData;
}

// sender.cpp:26-44: <= 2048 bytes, 8 ASCII bases per u64, little-endian (main.cpp:131-145)
static void upload(size_t token, Darwin::InitializeDRAMMessage& m, Darwin::InitializeDRAMMessageResponse& response) {
    response.status = Darwin::Status::OK;
    if (m.data.size() * 8 < m.num_bytes) { response.status = Darwin::Status::InvalidData; return; }
    (void)token;
    const std::vector<UploadSpan> span{UploadSpan{m.start_addr, reinterpret_cast<const char*>(m.data.data()), m.num_bytes}};
    for (int d = 0; d < g_gpus; d++)                                     // every GPU keeps ONE replica of the arena (its lanes share it)
        if (g_combiners[d]->upload(span, nullptr) != DARWIN_OK) response.status = Darwin::Status::InvalidData;
}
void InitializeReferenceMemory(size_t token, char*, Darwin::InitializeDRAMMessage& m, Darwin::InitializeDRAMMessageResponse& r) { upload(token, m, r); }
void InitializeReadMemory(size_t token, char*, Darwin::InitializeDRAMMessage& m, Darwin::InitializeDRAMMessageResponse& r) { upload(token, m, r); }

void BatchAlignmentSIMD(size_t token, char* /*dram*/, Darwin::BatchAlignmentInputFieldsDRAM& request,
                        Darwin::BatchAlignmentResultDRAM& result) {
    GpuCombiner& gc = combiner_for_token(token);
    const size_t n = request.requests.size();
    result.results.resize(n);
    if (n == 0) return;
    std::vector<DarwinTileReq> req(n);
    int max_tb = 1;
    for (size_t i = 0; i < n; i++) {
        const auto& r = request.requests[i];
        req[i] = DarwinTileReq{r.ref_bases_start_addr, r.query_bases_start_addr, r.score_threshold, r.index, r.ref_size,
                               r.query_size, r.max_tb_steps, r.align_fields, {0, 0, 0}};
        if (r.max_tb_steps > max_tb) max_tb = r.max_tb_steps;
    }
    const int words = max_tb / 16 + 2;
    std::vector<DarwinTileRes> res(n);
    std::vector<uint64_t> tb(request.do_traceback ? n * words : 1);
    std::string err;
    int rc = gc.tiles(request.do_traceback, req.data(), (int)n, res.data(), tb.data(), words, &err);
    if (rc != DARWIN_OK) fail_msg(rc, "BatchAlignmentSIMD", err);
    for (size_t i = 0; i < n; i++) {
        auto& o = result.results[i];
        o.index = res[i].index; o.score = (uint32_t)res[i].score;
        o.ref_offset = res[i].ref_offset; o.query_offset = res[i].query_offset;
        o.ref_max_pos = res[i].ref_max_pos; o.query_max_pos = res[i].query_max_pos;
        o.total_TB_pointers = res[i].total_TB_pointers;
        o.TB_pointers.clear();
        if (request.do_traceback)
            o.TB_pointers.assign(tb.begin() + i * words, tb.begin() + i * words + (res[i].total_TB_pointers + 31) / 32);
    }
}

void InstallProcessorTable() {
    g_InitializeScoringParameters = InitializeScoringParameters;
    g_InitializeReferenceMemory = InitializeReferenceMemory;
    g_InitializeReadMemory = InitializeReadMemory;
    g_BatchAlignmentSIMD = BatchAlignmentSIMD;
}

// arena range this thread's last gpu_filter_body call made resident (lets gpu_extender_body skip the same upload)
static thread_local uint64_t t_resident_addr = ~0ull, t_resident_end = 0, t_resident_print = 0;
static thread_local int t_resident_gpu = -1;         // the marker is only valid for the GPU (arena replica) it was sent to

static int gpu_of_token(size_t token) { return g_gpus > 0 ? (int)(token % (size_t)g_gpus) : 0; }
// cheap fingerprint of what the spans hold (the arena is reused after a wrap, main.cpp:652-655: same addresses, new reads)
static uint64_t span_fingerprint(const std::vector<UploadSpan>& spans) {
    uint64_t f = 1469598103934665603ull;
    for (const auto& s : spans) {
        const uint64_t n = s.n, step = n > 256 ? n / 32 : 1;
        for (uint64_t k = 0; k < n; k += step) { f ^= (unsigned char)s.ascii[k]; f *= 1099511628211ull; }
        if (n) { f ^= (unsigned char)s.ascii[n - 1]; f *= 1099511628211ull; }
        f ^= n; f *= 1099511628211ull;
    }
    return f;
}
static bool already_resident(size_t token, const std::vector<UploadSpan>& spans) {
    return !spans.empty() && t_resident_gpu == gpu_of_token(token) && t_resident_addr == spans[0].arena_addr &&
           t_resident_end == spans.back().arena_addr + spans.back().n && t_resident_print == span_fingerprint(spans);
}
static void mark_resident(size_t token, const std::vector<UploadSpan>& spans) {
    if (spans.empty()) return;
    t_resident_gpu = gpu_of_token(token); t_resident_addr = spans[0].arena_addr; t_resident_end = spans.back().arena_addr + spans.back().n;
    t_resident_print = span_fingerprint(spans);
}

// ExtendLocations (graph.h:83-91) -> DarwinAnchor (what makeForward/BackwardAlignment look up, extender.cpp:1067-1159)
static void to_anchor(const ExtendLocations& l, const Read& rd, int strand, std::vector<uint64_t>& pool, DarwinAnchor& a) {
    a = DarwinAnchor{};
    a.read_addr = (uint64_t)(rd.seq.data() - g_DRAM->buffer);
    a.reference_pos = l.reference_pos; a.query_pos = l.query_pos;
    a.chr_start = Index::chr_coord[l.chr_id]; a.ref_len = Index::chr_len[l.chr_id];
    a.read_len = (uint32_t)rd.seq.size(); a.read_num = l.read_num; a.chr_id = l.chr_id; a.score = l.score;
    a.left_hits_off = (uint32_t)pool.size(); a.left_hits_n = (uint32_t)l.left_hit_offsets.size();
    pool.insert(pool.end(), l.left_hit_offsets.begin(), l.left_hit_offsets.end());
    a.right_hits_off = (uint32_t)pool.size(); a.right_hits_n = (uint32_t)l.right_hit_offsets.size();
    pool.insert(pool.end(), l.right_hit_offsets.begin(), l.right_hit_offsets.end());
    a.strand = (uint8_t)strand;
}

// One extended anchor -> ExtendAlignments as extender_body would have pushed it (graph.h:97-121): offsets, strand,
// AlignmentScore and the gapped strings rebuilt from the op string (extender.cpp:287-323 / :434-458).
static void emit_alignment(const DarwinAnchor& a, const DarwinAlnRes& r, const std::vector<uint8_t>& ops, const std::vector<Read>& reads,
                           extend_data& output) {
    if (!(r.flags & DARWIN_ALN_EMITTED)) return;
    if (r.flags & DARWIN_ALN_OPS_OVERFLOW) fail_msg(DARWIN_ERR_CAPACITY, "gpu_extender_body", "op string overflow");
    const Read& rd = reads[a.read_num];
    const char* qchars = a.strand ? rd.rc_seq.data() : rd.seq.data();        // extender.cpp:243 / :758
    ExtendAlignments e;
    e.read_num = a.read_num; e.chr_id = a.chr_id;
    e.reference_start_offset = r.reference_start_offset; e.reference_end_offset = r.reference_end_offset;
    e.query_start_offset = r.query_start_offset; e.query_end_offset = r.query_end_offset;
    e.curr_reference_offset = r.reference_end_offset + 1; e.curr_query_offset = r.query_end_offset + 1;
    e.reference_start_addr = a.chr_start; e.query_start_addr = (uint32_t)a.read_addr;
    e.reference_length = a.ref_len; e.query_length = a.read_len;
    e.left_extension_done = 1; e.right_extension_done = 1;
    e.used_large_tile = false; e.do_print = true; e.strand = a.strand ? '-' : '+';
    e.score = r.score; e.chain_score = 0;
    e.aligned_reference_str.resize(r.n_ops); e.aligned_query_str.resize(r.n_ops);
    const uint8_t* o = ops.data() + r.ops_offset;
    uint32_t cr = a.reference_pos - a.chr_start, cq = a.query_pos;
    for (int64_t p = (int64_t)r.n_left_ops - 1; p >= 0; p--) {
        const uint8_t d = o[p];
        e.aligned_reference_str[p] = (d == DARWIN_OP_I) ? '-' : g_DRAM->buffer[a.chr_start + cr];
        e.aligned_query_str[p] = (d == DARWIN_OP_D) ? '-' : qchars[cq];
        if (d != DARWIN_OP_I && cr > 0) cr--;
        if (d != DARWIN_OP_D && cq > 0) cq--;
    }
    cr = a.reference_pos - a.chr_start + 1; cq = a.query_pos + 1;
    for (uint32_t p = r.n_left_ops; p < r.n_ops; p++) {
        const uint8_t d = o[p];
        e.aligned_reference_str[p] = (d == DARWIN_OP_I) ? '-' : g_DRAM->buffer[a.chr_start + cr];
        e.aligned_query_str[p] = (d == DARWIN_OP_D) ? '-' : qchars[cq];
        if (d != DARWIN_OP_I && cr < a.ref_len) cr++;
        if (d != DARWIN_OP_D && cq < a.read_len) cq++;
    }
    output.extend_alignments.push_back(std::move(e));
    extender_body::num_extend_tiles += (int)r.n_tiles;
    extender_body::num_active_tiles += (int)r.n_tiles;
    extender_body::num_large_tiles += (int)r.n_large_tiles;
}

void gpu_extender_body::operator()(extender_input input, extender_node::output_ports_type& op) {
    auto& payload = get<0>(input);
    auto& reads = get<0>(payload);
    auto& data = get<1>(payload);
    size_t token = get<1>(input);
    GpuCombiner& gc = combiner_for_token(token);
    extend_data output;

    std::vector<DarwinAnchor> anchors;
    std::vector<uint64_t> pool;
    for (int strand = 0; strand < 2; strand++)
        for (const auto& l : (strand ? data.rcLocations : data.fwLocations)) {
            anchors.emplace_back();
            to_anchor(l, reads[l.read_num], strand, pool, anchors.back());
        }
    const int n = (int)anchors.size();
    if (n > 0) {
        // the reads of this batch must be resident (the software reference reads g_DRAM directly); gpu_filter_body has
        // normally sent them already when it ran on this thread
        std::vector<UploadSpan> spans = read_spans(reads);
        if (already_resident(token, spans)) spans.clear();
        DarwinExtendParams prm{cfg.tile_size, cfg.tile_overlap, cfg.do_overlap, 0};
        std::vector<DarwinAlnRes> res(n);
        std::vector<uint8_t> ops;
        std::string err;
        int rc = gc.extend(prm, spans, anchors.data(), n, pool.data(), pool.size(), res.data(), &ops, &err);
        if (rc != DARWIN_OK) fail_msg(rc, "gpu_extender_body", err);
        for (int k = 0; k < n; k++) emit_alignment(anchors[k], res[k], ops, reads, output);
    }
    get<1>(op).try_put(token);                                                        // extender.cpp:1062-1063
    get<0>(op).try_put(printer_input(printer_payload(reads, output), token));
}

// SeedPosTable construction on the GPUs (main.cpp:323-341 minimizer pass + :508): every GPU builds the table from its own
// arena replica (the reference must have been uploaded); lanes share their GPU's table.
void BuildSeedIndex() {
    std::vector<DarwinChrom> chroms(Index::chr_id.size());
    for (size_t c = 0; c < chroms.size(); c++) { chroms[c].start = Index::chr_coord[c]; chroms[c].len_unpadded = (uint32_t)Index::chr_len_unpadded[c]; }
    DarwinSeedParams prm{cfg.seed_size, cfg.minimizer_window, (int32_t)cfg.bin_size, cfg.dsoft_threshold, cfg.num_seeds,
                         cfg.seed_occurence_multiple, cfg.max_stride, cfg.do_overlap};
    for (int d = 0; d < g_gpus; d++) {
        int rc = darwin_gpu_seed_index(g_handles[d], &prm, chroms.data(), (int)chroms.size(), g_DRAM->referenceSize);
        if (rc != DARWIN_OK) fail(g_handles[d], rc, "darwin_gpu_seed_index");
        for (int l = 1; l < g_lanes; l++) {
            rc = darwin_gpu_seed_index_share(g_handles[(size_t)l * g_gpus + d], g_handles[d]);
            if (rc != DARWIN_OK) fail(g_handles[(size_t)l * g_gpus + d], rc, "darwin_gpu_seed_index_share");
        }
    }
}

// seeder_body::operator() (seeder.cpp:6-55): SeedPosTable::DSOFT for both strands of every read of the batch on the GPU;
// the output has the reference's layout (anchors appended read by read, bucket boundaries per read).
filter_input gpu_seeder_body::operator()(seeder_input input) {
    reader_output& reads = get<0>(input);
    size_t token = get<1>(input);
    GpuCombiner& gc = combiner_for_token(token);
    seeder_data output;
    output.fwAnchorBuckets.push_back(0ull);
    output.rcAnchorBuckets.push_back(0ull);
    if (!reads.empty()) {
        const std::vector<UploadSpan> spans = read_spans(reads);
        std::vector<DarwinSeedRead> sr(reads.size());
        for (size_t r = 0; r < reads.size(); r++) sr[r] = DarwinSeedRead{(uint64_t)(reads[r].seq.data() - g_DRAM->buffer), (uint32_t)reads[r].seq.size(), 0};
        std::vector<uint32_t> begin; std::vector<DarwinSeedAnchor> anchors; std::vector<uint64_t> pool;
        std::string err;
        int rc = gc.seed(spans, sr.data(), (int)sr.size(), &begin, &anchors, &pool, &err);
        if (rc != DARWIN_OK) fail_msg(rc, "gpu_seeder_body", err);
        mark_resident(token, spans);
        for (size_t r = 0; r < reads.size(); r++)
            for (int strand = 0; strand < 2; strand++) {
                auto& dst = strand ? output.rcAnchors : output.fwAnchors;
                auto& buckets = strand ? output.rcAnchorBuckets : output.fwAnchorBuckets;
                for (uint32_t i = begin[2 * r + strand]; i < begin[2 * r + strand + 1]; i++) {
                    const DarwinSeedAnchor& a = anchors[i];
                    dst.emplace_back(a.hit_offset);
                    Anchors& o = dst.back();
                    o.left_chained_hits.assign(pool.begin() + a.left_off, pool.begin() + a.left_off + a.left_n);
                    o.right_chained_hits.assign(pool.begin() + a.right_off, pool.begin() + a.right_off + a.right_n);
                    o.num_chained_hits = (int)(a.left_n + a.right_n);
                    o.anchor_score = 0;                                 // computed but never read downstream (seed_pos_table.cpp:453,:481)
                }
                buckets.push_back(dst.size());
            }
    } 
    return filter_input(filter_payload(reads, output), token);
}

// seeder_body + filter_body + extender_body of one batch in ONE device call (darwin_gpu_align_reads): the chained hits stay
// in HBM between the stages, only the per-candidate first-tile results come up for the slope filter (restated inside the
// library, filter.cpp:227-289).  Output == what gpu_seeder_body -> gpu_filter_body -> gpu_extender_body produce.
void gpu_align_body::operator()(seeder_input input, extender_node::output_ports_type& op) {
    reader_output& reads = get<0>(input);
    size_t token = get<1>(input);
    GpuCombiner& gc = combiner_for_token(token);
    extend_data output;
    if (!reads.empty()) {
        const uint64_t t0 = now_ns();
        const std::vector<UploadSpan> spans = read_spans(reads);
        std::vector<DarwinSeedRead> sr(reads.size());
        for (size_t r = 0; r < reads.size(); r++) sr[r] = DarwinSeedRead{(uint64_t)(reads[r].seq.data() - g_DRAM->buffer), (uint32_t)reads[r].seq.size(), 0};
        DarwinAlignParams prm{};
        prm.filter = DarwinFilterParams{cfg.first_tile_size, cfg.first_tile_score_threshold, cfg.min_overlap, 0};
        prm.extend = DarwinExtendParams{cfg.tile_size, cfg.tile_overlap, cfg.do_overlap, 0};
        prm.slope_threshold = cfg.slope_threshold;
        std::vector<DarwinAnchor> anchors; std::vector<DarwinAlnRes> res; std::vector<uint8_t> ops;
        std::string err;
        const uint64_t t1 = now_ns();
        int rc = gc.align(prm, spans, sr.data(), (int)sr.size(), &anchors, &res, &ops, &err);
        if (rc != DARWIN_OK) fail_msg(rc, "gpu_align_body", err);
        const uint64_t t2 = now_ns();
        for (size_t k = 0; k < anchors.size(); k++) emit_alignment(anchors[k], res[k], ops, reads, output);
        const uint64_t t3 = now_ns();
        g_prof_ns[0] += t1 - t0; g_prof_ns[1] += t2 - t1; g_prof_ns[2] += t3 - t2;
    }
    get<1>(op).try_put(token);
    get<0>(op).try_put(printer_input(printer_payload(reads, std::move(output)), token));
}

// Reads in, SAM out (see the header).  Field by field what printer_body::sam_printer / AlignmentToSam print (printer.cpp:49-78,
// :206-306).
std::string gpu_sam_body::operator()(seeder_input input) {
    reader_output& reads = get<0>(input);
    size_t token = get<1>(input);
    GpuCombiner& gc = combiner_for_token(token);
    std::string out;
    if (reads.empty()) return out;
    const std::vector<UploadSpan> spans = read_spans(reads);
    std::vector<DarwinSeedRead> sr(reads.size());
    for (size_t r = 0; r < reads.size(); r++) sr[r] = DarwinSeedRead{(uint64_t)(reads[r].seq.data() - g_DRAM->buffer), (uint32_t)reads[r].seq.size(), 0};
    DarwinAlignParams prm{};
    prm.filter = DarwinFilterParams{cfg.first_tile_size, cfg.first_tile_score_threshold, cfg.min_overlap, 0};
    prm.extend = DarwinExtendParams{cfg.tile_size, cfg.tile_overlap, cfg.do_overlap, 0};
    prm.slope_threshold = cfg.slope_threshold;
    std::vector<DarwinAnchor> anchors; std::vector<DarwinAlnRes> res; std::vector<uint8_t> ops;
    std::string err;
    int rc = gc.align(prm, spans, sr.data(), (int)sr.size(), &anchors, &res, &ops, &err);
    if (rc != DARWIN_OK) fail_msg(rc, "gpu_sam_body", err);
    const size_t n = anchors.size();
    std::vector<uint32_t> order(n ? n : 1); std::vector<uint8_t> keep(n ? n : 1);
    uint64_t m = 0;
    rc = darwin_gpu_sam_select(anchors.data(), res.data(), n, order.data(), keep.data(), &m);
    if (rc != DARWIN_OK) fail_msg(rc, "gpu_sam_body", "darwin_gpu_sam_select");
    std::vector<char> cigar;
    for (uint64_t k = 0; k < m; k++) {
        if (!keep[k]) continue;
        const DarwinAnchor& a = anchors[order[k]];
        const DarwinAlnRes& r = res[order[k]];
        if (r.flags & DARWIN_ALN_OPS_OVERFLOW) fail_msg(DARWIN_ERR_CAPACITY, "gpu_sam_body", "op string overflow");
        extender_body::num_extend_tiles += (int)r.n_tiles;
        extender_body::num_active_tiles += (int)r.n_tiles;
        extender_body::num_large_tiles += (int)r.n_large_tiles;
        const Read& rd = reads[a.read_num];
        if (printer_body::done_header == 0) {                                // printer.cpp:55-62 (checked again under the lock)
            std::lock_guard<std::mutex> g(::io_lock);                   // software/printer.cpp:3, declared in graph.h:190
            if (printer_body::done_header == 0) {
                out += "@HD\tVN:1.6\tSO:coordinate\n";
                for (size_t c = 0; c < Index::chr_id.size(); c++)
                    out += "@SQ\tSN:" + Index::chr_id[c] + "\tLN:" + std::to_string(Index::chr_len_unpadded[c]) + "\n";
                printer_body::done_header = 1;
            }
        }
        uint64_t len = 0;
        cigar.resize(16 * ((size_t)r.n_ops + 4));
        rc = darwin_gpu_cigar(&r, ops.data(), a.read_len, cigar.data(), cigar.size(), &len);
        if (rc != DARWIN_OK) fail_msg(rc, "gpu_sam_body", "darwin_gpu_cigar");
        const bool minus = a.strand != 0;
        out += std::string(rd.description.c_str());                          // QNAME (:209)
        out += '\t'; out += std::to_string((minus ? 16 : 0) + 64);           // FLAG (:210-224)
        out += '\t'; out += std::string(Index::chr_id[a.chr_id].c_str());    // RNAME (:226)
        out += '\t'; out += std::to_string(1 + r.reference_start_offset);    // POS (:227)
        out += "\t60\t";                                                    // MAPQ (:230)
        out.append(cigar.data(), (size_t)len);
        out += "\t*\t0\t0\t";                                               // RNEXT, PNEXT, TLEN (:314-316)
        if (minus) out.append(rd.rc_seq.data(), rd.rc_seq.size()); else out.append(rd.seq.data(), rd.seq.size());   // SEQ (:212-222)
        out += "\t*\tAS:i:"; out += std::to_string(r.score);                // QUAL, AS (:317-318)
        out += "\tZS:i:"; out += std::to_string(r.score);                   // CHAIN_SCORE = score (:320)
        out += '\n';
    }
    return out;
}

// filter_body::operator() (filter.cpp:8-225) with every first tile of the batch -- both strands, all reads -- in ONE
// darwin_gpu_filter call instead of g_BatchAlignmentSIMD batches of 64 (filter.cpp:22, :75).  Request construction,
// score and overlap tests run on the device; the ExtendLocations (chr_id / read_num lookups, chained hits) and the
// slope filter -- the reference's own filter_body::slopeFilter, unchanged -- stay here.
extender_input gpu_filter_body::operator()(filter_input input) {
    auto& payload = get<0>(input);
    auto& reads = get<0>(payload);
    auto& data = get<1>(payload);
    size_t token = get<1>(input);
    GpuCombiner& gc = combiner_for_token(token);
    filter_data output;

    const size_t nf = data.fwAnchors.size(), nr = data.rcAnchors.size();
    std::vector<DarwinFilterCand> cands(nf + nr);
    std::vector<int> chr_of(nf + nr), read_of(nf + nr);
    for (int strand = 0; strand < 2; strand++) {
        const auto& anchors = strand ? data.rcAnchors : data.fwAnchors;
        const auto& buckets = strand ? data.rcAnchorBuckets : data.fwAnchorBuckets;
        const size_t base = strand ? nf : 0;
        for (size_t c = 0; c < anchors.size(); c++) {
            const uint32_t hit = (uint32_t)(anchors[c].hit_offset >> 32);
            const uint32_t offset = (uint32_t)((anchors[c].hit_offset << 32) >> 32);
            const size_t chr_id = std::upper_bound(Index::chr_coord.cbegin(), Index::chr_coord.cend(), hit) - Index::chr_coord.cbegin() - 1;   // filter.cpp:49
            const size_t read_num = std::upper_bound(buckets.cbegin(), buckets.cend(), c) - buckets.cbegin() - 1;                                // filter.cpp:53
            const Read& read = reads[read_num];
            DarwinFilterCand& k = cands[base + c];
            k = DarwinFilterCand{};
            k.read_addr = (uint64_t)(read.seq.data() - g_DRAM->buffer);
            k.hit = hit; k.offset = offset;
            k.chr_start = Index::chr_coord[chr_id]; k.chr_len = Index::chr_len[chr_id];
            k.read_len = (uint32_t)read.seq.size(); k.strand = (uint8_t)strand;
            chr_of[base + c] = (int)chr_id; read_of[base + c] = (int)read_num;
        }
    }
    std::vector<DarwinFilterRes> res(cands.size());
    if (!cands.empty()) {
        std::vector<UploadSpan> spans = read_spans(reads);                             // the reads of this batch must be resident
        if (already_resident(token, spans)) spans.clear();
        DarwinFilterParams prm{cfg.first_tile_size, cfg.first_tile_score_threshold, cfg.min_overlap, 0};
        std::string err;
        int rc = gc.filter(prm, spans, cands.data(), (int)cands.size(), res.data(), &err);
        if (rc != DARWIN_OK) fail_msg(rc, "gpu_filter_body", err);
        mark_resident(token, spans);
    }
    for (int strand = 0; strand < 2; strand++) {
        const auto& anchors = strand ? data.rcAnchors : data.fwAnchors;
        const size_t base = strand ? nf : 0;
        std::deque<ExtendLocations> read_extend_locations;
        filter_body::num_filter_tiles += (int)anchors.size();                          // filter.cpp:80
        for (size_t c = 0; c < anchors.size(); c++) {
            const DarwinFilterRes& r = res[base + c];
            if (!(r.flags & DARWIN_FILTER_SCORE_OK)) continue;                         // filter.cpp:87
            if (r.flags & DARWIN_FILTER_OVERLAP_OK) {                                  // filter.cpp:104
                ExtendLocations loc;
                loc.read_num = read_of[base + c];
                loc.chr_id = chr_of[base + c];
                loc.score = r.score;
                loc.reference_pos = r.reference_pos;
                loc.query_pos = r.query_pos;
                loc.left_hit_offsets.assign(anchors[c].left_chained_hits.begin(), anchors[c].left_chained_hits.end());
                loc.right_hit_offsets.assign(anchors[c].right_chained_hits.begin(), anchors[c].right_chained_hits.end());
                read_extend_locations.push_back(loc);
            }
            filter_body::num_extend_requests += 1;                                     // filter.cpp:117
        }
        filter_body().slopeFilter(read_extend_locations, strand ? output.rcLocations : output.fwLocations);   // filter.cpp:124 / :223
    }
    return extender_input(extender_payload(reads, output), token);
}

} // namespace darwin_gpu_host

// CPU unit test of darwin_b200/host/darwin_gpu_combiner.h: the merge / scatter logic of the cross-read batcher, driven
// by several host threads, with the oracle's CPU functions standing in for the GPU entry points.  Built and called by
// tests/test_combiner.py (g++ -shared, linked against oracle/libgact_oracle.so).  TEST CODE ONLY.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../darwin_b200/host/darwin_gpu_combiner.h"
#include "../../oracle/gact_oracle.h"
#include "../../oracle/dsoft_oracle.h"

using namespace darwin_gpu_host;

struct Fake { GactScoring sc; char* dram; uint64_t dram_bytes; std::atomic<int> device_calls; int fail_filter; DsoftIndex* ix; };
static Fake* fake(DarwinGpu* h) { return reinterpret_cast<Fake*>(h); }
static void device_latency() { std::this_thread::sleep_for(std::chrono::milliseconds(2)); }

static int f_upload(DarwinGpu* h, uint64_t addr, const char* a, uint64_t n) {
    if (addr + n > fake(h)->dram_bytes) return DARWIN_ERR_INVALID;
    memcpy(fake(h)->dram + addr, a, n);
    return 0;
}
static int f_tiles(DarwinGpu* h, int tb, const DarwinTileReq* req, int n, DarwinTileRes* res, uint64_t* w, int words) {
    fake(h)->device_calls++; device_latency();
    if (tb) memset(w, 0, sizeof(uint64_t) * (size_t)n * words);
    return gact_tiles(&fake(h)->sc, fake(h)->dram, tb, GACT_RULE_STREAM, req, n, res, tb ? w : nullptr, words, nullptr);
}
static int f_filter(DarwinGpu* h, const DarwinFilterParams* p, const DarwinFilterCand* c, int n, DarwinFilterRes* res) {
    fake(h)->device_calls++; device_latency();
    if (fake(h)->fail_filter == 2) throw std::bad_alloc();
    if (fake(h)->fail_filter) return DARWIN_ERR_CUDA;
    return gact_filter(&fake(h)->sc, fake(h)->dram, p, c, n, res);
}
static int f_extend(DarwinGpu* h, const DarwinExtendParams* p, const DarwinAnchor* a, int n, const uint64_t* pool, uint64_t,
                    DarwinAlnRes* res, uint8_t* ops, uint64_t cap) {
    fake(h)->device_calls++; device_latency();
    memset(res, 0, sizeof(DarwinAlnRes) * (size_t)n);
    return gact_extend(&fake(h)->sc, fake(h)->dram, p, GACT_RULE_STREAM, a, n, pool, res, ops, cap);
}
static char rc_char(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }
// darwin_gpu_seed stand-in: SeedPosTable::DSOFT of the oracle for both strands of every read, output in the C-ABI layout
static int f_seed(DarwinGpu* h, const DarwinSeedRead* reads, int n, uint32_t* begin, DarwinSeedAnchor* anchors, uint64_t acap, uint64_t* na,
                  uint64_t* pool, uint64_t pcap, uint64_t* np) {
    fake(h)->device_calls++; device_latency();
    uint64_t a = 0, p = 0;
    for (int r = 0; r < n; r++)
        for (int s = 0; s < 2; s++) {
            begin[2 * r + s] = (uint32_t)a;
            const uint32_t len = reads[r].read_len;
            std::vector<char> seq(len + 160, 'N');
            for (uint32_t q = 0; q < len; q++) seq[q] = s ? rc_char(fake(h)->dram[reads[r].read_addr + len - 1 - q]) : fake(h)->dram[reads[r].read_addr + q];
            uint64_t used = 0;
            const int got = dsoft_query(fake(h)->ix, seq.data(), len, 40, 20, 0, anchors + a, (int)(acap - a), pool + p, pcap - p, &used);
            if (got < 0) return DARWIN_ERR_CAPACITY;
            for (int i = 0; i < got; i++) { anchors[a + i].left_off += p; anchors[a + i].right_off += p; }
            a += (uint64_t)got; p += used;
        }
    begin[2 * n] = (uint32_t)a; *na = a; *np = p;
    return 0;
}
static const char* f_err(DarwinGpu*) { return "stand-in failure"; }

static bool same_seed(const DarwinSeedAnchor& x, const uint64_t* px, const DarwinSeedAnchor& y, const uint64_t* py) {
    return x.hit_offset == y.hit_offset && x.left_n == y.left_n && x.right_n == y.right_n &&
           !memcmp(px + x.left_off, py + y.left_off, 8ull * x.left_n) && !memcmp(px + x.right_off, py + y.right_off, 8ull * x.right_n);
}

static bool same_aln(const DarwinAlnRes& a, const uint8_t* oa, const DarwinAlnRes& b, const uint8_t* ob) {
    if (a.n_ops != b.n_ops || a.cells != b.cells || a.reference_start_offset != b.reference_start_offset ||
        a.reference_end_offset != b.reference_end_offset || a.query_start_offset != b.query_start_offset ||
        a.query_end_offset != b.query_end_offset || a.n_tiles != b.n_tiles || a.score != b.score || a.flags != b.flags) return false;
    if ((a.flags & DARWIN_ALN_EMITTED) && memcmp(oa + a.ops_offset, ob + b.ops_offset, a.n_ops)) return false;
    return true;
}

extern "C" int combiner_selftest(const DarwinScoring* s, const char* dram, uint64_t dram_bytes,
                                 const DarwinTileReq* req, int n_req, const DarwinFilterCand* cands, int n_cands,
                                 const DarwinAnchor* anchors, int n_anchors, const uint64_t* pool, uint64_t n_pool,
                                 int T, int O, int threads, uint64_t* stats_out) {
    Fake fk; gact_scoring_init(&fk.sc, s);
    std::vector<char> arena(dram_bytes, 'N');                         // the stand-in device starts empty: uploads must fill it
    fk.dram = arena.data(); fk.dram_bytes = dram_bytes; fk.device_calls = 0; fk.fail_filter = 0;
    DarwinGpu* h = reinterpret_cast<DarwinGpu*>(&fk);
    GpuCalls calls{f_upload, f_tiles, f_filter, f_extend, f_err, f_seed};
    GpuCombiner gc(h, calls);
    // a small seed position table (k = 8) over the first chromosome named by the anchors, and the reads to seed
    DsoftIndex ix;
    const uint32_t chr_start = anchors[0].chr_start, chr_len = anchors[0].ref_len;
    if (dsoft_index_build(&ix, dram, &chr_start, &chr_len, 1, chr_start + chr_len, 8, 3, 40, 64, 4)) return 4;
    fk.ix = &ix;
    std::vector<DarwinSeedRead> sreads(n_anchors);
    for (int i = 0; i < n_anchors; i++) sreads[i] = DarwinSeedRead{anchors[i].read_addr, anchors[i].read_len, 0};

    // ground truth: one direct call each on the real arena
    Fake truth; truth.sc = fk.sc; truth.dram = const_cast<char*>(dram); truth.dram_bytes = dram_bytes; truth.device_calls = 0; truth.fail_filter = 0; truth.ix = &ix;
    DarwinGpu* ht = reinterpret_cast<DarwinGpu*>(&truth);
    const int words = 50;
    std::vector<DarwinTileRes> t_res(n_req); std::vector<uint64_t> t_tb((size_t)n_req * words);
    if (f_tiles(ht, 1, req, n_req, t_res.data(), t_tb.data(), words)) return 1;
    DarwinFilterParams fp{128, 60, 1000, 0};
    std::vector<DarwinFilterRes> f_res(n_cands);
    if (f_filter(ht, &fp, cands, n_cands, f_res.data())) return 2;
    DarwinExtendParams ep{T, O, 0, 0};
    std::vector<DarwinAlnRes> a_res(n_anchors); uint64_t cap = 65536;
    for (int i = 0; i < n_anchors; i++) cap += 3ull * anchors[i].read_len;
    std::vector<uint8_t> a_ops(cap);
    if (f_extend(ht, &ep, anchors, n_anchors, pool, n_pool, a_res.data(), a_ops.data(), cap)) return 3;

    std::vector<uint32_t> s_begin(2 * n_anchors + 1); std::vector<DarwinSeedAnchor> s_anc(1 << 14); std::vector<uint64_t> s_pool(1 << 22);
    uint64_t s_na = 0, s_np = 0;
    if (f_seed(ht, sreads.data(), n_anchors, s_begin.data(), s_anc.data(), s_anc.size(), &s_na, s_pool.data(), s_pool.size(), &s_np)) return 5;
    if (s_na < (uint64_t)n_anchors) return 6;                    // the stand-in really finds the reads

    // the whole arena travels as upload spans attached to the FIRST request of every thread (in 3 pieces)
    std::atomic<int> bad(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        th.emplace_back([&, t] {
            std::string err;
            std::vector<UploadSpan> up;
            const uint64_t third = dram_bytes / 3;
            up.push_back(UploadSpan{0, dram, third}); up.push_back(UploadSpan{third, dram + third, third});
            up.push_back(UploadSpan{2 * third, dram + 2 * third, dram_bytes - 2 * third});
            if (gc.upload(up, &err)) { bad |= 1; return; }
            const std::vector<UploadSpan> none;
            for (int round = 0; round < 3; round++) {
                // this thread's slice of every kind of work
                {   const int lo = (int)((int64_t)n_cands * t / threads), hi = (int)((int64_t)n_cands * (t + 1) / threads);
                    std::vector<DarwinFilterRes> r(hi - lo);
                    if (gc.filter(fp, none, cands + lo, hi - lo, r.data(), &err)) bad |= 2;
                    else if (hi > lo && memcmp(r.data(), f_res.data() + lo, sizeof(DarwinFilterRes) * (hi - lo))) bad |= 4; }
                {   const int lo = (int)((int64_t)n_anchors * t / threads), hi = (int)((int64_t)n_anchors * (t + 1) / threads);
                    std::vector<DarwinAlnRes> r(hi - lo); std::vector<uint8_t> ops;
                    if (gc.extend(ep, none, anchors + lo, hi - lo, pool, n_pool, r.data(), &ops, &err)) bad |= 8;
                    else for (int i = lo; i < hi; i++) if (!same_aln(r[i - lo], ops.data(), a_res[i], a_ops.data())) bad |= 16; }
                {   const int lo = (int)((int64_t)n_anchors * t / threads), hi = (int)((int64_t)n_anchors * (t + 1) / threads);
                    std::vector<uint32_t> b; std::vector<DarwinSeedAnchor> a; std::vector<uint64_t> pl;
                    if (gc.seed(none, sreads.data() + lo, hi - lo, &b, &a, &pl, &err)) bad |= 256;
                    else {
                        if (b.size() != 2 * (size_t)(hi - lo) + 1) bad |= 512;
                        else for (int r = lo; r < hi; r++) for (int st = 0; st < 2; st++) {
                            const uint32_t g0 = s_begin[2 * r + st], g1 = s_begin[2 * r + st + 1], m0 = b[2 * (r - lo) + st], m1 = b[2 * (r - lo) + st + 1];
                            if (g1 - g0 != m1 - m0) { bad |= 512; continue; }
                            for (uint32_t i = 0; i < g1 - g0; i++) if (!same_seed(a[m0 + i], pl.data(), s_anc[g0 + i], s_pool.data())) bad |= 1024;
                        }
                    } }
                {   const int lo = (int)((int64_t)n_req * t / threads), hi = (int)((int64_t)n_req * (t + 1) / threads);
                    const int w = words - (t % 3);                      // callers may size their TB rows differently
                    std::vector<DarwinTileRes> r(hi - lo); std::vector<uint64_t> tb((size_t)(hi - lo) * w + 1);
                    if (gc.tiles(1, req + lo, hi - lo, r.data(), tb.data(), w, &err)) bad |= 32;
                    else for (int i = lo; i < hi; i++) {
                        if (memcmp(&r[i - lo], &t_res[i], sizeof(DarwinTileRes))) bad |= 64;
                        const int nw = (t_res[i].total_TB_pointers + 31) / 32;
                        if (memcmp(tb.data() + (size_t)(i - lo) * w, t_tb.data() + (size_t)i * words, sizeof(uint64_t) * nw)) bad |= 128;
                    } }
            }
        });
    }
    for (auto& x : th) x.join();
    CombinerStats st = gc.stats();
    for (int k = 0; k < 3; k++) { stats_out[k] = st.device_calls[k]; stats_out[3 + k] = st.requests[k]; stats_out[6 + k] = st.max_merged[k]; }
    stats_out[9] = st.device_calls[3]; stats_out[10] = st.requests[3]; stats_out[11] = st.max_merged[3];
    dsoft_index_free(&ix);
    if (bad.load()) return 100 + bad.load();
    // errors reach every merged caller
    fk.fail_filter = 1;
    std::atomic<int> errs(0);
    std::vector<std::thread> th2;
    for (int t = 0; t < threads; t++) th2.emplace_back([&] {
        std::string err; std::vector<DarwinFilterRes> r(n_cands); const std::vector<UploadSpan> none;
        if (gc.filter(fp, none, cands, n_cands, r.data(), &err) == DARWIN_ERR_CUDA && err.find("stand-in failure") != std::string::npos) errs++;
    });
    for (auto& x : th2) x.join();
    if (errs.load() != threads) return 50;
    // an exception inside the combining thread (host allocation failure while merging) fails the batch's callers and
    // leaves the combiner usable: nobody waits forever on a lane that stayed busy
    fk.fail_filter = 2;
    std::atomic<int> thrown(0);
    std::vector<std::thread> th3;
    for (int t = 0; t < threads; t++) th3.emplace_back([&] {
        std::string err; std::vector<DarwinFilterRes> r(n_cands); const std::vector<UploadSpan> none;
        const int rc = gc.filter(fp, none, cands, n_cands, r.data(), &err);
        if (rc == DARWIN_ERR_CAPACITY && err.find("combiner:") != std::string::npos) thrown++;
    });
    for (auto& x : th3) x.join();
    if (thrown.load() != threads) return 51;
    fk.fail_filter = 0;
    {   std::string err; std::vector<DarwinFilterRes> r(n_cands); const std::vector<UploadSpan> none;
        if (gc.filter(fp, none, cands, n_cands, r.data(), &err)) return 52;
        if (n_cands && memcmp(r.data(), f_res.data(), sizeof(DarwinFilterRes) * n_cands)) return 53; }
    return 0;
}

// ---- ALIGN requests (darwin_gpu_align_reads behind gpu_align_body / gpu_sam_body): merge + scatter with a synthetic
// stand-in whose output depends only on a read's arena address, laid out like the library's (forward-strand locations
// sorted by read, then the reverse-strand ones; dense op strings in that order).
static int n_fwd(uint64_t addr) { return (int)((addr / 128) % 3); }
static int n_rev(uint64_t addr) { return (int)((addr / 128 + 1) % 2); }
static uint32_t loc_ops(uint64_t addr, int strand, int j) { return (j == 1 && strand == 0) ? 0u : (uint32_t)(addr % 97) + 5u * (uint32_t)j + 1u; }
static uint8_t op_byte(uint64_t addr, int strand, int j, uint32_t p) { return (uint8_t)((addr / 128 + 7u * (uint32_t)strand + 3u * (uint32_t)j + p) % 3 + 1); }
static int f_align(DarwinGpu* h, const DarwinAlignParams*, const DarwinSeedRead* reads, int n, DarwinAnchor* an, DarwinAlnRes* res, uint64_t cap,
                   uint64_t* n_out, uint8_t* ops, uint64_t ops_cap) {
    fake(h)->device_calls++; device_latency();
    uint64_t total = 0, bytes = 0;
    for (int r = 0; r < n; r++) total += (uint64_t)(n_fwd(reads[r].read_addr) + n_rev(reads[r].read_addr));
    *n_out = total;
    if (total > cap) return DARWIN_ERR_CAPACITY;
    uint64_t at = 0;
    for (int strand = 0; strand < 2; strand++)
        for (int r = 0; r < n; r++) {
            const uint64_t addr = reads[r].read_addr;
            const int cnt = strand ? n_rev(addr) : n_fwd(addr);
            for (int j = 0; j < cnt; j++, at++) {
                an[at] = DarwinAnchor{}; res[at] = DarwinAlnRes{};
                an[at].read_addr = addr; an[at].read_len = reads[r].read_len; an[at].read_num = r; an[at].strand = (uint8_t)strand;
                an[at].reference_pos = (uint32_t)addr + (uint32_t)j;
                const uint32_t no = loc_ops(addr, strand, j);
                res[at].flags = no ? DARWIN_ALN_EMITTED : 0; res[at].n_ops = no; res[at].ops_offset = bytes; res[at].score = (int32_t)(addr % 1000) + j;
                if (bytes + no > ops_cap) return DARWIN_ERR_CAPACITY;
                for (uint32_t p = 0; p < no; p++) ops[bytes + p] = op_byte(addr, strand, j, p);
                bytes += no;
            }
        }
    return 0;
}

extern "C" int combiner_align_selftest(int threads, int reads_per_request, int rounds, uint64_t* stats_out) {
    Fake fk; fk.dram = nullptr; fk.dram_bytes = 0; fk.device_calls = 0; fk.fail_filter = 0; fk.ix = nullptr;
    DarwinGpu* h = reinterpret_cast<DarwinGpu*>(&fk);
    GpuCalls calls{f_upload, f_tiles, f_filter, f_extend, f_err, f_seed, f_align};
    GpuCombiner gc(h, calls);
    DarwinAlignParams prm{};
    prm.extend.tile_size = 384; prm.extend.tile_overlap = 64;
    std::atomic<int> bad(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        th.emplace_back([&, t] {
            const std::vector<UploadSpan> none;
            for (int round = 0; round < rounds; round++) {
                const int n = (t == 1 && round == 1) ? 0 : reads_per_request + (t % 3);       // one caller comes with no reads at all
                std::vector<DarwinSeedRead> reads(n);
                for (int i = 0; i < n; i++) reads[i] = DarwinSeedRead{128ull * (uint64_t)(1 + i + 1000 * t + 100000 * round), 500u + (uint32_t)i, 0};
                std::vector<DarwinAnchor> an; std::vector<DarwinAlnRes> res; std::vector<uint8_t> ops;
                std::string err;
                if (gc.align(prm, none, reads.data(), n, &an, &res, &ops, &err)) { bad |= 1; continue; }
                size_t at = 0; uint64_t bytes = 0;
                for (int strand = 0; strand < 2; strand++)
                    for (int i = 0; i < n; i++) {
                        const uint64_t addr = reads[i].read_addr;
                        const int cnt = strand ? n_rev(addr) : n_fwd(addr);
                        for (int j = 0; j < cnt; j++, at++) {
                            if (at >= an.size() || at >= res.size()) { bad |= 2; continue; }
                            const uint32_t no = loc_ops(addr, strand, j);
                            if (an[at].read_num != i || an[at].read_addr != addr || an[at].strand != strand || an[at].reference_pos != (uint32_t)addr + (uint32_t)j) bad |= 4;
                            if (res[at].n_ops != no || res[at].score != (int32_t)(addr % 1000) + j || (no != 0) != ((res[at].flags & DARWIN_ALN_EMITTED) != 0)) bad |= 8;
                            if (no) {
                                if (res[at].ops_offset != bytes || bytes + no > ops.size()) { bad |= 16; continue; }
                                for (uint32_t p = 0; p < no; p++) if (ops[bytes + p] != op_byte(addr, strand, j, p)) bad |= 32;
                                bytes += no;
                            }
                        }
                    }
                if (at != an.size() || at != res.size() || bytes != ops.size()) bad |= 64;
            }
        });
    }
    for (auto& x : th) x.join();
    const CombinerStats st = gc.stats();
    stats_out[0] = st.device_calls[4]; stats_out[1] = st.requests[4]; stats_out[2] = st.max_merged[4]; stats_out[3] = st.items[4];
    return bad.load() ? 100 + bad.load() : 0;
}

// Cross-read batching for the GPU stages ("flat combining" / group commit).
//
// The reference runs one read batch per TBB token (main.cpp:590-702: the reader emits a batch as soon as it holds more
// than readBufferLimit = 64 bytes, i.e. ONE read; cfg.num_threads tokens are in flight, main.cpp:615-624) and every
// stage body talks to the Processor on its own.  A GPU wants thousands of anchors per launch, not the one or two of a
// single read, and a DarwinGpu handle is single-threaded.  GpuCombiner sits between the stage bodies and the C-ABI:
// every host thread submits its own request (first-tile candidates, anchors, tiles -- plus the read spans that must be
// resident first) and blocks; whichever thread finds the device idle becomes the combiner, concatenates EVERYTHING
// queued so far into one darwin_gpu_filter / darwin_gpu_extend / darwin_gpu_tiles call, scatters the results and wakes
// the owners.  While that call runs the other threads keep seeding and queue up behind it, so the batch size adapts
// itself to the device latency; no timers, no extra thread.
//
// Only include/darwin_gpu.h is needed here (no reference headers), so the merge / scatter logic is unit-tested on the
// CPU with stand-in entry points (tests/cpp/test_combiner.cpp).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <mutex>
#include <string>
#include <vector>

#include "darwin_gpu.h"

namespace darwin_gpu_host {

// The C-ABI entry points the combiner drives; tests inject CPU stand-ins.
struct GpuCalls {
    int (*upload)(DarwinGpu*, uint64_t, const char*, uint64_t);
    int (*tiles)(DarwinGpu*, int, const DarwinTileReq*, int, DarwinTileRes*, uint64_t*, int);
    int (*filter)(DarwinGpu*, const DarwinFilterParams*, const DarwinFilterCand*, int, DarwinFilterRes*);
    int (*extend)(DarwinGpu*, const DarwinExtendParams*, const DarwinAnchor*, int, const uint64_t*, uint64_t,
                  DarwinAlnRes*, uint8_t*, uint64_t);
    const char* (*last_error)(DarwinGpu*);
    int (*seed)(DarwinGpu*, const DarwinSeedRead*, int, uint32_t*, DarwinSeedAnchor*, uint64_t, uint64_t*, uint64_t*, uint64_t, uint64_t*);
    int (*align)(DarwinGpu*, const DarwinAlignParams*, const DarwinSeedRead*, int, DarwinAnchor*, DarwinAlnRes*, uint64_t, uint64_t*,
                 uint8_t*, uint64_t);
    void* (*host_alloc)(uint64_t);    // page-locked buffers for the merged op strings (nullptr: plain malloc)
    void (*host_free)(void*);
    int (*upload_spans)(DarwinGpu*, const DarwinSpan*, int);   // all uploads of a merged call at once (nullptr: one by one)
    static GpuCalls library() {
        return GpuCalls{darwin_gpu_upload, darwin_gpu_tiles, darwin_gpu_filter, darwin_gpu_extend, darwin_gpu_last_error, darwin_gpu_seed,
                        darwin_gpu_align_reads, darwin_gpu_host_alloc, darwin_gpu_host_free, darwin_gpu_upload_spans};
    }
};

struct UploadSpan { uint64_t arena_addr; const char* ascii; uint64_t n; };

struct CombinerStats {
    uint64_t device_calls[5];   // [0] tiles, [1] filter, [2] extend, [3] seed, [4] align: calls that reached the device
    uint64_t requests[5];       // requests submitted by host threads
    uint64_t items[5];          // tiles / candidates / anchors / reads / reads
    uint64_t max_merged[5];     // largest number of requests served by one device call
    uint64_t phase_ns[3];       // where the combining threads' time went: [0] uploads, [1] inside the device call, [2] merging + scattering
};

class GpuCombiner {
public:
    GpuCombiner(DarwinGpu* h, const GpuCalls& calls) : h_(h), c_(calls) { memset(&st_, 0, sizeof(st_)); }
    ~GpuCombiner() { if (ops_buf_) { if (c_.host_free) c_.host_free(ops_buf_); else free(ops_buf_); } }
    GpuCombiner(const GpuCombiner&) = delete;
    GpuCombiner& operator=(const GpuCombiner&) = delete;

    // == g_BatchAlignmentSIMD for one caller; merged with other callers using the same do_traceback
    int tiles(int do_traceback, const DarwinTileReq* req, int n, DarwinTileRes* res, uint64_t* tb_words, int tb_words_per_req,
              std::string* err) {
        Request r; r.kind = TILES; r.do_tb = do_traceback; r.treq = req; r.n = n; r.tres = res; r.tb = tb_words; r.words = tb_words_per_req;
        return run(r, err);
    }
    // == the tile part of filter_body for one caller's candidates
    int filter(const DarwinFilterParams& p, const std::vector<UploadSpan>& up, const DarwinFilterCand* cands, int n,
               DarwinFilterRes* res, std::string* err) {
        Request r; r.kind = FILTER; r.fp = p; r.up = &up; r.cands = cands; r.n = n; r.fres = res;
        return run(r, err);
    }
    // == extender_body for one caller's anchors; `ops` receives this caller's op strings, res[i].ops_offset index into it
    int extend(const DarwinExtendParams& p, const std::vector<UploadSpan>& up, const DarwinAnchor* anchors, int n,
               const uint64_t* pool, uint64_t n_pool, DarwinAlnRes* res, std::vector<uint8_t>* ops, std::string* err) {
        Request r; r.kind = EXTEND; r.ep = p; r.up = &up; r.anchors = anchors; r.n = n; r.pool = pool; r.n_pool = n_pool;
        r.ares = res; r.ops = ops;
        return run(r, err);
    }
    // == seeder_body for one caller's reads: anchors of read r, strand s are anchors[begin[2r+s] .. begin[2r+s+1]),
    // their chained hits index into `pool`
    int seed(const std::vector<UploadSpan>& up, const DarwinSeedRead* reads, int n, std::vector<uint32_t>* begin,
             std::vector<DarwinSeedAnchor>* anchors, std::vector<uint64_t>* pool, std::string* err) {
        Request r; r.kind = SEED; r.up = &up; r.sreads = reads; r.n = n; r.sbegin = begin; r.sanchors = anchors; r.spool = pool;
        return run(r, err);
    }
    // == seeder_body + filter_body + extender_body for one caller's reads (darwin_gpu_align_reads); anchors[i].read_num
    // indexes the caller's reads, res[i].ops_offset its `ops`
    int align(const DarwinAlignParams& p, const std::vector<UploadSpan>& up, const DarwinSeedRead* reads, int n,
              std::vector<DarwinAnchor>* anchors, std::vector<DarwinAlnRes>* res, std::vector<uint8_t>* ops, std::string* err) {
        Request r; r.kind = ALIGN; r.ap = p; r.up = &up; r.sreads = reads; r.n = n; r.aanchors = anchors; r.ares_v = res; r.ops = ops;
        return run(r, err);
    }
    // == g_InitializeReferenceMemory / g_InitializeReadMemory from any thread (rides along with the next tile batch)
    int upload(const std::vector<UploadSpan>& up, std::string* err) {
        Request r; r.kind = TILES; r.do_tb = 0; r.up = &up; r.n = 0;
        return run(r, err);
    }
    CombinerStats stats() {
        std::lock_guard<std::mutex> g(m_);
        CombinerStats s = st_;
        for (int k = 0; k < 3; k++) s.phase_ns[k] = phase_ns_[k].load();
        return s;
    }
    DarwinGpu* handle() const { return h_; }

private:
    enum Kind { TILES = 0, FILTER = 1, EXTEND = 2, SEED = 3, ALIGN = 4 };
    // merged items per device call: tiles, candidates, anchors, reads (seeding), reads (resident pipeline)
    static uint64_t max_merged_items(int kind) { return kind <= 1 ? (1u << 22) : kind == 2 ? (1u << 18) : (1u << 15); }
    struct Request {
        Kind kind; int n = 0; bool done = false; int rc = 0; std::string err;
        const std::vector<UploadSpan>* up = nullptr;
        // tiles
        int do_tb = 0, words = 0; const DarwinTileReq* treq = nullptr; DarwinTileRes* tres = nullptr; uint64_t* tb = nullptr;
        // filter
        DarwinFilterParams fp{}; const DarwinFilterCand* cands = nullptr; DarwinFilterRes* fres = nullptr;
        // extend
        DarwinExtendParams ep{}; const DarwinAnchor* anchors = nullptr; const uint64_t* pool = nullptr; uint64_t n_pool = 0;
        DarwinAlnRes* ares = nullptr; std::vector<uint8_t>* ops = nullptr;
        // seed
        const DarwinSeedRead* sreads = nullptr; std::vector<uint32_t>* sbegin = nullptr;
        std::vector<DarwinSeedAnchor>* sanchors = nullptr; std::vector<uint64_t>* spool = nullptr;
        // align
        DarwinAlignParams ap{}; std::vector<DarwinAnchor>* aanchors = nullptr; std::vector<DarwinAlnRes>* ares_v = nullptr;
    };

    static bool mergeable(const Request& a, const Request& b) {
        if (a.kind != b.kind) return false;
        if (a.kind == TILES) return a.do_tb == b.do_tb;
        if (a.kind == FILTER) return memcmp(&a.fp, &b.fp, sizeof(a.fp)) == 0;
        if (a.kind == SEED) return true;
        if (a.kind == ALIGN) return memcmp(&a.ap, &b.ap, sizeof(a.ap)) == 0;
        return memcmp(&a.ep, &b.ep, sizeof(a.ep)) == 0;
    }

    int run(Request& r, std::string* err) {
        std::unique_lock<std::mutex> lk(m_);
        st_.requests[r.kind]++; st_.items[r.kind] += (uint64_t)r.n;
        q_.push_back(&r);
        while (!r.done) {
            if (busy_) { cv_.wait(lk); continue; }
            // become the combiner: take the oldest request and everything queued that can ride along with it
            busy_ = true;
            std::vector<Request*> batch;
            batch.push_back(q_.front()); q_.pop_front();
            // ... up to a bounded number of items per device call: the library indexes the seed hits of one call with
            // 32 bits (dsoft_host.cuh), and a call that large gains nothing from growing further
            uint64_t items = (uint64_t)batch[0]->n;
            for (auto it = q_.begin(); it != q_.end();) {
                if (mergeable(*batch[0], **it) && items + (uint64_t)(*it)->n <= max_merged_items(batch[0]->kind)) {
                    items += (uint64_t)(*it)->n; batch.push_back(*it); it = q_.erase(it);
                } else ++it;
            }
            lk.unlock();
            try { execute(batch); }
            catch (const std::exception& e) {                    // e.g. bad_alloc while merging: every caller of the batch fails, nobody hangs
                for (auto* b : batch) { b->rc = DARWIN_ERR_CAPACITY; b->err = std::string("combiner: ") + e.what(); }
            }
            lk.lock();
            const Kind k = batch[0]->kind;
            st_.device_calls[k]++;
            if (batch.size() > st_.max_merged[k]) st_.max_merged[k] = batch.size();
            for (auto* b : batch) b->done = true;
            busy_ = false;
            cv_.notify_all();
        }
        if (r.rc && err) *err = r.err;
        return r.rc;
    }

    // Reused landing buffer of the merged op strings (only the combining thread touches it): page-locked when the library
    // provides it, grown geometrically, never zero-filled.
    uint8_t* ops_buffer(uint64_t bytes) {
        if (bytes <= ops_cap_) return ops_buf_;
        if (ops_buf_) { if (c_.host_free) c_.host_free(ops_buf_); else free(ops_buf_); ops_buf_ = nullptr; ops_cap_ = 0; }
        const uint64_t want = bytes + bytes / 2;
        ops_buf_ = (uint8_t*)(c_.host_alloc ? c_.host_alloc(want) : malloc(want));
        if (ops_buf_) ops_cap_ = want;
        return ops_buf_;
    }

    void fail_all(std::vector<Request*>& batch, int rc, const char* what) {
        const std::string msg = std::string(what) + ": " + (c_.last_error ? c_.last_error(h_) : "");
        for (auto* b : batch) { b->rc = rc; b->err = msg; }
    }

    static uint64_t now_ns() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    // accounts the time between two marks of the combining thread to one of CombinerStats::phase_ns
    struct PhaseClock {
        std::atomic<uint64_t>* acc; uint64_t last;
        explicit PhaseClock(std::atomic<uint64_t>* a) : acc(a), last(now_ns()) {}
        void mark(int phase) { const uint64_t t = now_ns(); acc[phase] += t - last; last = t; }
    };

    void execute(std::vector<Request*>& batch) {
        PhaseClock clock(phase_ns_);
        struct Rest { PhaseClock& c; ~Rest() { c.mark(2); } } rest{clock};     // whatever follows the last mark is merge / scatter work
        if (c_.upload_spans) {
            std::vector<DarwinSpan> all;
            for (auto* b : batch)
                if (b->up) for (const auto& s : *b->up) all.push_back(DarwinSpan{s.arena_addr, s.ascii, s.n});
            if (!all.empty()) {
                const int rc = c_.upload_spans(h_, all.data(), (int)all.size());
                if (rc) { fail_all(batch, rc, "darwin_gpu_upload_spans"); return; }
            }
        } else {
            for (auto* b : batch)
                if (b->up)
                    for (const auto& s : *b->up) {
                        const int rc = c_.upload(h_, s.arena_addr, s.ascii, s.n);
                        if (rc) { fail_all(batch, rc, "darwin_gpu_upload"); return; }
                    }
        }
        clock.mark(0);
        size_t total = 0;
        for (auto* b : batch) total += (size_t)b->n;
        const Kind k = batch[0]->kind;
        if (k == TILES) {
            int words = 1;
            for (auto* b : batch) if (b->words > words) words = b->words;
            const bool tb = batch[0]->do_tb != 0;
            if (batch.size() == 1) {
                Request* b = batch[0];
                b->rc = b->n ? c_.tiles(h_, b->do_tb, b->treq, b->n, b->tres, b->tb, b->words) : 0;
                if (b->rc) fail_all(batch, b->rc, "darwin_gpu_tiles");
                return;
            }
            std::vector<DarwinTileReq> req; req.reserve(total);
            for (auto* b : batch) req.insert(req.end(), b->treq, b->treq + b->n);
            std::vector<DarwinTileRes> res(total);
            std::vector<uint64_t> tbw(tb ? total * (size_t)words : 1);
            const int rc = total ? c_.tiles(h_, tb, req.data(), (int)total, res.data(), tb ? tbw.data() : nullptr, words) : 0;
            if (rc) { fail_all(batch, rc, "darwin_gpu_tiles"); return; }
            // the merged call ran with the widest TB row of the batch: a caller whose own row is narrower than one of its
            // tiles' tracebacks gets what the unmerged call would have given it (status 2, DARWIN_ERR_CAPACITY)
            size_t at = 0;
            for (auto* b : batch) {
                for (int i = 0; i < b->n; i++) {
                    b->tres[i] = res[at + i];
                    if (tb) {
                        memcpy(b->tb + (size_t)i * b->words, tbw.data() + (at + i) * (size_t)words, sizeof(uint64_t) * (size_t)b->words);
                        if (((size_t)res[at + i].total_TB_pointers + 31) / 32 > (size_t)b->words) {
                            b->tres[i].status = (uint8_t)((b->tres[i].status & 0xF0) | 2);
                            b->rc = DARWIN_ERR_CAPACITY; b->err = "darwin_gpu_tiles: tb_words_per_req too small";
                        }
                    }
                }
                at += (size_t)b->n;
            }
        } else if (k == FILTER) {
            if (batch.size() == 1) {
                Request* b = batch[0];
                b->rc = c_.filter(h_, &b->fp, b->cands, b->n, b->fres);
                if (b->rc) fail_all(batch, b->rc, "darwin_gpu_filter");
                return;
            }
            std::vector<DarwinFilterCand> cands; cands.reserve(total);
            for (auto* b : batch) cands.insert(cands.end(), b->cands, b->cands + b->n);
            std::vector<DarwinFilterRes> res(total);
            const int rc = c_.filter(h_, &batch[0]->fp, cands.data(), (int)total, res.data());
            if (rc) { fail_all(batch, rc, "darwin_gpu_filter"); return; }
            size_t at = 0;
            for (auto* b : batch) { if (b->n) memcpy(b->fres, res.data() + at, sizeof(DarwinFilterRes) * (size_t)b->n); at += (size_t)b->n; }
        } else if (k == ALIGN) {
            std::vector<DarwinSeedRead> reads; reads.reserve(total);
            uint64_t bases = 0;
            for (auto* b : batch) { reads.insert(reads.end(), b->sreads, b->sreads + b->n); for (int i = 0; i < b->n; i++) bases += b->sreads[i].read_len; }
            uint64_t cap = std::max<uint64_t>(64, 8 * total), ops_cap = std::max<uint64_t>(ops_cap_, 6 * bases + 65536), n_out = 0;
            std::vector<DarwinAnchor> anchors; std::vector<DarwinAlnRes> res;
            uint8_t* ops = nullptr;
            int rc = DARWIN_ERR_CAPACITY;
            for (int attempt = 0; attempt < 4 && rc == DARWIN_ERR_CAPACITY; attempt++) {
                anchors.resize(cap); res.resize(cap);
                ops = ops_buffer(ops_cap);
                if (!ops) { fail_all(batch, DARWIN_ERR_CAPACITY, "host buffer for the op strings"); return; }
                clock.mark(2);
                rc = c_.align(h_, &batch[0]->ap, reads.data(), (int)total, anchors.data(), res.data(), cap, &n_out, ops, ops_cap);
                clock.mark(1);
                if (rc == DARWIN_ERR_CAPACITY) { if (n_out > cap) cap = n_out; else ops_cap *= 2; }
            }
            if (rc) { fail_all(batch, rc, "darwin_gpu_align_reads"); return; }
            // every caller's locations are contiguous inside the forward part and inside the reverse part (sorted by read)
            // (one pass over the locations finds each caller's share, so that its vectors are sized once: the op strings are
            // ~11 kB per location and this copy runs while the lane's device is not being fed)
            std::vector<size_t> first(batch.size() + 1, 0);
            for (size_t k2 = 0; k2 < batch.size(); k2++) first[k2 + 1] = first[k2] + (size_t)batch[k2]->n;
            std::vector<uint64_t> n_loc(batch.size(), 0), n_ops(batch.size(), 0);
            std::vector<uint32_t> owner(n_out);
            for (uint64_t i = 0; i < n_out; i++) {
                const size_t rn = (size_t)anchors[i].read_num;
                owner[i] = UINT32_MAX;
                if (anchors[i].read_num < 0 || rn >= total) continue;            // not a read of this call: nobody's
                const size_t k2 = (size_t)(std::upper_bound(first.begin(), first.end(), rn) - first.begin()) - 1;
                owner[i] = (uint32_t)k2;
                n_loc[k2]++;
                const DarwinAlnRes& r = res[i];
                if ((r.flags & DARWIN_ALN_EMITTED) && !(r.flags & DARWIN_ALN_OPS_OVERFLOW)) n_ops[k2] += r.n_ops;
            }
            for (size_t k2 = 0; k2 < batch.size(); k2++) {
                Request* b = batch[k2];
                b->aanchors->clear(); b->ares_v->clear(); b->ops->clear();
                b->aanchors->reserve(n_loc[k2]); b->ares_v->reserve(n_loc[k2]); b->ops->reserve(n_ops[k2]);
            }
            for (uint64_t i = 0; i < n_out; i++) {                 // location order is kept per caller
                const size_t k2 = owner[i];
                if (k2 == UINT32_MAX) continue;
                Request* b = batch[k2];
                DarwinAnchor a = anchors[i]; a.read_num -= (int)first[k2];
                DarwinAlnRes r = res[i];
                const bool has = (r.flags & DARWIN_ALN_EMITTED) && !(r.flags & DARWIN_ALN_OPS_OVERFLOW) && r.n_ops;
                const uint64_t off = b->ops->size();
                if (has) b->ops->insert(b->ops->end(), ops + r.ops_offset, ops + r.ops_offset + r.n_ops);
                r.ops_offset = off;
                b->aanchors->push_back(a); b->ares_v->push_back(r);
            }
        } else if (k == SEED) {
            std::vector<DarwinSeedRead> reads; reads.reserve(total);
            for (auto* b : batch) reads.insert(reads.end(), b->sreads, b->sreads + b->n);
            std::vector<uint32_t> begin(2 * total + 1);
            std::vector<DarwinSeedAnchor> anchors(std::max<size_t>(64, 16 * total));
            std::vector<uint64_t> pool(std::max<size_t>(1 << 16, 8192 * total));
            uint64_t na = 0, np = 0;
            int rc = DARWIN_ERR_CAPACITY;
            for (int attempt = 0; attempt < 3 && rc == DARWIN_ERR_CAPACITY; attempt++) {
                rc = c_.seed(h_, reads.data(), (int)total, begin.data(), anchors.data(), anchors.size(), &na, pool.data(), pool.size(), &np);
                if (rc == DARWIN_ERR_CAPACITY) { anchors.resize(std::max<size_t>(anchors.size(), na)); pool.resize(std::max<size_t>(pool.size(), np)); }
            }
            if (rc) { fail_all(batch, rc, "darwin_gpu_seed"); return; }
            size_t at = 0;                                        // reads before this caller
            for (auto* b : batch) {
                const uint32_t a0 = begin[2 * at], a1 = begin[2 * (at + (size_t)b->n)];
                uint64_t lo = UINT64_MAX, hi = 0;
                for (uint32_t i = a0; i < a1; i++) {
                    lo = std::min(lo, std::min(anchors[i].left_off, anchors[i].right_off));
                    hi = std::max(hi, std::max(anchors[i].left_off + anchors[i].left_n, anchors[i].right_off + anchors[i].right_n));
                }
                if (lo == UINT64_MAX) { lo = 0; hi = 0; }
                b->sbegin->resize(2 * (size_t)b->n + 1);
                for (size_t i = 0; i <= 2 * (size_t)b->n; i++) (*b->sbegin)[i] = begin[2 * at + i] - a0;
                b->sanchors->assign(anchors.begin() + a0, anchors.begin() + a1);
                for (auto& a : *b->sanchors) { a.left_off -= lo; a.right_off -= lo; }
                b->spool->assign(pool.begin() + lo, pool.begin() + hi);
                at += (size_t)b->n;
            }
        } else {
            // anchors of all callers back to back; their hit lists are rebased into one pool
            std::vector<DarwinAnchor> anchors; anchors.reserve(total);
            std::vector<uint64_t> pool;
            uint64_t cap = 65536;
            for (auto* b : batch) {
                const uint64_t base = pool.size();
                if (b->n_pool) pool.insert(pool.end(), b->pool, b->pool + b->n_pool);
                for (int i = 0; i < b->n; i++) {
                    DarwinAnchor a = b->anchors[i];
                    a.left_hits_off += (uint32_t)base; a.right_hits_off += (uint32_t)base;
                    cap += 3ull * a.read_len;
                    anchors.push_back(a);
                }
            }
            std::vector<DarwinAlnRes> res(total);
            uint8_t* ops = nullptr;
            int rc = DARWIN_ERR_CAPACITY;
            for (int attempt = 0; attempt < 4 && rc == DARWIN_ERR_CAPACITY; attempt++) {     // like ALIGN: grow the op pool and retry
                ops = ops_buffer(cap);
                if (!ops) { fail_all(batch, DARWIN_ERR_CAPACITY, "host buffer for the op strings"); return; }
                clock.mark(2);
                rc = c_.extend(h_, &batch[0]->ep, anchors.data(), (int)total, pool.empty() ? nullptr : pool.data(), pool.size(),
                               res.data(), ops, cap);
                clock.mark(1);
                if (rc == DARWIN_ERR_CAPACITY) cap *= 2;
            }
            if (rc) { fail_all(batch, rc, "darwin_gpu_extend"); return; }
            // op strings are dense and in anchor order: each caller owns one contiguous slice of the pool
            size_t at = 0;
            for (auto* b : batch) {
                uint64_t lo = UINT64_MAX, hi = 0;
                for (int i = 0; i < b->n; i++) {
                    const DarwinAlnRes& r = res[at + i];
                    if (!(r.flags & DARWIN_ALN_EMITTED) || (r.flags & DARWIN_ALN_OPS_OVERFLOW) || r.n_ops == 0) continue;
                    if (r.ops_offset < lo) lo = r.ops_offset;
                    if (r.ops_offset + r.n_ops > hi) hi = r.ops_offset + r.n_ops;
                }
                if (lo == UINT64_MAX) { lo = 0; hi = 0; }
                if (b->ops) b->ops->assign(ops + lo, ops + hi);
                for (int i = 0; i < b->n; i++) {
                    b->ares[i] = res[at + i];
                    b->ares[i].ops_offset = (res[at + i].ops_offset >= lo) ? res[at + i].ops_offset - lo : 0;
                }
                at += (size_t)b->n;
            }
        }
    }

    DarwinGpu* h_;
    GpuCalls c_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Request*> q_;
    bool busy_ = false;
    CombinerStats st_;
    std::atomic<uint64_t> phase_ns_[3] = {};
    uint8_t* ops_buf_ = nullptr; uint64_t ops_cap_ = 0;
};

} // namespace darwin_gpu_host

// Output stage of the GACT path, host side of libdarwin_gact.so: what printer_body needs from an alignment, computed from
// the op string the extension kernels return instead of from two gapped strings of ~11 k characters per alignment.
//
//   darwin_gpu_cigar       == the CIGAR construction of printer_body::AlignmentToSam (software/printer.cpp:236-301)
//   darwin_gpu_sam_select  == the ordering + overlap suppression of printer_body::sam_printer (printer.cpp:15-47)
//
// Plain C++ (no device code): these run on the host thread that received the results of darwin_gpu_extend /
// darwin_gpu_align_reads.  The SAM line itself is assembled by the host adapter (darwin_b200/host, gpu_sam_body).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/darwin_gpu.h"

namespace {

// appends "<count><op>" at out[pos...]; returns false when it does not fit (pos keeps counting so the caller learns the size).
// Hand-rolled decimal conversion: a 10 kbp alignment at 15 % error has ~3 000 runs, and snprintf per run was the largest
// host cost of the SAM stage.
inline bool put_run(char* out, uint64_t cap, uint64_t& pos, uint64_t count, char op) {
    char buf[24];
    int n = 0;
    do { buf[n++] = (char)('0' + count % 10); count /= 10; } while (count);
    const bool fits = pos + (uint64_t)n + 1 <= cap;
    if (fits && out) {
        char* w = out + pos;
        for (int k = n - 1; k >= 0; k--) *w++ = buf[k];
        *w = op;
    }
    pos += (uint64_t)n + 1;
    return fits;
}

}  // namespace

extern "C" {

int darwin_gpu_cigar(const DarwinAlnRes* r, const uint8_t* ops_pool, uint32_t query_length, char* out, uint64_t cap, uint64_t* len) {
    if (!r || !len || (r->n_ops && !ops_pool)) return DARWIN_ERR_INVALID;
    uint64_t pos = 0;
    bool ok = true;
    if (r->query_start_offset > 0) ok &= put_run(out, cap, pos, r->query_start_offset, 'S');             // printer.cpp:245-253
    // one CIGAR run per maximal run of equal ops: reference '-' (op I) -> 'I', query '-' (op D) -> 'D', else 'M' (:257-297)
    const uint8_t* o = ops_pool + r->ops_offset;
    static const char kOp[4] = {'M', 'I', 'D', 'M'};
    uint32_t p = 0;
    while (p < r->n_ops) {
        const uint8_t d = o[p];
        uint32_t q = p + 1;
        while (q < r->n_ops && o[q] == d) q++;
        ok &= put_run(out, cap, pos, q - p, kOp[d & 3]);
        p = q;
    }
    // size_t arithmetic in the reference (:299): query_length - query_end_offset - 1
    const uint64_t tail = (uint64_t)query_length - (uint64_t)r->query_end_offset - 1ull;
    if (tail > 0) ok &= put_run(out, cap, pos, tail, 'S');
    if (pos == 0) { ok = cap >= 1; if (ok && out) out[0] = '*'; pos = 1; }                               // :310
    *len = pos;
    return ok ? DARWIN_OK : DARWIN_ERR_CAPACITY;
}

int darwin_gpu_sam_select(const DarwinAnchor* anchors, const DarwinAlnRes* res, uint64_t n, uint32_t* order, uint8_t* keep,
                          uint64_t* n_order) try {
    if (!n_order || (n && (!anchors || !res || !order || !keep)) || n > 0xFFFFFFFFull) return DARWIN_ERR_INVALID;
    std::vector<uint32_t> idx;
    idx.reserve(n);
    for (uint64_t i = 0; i < n; i++) if (res[i].flags & DARWIN_ALN_EMITTED) idx.push_back((uint32_t)i);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {                              // printer.cpp:18-21
        return (anchors[a].read_num < anchors[b].read_num) || ((anchors[a].read_num == anchors[b].read_num) && (res[a].score > res[b].score));
    });
    const size_t m = idx.size();
    std::fill(keep, keep + m, (uint8_t)1);
    for (size_t i = 0; i < m; i++) {                                                                     // printer.cpp:23-47
        if (!keep[i]) continue;
        const uint32_t s1 = res[idx[i]].query_start_offset, e1 = res[idx[i]].query_end_offset;
        for (size_t j = i + 1; j < m; j++) {
            if (!keep[j]) continue;
            if (anchors[idx[j]].read_num != anchors[idx[i]].read_num) break;
            const uint32_t s2 = res[idx[j]].query_start_offset, e2 = res[idx[j]].query_end_offset;
            const uint32_t s = std::max(s1, s2), e = std::min(e1, e2);
            const uint32_t overlap = e > s ? e - s : 0;
            if (2 * overlap > (e2 - s2)) keep[j] = 0;                                                    // uint32 arithmetic, as the reference
        }
    }
    std::copy(idx.begin(), idx.end(), order);
    *n_order = m;
    return DARWIN_OK;
} catch (const std::exception&) { return DARWIN_ERR_NOMEM; }

}  // extern "C"

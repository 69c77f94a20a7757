// GACT anchor state machine on the device: one persistent warp owns one anchor and walks its tiles
// in-kernel (left extension, then right extension, large-tile fallback, chained-hit popping).
//
// Follows extender_body::operator() of the reference (software/extender.cpp:45-530 forward strand,
// :557-1051 reverse strand) and makeForward/BackwardAlignment (:1067-1159); CPU twin:
// oracle/gact_oracle.c extend_one().  The traceback of each tile feeds the consumption rule
// directly (including the word-granular `break` quirk, extender.cpp:327-329), so no TB words and no
// per-tile round trip to the host exist on this path.
#pragma once
#include "gact_common.cuh"

namespace gact {

struct ExtendArgs {
    const uint8_t*  arena;
    const DarwinAnchor* anchors;
    const uint64_t* hit_pool;
    DarwinAlnRes*   res;
    uint8_t*        ops;          // device op slots
    const uint64_t* slot_base;    // per anchor: first byte of its slot
    const uint32_t* slot_left;    // per anchor: capacity of the left part (right part = slot_size - left)
    const uint32_t* slot_size;
    int n;
    int T, O, do_overlap;
    const uint32_t* order;        // work queue order: longest expected anchor first (LPT), so the last wave is short
    unsigned int* counter;        // work queue head
};

// Anchor registers shared by the whole warp (all lanes hold identical copies).
struct AnchorState {
    uint32_t cr, cq, rso, reo, qso, qeo;
    uint32_t RL, QL;
    uint64_t rsa, read_addr;
    int large, ldone, rdone, emit, rc;
    int64_t nl, nr;
    const uint64_t* lh; const uint64_t* rh;
    uint32_t nleft, nright, n_tiles, n_large, flags;
    uint64_t cells;
};

// The traceback (one lane) only records the tile's ops, 2 bits each (16 per 32-bit word, the reference's TB-word packing,
// Processor.cpp:568-582), in shared memory ...
struct SmemOpSink {
    uint32_t* wptr; int n; int cap; int overflow; uint32_t cur; int shift; bool wr;     // wr: this lane stores (lane 0)
    __device__ __forceinline__ void flush_word() {
        if (n <= cap) { if (wr) *wptr = cur; } else overflow = 1;
        wptr++; cur = 0; shift = 0;
    }
    __device__ __forceinline__ void operator()(uint32_t d) {
        cur |= d << shift;
        shift += 2; n++;
        if (shift == 32) flush_word();
    }
    __device__ __forceinline__ void run_m(int count) {          // `count` times M (0b11)
        while (count > 0) {
            const int take = min((32 - shift) >> 1, count);
            cur |= (0xFFFFFFFFu >> (32 - 2 * take)) << shift;
            shift += 2 * take; n += take; count -= take;
            if (shift == 32) flush_word();
        }
    }
    __device__ __forceinline__ int count() const { return n; }
    __device__ __forceinline__ void finish() { if (shift) { if (n <= cap) *wptr = cur; else overflow = 1; } }
};
constexpr int kOpsSmemBytes = 1024;                 // K = 0 kernels: 4096 ops >= i_steps + j_steps of the largest tile (1984 + 960)
// K > 0 kernels keep a small dedicated buffer -- as many ops as the largest single-strip tile of the geometry can produce
// (Q + R <= 128 K) -- so that 12 warps fit an SM (K <= 6: 18 096 + 192 B of dynamic + 128 B of static + 1 KB of reserved shared
// memory per warp); larger tiles (tile_size 1024, the 1984x960 stall tiles) record their ops in the big raw TMA windows of the
// warp, which are dead once the tile has been staged and which none of their paths (multi-strip, packed / unpacked exact)
// touches again
template <int K> struct OpsSmall { static constexpr int kOps = 128 * (K == 0 ? 1 : K), kBytes = kOps / 4; };

// ... and the whole warp consumes them (extender.cpp:280-331 / :427-466 and the rc twins): one 32-op TB word per
// iteration, one lane per op.  The reference's `break` leaves only the 32-op loop, so inside word w the ops up to and
// including the first M at or after step S are taken; ballots find that M, pop-counts give the number of
// reference / query bases consumed, and the taken ops are stored with one coalesced store per word.
struct ConsumeResult { uint32_t consumed, ref_steps, qry_steps; };

__device__ __forceinline__ ConsumeResult consume_ops_warp(const uint32_t* ops, int total, int S, bool left,
                                                          uint8_t* slot, uint32_t lcap, uint32_t rcap,
                                                          uint32_t nleft, uint32_t nright, uint32_t& overflow) {
    const int lane = lane_id();
    int steps = 0;
    uint32_t consumed = 0, ref_c = 0, qry_c = 0;
    for (int w0 = 0; w0 < total; w0 += 32) {
        const int k = w0 + lane;
        const bool valid = k < total;
        const uint32_t d = valid ? ((ops[k >> 4] >> (2 * (k & 15))) & 3u) : 0u;
        const int np = min(32, total - w0);
        const uint32_t mM = __ballot_sync(0xffffffffu, valid && d == DARWIN_OP_M);
        const int thr = S - steps - 1;                              // first position p with steps + p + 1 >= S
        const uint32_t cand = thr <= 0 ? mM : (thr >= 32 ? 0u : (mM & (0xFFFFFFFFu << thr)));
        const int c = cand ? __ffs(cand) : np;                      // ops taken from this word
        const uint32_t low = c >= 32 ? 0xFFFFFFFFu : ((1u << c) - 1u);
        const uint32_t mR = __ballot_sync(0xffffffffu, valid && d != DARWIN_OP_I) & low;
        const uint32_t mQ = __ballot_sync(0xffffffffu, valid && d != DARWIN_OP_D) & low;
        if (lane < c) {
            const uint32_t pos = consumed + (uint32_t)lane;
            if (left) { if (nleft + pos < lcap) slot[lcap - nleft - 1 - pos] = (uint8_t)d; else overflow = 1; }   // prepended
            else      { if (nright + pos < rcap) slot[lcap + nright + pos] = (uint8_t)d; else overflow = 1; }
        }
        steps += c; consumed += (uint32_t)c; ref_c += __popc(mR); qry_c += __popc(mQ);
    }
    overflow = __any_sync(0xffffffffu, overflow != 0) ? 1u : 0u;
    return ConsumeResult{consumed, ref_c, qry_c};
}

// Build the next tile request of an anchor (extender.cpp:58-207 / :573-722).  Returns rt/qt through refs.
__device__ __forceinline__ void next_tile(const AnchorState& a, int T, TileJob& t, int& rt, int& qt) {
    const bool left = !a.ldone;
    rt = T; qt = T;
    if (a.large) {
        const uint64_t ho = left ? a.lh[a.nl - 1] : a.rh[a.nr - 1];
        const uint64_t h1 = a.rsa + a.cr, o1 = a.cq, h2 = ho >> 32, o2 = (ho << 32) >> 32;
        const bool wide = left ? ((h1 - h2) > (o1 - o2)) : ((h2 - h1) > (o2 - o1));   // uint64 arithmetic, as the reference
        rt = wide ? 1984 : 960; qt = wide ? 960 : 1984;
    }
    if (left) {
        t.R = (int)min((uint64_t)a.cr + 1, (uint64_t)rt);
        t.Q = (int)min((uint64_t)a.cq + 1, (uint64_t)qt);
        t.ra = a.rsa + (a.cr >= (uint32_t)rt ? a.cr - rt + 1 : 0);
        const uint32_t qoff = (a.cq >= (uint32_t)qt ? a.cq - qt + 1 : 0);
        t.qa = a.rc ? a.read_addr + a.QL - t.Q - qoff : a.read_addr + qoff;
        t.flags = a.rc ? (DARWIN_REVERSE_QUERY | DARWIN_COMPLEMENT_QUERY | DARWIN_START_END) : DARWIN_START_END;
    } else {
        t.R = (int)min(a.RL - a.cr, (uint32_t)rt);
        t.Q = (int)min(a.QL - a.cq, (uint32_t)qt);
        t.ra = a.rsa + a.cr;
        t.qa = a.rc ? a.read_addr + a.QL - t.Q - a.cq : a.read_addr + a.cq;
        t.flags = a.rc ? (DARWIN_REVERSE_REF | DARWIN_COMPLEMENT_QUERY | DARWIN_START_END)
                       : (DARWIN_REVERSE_REF | DARWIN_REVERSE_QUERY | DARWIN_START_END);
    }
}

// State transition after a tile has been consumed (extender.cpp:336-394 / :472-524 and rc twins).
__device__ __forceinline__ void after_tile(AnchorState& a, int len) {
    if (!a.ldone) {
        while (a.nl > 0) {
            const uint64_t ho = a.lh[a.nl - 1], hit = ho >> 32, off = (ho << 32) >> 32;
            if (hit < a.rsa + a.cr && off < a.cq) break;
            a.nl--;
        }
        const bool stall = a.rc ? (len == 0 || a.rso == 0 || a.qso == 0)
                                : (len == 0 || a.nl == 0 || a.rso == 0 || a.qso == 0);
        if (stall) {
            if (a.large || a.nl == 0 || a.rso == 0 || a.qso == 0) {
                a.ldone = 1;
                if (a.rso > 0) a.rso = a.cr + 1;
                if (a.qso > 0) a.qso = a.cq + 1;
                if ((a.cr + 1 < a.RL) && (a.cq + 1 < a.QL) && !a.rdone) { a.cr = a.reo + 1; a.cq = a.qeo + 1; }
                else { a.rdone = 1; if (a.rc) a.emit = 1; }
            } else a.large = 1;
        } else a.large = 0;
    } else {
        while (a.nr > 0) {
            const uint64_t ho = a.rh[a.nr - 1], hit = ho >> 32, off = (ho << 32) >> 32;
            if (hit > a.rsa + a.cr && off > a.cq) break;
            a.nr--;
        }
        if (len == 0 || a.cr == a.RL || a.cq == a.QL) {
            if (a.large || a.nr == 0 || a.cr == a.RL || a.cq == a.QL) { a.reo = a.cr - 1; a.qeo = a.cq - 1; a.emit = 1; a.rdone = 1; }
            else a.large = 1;
        } else a.large = 0;
    }
}

} // namespace gact

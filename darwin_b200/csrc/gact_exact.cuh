// Exact GACT tile: one warp per tile, anti-diagonal wavefront, any shape up to 1984 x 1984.
//
// Computes the same results as the reference's AVX2 striped kernel + lazy-F loop
// (software/Processor.cpp:164-566) including its tie-breaking, using the streaming closed form
// (CPU twin: oracle/gact_oracle.c tile_scalar(exact=1)):
// lane l owns KX consecutive query rows of a 32*KX-row strip and lags lane l-1 by one column;
// the vertical chain state (true F/F_L, the striped kernel's own-lane chains f0/fl0 and carried
// chains fc/flc with their lane-distance tags) crosses lanes by __shfl_up_sync once per step.
// Strips are chained through a per-warp global boundary row.  One trace byte per cell goes to a
// per-warp global scratch; the traceback (Processor.cpp:585-716) is walked by lane 0.
//
// This is the general path: arbitrary 5x5 substitution matrix, N bases, max-cell mode (filter
// tiles) and the 1984x960 "large tiles".  The packed fast path lives in gact_fast.cuh.
#pragma once
#include "gact_common.cuh"

namespace gact {

struct ExactSmem {                       // per warp
    uint8_t  ref[kSeqSmem];              // tile reference bases after reverse/complement, codes 0..4
    uint8_t  qry[kSeqSmem];
    ChainRec rec[2][32];                 // boundary records of the previous strip, 32 columns at a time
};

// ---- TMA staging of the packed sequences -----------------------------------------------------------------------
// The 4-bit packed windows of the tile's reference and query are brought into shared memory by two bulk async copies
// (cp.async.bulk, completion on an mbarrier); the lanes then unpack / reverse / complement out of shared memory.
constexpr int kRawBig = (kMaxTile / 2 + 48 + 15) & ~15;          // bytes of one packed window of <= 1984 bases (16-byte rounded)

struct TmaStage {
    uint8_t*  raw;        // two windows, `stride` bytes apart, 16-byte aligned (shared memory)
    int       stride;
    uint64_t* mbar;       // shared-memory mbarrier of this warp
    uint32_t  phase;      // parity of the next completion
};

__device__ __forceinline__ void tma_stage_init(uint64_t* mbar) {
    if (lane_id() == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(a));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
}

// Stage one tile's sequences into shared memory (Processor.cpp:105-106 and :276-277 index rules).
// Returns (warp-uniform) whether any staged base is N.
__device__ __forceinline__ bool stage_sequences(const uint8_t* __restrict__ arena, const TileJob& t,
                                                uint8_t* sref, uint8_t* sqry, TmaStage& ts) {
    const int lane = lane_id();
    const bool rr = t.flags & DARWIN_REVERSE_REF, cr = t.flags & DARWIN_COMPLEMENT_REF;
    const bool rq = t.flags & DARWIN_REVERSE_QUERY, cq = t.flags & DARWIN_COMPLEMENT_QUERY;
    // 16-byte aligned packed windows [r0, r1) and [q0, q1) (byte offsets into the packed arena)
    const uint64_t r0 = (t.ra >> 1) & ~(uint64_t)15, r1 = ((((t.ra + (uint64_t)t.R - 1) >> 1) + 16) & ~(uint64_t)15);
    const uint64_t q0 = (t.qa >> 1) & ~(uint64_t)15, q1 = ((((t.qa + (uint64_t)t.Q - 1) >> 1) + 16) & ~(uint64_t)15);
    __syncwarp();                                                   // nobody still reads the previous tile's windows
    if (lane == 0) {
        const uint32_t mb = (uint32_t)__cvta_generic_to_shared(ts.mbar);
        const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(ts.raw), d1 = d0 + (uint32_t)ts.stride;
        const uint32_t nr = (uint32_t)(r1 - r0), nq = (uint32_t)(q1 - q0);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads of raw[] before the async writes
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mb), "r"(nr + nq) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(d0), "l"(arena + r0), "r"(nr), "r"(mb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(d1), "l"(arena + q0), "r"(nq), "r"(mb) : "memory");
    }
    {   // every lane waits for the completion of this phase (bounded: a lost copy traps instead of hanging the GPU)
        const uint32_t mb = (uint32_t)__cvta_generic_to_shared(ts.mbar);
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 24) && !done; spin++)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(mb), "r"(ts.phase) : "memory");
        if (!done) __trap();
        ts.phase ^= 1u;
    }
    const uint8_t* wr = ts.raw;
    const uint8_t* wq = ts.raw + ts.stride;
    uint32_t seen = 0;
    for (int k = lane; k < t.R; k += 32) {
        const uint64_t a = rr ? t.ra + (uint64_t)(t.R - 1 - k) : t.ra + (uint64_t)k;
        uint32_t c = (wr[(a >> 1) - r0] >> ((a & 1) * 4)) & 0xFu;
        if (cr && c < 4) c = 3 - c;
        seen |= c;
        sref[k] = (uint8_t)c;
    }
    for (int k = lane; k < t.Q; k += 32) {
        const uint64_t a = rq ? t.qa + (uint64_t)(t.Q - 1 - k) : t.qa + (uint64_t)k;
        uint32_t c = (wq[(a >> 1) - q0] >> ((a & 1) * 4)) & 0xFu;
        if (cq && c < 4) c = 3 - c;
        seen |= c;
        sqry[k] = (uint8_t)c;
    }
    const bool has_n = __any_sync(0xffffffffu, seen >= 4u);     // also the warp barrier that publishes the staging
    return has_n;
}

// Forward pass.  ssub = 25-entry substitution table in shared memory.
template <bool START_END, bool TRACE>
__device__ void exact_forward(const int* __restrict__ ssub, int go, int ge, int lgo, int lge,
                              const TileJob& t, ExactSmem* sm, const WarpScratch& ws, TileOut& out) {
    const int lane = lane_id();
    const int Q = t.Q, R = t.R;
    const int segLen = (Q + 15) >> 4;                       // Processor.cpp:174
    const int nstrips = (Q + kStrip - 1) / kStrip;
    const int steps = R + 31;
    int best = 0, bj = 0, bi = 0, corner = 0;

    for (int strip = 0; strip < nstrips; strip++) {
        const int i0 = strip * kStrip + lane * KX;
        const int nrows = min(max(Q - i0, 0), KX);
        int Ep[KX], ELp[KX], Ea[KX], ELa[KX], Hl[KX];
        uint32_t eo = 0xFF, elo = 0xFF;                     // Eopen / ELopen of the current column, bit per row
        uint32_t qpack = 0, bmask = 0;
#pragma unroll
        for (int r = 0; r < KX; r++) {
            Ep[r] = Ea[r] = go; ELp[r] = ELa[r] = lgo; Hl[r] = 0;
            const int i = i0 + r;
            if (i < Q) {
                qpack |= (uint32_t)sm->qry[i] << (4 * r);
                if (i > 0 && (i % segLen) == 0) bmask |= 1u << r;
            }
        }
        ChainRec o;                                         // my state below my last row, column of the previous step
        o.hbot = 0; o.F = go; o.FL = lgo; o.f0 = go; o.fl0 = lgo; o.fc = kNegInf; o.flc = kNegInf; o.misc = 3 << 16;
        int diag_prev = 0;                                  // H(i0-1, j-1)
        int4 pre0 = make_int4(0, 0, 0, 0), pre1 = make_int4(0, 0, 0, 0);   // prefetched boundary record
        uint8_t* tr = ws.trace + (size_t)strip * steps * (32 * KX) + lane * KX;

        for (int s = 0; s < steps; s++) {
            if (strip > 0 && (s & 31) == 0) {
                // Boundary records of the previous strip, 32 columns per window.  Each lane keeps the record of
                // column 32*(w+1)+lane in registers one window ahead (ld.global.cg issued a window early, so its
                // latency hides behind 32 steps) and publishes it to shared memory when the window starts.
                // (cp.async.cg is NOT used here: on sm_100a it returned stale lines for addresses this SM had
                // re-written with st.global.cg since the previous tile -- measured, see DESIGN.md.)
                const int w = s >> 5;
                if (s == 0) { const int jj = lane; if (jj < R) { const int4* src = reinterpret_cast<const int4*>(ws.bound + jj); pre0 = __ldcg(src); pre1 = __ldcg(src + 1); } }
                int4* dst = reinterpret_cast<int4*>(&sm->rec[w & 1][lane]);
                dst[0] = pre0; dst[1] = pre1;
                const int jn = (w + 1) * 32 + lane;
                if (jn < R) { const int4* src = reinterpret_cast<const int4*>(ws.bound + jn); pre0 = __ldcg(src); pre1 = __ldcg(src + 1); }
                __syncwarp();
            }
            ChainRec in;
            in.hbot = __shfl_up_sync(0xffffffffu, o.hbot, 1);
            in.F    = __shfl_up_sync(0xffffffffu, o.F, 1);
            in.FL   = __shfl_up_sync(0xffffffffu, o.FL, 1);
            in.f0   = __shfl_up_sync(0xffffffffu, o.f0, 1);
            in.fl0  = __shfl_up_sync(0xffffffffu, o.fl0, 1);
            in.fc   = __shfl_up_sync(0xffffffffu, o.fc, 1);
            in.flc  = __shfl_up_sync(0xffffffffu, o.flc, 1);
            in.misc = __shfl_up_sync(0xffffffffu, o.misc, 1);
            if (lane == 0) {
                if (strip == 0) {                           // column start
                    in.hbot = 0; in.F = go; in.FL = lgo; in.f0 = go; in.fl0 = lgo;
                    in.fc = kNegInf; in.flc = kNegInf; in.misc = 3 << 16;
                } else if (s < R) {
                    in = sm->rec[(s >> 5) & 1][s & 31];
                }
            }
            const int j = s - lane;
            if (j >= 0 && j < R && nrows > 0) {
                int f0 = in.f0, fl0 = in.fl0, fc = in.fc, flc = in.flc, F = in.F, FL = in.FL;
                int kf = in.misc & 0xFF, kfl = (in.misc >> 8) & 0xFF;
                uint32_t Fo = (in.misc >> 16) & 1, FLo = (in.misc >> 17) & 1;
                int d = diag_prev;
                const int rb5 = (int)sm->ref[j] * 5;
                uint32_t w0 = 0, w1 = 0, eo_n = 0, elo_n = 0;
#pragma unroll
                for (int r = 0; r < KX; r++) {
                    if (r < nrows) {
                        if ((bmask >> r) & 1) {             // lane boundary of the striped layout
                            if (fc >= f0) kf += 1; else { fc = f0; kf = 1; }
                            if (flc >= fl0) kfl += 1; else { flc = fl0; kfl = 1; }
                            f0 = go; fl0 = lgo;
                        }
                        const int qb = (qpack >> (4 * r)) & 7;
                        const int hd = max(0, d + ssub[rb5 + qb]);
                        const int hm = max(max(max(hd, Ep[r]), max(f0, ELp[r])), fl0);
                        const int h = max(hm, max(fc, flc));
                        uint32_t T;
                        if (TRACE) {
                            const bool cs = (fc == h), cl = (flc == h);
                            const uint32_t Tm = (ELp[r] == h) ? XT_DEL_L : (fl0 == h) ? XT_INS_L : 0xFFu;
                            if (h == hd) T = (Tm != 0xFFu) ? Tm : (h == 0 ? XT_ZERO : XT_DIAG);
                            else if (cs || cl) T = (cs && (!cl || kf >= kfl)) ? XT_INS : XT_INS_L;
                            else T = (Tm != 0xFFu) ? Tm : (f0 == h) ? XT_INS : XT_DEL;
                            const uint32_t code = T | (((eo >> r) & 1) << 3) | (Fo << 4) | (((elo >> r) & 1) << 5) | (FLo << 6);
                            if (r < 4) w0 |= code << (8 * r); else w1 |= code << (8 * (r - 4));
                        }
                        Ep[r] = max(hm + go, Ep[r] + ge); ELp[r] = max(hm + lgo, ELp[r] + lge);
                        f0 = max(hm + go, f0 + ge); fl0 = max(hm + lgo, fl0 + lge); fc += ge; flc += lge;
                        const int ho = h + go, hlo = h + lgo;
                        const int ee = Ea[r] + ge, ele = ELa[r] + lge, fe = F + ge, fle = FL + lge;
                        if (TRACE) {
                            eo_n |= (uint32_t)(ho > ee) << r; elo_n |= (uint32_t)(hlo > ele) << r;
                            Fo = (ho > fe); FLo = (hlo > fle);
                        }
                        Ea[r] = max(ho, ee); ELa[r] = max(hlo, ele); F = max(ho, fe); FL = max(hlo, fle);
                        d = Hl[r]; Hl[r] = h;
                        if (!START_END) {
                            const int i = i0 + r;
                            if (h > best || (h == best && (j < bj || (j == bj && i < bi)))) { best = h; bj = j; bi = i; }
                        } else if (i0 + r == Q - 1 && j == R - 1) corner = h;
                    }
                }
                eo = eo_n; elo = elo_n;
                if (TRACE) __stcg(reinterpret_cast<uint2*>(tr + (size_t)s * (32 * KX)), make_uint2(w0, w1));
                o.hbot = Hl[KX - 1]; o.F = F; o.FL = FL; o.f0 = f0; o.fl0 = fl0; o.fc = fc; o.flc = flc;
                o.misc = kf | (kfl << 8) | (Fo << 16) | (FLo << 17);
                if (lane == 31 && strip + 1 < nstrips) {    // boundary row for the next strip
                    int4* dst = reinterpret_cast<int4*>(ws.bound + j);
                    __stcg(dst, make_int4(o.hbot, o.F, o.FL, o.f0));
                    __stcg(dst + 1, make_int4(o.fl0, o.fc, o.flc, o.misc));
                }
            }
            diag_prev = in.hbot;
        }
        __syncwarp();
    }
    if (START_END) {
        // the lane that owns row Q-1 holds the corner (Processor.cpp:514-517)
        const int owner = ((Q - 1) % kStrip) / KX;
        out.score = __shfl_sync(0xffffffffu, corner, owner);
        out.query_max_pos = Q - 1; out.ref_max_pos = R - 1;              // Processor.cpp:544-547
    } else {
        // global max; first column, then smallest row (Processor.cpp:502-509, :528-541)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && (oj < bj || (oj == bj && oi < bi)))) { best = ob; bj = oj; bi = oi; }
        }
        out.score = best; out.ref_max_pos = bj; out.query_max_pos = bi;
    }
}

__device__ __forceinline__ uint32_t exact_trace_at(const uint8_t* trace, int steps, int i, int j) {
    const int strip = i / kStrip, rem = i - strip * kStrip;
    const int v = rem / KX, r = rem - v * KX;
    const size_t off = ((size_t)strip * steps + (size_t)(j + v)) * (32 * KX) + v * KX + r;
    return __ldcg(trace + off);
}

// Traceback (Processor.cpp:585-716), executed by ONE lane.  sink(op) receives I/D/M in emission order.
template <class Sink>
__device__ void exact_traceback(const uint8_t* trace, int Q, int R, int i, int j, int max_tb, TileOut& out, Sink& sink) {
    const int steps = R + 31;
    int is = 0, js = 0, total = 0;
    uint32_t where = XT_DIAG, tfl = 0;
    (void)Q;
    while (i >= 0 && j >= 0) {
        if (is == max_tb || js == max_tb) break;
        const uint32_t c = exact_trace_at(trace, steps, i, j);
        if (where == XT_DIAG) {
            const uint32_t T = c & 7;
            if (T == XT_DIAG) { sink(DARWIN_OP_M); total++; i--; j--; is++; js++; }
            else if (T == XT_ZERO) break;
            else { where = T; if (T == XT_INS_L) tfl |= 1; if (T >= XT_DEL_L) tfl |= 2; }
        } else if (where == XT_DEL) {
            sink(DARWIN_OP_D); total++; j--; js++; where = (c & XB_EOPEN) ? XT_DIAG : XT_DEL;
        } else if (where == XT_INS) {
            sink(DARWIN_OP_I); total++; i--; is++; where = (c & XB_FOPEN) ? XT_DIAG : XT_INS;
        } else if (where == XT_INS_L) {
            sink(DARWIN_OP_I); total++; i--; is++; where = (c & XB_FLOPEN) ? XT_DIAG : XT_INS_L;
        } else {
            sink(DARWIN_OP_D); total++; j--; js++; where = (c & XB_ELOPEN) ? XT_DIAG : XT_DEL_L;
        }
    }
    out.query_offset = is; out.ref_offset = js; out.total = total; out.tflags = tfl;
}

} // namespace gact

// Packed EXACT path: the lazy-F-faithful rule of the reference (oracle/gact_oracle.c rule STREAM)
// on the 64-virtual-lane packed wavefront of gact_fast.cuh.  Used when the clean fast path cannot decide a tile
// (long-gap candidate on the traceback path, path outside the band) and for the 1984x960 / 960x1984 large tiles,
// whose paths cross a long gap by construction.  Anything it cannot take (N bases, non-uniform matrices, tiles whose
// striped segment length is not a multiple of 4, score range beyond 11 bits) goes to the unpacked path (gact_exact.cuh).
//
// Geometry: K = 4 rows per virtual lane, strips of 256 rows; with segLen = ceil(Q/16) a multiple of 4 every lane
// boundary of the reference's striped layout (rows i % segLen == 0, Processor.cpp:96-111) falls on row 0 of a virtual
// lane, so the boundary rule is applied once per step under a per-half mask.
//
// Chains per column (all tagged (score+bias)*32 + tag, unsigned 16-bit halves):
//   main-loop group -- what the AVX2 main loop sees (Processor.cpp:293-380): hd (tag DIAG/ZERO), E' (DEL), f0 (INS),
//       EL' (DEL_L), fl0 (INS_L); one packed max gives Hm with the priority EL' > FL0 > diag > F0 > E' (:309-325);
//   carried group -- pure-extension chains entering from lanes above (lazy-F passes, :385-408): fc, flc, tag =
//       lane distance k*2 + (1 for the short chain), so max(fc, flc) picks the farthest pass and F over F_L on equal k
//       (later passes override earlier ones, F before F_L inside a pass, :424-431);
//   true chains -- E, F, E_L, F_L of the recurrence proper, only their open/extend markers are stored (:353-372, :442-474).
// T = carried winner if a carried chain attains H and the diagonal does not (:427-431), else the main-loop winner.
// Trace: 7 bits per cell (T, 4 markers), one 32-bit word per virtual lane and step, full matrix, in the warp's global scratch.
#pragma once
#include "gact_fast.cuh"

namespace gact {

constexpr int XK = 4;                                   // rows per virtual lane
constexpr uint32_t X_DEL = 0, X_INS = 1, X_DIAG = 2, X_ZERO = 3, X_INSL = 4, X_DELL = 5;

struct XConst {
    uint32_t zeroc, hm_init, pkc32, negc32, diaga;
    uint32_t ep_init, f0_init, elp_init, fl0_init;      // main-loop chain starts (tags DEL / INS / DEL_L / INS_L)
    uint32_t gop_e, gop_f0, lgop_el, lgop_fl0;          // open addends incl. the chain's tag (x 65537)
    uint32_t go_plain, lgo_plain;                       // open addends of the true chains (no tag)
    uint32_t e_true_init, el_true_init;                 // (B+go)<<5, (B+lgo)<<5: E(i,0), E_L(i,0), F(0,j), F_L(0,j)
    uint32_t geh, lgeh;                                 // extension addends as two's-complement halves
    uint32_t minc, mincl;                               // floors of the carried chains ("-inf" that cannot wrap)
    uint32_t one[4];
    int32_t bias, max_score, eligible, match;
};

__host__ inline XConst make_xconst(const DevScoring& sc, const FastConst& f) {
    XConst x{};
    const int m = sc.match, mm = sc.mismatch, go = sc.go, ge = sc.ge, lgo = sc.lgo, lge = sc.lge, B = f.bias;
    auto pk = [](int v) { return (uint32_t)(v & 0xFFFF) * 0x00010001u; };
    auto add = [](int v) { return (uint32_t)(v * 65537); };
    x.zeroc = pk((B << 5) | (X_ZERO << 2));
    x.hm_init = pk(((B + mm) << 5) | (X_DIAG << 2));
    x.pkc32 = pk((m - mm) << 5); x.negc32 = (uint32_t)(-((m - mm) << 5));
    x.diaga = add(mm * 32 + (int)(X_DIAG << 2));
    x.ep_init = pk(((B + go) << 5) | (X_DEL << 2)); x.f0_init = pk(((B + go) << 5) | (X_INS << 2));
    x.elp_init = pk(((B + lgo) << 5) | (X_DELL << 2)); x.fl0_init = pk(((B + lgo) << 5) | (X_INSL << 2));
    x.gop_e = add(go * 32 + (int)(X_DEL << 2)); x.gop_f0 = add(go * 32 + (int)(X_INS << 2));
    x.lgop_el = add(lgo * 32 + (int)(X_DELL << 2)); x.lgop_fl0 = add(lgo * 32 + (int)(X_INSL << 2));
    x.go_plain = add(go * 32); x.lgo_plain = add(lgo * 32);
    x.e_true_init = pk((B + go) << 5); x.el_true_init = pk((B + lgo) << 5);
    x.geh = pk(ge * 32); x.lgeh = pk(lge * 32);
    x.minc = pk(-ge * 32); x.mincl = pk((lge ? -lge : 1) * 32);
    x.one[0] = x.one[1] = x.one[2] = x.one[3] = 1;
    x.bias = B; x.max_score = f.max_score; x.match = m;
    x.eligible = f.eligible && B >= 2;
    return x;
}

struct XSmemView {
    uint8_t* sref; uint8_t* sqry; uint4* rec;           // rec[2][32]: boundary records of the previous strip
    static constexpr size_t kBytes = 2 * kSeqSmem + 2 * 32 * 16;
    __device__ explicit XSmemView(unsigned char* base)
        : sref(base), sqry(base + kSeqSmem), rec(reinterpret_cast<uint4*>(base + 2 * kSeqSmem)) {}
};

__host__ __device__ inline size_t xfast_trace_bytes(int Q, int R) {
    return (size_t)((Q + 64 * XK - 1) / (64 * XK)) * (size_t)(R + 63) * 64 * 4;
}

// Tile shapes the packed exact path takes: every lane boundary of the striped layout must fall on a virtual-lane start.
__host__ __device__ inline bool xfast_shape_ok(int Q) { return Q >= 16 && (((Q + 15) >> 4) % XK) == 0; }

// Forward pass; returns the corner score.  trace: warp's global scratch, bound: >= R records of 16 bytes.
static __device__ int xfast_forward(const XConst& xc, const XSmemView& v, uint32_t* trace, uint4* bound, int Q, int R) {
    constexpr int K = XK;
    constexpr uint32_t CLEAN = 0xFFE0FFE0u;
    const int lane = lane_id();
    const int segLen = (Q + 15) >> 4;
    const int nstrips = (Q + 64 * K - 1) / (64 * K);
    const int vc = (Q - 1) / K, rc = (Q - 1) - vc * K;
    const int sc_step = R - 1 + (vc & 63);
    const int src = (lane + 31) & 31;
    const int steps = R + 63;
    const uint32_t zeroc = xc.zeroc, pkc32 = xc.pkc32, negc32 = xc.negc32, diaga = xc.diaga;
    const uint32_t geh = xc.geh, lgeh = xc.lgeh, minc = xc.minc, mincl = xc.mincl;
    const uint32_t one0 = xc.one[0], one1 = xc.one[1], one2 = xc.one[2], one3 = xc.one[3];
    uint32_t corner = 0;

    for (int strip = 0; strip < nstrips; strip++) {
        const int row0 = strip * 64 * K;
        uint32_t qq[K], Hm[K], Ep[K], ELp[K], Ea[K], ELa[K];
#pragma unroll
        for (int r = 0; r < K; r++) {
            const int ilo = row0 + K * lane + r, ihi = row0 + K * (lane + 32) + r;
            qq[r] = (ilo < Q ? (uint32_t)v.sqry[ilo] : 6u) | ((ihi < Q ? (uint32_t)v.sqry[ihi] : 6u) << 16);
            Hm[r] = xc.hm_init; Ep[r] = xc.ep_init; ELp[r] = xc.elp_init; Ea[r] = xc.e_true_init; ELa[r] = xc.el_true_init;
        }
        // lane-boundary mask of my two virtual lanes (rows K*v with (K*v) % segLen == 0, v > 0)
        const int vlo = strip * 64 + lane, vhi = vlo + 32;
        const uint32_t bm = ((vlo > 0 && (K * vlo) % segLen == 0) ? 0x0000FFFFu : 0u) |
                            (((K * vhi) % segLen == 0) ? 0xFFFF0000u : 0u);
        // state below my last row (what the next virtual lane receives)
        uint32_t sH = xc.hm_init, sF = xc.e_true_init, sFL = xc.el_true_init, sf0 = xc.f0_init, sfl0 = xc.fl0_init,
                 sfc = minc, sflc = mincl;
        uint32_t diag_in = xc.hm_init;
        uint32_t* tr = trace + (size_t)strip * steps * 64 + lane;
        uint32_t rlo = (lane == 0 && R > 0) ? v.sref[0] : 5u, rhi = 5u;

        // the corner is the last valid cell of the last strip: its loop ends there (nothing later is ever read)
        const int strip_steps = (strip == nstrips - 1) ? sc_step + 1 : steps;
        for (int s = 0; s < strip_steps; s++) {
            if (strip > 0 && (s & 31) == 0) {                                // next 32 boundary records -> shared memory
                const int jj = s + lane;
                if (jj < R) v.rec[((s >> 5) & 1) * 32 + lane] = __ldcg(bound + jj);
                __syncwarp();
            }
            uint32_t inH = __shfl_sync(0xffffffffu, sH, src);
            uint32_t F   = __shfl_sync(0xffffffffu, sF, src);
            uint32_t FL  = __shfl_sync(0xffffffffu, sFL, src);
            uint32_t f0  = __shfl_sync(0xffffffffu, sf0, src);
            uint32_t fl0 = __shfl_sync(0xffffffffu, sfl0, src);
            uint32_t fc  = __shfl_sync(0xffffffffu, sfc, src);
            uint32_t flc = __shfl_sync(0xffffffffu, sflc, src);
            if (lane == 0) {
                uint32_t tH = xc.hm_init, tF = xc.e_true_init, tFL = xc.el_true_init, tf0 = xc.f0_init, tfl0 = xc.fl0_init,
                         tfc = minc, tflc = mincl;
                if (strip > 0 && s < R) {
                    const uint4 q = v.rec[((s >> 5) & 1) * 32 + (s & 31)];
                    tH = q.x; tF = q.x >> 16; tFL = q.y; tf0 = q.y >> 16; tfl0 = q.z; tfc = q.z >> 16; tflc = q.w;
                }
                inH = __byte_perm(tH, inH, 0x5410); F = __byte_perm(tF, F, 0x5410); FL = __byte_perm(tFL, FL, 0x5410);
                f0 = __byte_perm(tf0, f0, 0x5410); fl0 = __byte_perm(tfl0, fl0, 0x5410);
                fc = __byte_perm(tfc, fc, 0x5410); flc = __byte_perm(tflc, flc, 0x5410);
            }
            // lane boundary of the striped layout: the own-lane chains become "carried, distance 1" (:385-408)
            {
                const uint32_t cF = __vmaxu2(fc + 0x00020002u, (f0 & CLEAN) | 0x00030003u);
                const uint32_t cL = __vmaxu2(flc + 0x00020002u, (fl0 & CLEAN) | 0x00020002u);
                fc = (cF & bm) | (fc & ~bm); flc = (cL & bm) | (flc & ~bm);
                f0 = (xc.f0_init & bm) | (f0 & ~bm); fl0 = (xc.fl0_init & bm) | (fl0 & ~bm);
            }
            const uint32_t rq = rlo | (rhi << 16);
            {
                const int jl = s + 1 - lane, jh = jl - 32;
                rlo = ((unsigned)jl < (unsigned)R) ? v.sref[jl] : 5u;
                rhi = ((unsigned)jh < (unsigned)R) ? v.sref[jh] : 5u;
            }
            uint32_t d = diag_in, acc0 = 0, acc1 = 0;
#pragma unroll
            for (int r = 0; r < K; r++) {
                const uint32_t x  = rq ^ qq[r];
                const uint32_t t  = __vminu2(x, 0x00010001u);
                const uint32_t sb = t * negc32 + pkc32;
                const uint32_t hd = __viaddmax_u16x2(d, sb, zeroc);
                const uint32_t M1 = __vimax3_u16x2(__vimax3_u16x2(hd, Ep[r], f0), ELp[r], fl0);    // Hm with its source tag
                const uint32_t cC = __vmaxu2(fc, flc);                                            // best carried chain
                const uint32_t S1 = M1 & CLEAN, SC = cC & CLEAN, hdS = hd & CLEAN;
                const uint32_t Hc = __vmaxu2(S1, SC);                                             // true H
                const uint32_t t1 = __vminu2(SC ^ Hc, 0x00010001u);       // 1: carried chain below H
                const uint32_t t2 = __vminu2(hdS ^ Hc, 0x00010001u);      // 1: diagonal below H
                const uint32_t use = t2 & ~t1;                            // 1: a carried chain decides T (:427-431)
                const uint32_t msk = use * 0xFFFFu;
                const uint32_t Tm = (M1 & 0x001C001Cu) * 4u;              // main-loop pointer at bits 6:4
                const uint32_t Tc = (cC & 0x00010001u) * 0xFFFFFFD0u + 0x00400040u;   // INS_L (4<<4) or INS (1<<4): 64 - 48*b
                const uint32_t marks = ((Ea[r] | F | ELa[r]) | FL) & 0x000F000Fu;
                const uint32_t code = ((Tc & msk) | (Tm & ~msk)) | marks;
                d = Hm[r];
                Hm[r] = Hc * one0 + diaga;
                // main-loop chains continue from the UNcorrected Hm (:332-341, :363-366)
                Ep[r]  = __viaddmax_u16x2(Ep[r], geh, S1 * one1 + xc.gop_e);
                ELp[r] = __viaddmax_u16x2(ELp[r], lgeh, S1 * one2 + xc.lgop_el);
                f0     = __viaddmax_u16x2(f0, geh, S1 * one3 + xc.gop_f0);
                fl0    = __viaddmax_u16x2(fl0, lgeh, S1 * one0 + xc.lgop_fl0);
                fc     = __viaddmax_u16x2(fc, geh, minc);
                flc    = __viaddmax_u16x2(flc, lgeh, mincl);
                // true chains and their open/extend markers (:442-474), from the corrected H
                const uint32_t Ho = Hc * one1 + xc.go_plain, HoL = Hc * one2 + xc.lgo_plain;
                Ea[r]  = __viaddmax_u16x2(Ea[r] | 0x00010001u, geh, Ho);
                F      = __viaddmax_u16x2(F | 0x00020002u, geh, Ho);
                ELa[r] = __viaddmax_u16x2(ELa[r] | 0x00040004u, lgeh, HoL);
                FL     = __viaddmax_u16x2(FL | 0x00080008u, lgeh, HoL);
                if (r < 2) acc0 += code << (7 * r); else acc1 += code << (7 * (r - 2));
            }
            diag_in = inH;
            sH = Hm[K - 1]; sF = F; sFL = FL; sf0 = f0; sfl0 = fl0; sfc = fc; sflc = flc;
            if (lane == 31 && strip + 1 < nstrips && (unsigned)(s - 63) < (unsigned)R) {       // bottom row of the strip
                uint4 q;
                q.x = __byte_perm(sH, sF, 0x7632); q.y = __byte_perm(sFL, sf0, 0x7632);
                q.z = __byte_perm(sfl0, sfc, 0x7632); q.w = sflc >> 16;
                __stcg(bound + (s - 63), q);
            }
            if ((unsigned)(s - lane) < (unsigned)R) __stcg(tr, __byte_perm(acc0, acc1, 0x5410));
            if ((unsigned)(s - lane - 32) < (unsigned)R) __stcg(tr + 32, __byte_perm(acc0, acc1, 0x7632));
            tr += 64;
        }
        __syncwarp();
        if (strip == nstrips - 1) {
#pragma unroll
            for (int r = 0; r < K; r++) if (r == rc) corner = Hm[r] - diaga;
        }
    }
    const int vl = vc & 63;
    uint32_t cw = __shfl_sync(0xffffffffu, corner, vl & 31);
    cw = (vl >= 32) ? (cw >> 16) : (cw & 0xFFFFu);
    return (int)(cw >> 5) - xc.bias;
}

// Traceback over the full trace (Processor.cpp:585-716), executed by the whole warp with identical state in every lane
// (see fast_traceback): in the DIAG state lane k probes cell (i-k, j-k), a ballot gives the length of the run of DIAG
// pointers, which is emitted and skipped at once -- on the 1984x960 tiles that replaces thousands of dependent global
// loads on one lane by a few dozen warp-wide ones.  Every other cell goes through the generic step.
template <class Sink>
__device__ void xfast_traceback(const uint32_t* trace, int Q, int R, int max_tb, TileOut& out, Sink& sink) {
    constexpr int K = XK;
    const int lane = lane_id();
    const int steps = R + 63;
    int i = Q - 1, j = R - 1;
    const int lim_i = min(Q, max_tb), lim_j = min(R, max_tb);
    int left_i = lim_i, left_j = lim_j;
    uint32_t where = X_DIAG, tfl = 0;
    for (;;) {
        const int lim = min(left_i, left_j);
        if (lim <= 0) break;
        if (where == X_DIAG) {
            const bool ok = lane < lim;                                          // implies i - lane >= 0 and j - lane >= 0
            const int ii = ok ? i - lane : 0, jj = ok ? j - lane : 0;
            const int vg = ii / K, r = ii - vg * K;
            const int strip = vg >> 6, vl = vg & 63;
            uint32_t w = 0;
            if (ok) w = __ldcg(trace + ((size_t)strip * steps + (size_t)(jj + vl)) * 64 + vl);
            const uint32_t code = (w >> ((r & 1) * 7 + (r >> 1) * 16)) & 127u;
            const uint32_t is_m = __ballot_sync(0xffffffffu, ok && (code >> 4) == X_DIAG);
            const int run = (is_m == 0xffffffffu) ? 32 : __ffs(~is_m) - 1;
            if (run > 0) {
                sink.run_m(run);
                i -= run; j -= run; left_i -= run; left_j -= run;
                continue;
            }
        }
        const int vg = i / K, r = i - vg * K;
        const int strip = vg >> 6, vl = vg & 63;
        const uint32_t w = __ldcg(trace + ((size_t)strip * steps + (size_t)(j + vl)) * 64 + vl);
        const uint32_t code = (w >> ((r & 1) * 7 + (r >> 1) * 16)) & 127u;
        uint32_t st = where;
        if (where == X_DIAG) {
            st = code >> 4;
            if (st == X_ZERO) break;
            if (st == X_INSL) tfl |= 1;
            if (st >= X_INSL) tfl |= 2;
        }
        bool up, left;
        if (st == X_DIAG) { sink(DARWIN_OP_M); up = true; left = true; where = X_DIAG; }
        else if (st == X_DEL)  { sink(DARWIN_OP_D); up = false; left = true; where = (code & 1u) ? X_DEL : X_DIAG; }
        else if (st == X_INS)  { sink(DARWIN_OP_I); up = true; left = false; where = (code & 2u) ? X_INS : X_DIAG; }
        else if (st == X_DELL) { sink(DARWIN_OP_D); up = false; left = true; where = (code & 4u) ? X_DELL : X_DIAG; }
        else                   { sink(DARWIN_OP_I); up = true; left = false; where = (code & 8u) ? X_INSL : X_DIAG; }
        if (left) { j--; left_j--; }
        if (up) { i--; left_i--; }
    }
    out.query_offset = lim_i - left_i; out.ref_offset = lim_j - left_j; out.total = sink.count(); out.tflags = tfl;
}

} // namespace gact

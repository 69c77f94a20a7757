// Packed fast path of the GACT tile: one warp per tile, 64 virtual lanes (two int16 cells per 32-bit
// register), tagged scores, traceback band in shared memory.
//
// What is computed: the textbook two-piece-affine local recurrence of the reference (SURVEY A.1,
// software/Processor.cpp:293-366) with the "clean" pointer rule.  That rule equals the
// reference's lazy-F tie-breaking whenever the traceback never meets a long-gap candidate while in the
// DIAG state; when it does (code T_L), or when the path leaves the stored band, the tile is recomputed by
// the exact path (gact_exact.cuh).  CPU twin: oracle/gact_oracle.c rule CLEAN + rerun.
//
// Layout.  Physical lane l holds virtual lane l in the low halves and virtual lane l+32 in the high halves
// of its registers; virtual lane v owns query rows [K*v, K*v+K) and at step s works on reference column
// s - v.  The vertical chain (F, F_L, H of the row above) moves one virtual lane per step with ONE rotate
// shuffle per value; lane 0 moves the low half of lane 31 into its high half.
//
// Tagged scores.  Every DP value is stored as (score + bias) * 32 + tag in an unsigned 16-bit half, so one
// packed max both compares scores and resolves ties in the reference's priority order:
//     tag bits 4:2 = source   L(4) > ZERO(3) > DIAG(2) > F/INS(1) > E/DEL(0)      (Processor.cpp:309-325)
//     bit 0 = "E was extended", bit 1 = "F was extended"  (strict > opens, ties extend; :353, :369)
// so the traceback pointer T and the open/extend bits drop out of the max for free (5 bits per cell).
//
// Fast-path preconditions (checked on the host / per tile, otherwise the exact path runs):
//   uniform match/mismatch matrix, match > 0 > mismatch, gap_open <= gap_extend < 0, long gaps <= 0,
//   (match * min(Q,R) + bias) < 2048, no N in the tile, Q <= 64*K, R <= 64*K, corner traceback.
#pragma once
#include "gact_common.cuh"

namespace gact {

constexpr int kBandHalf = 48;                       // +-rows around the corner diagonal kept in shared memory

template <int K> struct FastGeom {
    static constexpr int kRows  = 64 * K;           // max query rows (and reference columns) of a fast tile
    static constexpr int kSteps = kRows + 63;
    static constexpr int kNB    = 2 * kBandHalf / (K + 1) + 2;      // band slots (virtual lanes) per step
    static constexpr int kPWords = kRows + 128;     // packed reference pairs P[j + 32] = r[j] | r[j-32] << 16
    static constexpr size_t kBandWords = (size_t)kSteps * kNB;
    static constexpr size_t kSmemBytes = (kBandWords + kPWords) * 4 + 2 * kRows;    // + staged byte sequences
};

struct FastConst {                                  // packed constants derived from the scoring (both halves equal)
    uint32_t zeroc;       // (B << 5) | ZERO tag        -- the clamp of the local alignment
    uint32_t hm_init;     // (B << 5) + mismatch*32 + DIAG tag   -- "H = 0" as a diagonal source
    uint32_t e_init;      // ((B + go) << 5)                       E(i,0), marker "opened"
    uint32_t el_init;     // ((B + lgo) << 5) | L tag
    uint32_t f_top;       // ((B + go) << 5) | INS tag             F(0,j)
    uint32_t fl_top;      // ((B + lgo) << 5) | L tag
    uint32_t pkc32;       // (match - mismatch) * 32, both halves
    int32_t  negc32;      // -(match - mismatch) * 32            (multiplier of the packed mismatch flags)
    int32_t  diaga;       // (mismatch*32 + DIAG tag) * 65537     addends: value * 65537 adds to both halves
    int32_t  goa, gofa;   // go*32 * 65537, (go*32 + INS tag) * 65537
    int32_t  gea;         // ge*32 * 65537
    int32_t  lgoa, lgea;  // (lgo*32 + L tag) * 65537, lge*32 * 65537
    int32_t  bias;        // B
    int32_t  max_score;   // largest corner score representable: 2047 - B
    int32_t  eligible;    // scoring admits the fast path
    int32_t  match;
};

constexpr uint32_t FT_DEL = 0, FT_INS = 1, FT_DIAG = 2, FT_ZERO = 3, FT_L = 4;
constexpr uint32_t kMaskT = 0x001C001Cu, kMaskM = 0x00030003u, kMaskClean = 0xFFE0FFE0u;

__host__ inline FastConst make_fast_const(const DevScoring& sc) {
    FastConst f{};
    const int m = sc.match, mm = sc.mismatch, go = sc.go, ge = sc.ge, lgo = sc.lgo, lge = sc.lge;
    int B = -mm;
    if (-(go + ge) > B) B = -(go + ge);
    if (-(lgo + lge) > B) B = -(lgo + lge);
    B += 1;
    f.eligible = sc.uniform && m > 0 && mm < 0 && go <= ge && ge < 0 && lgo <= lge && lge <= 0 && B < 512;
    auto pk = [](int v) { return (uint32_t)(v & 0xFFFF) * 0x00010001u; };
    f.bias = B; f.match = m; f.max_score = 2047 - B - m;
    f.zeroc = pk((B << 5) | (FT_ZERO << 2));
    f.hm_init = pk(((B + mm) << 5) | (FT_DIAG << 2));
    f.e_init = pk((B + go) << 5);
    f.el_init = pk(((B + lgo) << 5) | (FT_L << 2));
    f.f_top = pk(((B + go) << 5) | (FT_INS << 2));
    f.fl_top = pk(((B + lgo) << 5) | (FT_L << 2));
    f.pkc32 = pk((m - mm) << 5);
    f.negc32 = -((m - mm) << 5);
    f.diaga = (mm * 32 + (int)(FT_DIAG << 2)) * 65537;
    f.goa = (go * 32) * 65537; f.gofa = (go * 32 + (int)(FT_INS << 2)) * 65537;
    f.gea = (ge * 32) * 65537;
    f.lgoa = (lgo * 32 + (int)(FT_L << 2)) * 65537; f.lgea = (lge * 32) * 65537;
    return f;
}

// Per-warp shared memory of the fast path, carved from the dynamic shared memory of the CTA.
template <int K> struct FastSmemView {
    uint32_t* band;       // [kSteps][kNB]
    uint32_t* P;          // [kPWords]
    uint8_t*  sref;       // [kRows]
    uint8_t*  sqry;       // [kRows]
    __device__ explicit FastSmemView(unsigned char* base) {
        band = reinterpret_cast<uint32_t*>(base);
        P = band + FastGeom<K>::kBandWords;
        sref = reinterpret_cast<uint8_t*>(P + FastGeom<K>::kPWords);
        sqry = sref + FastGeom<K>::kRows;
    }
};

// Band addressing shared by the forward pass and the traceback: slot of virtual lane v at step s.
template <int K> struct BandMap {
    int off, c;           // qd(s) = (s + off) / (K+1);  slot = v - qd + c
    __device__ BandMap(int Q, int R) {
        const int vc = (Q - 1) / K;                   // virtual lane of the corner row
        // u(s) = s - s_c + (K+1) * v_c with s_c = R - 1 + v_c; +(K+1)*512 keeps the dividend positive
        off = -(R - 1 + vc) + (K + 1) * vc + (K + 1) * 512;
        c = 512 + FastGeom<K>::kNB / 2;
    }
    __device__ __forceinline__ int slot(int s, int v) const { return v - (s + off) / (K + 1) + c; }
};

// Forward pass of one tile.  Sequences must already be staged (codes 0..3) in v.sref / v.sqry.
// Returns the corner score H(Q-1, R-1) in all lanes.
template <int K>
__device__ int fast_forward(const FastConst& fc, const FastSmemView<K>& v, int Q, int R) {
    using G = FastGeom<K>;
    const int lane = lane_id();
    // packed reference pairs: P[32 + j] = r[j] | r[j-32] << 16, dummy base 5 outside [0,R)
    for (int k = lane; k < G::kPWords; k += 32) {
        const int j = k - 32;
        const uint32_t lo = (j >= 0 && j < R) ? v.sref[j] : 5u;
        const uint32_t hi = (j - 32 >= 0 && j - 32 < R) ? v.sref[j - 32] : 5u;
        v.P[k] = lo | (hi << 16);
    }
    uint32_t qq[K], Hm[K], E[K], EL[K];
#pragma unroll
    for (int r = 0; r < K; r++) {
        const int ilo = K * lane + r, ihi = K * (lane + 32) + r;
        qq[r] = (ilo < Q ? (uint32_t)v.sqry[ilo] : 6u) | ((ihi < Q ? (uint32_t)v.sqry[ihi] : 6u) << 16);
        Hm[r] = fc.hm_init; E[r] = fc.e_init; EL[r] = fc.el_init;
    }
    __syncwarp();
    const BandMap<K> bm(Q, R);
    const int vc = (Q - 1) / K, rc = (Q - 1) - vc * K, sc_step = R - 1 + vc;
    uint32_t sendH = fc.hm_init, sendF = fc.f_top, sendFL = fc.fl_top;   // state below my last row (previous step)
    uint32_t diag_in = fc.hm_init;                                       // Hm(row above, previous column)
    uint32_t corner = 0;
    const int src = (lane + 31) & 31;
    const int steps = R + 63;
    uint32_t rq_next = v.P[32 - lane];                                   // step 0: j = -lane (dummy unless lane 0)

    for (int s = 0; s < steps; s++) {
        // values from the virtual lane above: rotate by one lane; lane 0 shifts lane 31's low half up and
        // takes the top boundary in its low half
        uint32_t inH = __shfl_sync(0xffffffffu, sendH, src);
        uint32_t F   = __shfl_sync(0xffffffffu, sendF, src);
        uint32_t FL  = __shfl_sync(0xffffffffu, sendFL, src);
        if (lane == 0) {
            inH = __byte_perm(fc.hm_init, inH, 0x5410);
            F   = __byte_perm(fc.f_top, F, 0x5410);
            FL  = __byte_perm(fc.fl_top, FL, 0x5410);
        }
        const uint32_t rq = rq_next;
        rq_next = v.P[32 + s + 1 - lane];                                // prefetch next step's reference pair
        uint32_t d = diag_in;
        uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
        for (int r = 0; r < K; r++) {
            const uint32_t x  = rq ^ qq[r];
            const uint32_t t  = __vminu2(x, 0x00010001u);
            const uint32_t sb = t * (uint32_t)fc.negc32 + fc.pkc32;
            const uint32_t hd = __viaddmax_u16x2(d, sb, fc.zeroc);
            const uint32_t h1 = __vimax3_u16x2(hd, E[r], F);
            const uint32_t Hk = __vmaxu2(h1, __vmaxu2(EL[r], FL));
            const uint32_t code = (Hk & kMaskT) | ((E[r] | F) & kMaskM);
            const uint32_t Hc = Hk & kMaskClean;
            d = Hm[r];
            Hm[r] = Hc + (uint32_t)fc.diaga;
            const uint32_t Ho = Hc + (uint32_t)fc.goa, HoF = Hc + (uint32_t)fc.gofa, HoL = Hc + (uint32_t)fc.lgoa;
            E[r]  = __vmaxu2((E[r] | 0x00010001u) + (uint32_t)fc.gea, Ho);
            F     = __vmaxu2((F | 0x00020002u) + (uint32_t)fc.gea, HoF);
            EL[r] = __vmaxu2(EL[r] + (uint32_t)fc.lgea, HoL);
            FL    = __vmaxu2(FL + (uint32_t)fc.lgea, HoL);
            if (r < 3) acc0 += code << (5 * r); else acc1 += code << (5 * (r - 3));
            if (r == rc && s == sc_step) corner = Hc;                    // s, rc warp-uniform: taken at one step only
        }
        diag_in = inH;
        sendH = Hm[K - 1]; sendF = F; sendFL = FL;
        // band store: one word per (virtual lane, step): rows 0-2 in bits 0-14, rows 3-5 in bits 16-30
        const int qd = (s + bm.off) / (K + 1);
        const int slot_lo = lane - qd + bm.c, slot_hi = slot_lo + 32;
        if ((unsigned)slot_lo < (unsigned)G::kNB && (unsigned)(s - lane) < (unsigned)R)
            v.band[s * G::kNB + slot_lo] = __byte_perm(acc0, acc1, 0x5410);
        if ((unsigned)slot_hi < (unsigned)G::kNB && (unsigned)(s - lane - 32) < (unsigned)R)
            v.band[s * G::kNB + slot_hi] = __byte_perm(acc0, acc1, 0x7632);
    }
    __syncwarp();
    // corner owner: virtual lane vc -> physical lane vc & 31, half vc >> 5
    uint32_t cw = __shfl_sync(0xffffffffu, corner, vc & 31);
    cw = (vc >= 32) ? (cw >> 16) : (cw & 0xFFFFu);
    return (int)(cw >> 5) - fc.bias;
}

enum : int { FAST_OK = 0, FAST_LFLAG = 1, FAST_BAND = 2 };

// Traceback over the shared-memory band (Processor.cpp:585-716 with the clean rule), ONE lane.
// Returns FAST_OK, or the reason the tile must be recomputed by the exact path.
template <int K, class Sink>
__device__ int fast_traceback(const FastSmemView<K>& vw, int Q, int R, int max_tb, TileOut& out, Sink& sink) {
    using G = FastGeom<K>;
    const BandMap<K> bm(Q, R);
    int i = Q - 1, j = R - 1;
    int v = i / K, r = i - v * K;
    int is = 0, js = 0, total = 0;
    uint32_t where = FT_DIAG;
    while (i >= 0 && j >= 0) {
        if (is == max_tb || js == max_tb) break;
        const int s = j + v;
        const int slot = bm.slot(s, v);
        if ((unsigned)slot >= (unsigned)G::kNB) return FAST_BAND;
        const uint32_t w = vw.band[s * G::kNB + slot];
        const uint32_t code = (w >> (r < 3 ? 5 * r : 16 + 5 * (r - 3))) & 31u;
        if (where == FT_DIAG) {
            const uint32_t T = code >> 2;
            if (T == FT_DIAG) {
                sink(DARWIN_OP_M); total++; i--; j--; is++; js++;
                if (r == 0) { r = K - 1; v--; } else r--;
            } else if (T == FT_ZERO) break;
            else if (T == FT_L) return FAST_LFLAG;
            else where = T;                                   // FT_DEL / FT_INS
        } else if (where == FT_DEL) {
            sink(DARWIN_OP_D); total++; j--; js++;
            where = (code & 1u) ? FT_DEL : FT_DIAG;           // bit 0: E was extended
        } else {
            sink(DARWIN_OP_I); total++; i--; is++;
            if (r == 0) { r = K - 1; v--; } else r--;
            where = (code & 2u) ? FT_INS : FT_DIAG;           // bit 1: F was extended
        }
    }
    out.query_offset = is; out.ref_offset = js; out.total = total; out.tflags = 0;
    return FAST_OK;
}

} // namespace gact

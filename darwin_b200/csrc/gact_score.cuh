// Score-only pre-pass for the 1984x960 / 960x1984 "large tiles" (extender.cpp:61-78, :385).
//
// A large tile is requested after a normal tile stalled; for spurious anchors (each burns two of them) and for real
// alignment ends it usually comes back with ZERO traceback pointers -- SURVEY 7: ~40 % of the large tiles, which are ~40 %
// of all DP cells end to end.  total_TB_pointers == 0  <=>  T(corner) == ZERO (the walk starts at the corner in DIAG
// state, Processor.cpp:613-642), and by the exact rule that is
//       H(corner) == 0  and  EL'(corner) != 0  and  FL0(corner) != 0.
// EL' (built from the not-yet-corrected H, Processor.cpp:336-341) and FL0 (own-lane chain) are bounded above by the true
// E_L and F_L of the textbook recurrence, so  H == 0 && E_L < 0 && F_L < 0  at the corner PROVES the zero result without any
// trace: no pointer is computed or stored, no tags are carried, and the cell update shrinks to 9 ALU-pipe + 4 FMA-pipe
// instructions per cell PAIR (fast path: 15 + 6, packed exact path: ~35).  Anything else (H > 0, or one of the long chains
// exactly 0) is inconclusive and the tile takes the traced path as before.
//
// Geometry: the 64-virtual-lane wavefront of gact_fast.cuh with K = 8 rows per virtual lane (strips of 512 rows: 2 strips
// for 960 rows, 4 for 1984), untagged values (score + bias) in unsigned 16-bit halves.
#pragma once
#include "gact_fast.cuh"

namespace gact {

constexpr int SK = 8;                                   // rows per virtual lane of the score-only pass

struct ScoreConst {
    uint32_t zeroc;       // B                              (clamp of the local alignment)
    uint32_t hm_init;     // B + mismatch                   ("H = 0" as a diagonal source)
    uint32_t e_init;      // B + go
    uint32_t el_init;     // B + lgo
    uint32_t pkc;         // (match - mismatch), both halves
    uint32_t negc;        // -(match - mismatch)            (multiplier of the packed mismatch flags)
    uint32_t mma, goa, lgoa;   // mismatch, go, lgo as addends (value * 65537)
    uint32_t geh, lgeh;   // ge, lge as two's-complement halves
    uint32_t one[3];
    int32_t  bias, eligible;
};

__host__ inline ScoreConst make_score_const(const DevScoring& sc, const FastConst& f) {
    ScoreConst s{};
    auto pk = [](int v) { return (uint32_t)(v & 0xFFFF) * 0x00010001u; };
    const int B = f.bias;
    s.zeroc = pk(B); s.hm_init = pk(B + sc.mismatch); s.e_init = pk(B + sc.go); s.el_init = pk(B + sc.lgo);
    s.pkc = pk(sc.match - sc.mismatch); s.negc = (uint32_t)(-(sc.match - sc.mismatch));
    s.mma = (uint32_t)(sc.mismatch * 65537); s.goa = (uint32_t)(sc.go * 65537); s.lgoa = (uint32_t)(sc.lgo * 65537);
    s.geh = pk(sc.ge); s.lgeh = pk(sc.lge);
    s.one[0] = s.one[1] = s.one[2] = 1;
    s.bias = B;
    // same preconditions as the fast path; scores of a 1984-row tile plus the bias must fit 16 bits
    s.eligible = f.eligible && (sc.match * kMaxTile + B < 65000);
    return s;
}

// true (warp-uniform) when the tile provably returns zero traceback pointers.  Sequences staged in v.sref / v.sqry.
static __device__ bool score_only_corner_is_zero(const ScoreConst& sc, const MultiSmemView& v, int Q, int R) {
    constexpr int K = SK;
    const int lane = lane_id();
    const uint32_t zeroc = sc.zeroc, pkc = sc.pkc, negc = sc.negc, mma = sc.mma, goa = sc.goa, lgoa = sc.lgoa;
    const uint32_t geh = sc.geh, lgeh = sc.lgeh, one0 = sc.one[0], one1 = sc.one[1], one2 = sc.one[2];
    const int nstrips = (Q + 64 * K - 1) / (64 * K);
    const int vc = (Q - 1) / K, rc = (Q - 1) - vc * K;                   // global virtual lane / row of the corner
    const int sc_step = R - 1 + (vc & 63);                               // step of the corner inside the last strip
    const int src = (lane + 31) & 31;
    const int steps = R + 63;
    uint32_t cH = 0, cEL = 0, cFL = 0;                                   // corner: H, and E_L / F_L as they ENTER its max

    for (int strip = 0; strip < nstrips; strip++) {
        const int row0 = strip * 64 * K;
        uint32_t qq[K], Hm[K], E[K], EL[K];
#pragma unroll
        for (int r = 0; r < K; r++) {
            const int ilo = row0 + K * lane + r, ihi = row0 + K * (lane + 32) + r;
            qq[r] = (ilo < Q ? (uint32_t)v.sqry[ilo] : 6u) | ((ihi < Q ? (uint32_t)v.sqry[ihi] : 6u) << 16);
            Hm[r] = sc.hm_init; E[r] = sc.e_init; EL[r] = sc.el_init;
        }
        uint32_t sendH = sc.hm_init, sendF = sc.e_init, sendFL = sc.el_init;
        uint32_t diag_in = sc.hm_init;
        uint32_t topHF = (sc.hm_init & 0xFFFFu) | (sc.e_init << 16), topFL = sc.el_init & 0xFFFFu;
        uint32_t nextHF = topHF, nextFL = topFL;
        if (strip > 0 && lane == 0) { nextHF = v.bHF[0]; nextFL = v.bFL[0]; }
        uint32_t rlo = (lane == 0 && R > 0) ? v.sref[0] : 5u, rhi = 5u;
        const bool last = strip == nstrips - 1;
        const int strip_steps = last ? sc_step + 1 : steps;

        for (int s = 0; s < strip_steps; s++) {
            uint32_t inH = __shfl_sync(0xffffffffu, sendH, src);
            uint32_t F   = __shfl_sync(0xffffffffu, sendF, src);
            uint32_t FL  = __shfl_sync(0xffffffffu, sendFL, src);
            if (lane == 0) {
                if (strip > 0) { topHF = nextHF; topFL = nextFL; if (s + 1 < R) { nextHF = v.bHF[s + 1]; nextFL = v.bFL[s + 1]; } }
                inH = __byte_perm(topHF, inH, 0x5410);
                F   = __byte_perm(topHF, F, 0x5432);
                FL  = __byte_perm(topFL, FL, 0x5410);
            }
            const uint32_t rq = rlo | (rhi << 16);
            {
                const int jl = s + 1 - lane, jh = jl - 32;
                rlo = ((unsigned)jl < (unsigned)R) ? v.sref[jl] : 5u;
                rhi = ((unsigned)jh < (unsigned)R) ? v.sref[jh] : 5u;
            }
            const bool cap = last && s == sc_step;                       // warp-uniform: the peeled corner step
            uint32_t d = diag_in;
#pragma unroll
            for (int r = 0; r < K; r++) {
                const uint32_t x  = rq ^ qq[r];
                const uint32_t t  = __vminu2(x, 0x00010001u);
                const uint32_t sb = t * negc + pkc;                      // IMAD
                const uint32_t hd = __viaddmax_u16x2(d, sb, zeroc);
                const uint32_t h1 = __vimax3_u16x2(hd, E[r], F);
                const uint32_t H  = __vimax3_u16x2(h1, EL[r], FL);
                if (cap && r == rc) { cH = H; cEL = EL[r]; cFL = FL; }
                d = Hm[r];
                Hm[r] = H * one0 + mma;                                  // IMADs keep the adds off the ALU pipe
                const uint32_t Ho = H * one1 + goa, HoL = H * one2 + lgoa;
                E[r]  = __viaddmax_u16x2(E[r], geh, Ho);
                F     = __viaddmax_u16x2(F, geh, Ho);
                EL[r] = __viaddmax_u16x2(EL[r], lgeh, HoL);
                FL    = __viaddmax_u16x2(FL, lgeh, HoL);
            }
            diag_in = inH;
            sendH = Hm[K - 1]; sendF = F; sendFL = FL;
            if (lane == 31 && !last && (unsigned)(s - 63) < (unsigned)R) {      // bottom row of the strip
                v.bHF[s - 63] = __byte_perm(sendH, sendF, 0x7632);
                v.bFL[s - 63] = (uint16_t)(sendFL >> 16);
            }
        }
        __syncwarp();
    }
    const int vl = vc & 63;
    uint32_t h = __shfl_sync(0xffffffffu, cH, vl & 31), el = __shfl_sync(0xffffffffu, cEL, vl & 31), fl = __shfl_sync(0xffffffffu, cFL, vl & 31);
    if (vl >= 32) { h >>= 16; el >>= 16; fl >>= 16; } else { h &= 0xFFFFu; el &= 0xFFFFu; fl &= 0xFFFFu; }
    const uint32_t B = (uint32_t)sc.bias;
    return h == B && el < B && fl < B;
}

} // namespace gact

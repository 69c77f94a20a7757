// Two tiles per warp ("pair" geometry of the packed fast path): tile A lives in the low 16-bit halves of every register,
// tile B in the high halves -- the two halves of a packed DPX operand are independent cells of two DIFFERENT tiles of the
// same shape (kernel design notes), instead of two virtual lanes of one tile (gact_fast.cuh).
//
// Why: with 32 virtual lanes of K2 = 2K rows each instead of 64 of K rows
//   * the wavefront of a T x T tile takes R + 31 steps instead of R + 63 (fill / drain 9 % instead of 16 % at T = 320),
//   * the per-step overhead (3 shuffles, band stores, reference fetch, loop control) is paid once per 2*K2*32 cells,
//   * lane 0 takes the top boundary in BOTH halves: no cross-half PRMT of lane 31's values.
// The recurrence, the tagged scores and the 5-bit pointer are exactly fast_cell's; the band of each tile keeps the same
// per-virtual-lane window, two 32-bit words per (virtual lane, step): rows 0-5 in word 0, rows 6-11 in word 1, written with
// one 64-bit store per tile and step.
//
// Pairing is a scheduling decision of the kernel (same Q and R, both tiles eligible for the single-strip fast path, no N);
// the traceback of each tile runs on its own band, and a tile whose clean traceback is refused is recomputed alone.
#pragma once
#include "gact_fast.cuh"

namespace gact {

template <int K2, int BH = kBandHalf> struct PairGeom {
    static constexpr int kRows  = 32 * K2;          // max query rows and reference columns of a paired tile
    static constexpr int kW     = 2;                // band words per (virtual lane, step)
    static constexpr int kL     = 2 * BH + 1;
    // lane stride (in steps) minus the K2+1 steps a lane lags its neighbour must be odd: the 64-bit stores of a half-warp then
    // fall into 16 different 8-byte banks
    static constexpr int kLp    = kL + (((kL - K2 - 1) & 1) ? 0 : 1);
    static constexpr int kPWords = kRows + 64;      // P[32 + j] = refA[j] | refB[j] << 16
    static constexpr size_t kBandWords = (size_t)32 * kLp * kW;                 // per tile
    static constexpr size_t kSmemBytes = (2 * kBandWords + kPWords) * 4 + 4 * kRows;   // + staged byte sequences of both tiles
    using Trace = TraceGeo<K2, BH, kLp, kW>;
};

template <int K2, int BH = kBandHalf> struct PairSmemView {
    uint32_t* band[2];    // [32][kLp][2] each
    uint32_t* P;          // [kPWords]
    uint8_t*  sref[2];    // [kRows] each
    uint8_t*  sqry[2];
    __device__ explicit PairSmemView(unsigned char* base) {
        using G = PairGeom<K2, BH>;
        band[0] = reinterpret_cast<uint32_t*>(base);
        band[1] = band[0] + G::kBandWords;
        P = band[1] + G::kBandWords;
        sref[0] = reinterpret_cast<uint8_t*>(P + G::kPWords);
        sqry[0] = sref[0] + G::kRows;
        sref[1] = sqry[0] + G::kRows;
        sqry[1] = sref[1] + G::kRows;
    }
};

// Forward pass of two Q x R tiles.  Sequences staged (codes 0..3) in v.sref[t] / v.sqry[t].
// Returns the corner scores: low 16 bits tile A, high 16 bits tile B (biased by 32768 each), in all lanes.
template <int K2, int BH = kBandHalf>
__device__ void pair_forward(const FastConst& fc, const PairSmemView<K2, BH>& v, int Q, int R, int& scoreA, int& scoreB) {
    using G = PairGeom<K2, BH>;
    constexpr int NACC = (K2 + 2) / 3;              // three 5-bit pointers per 16-bit half of an accumulator
    const int lane = lane_id();
    for (int k = lane; k < G::kPWords; k += 32) {
        const int j = k - 32;
        const bool in = j >= 0 && j < R;
        v.P[k] = (in ? (uint32_t)v.sref[0][j] : kDummyRef) | ((in ? (uint32_t)v.sref[1][j] : kDummyRef) << 16);
    }
    uint32_t qq[K2], Hm[K2], E[K2], EL[K2];
#pragma unroll
    for (int r = 0; r < K2; r++) {
        const int i = K2 * lane + r;
        qq[r] = (i < Q) ? ((uint32_t)v.sqry[0][i] | ((uint32_t)v.sqry[1][i] << 16)) : (kDummyQry | (kDummyQry << 16));
        Hm[r] = fc.hm_init; E[r] = fc.e_init; EL[r] = fc.el_init;
    }
    __syncwarp();
    const BandMap<K2, BH> bm(Q, R);
    const int vc = (Q - 1) / K2, rc = (Q - 1) - vc * K2;
    const int steps = R + vc;                                            // the corner step is the last one (see fast_forward)
    const FastRegs kr(fc);
    uint32_t sendH = fc.hm_init, sendF = fc.f_top, sendFL = fc.fl_top;
    uint32_t diag_in = fc.hm_init;
    uint32_t rq_next = v.P[32 - lane];
    int t = -(K2 + 1) * lane + bm.c1;                                    // window position of my virtual lane at s = 0
    uint2* bpA = reinterpret_cast<uint2*>(v.band[0]) + lane * G::kLp + t;
    uint2* bpB = reinterpret_cast<uint2*>(v.band[1]) + lane * G::kLp + t;

    for (int s = 0; s < steps; s++) {
        uint32_t inH = __shfl_up_sync(0xffffffffu, sendH, 1);
        uint32_t F   = __shfl_up_sync(0xffffffffu, sendF, 1);
        uint32_t FL  = __shfl_up_sync(0xffffffffu, sendFL, 1);
        if (lane == 0) { inH = fc.hm_init; F = fc.f_top; FL = fc.fl_top; }   // top boundary of both tiles
        const uint32_t rq = rq_next;
        rq_next = v.P[32 + s + 1 - lane];
        uint32_t d = diag_in;
        uint32_t acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; a++) acc[a] = 0;
#pragma unroll
        for (int r = 0; r < K2; r++) {
            const uint32_t code = fast_cell<5, false>(kr, rq, qq[r], d, Hm[r], E[r], EL[r], F, FL);
            acc[r / 3] += code << (5 * (r % 3));
        }
        diag_in = inH;
        sendH = Hm[K2 - 1]; sendF = F; sendFL = FL;
        if ((unsigned)t < (unsigned)G::kL) {
            const uint32_t a2 = NACC > 2 ? acc[2] : 0u, a3 = NACC > 3 ? acc[3] : 0u;
            *bpA = make_uint2(__byte_perm(acc[0], acc[1], 0x5410), __byte_perm(a2, a3, 0x5410));
            *bpB = make_uint2(__byte_perm(acc[0], acc[1], 0x7632), __byte_perm(a2, a3, 0x7632));
        }
        t++; bpA++; bpB++;
    }
    __syncwarp();
    uint32_t corner = 0;
#pragma unroll
    for (int r = 0; r < K2; r++) if (r == rc) corner = Hm[r] - kr.diaga;
    const uint32_t cw = __shfl_sync(0xffffffffu, corner, vc);
    scoreA = (int)((cw & 0xFFFFu) >> 5) - fc.bias;
    scoreB = (int)(cw >> 21) - fc.bias;
}

} // namespace gact

// Packed score-only path of the first-tile filter (software/filter.cpp:28-122, :131-223): the 128x128 tiles that
// filter_body sends through g_BatchAlignmentSIMD with do_traceback = 0 and align_fields without start_end, i.e. the
// "max-cell" mode of DualAlignSIMD (software/Processor.cpp:502-552): score = global maximum of H, ref_max_pos = the
// first column that holds it, query_max_pos = the smallest row of that column that holds it.
//
// One warp computes TWO tiles at once: tile A lives in the low 16-bit halves of every register, tile B in the high
// halves.  Lane l owns query rows [4l, 4l+4) of both tiles and works on reference column s - l at step s (anti-diagonal
// wavefront, R + 31 steps); H / F / F_L of the row above cross lanes with one shuffle each per step.
//
// No traceback, so no tags: a value is (score + bias) << 7 in an unsigned half, the low 7 bits are zero throughout the
// recurrence (every addend is a multiple of 128).  The maximum is tracked per ROW as max over columns of
// (H | (127 - column)): the packed max then yields the row's best score AND the first column reaching it; the
// warp-wide max of those gives score and first column, the smallest row holding that value is the reference's row.
// Per packed cell pair: 10 ALU-pipe instructions (XOR, VIMNMX, 5 VIADDMNMX, 2 VIMNMX3, VIMNMX) + 5 IMAD.
//
// Preconditions (otherwise the tile goes to the exact path): uniform match/mismatch matrix, match > 0 > mismatch,
// Q, R <= 128, match * min(Q,R) + bias <= 511, no N, max-cell mode.  CPU twin: oracle/gact_oracle.c (score-only).
#pragma once
#include "gact_common.cuh"
#include "gact_exact.cuh"

namespace gact {

constexpr int kFiltMax = 128;                       // largest Q / R of a packed filter tile
constexpr int kFiltK   = 4;                         // rows per lane
constexpr int kFiltRaw = 96;                        // bytes of one packed 16-byte-rounded window of <= 128 bases

struct FilterConst {
    uint32_t zero;        // (B << 7), both halves            -- the clamp of the local alignment
    uint32_t hm_init;     // (B + mismatch) << 7                "H = 0" as a diagonal source
    uint32_t e_init;      // (B + go) << 7                      E(i,0) / F(0,j)
    uint32_t el_init;     // (B + lgo) << 7
    uint32_t pkc;         // (match - mismatch) << 7
    int32_t  negc;        // -((match - mismatch) << 7)         multiplier of the packed mismatch flags
    int32_t  mma;         // (mismatch << 7) * 65537            addends: value * 65537 adds to both halves
    int32_t  goa, lgoa;
    uint32_t geh, lgeh;   // ge << 7, lge << 7 as two's-complement halves (operand b of VIADDMNMX.U16x2)
    int32_t  bias, match, eligible;
    uint32_t one[4];      // all 1, opaque to the compiler (keeps the adds on the FMA pipe as IMADs)
};

__host__ inline FilterConst make_filter_const(const DevScoring& sc) {
    FilterConst f{};
    const int m = sc.match, mm = sc.mismatch, go = sc.go, ge = sc.ge, lgo = sc.lgo, lge = sc.lge;
    int B = -mm;
    if (-(go + ge) > B) B = -(go + ge);
    if (-(lgo + lge) > B) B = -(lgo + lge);
    B += 1;
    f.eligible = sc.uniform && m > 0 && mm < 0 && go <= ge && ge <= 0 && lgo <= lge && lge <= 0 && B + m <= 511;
    auto pk = [](int v) { return (uint32_t)(v & 0xFFFF) * 0x00010001u; };
    f.bias = B; f.match = m; f.one[0] = f.one[1] = f.one[2] = f.one[3] = 1;
    f.zero = pk(B << 7); f.hm_init = pk((B + mm) << 7); f.e_init = pk((B + go) << 7); f.el_init = pk((B + lgo) << 7);
    f.pkc = pk((m - mm) << 7); f.negc = -((m - mm) << 7);
    f.mma = (mm * 128) * 65537; f.goa = (go * 128) * 65537; f.lgoa = (lgo * 128) * 65537;
    f.geh = pk(ge * 128); f.lgeh = pk(lge * 128);
    return f;
}

struct __align__(16) FilterSmem {                   // per warp
    uint8_t  sref[2][kFiltMax];                     // staged codes (after reverse / complement) of tile A and tile B
    uint8_t  sqry[2][kFiltMax];
    uint32_t P[kFiltMax + 64];                      // P[32 + j] = refA[j] | refB[j] << 16, dummy base 5 outside [0, R)
    uint8_t  raw[2][kFiltRaw];                      // TMA landing windows (packed arena bytes)
    uint64_t mbar;
    uint64_t pad;
};

__device__ __forceinline__ bool filter_tile_ok(const FilterConst& fc, const TileJob& t) {
    return fc.eligible && !(t.flags & DARWIN_START_END) && t.Q >= 1 && t.R >= 1 && t.Q <= kFiltMax && t.R <= kFiltMax &&
           fc.match * min(t.Q, t.R) + fc.bias <= 511;
}

struct FilterHit { int score, ref_max_pos, query_max_pos; };

// Forward pass of one tile pair (same Q and R; sequences staged in sm).  Results are warp-uniform.
__device__ __forceinline__ void filter_pair_forward(const FilterConst& fc, FilterSmem& sm, int Q, int R,
                                                    FilterHit& ha, FilterHit& hb) {
    constexpr int K = kFiltK;
    const int lane = lane_id();
    for (int k = lane; k < kFiltMax + 64; k += 32) {
        const int j = k - 32;
        const bool in = (unsigned)j < (unsigned)R;
        sm.P[k] = (in ? (uint32_t)sm.sref[0][j] : 5u) | ((in ? (uint32_t)sm.sref[1][j] : 5u) << 16);
    }
    uint32_t qq[K], Hm[K], E[K], EL[K], best[K];
#pragma unroll
    for (int r = 0; r < K; r++) {
        const int i = K * lane + r;
        qq[r] = (i < Q) ? ((uint32_t)sm.sqry[0][i] | ((uint32_t)sm.sqry[1][i] << 16)) : 0x00060006u;
        Hm[r] = fc.hm_init; E[r] = fc.e_init; EL[r] = fc.el_init; best[r] = 0;
    }
    __syncwarp();
    const uint32_t zero = fc.zero, pkc = fc.pkc, negc = (uint32_t)fc.negc, mma = (uint32_t)fc.mma, goa = (uint32_t)fc.goa,
                   lgoa = (uint32_t)fc.lgoa, geh = fc.geh, lgeh = fc.lgeh;
    const uint32_t one0 = fc.one[0], one1 = fc.one[1], one2 = fc.one[2], one3 = fc.one[3];
    uint32_t sendH = fc.hm_init, sendF = fc.e_init, sendFL = fc.el_init;     // state below my last row (previous step)
    uint32_t diag_in = fc.hm_init;
    const int src = (lane + 31) & 31;
    const int steps = R + 31;
    uint32_t rq_next = sm.P[32 - lane];
    uint32_t codepk = (uint32_t)(127 + lane) * 0x00010001u;                  // 127 - column, column = s - lane
    int col = -lane;
    for (int s = 0; s < steps; s++) {
        uint32_t inH = __shfl_sync(0xffffffffu, sendH, src);
        uint32_t F   = __shfl_sync(0xffffffffu, sendF, src);
        uint32_t FL  = __shfl_sync(0xffffffffu, sendFL, src);
        if (lane == 0) { inH = fc.hm_init; F = fc.e_init; FL = fc.el_init; }
        const uint32_t rq = rq_next;
        rq_next = sm.P[32 + s + 1 - lane];
        const bool valid = (unsigned)col < (unsigned)R;                      // columns outside [0, R) track nothing:
        const uint32_t tmul = valid ? one3 : 0u, tadd = valid ? codepk : 0u; // H * 0 + 0 never beats a real entry
        uint32_t d = diag_in;
#pragma unroll
        for (int r = 0; r < K; r++) {
            const uint32_t x  = rq ^ qq[r];
            const uint32_t t  = __vminu2(x, 0x00010001u);                    // 1 = mismatch, per half
            const uint32_t sb = t * negc + pkc;                              // IMAD: (match - mismatch) << 7 or 0
            const uint32_t hd = __viaddmax_u16x2(d, sb, zero);               // max(Hdiag + s, 0)    Processor.cpp:298-299
            const uint32_t h1 = __vimax3_u16x2(hd, E[r], F);
            const uint32_t H  = __vimax3_u16x2(h1, EL[r], FL);               //                      :300-303
            d = Hm[r];
            Hm[r] = H * one0 + mma;
            const uint32_t Ho = H * one1 + goa, HoL = H * one2 + lgoa;
            E[r]  = __viaddmax_u16x2(E[r], geh, Ho);                         //                      :336-337
            F     = __viaddmax_u16x2(F, geh, Ho);                            //                      :363-364
            EL[r] = __viaddmax_u16x2(EL[r], lgeh, HoL);                      //                      :339-340
            FL    = __viaddmax_u16x2(FL, lgeh, HoL);                         //                      :365-366
            const uint32_t trk = H * tmul + tadd;                            // H | (127 - column)
            best[r] = __vmaxu2(best[r], trk);                                // vMaxH, :347 / :497-500
        }
        diag_in = inH;
        sendH = Hm[K - 1]; sendF = F; sendFL = FL;
        codepk -= 0x00010001u; col++;
    }
    // rows beyond Q do not exist; warp-wide maximum of the packed (score, 127 - first column) words
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < K; r++) { if (K * lane + r >= Q) best[r] = 0; m = __vmaxu2(m, best[r]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = __vmaxu2(m, __shfl_xor_sync(0xffffffffu, m, o));
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t mh = half ? (m >> 16) : (m & 0xFFFFu);
        uint32_t mine = 0;
#pragma unroll
        for (int r = 0; r < K; r++) mine |= ((half ? (best[r] >> 16) : (best[r] & 0xFFFFu)) == mh) ? (1u << r) : 0u;
        const uint32_t ball = __ballot_sync(0xffffffffu, mine != 0);
        const int l0 = __ffs(ball) - 1;                                      // smallest row wins (:538-548)
        const int r0 = __ffs(__shfl_sync(0xffffffffu, mine, l0)) - 1;
        FilterHit h;
        h.score = (int)(mh >> 7) - fc.bias; h.ref_max_pos = 127 - (int)(mh & 127u); h.query_max_pos = K * l0 + r0;
        if (half) hb = h; else ha = h;
    }
}

// One candidate of filter_body (a D-SOFT hit of one read) and what the kernels make of it (include/darwin_gpu.h).
// Request construction: filter.cpp:44-71 (forward), :154-181 (reverse complement).
__device__ __forceinline__ void filter_make_request(const DarwinFilterCand& c, int fts, DarwinTileReq& rq,
                                                    uint32_t& ref_tile_start, uint32_t& query_tile_start) {
    const uint32_t chr_end = c.chr_start + c.chr_len;
    const uint32_t ufts = (uint32_t)fts;
    ref_tile_start = ((uint32_t)(c.hit + ufts) < chr_end) ? c.hit : ((chr_end > ufts) ? chr_end - ufts : 0u);
    query_tile_start = ((uint64_t)(uint32_t)(c.offset + ufts) < (uint64_t)c.read_len) ? c.offset
                       : ((c.read_len > ufts) ? c.read_len - ufts : 0u);
    const uint32_t ref_tile_size = min(ufts, c.chr_len);
    const uint32_t query_tile_size = min(ufts, c.read_len);
    rq.ref_bases_start_addr = ref_tile_start;
    rq.query_bases_start_addr = c.strand ? c.read_addr + c.read_len - (query_tile_start + query_tile_size)
                                         : c.read_addr + query_tile_start;
    rq.score_threshold = 0; rq.index = 0;
    rq.ref_size = (uint16_t)ref_tile_size; rq.query_size = (uint16_t)query_tile_size;
    rq.max_tb_steps = (uint16_t)(2 * fts);
    rq.align_fields = c.strand ? (DARWIN_REVERSE_QUERY | DARWIN_COMPLEMENT_QUERY) : 0;
    rq.reserved[0] = rq.reserved[1] = rq.reserved[2] = 0;
}

} // namespace gact

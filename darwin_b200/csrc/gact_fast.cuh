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
//   (match * min(Q,R) + bias) < 2048, Q <= 64*K, R <= 64*K, corner traceback.  Tiles containing N run the HASN
//   instantiation (one extra ALU-pipe and one FMA-pipe instruction per cell pair) when mismatch <= sub_N <= 0.
#pragma once
#include <type_traits>
#include "gact_common.cuh"

namespace gact {

constexpr int kBandHalf = 32;                       // +-steps around the corner diagonal kept per virtual lane
constexpr int kBandHalfWide = 96;                   // multi-strip tiles beyond 512 (T = 1024): indel drift grows with the tile

// Band layout: virtual lane v keeps the trace words of the 2*kBandHalf+1 steps centred on the step at which it
// crosses the corner diagonal, contiguously: word(v, s) = band[v * kLp + t],  t = s - (K+1)*v - c2  in [0, kL).
// For a cell (i, j), i = K*v + r:  t = j - K*v + c1  -- linear in v and j, so the traceback needs no division.
template <int K, int BH = kBandHalf> struct FastGeom {
    static constexpr int kRows  = 64 * K;           // max query rows (and reference columns) of a fast tile
    static constexpr int kSteps = kRows + 63;
    static constexpr int kL     = 2 * BH + 1;
    static constexpr int kLp    = kL + (((kL - K - 1) & 1) ? 0 : 1);   // lane stride kLp-(K+1) odd -> conflict-free stores
    static constexpr int kW     = (K + 5) / 6;      // band words per (virtual lane, step): six 5-bit pointers per word (K = 8: two)
    static constexpr int kPWords = kRows + 128;     // one 16-bit word per reference column pair (j, j - 32): P[j + 32], see fast_forward
    static constexpr size_t kBandWords = (size_t)64 * kLp * kW;
    // the staged byte sequences alias the band: they are dead once P[] and the query tables are built, before the first band store
    static constexpr int kRawStride = (32 * K + 32 + 15) & ~15;     // one raw packed TMA window (<= kRows bases, 16-byte rounded)
    static constexpr int kSeqOff = 2 * kRawStride;  // behind the two raw TMA windows at the start of the region (KernelGeom)
    static constexpr size_t kSmemBytes = kBandWords * 4 + kPWords * 2;
    static_assert(kSeqOff + 2 * kRows <= kBandWords * 4, "staged sequences fit inside the band");
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
    uint32_t geh, lgeh;   // ge*32 and lge*32 as two's-complement halves (operand b of VIADDMNMX.U16x2)
    uint32_t goh, gofh, lgoh;   // go*32, go*32 + INS tag, lgo*32 + L tag as halves: the OPEN candidates as VIADDMNMX addends
    uint32_t fmark;       // F's "extended" marker, both halves
    int32_t  bias;        // B
    int32_t  max_score;   // largest corner score representable: 2047 - B
    int32_t  eligible;    // scoring admits the fast path
    int32_t  match;
    uint32_t nmul;        // (sub_N - mismatch) * U / 4: multiplier of the packed "N involved" flags (value 4 per half)
    int32_t  n_ok;        // tiles containing N may stay on the packed path (mismatch <= sub_N <= 0)
    uint32_t one[4];      // all 1, opaque to the compiler: `x * one[k] + c` stays an IMAD (fma pipe) instead of an ALU add
    uint32_t shl16;       // 65536, opaque likewise: lane 0's `x * shl16 + boundary` stays an IMAD instead of a shift + add
};

constexpr uint32_t FT_DEL = 0, FT_INS = 1, FT_DIAG = 2, FT_ZERO = 3, FT_L = 4;
constexpr uint32_t kMaskT = 0x001C001Cu, kMaskM = 0x00030003u, kMaskClean = 0xFFE0FFE0u;

// S = tag bits below the score: 5 (bits 4:2 source, 1 "F extended", 0 "E extended"; scores up to 2047 - bias) or
// 4 (bits 3:1 source, bit 0 "this gap chain was extended"; scores up to 4095 - bias: T = 1024 tiles at match = 2).
__host__ inline FastConst make_fast_const(const DevScoring& sc, int S = 5) {
    FastConst f{};
    const int m = sc.match, mm = sc.mismatch, go = sc.go, ge = sc.ge, lgo = sc.lgo, lge = sc.lge;
    int B = -mm;
    if (-(go + ge) > B) B = -(go + ge);
    if (-(lgo + lge) > B) B = -(lgo + lge);
    B += 1;
    f.eligible = sc.uniform && m > 0 && mm < 0 && go <= ge && ge < 0 && lgo <= lge && lge <= 0 && B < 512;
    auto pk = [](int v) { return (uint32_t)(v & 0xFFFF) * 0x00010001u; };
    const int U = 1 << S, ts = S - 3;                          // score unit, position of the 3-bit source field
    f.bias = B; f.match = m; f.max_score = (65536 / U - 1) - B - m; f.one[0] = f.one[1] = f.one[2] = f.one[3] = 1; f.shl16 = 65536u;
    f.zeroc = pk((B << S) | (FT_ZERO << ts));
    f.hm_init = pk(((B + mm) << S) | (FT_DIAG << ts));
    f.e_init = pk((B + go) << S);
    f.el_init = pk(((B + lgo) << S) | (FT_L << ts));
    f.f_top = pk(((B + go) << S) | (FT_INS << ts));
    f.fl_top = pk(((B + lgo) << S) | (FT_L << ts));
    f.pkc32 = pk((m - mm) << S);
    f.negc32 = -((m - mm) << S);
    f.diaga = (mm * U + (int)(FT_DIAG << ts)) * 65537;
    f.goa = (go * U) * 65537; f.gofa = (go * U + (int)(FT_INS << ts)) * 65537;
    f.gea = (ge * U) * 65537;
    f.lgoa = (lgo * U + (int)(FT_L << ts)) * 65537; f.lgea = (lge * U) * 65537;
    f.geh = pk(ge * U); f.lgeh = pk(lge * U);
    f.goh = pk(go * U); f.gofh = pk(go * U + (int)(FT_INS << ts)); f.lgoh = pk(lgo * U + (int)(FT_L << ts));
    f.fmark = S == 5 ? 0x00020002u : 0x00010001u;
    // N bases (Nt2Int code 4, Processor.cpp:21-46; sub_N for N against anything, :50-74): see fast_cell<S, true>
    f.n_ok = f.eligible && sc.subn <= 0 && sc.subn >= mm;
    f.nmul = (uint32_t)((sc.subn - mm) * U / 4);
    return f;
}

// Per-warp shared memory of the fast path, carved from the dynamic shared memory of the CTA.
template <int K> struct FastSmemView {
    uint32_t* band;       // [kSteps][kNB]
    uint16_t* P;          // [kPWords]
    uint8_t*  sref;       // [kRows]   (inside the band, see FastGeom)
    uint8_t*  sqry;       // [kRows]
    __device__ explicit FastSmemView(unsigned char* base) {
        band = reinterpret_cast<uint32_t*>(base);
        P = reinterpret_cast<uint16_t*>(band + FastGeom<K>::kBandWords);
        sref = base + FastGeom<K>::kSeqOff;
        sqry = sref + FastGeom<K>::kRows;
    }
};

// Band addressing shared by the forward pass and the traceback.
template <int K, int BH = kBandHalf> struct BandMap {
    int c1;               // t(i,j) = j - K*v + c1
    __device__ BandMap(int Q, int R, int shift = 0) {
        // virtual lane v crosses the corner diagonal (i - j = Q - R) at its middle row K*v + K/2:
        // step s_c(v) = (K+1)*v + K/2 - (Q - R); window t = s - s_c(v) + BH.
        // The band covers the diagonals (i - j) - (Q - R) in [-BH - K/2 - shift, BH + K/2 - 1 - shift]: shift > 0 moves it towards
        // paths that consume more query than reference on their way from the corner to the origin (insertion-rich reads).
        c1 = (Q - R) - K / 2 + BH - shift;
    }
    __device__ __forceinline__ int t_of(int j, int v) const { return j - K * v + c1; }
};

// Scoring constants of one tile in registers.
struct FastRegs {
    uint32_t zeroc, pkc32, negc32, diaga, gea, lgea, goh, gofh, lgoh, fmark, one0, one1, one2, one3, nmul;
    __device__ explicit FastRegs(const FastConst& fc)
        : zeroc(fc.zeroc), pkc32(fc.pkc32), negc32((uint32_t)fc.negc32), diaga((uint32_t)fc.diaga), gea((uint32_t)fc.gea),
          lgea((uint32_t)fc.lgea), goh(fc.goh), gofh(fc.gofh), lgoh(fc.lgoh), fmark(fc.fmark),
          one0(fc.one[0]), one1(fc.one[1]), one2(fc.one[2]), one3(fc.one[3]), nmul(fc.nmul) {}
};

// One packed cell pair (two int16 cells): recurrence of Processor.cpp:293-366 on tagged scores.
// d = Hm of the row above at the previous column (in), Hm of this row at the previous column (out).
// S = 5: the layout described above.  S = 4 (wide scores): source in bits 3:1, one shared "extended" bit 0; the 5-bit
// trace code (same as S = 5) is assembled from Hk's source, E's bit 0 and F's bit 0 (one more ALU op per cell pair).
//
// HASN: the tile contains N.  Reference N is code 4, query N is remapped to 12 when the rows are loaded, so an N on either
// side always counts as a "mismatch" (x != 0, also N against N) and bit 2 of (rq | qq) says "N involved"; the substitution
// score then becomes sub_N: sb + 4 * nmul = (sub_N - mismatch) * U on top of the mismatch already folded into d.
// core: sb = substitution addend of the cell pair ((match - mismatch) * U where the bases match, 0 otherwise; the mismatch
// score itself is folded into d)
// (Measured and dropped: carrying F's extend candidate Fx = (F | marker) + extend down the rows as its own max chain, which
// takes the OR and the IMAD off the F chain at the price of one more VIADDMNMX per cell pair: no gain on the tile kernel,
// -2 % on the extension kernel.)
template <int S = 5>
__device__ __forceinline__ uint32_t fast_cell_core(const FastRegs& k, uint32_t sb, uint32_t& d, uint32_t& Hm,
                                                   uint32_t& E, uint32_t& EL, uint32_t& F, uint32_t& FL) {
    const uint32_t hd = __viaddmax_u16x2(d, sb, k.zeroc);        // max(Hdiag + s, 0)          :298-299
    // (the row's own E and E_L first: F and F_L arrive through the dependent chain of the rows above)
    const uint32_t h1 = __vimax3_u16x2(hd, E, EL);
    const uint32_t Hk = __vimax3_u16x2(h1, F, FL);               // H with the winner's tag      :300-303
    uint32_t code, Hc;
    if (S == 5) {
        const uint32_t em = E | F;
        code = (Hk & kMaskT) | (em & kMaskM);
        Hc = Hk & kMaskClean;
    } else {
        const uint32_t fm = F & 0x00010001u;                     // masked first: F's top score bit must not spill over
        const uint32_t f2 = fm * k.one0 + fm, hk2 = Hk * k.one1 + Hk;          // x2 as IMADs
        const uint32_t em = (E | f2) & kMaskM;                   // E's source bits are 0 (DEL): bit 0 is its marker
        code = (hk2 & kMaskT) | em;
        Hc = Hk & 0xFFF0FFF0u;
    }
    d = Hm;
    Hm = Hc * k.one0 + k.diaga;                                  // IMADs: keeps the adds off the ALU pipe
    // gap updates max(gap + extend, H + open): the OPEN candidate takes the add of VIADDMNMX (H -> clean -> VIADDMNMX stays on
    // the ALU pipe: this is the dependent chain that runs down the rows of a step through F and F_L), the EXTEND candidate
    // is formed by an IMAD off that chain
    const uint32_t Ee = (E | 0x00010001u) * k.one1 + k.gea;      // ties extend                  :336-337,:353
    const uint32_t Fe = (F | k.fmark) * k.one2 + k.gea;          //                              :363-364,:369
    const uint32_t ELe = EL * k.one3 + k.lgea, FLe = FL * k.one3 + k.lgea;             //        :339-340, :365-366
    E  = __viaddmax_u16x2(Hc, k.goh, Ee);
    F  = __viaddmax_u16x2(Hc, k.gofh, Fe);
    EL = __viaddmax_u16x2(Hc, k.lgoh, ELe);
    FL = __viaddmax_u16x2(Hc, k.lgoh, FLe);
    return code;
}

template <int S = 5, bool HASN = false>
__device__ __forceinline__ uint32_t fast_cell(const FastRegs& k, uint32_t rq, uint32_t qq, uint32_t& d, uint32_t& Hm,
                                              uint32_t& E, uint32_t& EL, uint32_t& F, uint32_t& FL) {
    const uint32_t x  = rq ^ qq;
    const uint32_t t  = __vminu2(x, 0x00010001u);                // 1 = mismatch, per half
    uint32_t sb = t * k.negc32 + k.pkc32;                        // IMAD: (match-mismatch)*32 or 0
    if (HASN) sb = ((rq | qq) & 0x00040004u) * k.nmul + sb;      // LOP3 + IMAD: N involved -> sub_N
    return fast_cell_core<S>(k, sb, d, Hm, E, EL, F, FL);
}

// ---- table look-up of the mismatch flags (tiles without N) -----------------------------------------------------------------
// The query rows of a virtual lane are fixed for the whole tile, the reference base changes every step: each row keeps a
// 4-byte table "does reference base b differ from my query base" (one register per half), and ONE PRMT per cell pair picks
// the flag of the low half's row from table a and of the high half's row from table b -- instead of XOR + VMIN.  The
// selector is stored per reference column, ready made, in P[] (see fast_forward): nibble 0 = base of column j, nibble 2 =
// 4 + base of column j - 32, nibbles 1 and 3 = 8 (replicate the sign of a 0/1 byte: zero).  Columns outside the tile (the
// dummy base that mismatches everything) matter only while the wavefront fills: there the flag is derived from the step.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
// ptxas likes to re-derive loop-invariant per-lane values (lane tests, masks) inside the loop instead of keeping them in a
// register, on the very ALU pipe the loop is bound by; an empty asm makes the value opaque
template <class T> __device__ __forceinline__ void keep_in_register(T& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ uint32_t qry_table(uint32_t q) { return q < 4u ? (0x01010101u ^ (1u << (8u * q))) : 0x01010101u; }
__device__ __forceinline__ uint32_t ref_selector(uint32_t lo, bool in_lo, uint32_t hi, bool in_hi) {
    return 0x8480u + (in_lo ? lo : 0u) + ((in_hi ? hi : 0u) << 8);
}

// Forward pass of one tile.  Sequences must already be staged (codes 0..3) in v.sref / v.sqry.
// Returns the corner score H(Q-1, R-1) in all lanes.
// dummy bases outside the tile: they mismatch everything and (HASN) carry no "N involved" bit
constexpr uint32_t kDummyRef = 16u, kDummyQry = 32u;
// query base as the cell update wants it: N (4) becomes 12 in tiles that contain N (see fast_cell)
template <bool HASN> __device__ __forceinline__ uint32_t qry_code(uint32_t c) { return HASN ? (c | ((c & 4u) << 1)) : c; }

// State of one warp's wavefront between steps.
template <int K> struct FwdState {
    uint32_t qa[K], qb[K];            // !HASN: mismatch tables of my low / high virtual lane's rows; HASN: qa = packed query codes
    uint32_t Hm[K], E[K], EL[K];
    uint32_t sendH, sendF, sendFL;    // state below my last row (previous step)
    uint32_t diag_in;                 // Hm(row above, previous column)
    uint32_t rq_next;                 // P word of the next step
    int t_lo;                         // band window position of my low virtual lane
    uint32_t* bp;
};

// Steps [s0, s1) of the wavefront.  FILL: some virtual lanes are still left of column 0 (dummy base: mismatch forced).
template <int K, bool HASN, bool FILL>
__device__ __forceinline__ void fast_steps(const FastRegs& kr, FwdState<K>& st, const uint16_t* Pl, int s0, int s1,
                                           int src, uint32_t mulL, uint32_t addH, uint32_t addF, uint32_t addFL) {
    using G = FastGeom<K>;
    constexpr int kHiOff = 32 * (G::kLp - (K + 1)) * G::kW;
    const int lane = lane_id();
    for (int s = s0; s < s1; s++) {
        // values from the virtual lane above: rotate by one lane; lane 0 shifts lane 31's low half up and takes the top
        // boundary in its low half -- as one IMAD per value (x * 65536 + boundary in lane 0, x * 1 + 0 elsewhere)
        const uint32_t inH = __shfl_sync(0xffffffffu, st.sendH, src) * mulL + addH;
        uint32_t F  = __shfl_sync(0xffffffffu, st.sendF, src) * mulL + addF;
        uint32_t FL = __shfl_sync(0xffffffffu, st.sendFL, src) * mulL + addFL;
        const uint32_t rq = HASN ? __byte_perm(st.rq_next, 0u, 0x4140) : st.rq_next;   // HASN: codes of (j, j - 32) into the two halves
        st.rq_next = Pl[s + 1];                                          // prefetch next step's reference word
        uint32_t dm = 0;
        if (!HASN && FILL) dm = (s < lane ? 1u : 0u) | (s < lane + 32 ? 0x10000u : 0u);   // my column is left of the tile
        uint32_t d = st.diag_in;
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int r = 0; r < K; r++) {
            uint32_t code;
            if (HASN) {
                code = fast_cell<5, true>(kr, rq, st.qa[r], d, st.Hm[r], st.E[r], st.EL[r], F, FL);
            } else {
                uint32_t t = prmt(st.qa[r], st.qb[r], rq);               // 1 = mismatch, per half
                if (FILL) t |= dm;
                const uint32_t sb = t * kr.negc32 + kr.pkc32;            // IMAD: (match-mismatch)*32 or 0
                code = fast_cell_core<5>(kr, sb, d, st.Hm[r], st.E[r], st.EL[r], F, FL);
            }
            acc[r / 3] += code << (5 * (r % 3));
        }
        st.diag_in = inH;
        st.sendH = st.Hm[K - 1]; st.sendF = F; st.sendFL = FL;
        // band store: kW words per (virtual lane, step): rows 0-2 in bits 0-14, rows 3-5 in bits 16-30 (word 1: rows 6-11)
        if (G::kW == 1) {
            if ((unsigned)st.t_lo < (unsigned)G::kL) st.bp[0] = __byte_perm(acc[0], acc[1], 0x5410);
            if ((unsigned)(st.t_lo - 32 * (K + 1)) < (unsigned)G::kL) st.bp[kHiOff] = __byte_perm(acc[0], acc[1], 0x7632);
        } else {
            if ((unsigned)st.t_lo < (unsigned)G::kL)
                *reinterpret_cast<uint2*>(st.bp) = make_uint2(__byte_perm(acc[0], acc[1], 0x5410), __byte_perm(acc[2], acc[3], 0x5410));
            if ((unsigned)(st.t_lo - 32 * (K + 1)) < (unsigned)G::kL)
                *reinterpret_cast<uint2*>(st.bp + kHiOff) = make_uint2(__byte_perm(acc[0], acc[1], 0x7632), __byte_perm(acc[2], acc[3], 0x7632));
        }
        st.t_lo++; st.bp += G::kW;
    }
}

template <int K, bool HASN = false>
__device__ int fast_forward(const FastConst& fc, const FastSmemView<K>& v, int Q, int R, int band_shift = 0) {
    using G = FastGeom<K>;
    const int lane = lane_id();
    // per reference column: HASN: code pair P[32 + j] = r[j] | r[j-32] << 8, dummy base outside [0,R); otherwise the PRMT
    // selector of the pair (ref_selector)
    for (int k = lane; k < G::kPWords; k += 32) {
        const int j = k - 32;
        const bool in_lo = (unsigned)j < (unsigned)R, in_hi = (unsigned)(j - 32) < (unsigned)R;
        const uint32_t lo = in_lo ? v.sref[j] : kDummyRef, hi = in_hi ? v.sref[j - 32] : kDummyRef;
        v.P[k] = (uint16_t)(HASN ? (lo | (hi << 8)) : ref_selector(lo, in_lo, hi, in_hi));
    }
    FwdState<K> st;
#pragma unroll
    for (int r = 0; r < K; r++) {
        const int ilo = K * lane + r, ihi = K * (lane + 32) + r;
        const uint32_t qlo = ilo < Q ? v.sqry[ilo] : kDummyQry, qhi = ihi < Q ? v.sqry[ihi] : kDummyQry;
        if (HASN) { st.qa[r] = qry_code<true>(qlo) | (qry_code<true>(qhi) << 16); st.qb[r] = 0; }
        else      { st.qa[r] = qry_table(qlo); st.qb[r] = qry_table(qhi); }
        st.Hm[r] = fc.hm_init; st.E[r] = fc.e_init; st.EL[r] = fc.el_init;
    }
    __syncwarp();
    const BandMap<K> bm(Q, R, band_shift);
    const int vc = (Q - 1) / K, rc = (Q - 1) - vc * K, sc_step = R - 1 + vc;
    const FastRegs kr(fc);                                               // scoring constants in registers
    st.sendH = fc.hm_init; st.sendF = fc.f_top; st.sendFL = fc.fl_top;
    st.diag_in = fc.hm_init;
    int src = (lane + 31) & 31;
    uint32_t mulL = lane == 0 ? fc.shl16 : fc.one[0];             // both opaque to the compiler
    uint32_t addH = lane == 0 ? (fc.hm_init & 0xFFFFu) : 0u, addF = lane == 0 ? (fc.f_top & 0xFFFFu) : 0u,
             addFL = lane == 0 ? (fc.fl_top & 0xFFFFu) : 0u;
    keep_in_register(src); keep_in_register(mulL); keep_in_register(addH); keep_in_register(addF); keep_in_register(addFL);
    // the corner cell (Q-1, R-1) is the last valid cell of the wavefront (step R-1+vc): nothing after it is ever read,
    // so the loop ends there and the corner score is simply what its owner holds afterwards
    const int steps = sc_step + 1;
    const uint16_t* Pl = v.P + 32 - lane;                                // step s: column j = s - lane (low), j - 32 (high)
    st.rq_next = Pl[0];
    // band window of my two virtual lanes: t = s - (K+1)*v + c1 (t(i,j) with j = s - v)
    st.t_lo = -(K + 1) * lane + bm.c1;                                   // at s = 0
    st.bp = v.band + (lane * G::kLp + st.t_lo) * G::kW;                  // &band[(v_lo * kLp + t_lo) * kW]; hi: + 32*(kLp - (K+1)) steps
    if (HASN) {
        fast_steps<K, true, false>(kr, st, Pl, 0, steps, src, mulL, addH, addF, addFL);
    } else {
        // fill: until virtual lane 63 reaches column 0 some lanes chew the dummy base (flag OR-ed into the PRMT result);
        // afterwards every virtual lane is inside the tile or past its last column, where nothing it computes is ever read.
        // (Measured and dropped: unrolling the step loop by two and step loops specialised by band-window phase save 3 of 73
        // ALU-pipe instructions per step but gain nothing -- the loop is bound by its dependent chain and the issue rate --
        // and the larger code costs the extension kernel 12 %.)
        const int fill = min(steps, 63);
        fast_steps<K, false, true>(kr, st, Pl, 0, fill, src, mulL, addH, addF, addFL);
        fast_steps<K, false, false>(kr, st, Pl, fill, steps, src, mulL, addH, addF, addFL);
    }
    __syncwarp();
    uint32_t corner = 0;
#pragma unroll
    for (int r = 0; r < K; r++) if (r == rc) corner = st.Hm[r] - kr.diaga;
    // corner owner: virtual lane vc -> physical lane vc & 31, half vc >> 5
    uint32_t cw = __shfl_sync(0xffffffffu, corner, vc & 31);
    cw = (vc >= 32) ? (cw >> 16) : (cw & 0xFFFFu);
    return (int)(cw >> 5) - fc.bias;
}

// ---- multi-strip variant: tiles with more than 64*K rows or columns (T = 512, the 1984x960 / 960x1984 large tiles) ----
// Strips of 64*K query rows run one after the other; the bottom row of a strip (Hm, F, F_L of virtual lane 63) is kept
// per reference column in shared memory and becomes the top boundary of the next strip.  The traceback band goes to
// the warp's global scratch (same per-virtual-lane window layout, virtual lanes numbered across strips).
struct MultiSmemView {
    uint8_t*  sref;       // [kSeqSmem]
    uint8_t*  sqry;       // [kSeqSmem]
    uint32_t* bHF;        // [kMaxTile]  Hm (low half) | F (high half) below the previous strip, per column
    uint16_t* bFL;        // [kMaxTile]
    static constexpr size_t kBytes = 2 * kSeqSmem + kMaxTile * 4 + kMaxTile * 2;
    __device__ explicit MultiSmemView(unsigned char* base) {
        sref = base; sqry = base + kSeqSmem;
        bHF = reinterpret_cast<uint32_t*>(base + 2 * kSeqSmem);
        bFL = reinterpret_cast<uint16_t*>(base + 2 * kSeqSmem + kMaxTile * 4);
    }
};

template <int K, int BH = kBandHalf> __host__ __device__ inline size_t multi_band_bytes(int Q) {
    return (size_t)((Q + K - 1) / K + 64) * FastGeom<K, BH>::kLp * 4;
}

template <int K, int S = 5, int BH = kBandHalf, bool HASN = false>
__device__ int fast_forward_multi(const FastConst& fc, const MultiSmemView& v, uint32_t* gband, int Q, int R) {
    using G = FastGeom<K, BH>;
    const int lane = lane_id();
    const FastRegs kr(fc);
    const BandMap<K, BH> bm(Q, R);
    const int nstrips = (Q + 64 * K - 1) / (64 * K);
    const int vc = (Q - 1) / K, rc = (Q - 1) - vc * K;                   // global virtual lane / row of the corner
    const int sc_step = R - 1 + (vc & 63);                               // step of the corner inside the last strip
    int src = (lane + 31) & 31;
    uint32_t mulL = lane == 0 ? fc.shl16 : fc.one[0];                    // lane 0: x * 65536 + boundary (see fast_steps)
    keep_in_register(src); keep_in_register(mulL);
    const int steps = R + 63;
    constexpr int kHiOff = 32 * (G::kLp - (K + 1));
    const uint16_t* bHF16 = reinterpret_cast<const uint16_t*>(v.bHF);    // [2j] = Hm, [2j + 1] = F below the previous strip
    uint32_t corner = 0;
    if (!HASN) {
        // virtual lanes past the last column read up to 63 bytes behind the reference (nothing they compute is ever read):
        // keep those bytes valid PRMT indices
        for (int k = R + lane; k < R + 64; k += 32) v.sref[k] = 0;
        __syncwarp();
    }

    for (int strip = 0; strip < nstrips; strip++) {
        const int row0 = strip * 64 * K;
        uint32_t qa[K], qb[K], Hm[K], E[K], EL[K];
#pragma unroll
        for (int r = 0; r < K; r++) {
            const int ilo = row0 + K * lane + r, ihi = row0 + K * (lane + 32) + r;
            const uint32_t qlo = ilo < Q ? v.sqry[ilo] : kDummyQry, qhi = ihi < Q ? v.sqry[ihi] : kDummyQry;
            if (HASN) { qa[r] = qry_code<true>(qlo) | (qry_code<true>(qhi) << 16); qb[r] = 0; }
            else      { qa[r] = qry_table(qlo); qb[r] = qry_table(qhi); }
            Hm[r] = fc.hm_init; E[r] = fc.e_init; EL[r] = fc.el_init;
        }
        uint32_t sendH = fc.hm_init, sendF = fc.f_top, sendFL = fc.fl_top;
        uint32_t diag_in = fc.hm_init;
        // top boundary of this strip: addends of lane 0's low half (zero in the other lanes); strips below the first read
        // them per column from shared memory, one step ahead
        const bool top_smem = strip > 0 && lane == 0;
        uint32_t addH = lane == 0 ? (fc.hm_init & 0xFFFFu) : 0u, addF = lane == 0 ? (fc.f_top & 0xFFFFu) : 0u,
                 addFL = lane == 0 ? (fc.fl_top & 0xFFFFu) : 0u;
        uint32_t nH = addH, nF = addF, nFL = addFL;
        if (top_smem) { nH = bHF16[0]; nF = bHF16[1]; nFL = v.bFL[0]; }
        const int vg = strip * 64 + lane;                                // my low-half global virtual lane
        int t_lo = -lane - K * vg + bm.c1;                               // t(i,j) at s = 0 (j = -lane)
        uint32_t* bp = gband + (size_t)vg * G::kLp + t_lo;
        uint32_t rlo = (lane == 0 && R > 0) ? v.sref[0] : (HASN ? kDummyRef : 0u), rhi = HASN ? kDummyRef : 0u;   // bases of step 0

        // the corner is the last valid cell of the last strip: its loop ends there (nothing later is ever read)
        const int strip_steps = (strip == nstrips - 1) ? sc_step + 1 : steps;
        // FILL: the first 63 steps of a strip, while some virtual lanes are still left of column 0 (dummy base)
        auto run = [&](auto fill_c, int s0, int s1) {
            constexpr bool FILL = decltype(fill_c)::value;
            for (int s = s0; s < s1; s++) {
                if (top_smem) {
                    addH = nH; addF = nF; addFL = nFL;
                    if (s + 1 < R) { nH = bHF16[2 * (s + 1)]; nF = bHF16[2 * (s + 1) + 1]; nFL = v.bFL[s + 1]; }
                }
                const uint32_t inH = __shfl_sync(0xffffffffu, sendH, src) * mulL + addH;
                uint32_t F  = __shfl_sync(0xffffffffu, sendF, src) * mulL + addF;
                uint32_t FL = __shfl_sync(0xffffffffu, sendFL, src) * mulL + addFL;
                const uint32_t rq = HASN ? (rlo | (rhi << 16)) : (rhi * 256u + rlo + 0x8480u);   // !HASN: PRMT selector (ref_selector)
                uint32_t dm = 0;
                if (!HASN && FILL) dm = (s < lane ? 1u : 0u) | (s < lane + 32 ? 0x10000u : 0u);
                {   // prefetch the reference bases of step s+1: columns s+1-lane and s+1-lane-32
                    const int jl = s + 1 - lane, jh = jl - 32;
                    if (HASN || FILL) {
                        rlo = ((unsigned)jl < (unsigned)R) ? v.sref[jl] : (HASN ? kDummyRef : 0u);
                        rhi = ((unsigned)jh < (unsigned)R) ? v.sref[jh] : (HASN ? kDummyRef : 0u);
                    } else {
                        rlo = v.sref[jl]; rhi = v.sref[jh];               // both >= 0 after the fill; tail bytes zeroed above
                    }
                }
                uint32_t d = diag_in;
                uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
                for (int r = 0; r < K; r++) {
                    uint32_t code;
                    if (HASN) {
                        code = fast_cell<S, true>(kr, rq, qa[r], d, Hm[r], E[r], EL[r], F, FL);
                    } else {
                        uint32_t t = prmt(qa[r], qb[r], rq);             // 1 = mismatch, per half
                        if (FILL) t |= dm;
                        const uint32_t sb = t * kr.negc32 + kr.pkc32;
                        code = fast_cell_core<S>(kr, sb, d, Hm[r], E[r], EL[r], F, FL);
                    }
                    if (r < 3) acc0 += code << (5 * r); else acc1 += code << (5 * (r - 3));
                }
                diag_in = inH;
                sendH = Hm[K - 1]; sendF = F; sendFL = FL;
                if (lane == 31 && strip + 1 < nstrips && (unsigned)(s - 63) < (unsigned)R) {   // bottom row of the strip
                    v.bHF[s - 63] = __byte_perm(sendH, sendF, 0x7632);
                    v.bFL[s - 63] = (uint16_t)(sendFL >> 16);
                }
                if ((unsigned)t_lo < (unsigned)G::kL) __stcg(bp, __byte_perm(acc0, acc1, 0x5410));
                if ((unsigned)(t_lo - 32 * (K + 1)) < (unsigned)G::kL) __stcg(bp + kHiOff, __byte_perm(acc0, acc1, 0x7632));
                t_lo++; bp++;
            }
        };
        const int fill = min(strip_steps, 63);
        run(std::true_type{}, 0, fill);
        run(std::false_type{}, fill, strip_steps);
        __syncwarp();
        if (strip == nstrips - 1) {
#pragma unroll
            for (int r = 0; r < K; r++) if (r == rc) corner = Hm[r] - kr.diaga;
        }
    }
    const int vl = vc & 63;
    uint32_t cw = __shfl_sync(0xffffffffu, corner, vl & 31);
    cw = (vl >= 32) ? (cw >> 16) : (cw & 0xFFFFu);
    return (int)(cw >> S) - fc.bias;
}

enum : int { FAST_OK = 0, FAST_LFLAG = 1, FAST_BAND = 2 };

// Traceback over the band (Processor.cpp:585-716 with the clean rule), executed by the WHOLE warp with identical state in
// every lane.  GLOBAL: the band lives in the warp's global scratch (multi-strip tiles) instead of shared memory.
//
// The walk itself is a dependent chain, but most of it consists of runs of M along a diagonal (85 % of the ops at 15 %
// error).  In the DIAG state lane k therefore probes the cell k steps up the diagonal, (i-k, j-k); one ballot gives the
// length of the run of cells whose pointer is DIAG, and the whole run is emitted and skipped at once.  Cells that are not
// M (gap steps, the end of the path, band exits) are handled one at a time by the generic step below, which every lane
// executes redundantly.  Only lane 0's sink writes.
// Returns FAST_OK, or the reason the tile must be recomputed by the exact path (warp-uniform).
// Geometry of a stored band as the traceback sees it: K rows per virtual lane, window of 2*BH+1 steps per virtual lane with
// lane stride LP (in steps), W words per (virtual lane, step) -- rows 0-5 in word 0, rows 6-11 in word 1.
template <int K_, int BH_, int LP_, int W_ = 1> struct TraceGeo {
    static constexpr int K = K_, BH = BH_, kL = 2 * BH_ + 1, kLp = LP_, W = W_;
};

template <class GEO, bool GLOBAL, class Sink>
__device__ int fast_traceback_g(const uint32_t* band, int Q, int R, int max_tb, TileOut& out, Sink& sink, int band_shift);

template <int K, bool GLOBAL, class Sink, int BH = kBandHalf>
__device__ __forceinline__ int fast_traceback(const uint32_t* band, int Q, int R, int max_tb, TileOut& out, Sink& sink, int band_shift = 0) {
    return fast_traceback_g<TraceGeo<K, BH, FastGeom<K, BH>::kLp, GLOBAL ? 1 : FastGeom<K, BH>::kW>, GLOBAL, Sink>(band, Q, R, max_tb, out, sink, band_shift);
}

template <class GEO, bool GLOBAL, class Sink>
__device__ int fast_traceback_g(const uint32_t* band, int Q, int R, int max_tb, TileOut& out, Sink& sink, int band_shift) {
    using G = GEO;
    constexpr int K = GEO::K, BH = GEO::BH;
    const BandMap<K, BH> bm(Q, R, band_shift);
    const int lane = lane_id();
    // 5-bit pointer of cell (row r of virtual lane v, window position t)
    auto code_at = [&](int v, int r, int t) -> uint32_t {
        constexpr uint32_t kSh = 0u | (5u << 5) | (10u << 10) | (16u << 15) | (21u << 20) | (26u << 25);
        const int hi = (GEO::W > 1 && r >= 6) ? 1 : 0;
        uint32_t w;
        if (GLOBAL) w = __ldcg(band + ((size_t)v * G::kLp + (size_t)t) * GEO::W + hi);
        else        w = band[(v * G::kLp + t) * GEO::W + hi];                // shared memory: 32-bit index arithmetic
        return (w >> ((kSh >> (5 * (r - 6 * hi))) & 31u)) & 31u;
    };
    // (bit position of row r inside a band word: rows 0-2 in bits 0-14, rows 3-5 in bits 16-30 -- see code_at)
    // i = i0 - is, j = j0 - js: the loop of Processor.cpp:613-618 runs while is < min(Q, max_tb) and js < min(R, max_tb)
    const int lim_i = min(Q, max_tb), lim_j = min(R, max_tb);
    int i = Q - 1, j = R - 1;
    int left_i = lim_i, left_j = lim_j;           // steps still allowed in each direction
    uint32_t where = FT_DIAG, st = FT_DIAG;
    bool off_band = false;
    constexpr uint32_t kOffBand = 0x80000000u;                                       // "this cell is outside the stored band"
    for (;;) {
        const int lim = min(left_i, left_j);
        if (lim <= 0) break;
        uint32_t code;                                                               // pointer of the cell the generic step handles
        if (where == FT_DIAG) {
            // probe the diagonal: lane k looks at cell (i - k, j - k)
            const int ii = i - lane, jj = j - lane;
            const bool ok = lane < lim;                                              // implies ii >= 0 and jj >= 0
            const unsigned uii = ok ? (unsigned)ii : 0u;
            const int v = (int)(uii / (unsigned)K), r = (int)uii - v * K;
            const int t = bm.t_of(jj, v);
            const uint32_t c = !ok ? 0u : ((unsigned)t < (unsigned)G::kL) ? code_at(v, r, t) : kOffBand;
            const uint32_t is_m = __ballot_sync(0xffffffffu, ok && (c >> 2) == FT_DIAG);
            const int run = (is_m == 0xffffffffu) ? 32 : __ffs(~is_m) - 1;
            if (run > 0) {                                                           // DIAG pointers in DIAG state: M, stay in DIAG
                sink.run_m(run);
                i -= run; j -= run; left_i -= run; left_j -= run;
            }
            if (run >= lim || run == 32) continue;                                   // step limit reached / the diagonal goes on
            // lane `run` has already loaded the cell that ends the run: it is handled right here instead of being probed again
            code = __shfl_sync(0xffffffffu, c, run);
            if (code == kOffBand) { off_band = true; break; }
        } else {
            // gap states: one cell at a time, same in every lane
            const int v = i / K, r = i - v * K;
            const int t = bm.t_of(j, v);
            if ((unsigned)t >= (unsigned)G::kL) { off_band = true; break; }
            code = code_at(v, r, t);
        }
        // a DIAG-state cell whose pointer is DEL/INS switches state and is re-read by the reference (:628-633):
        // nothing moves in between, so the gap step is taken right away
        st = (where == FT_DIAG) ? (code >> 2) : where;
        if (st >= FT_ZERO) break;                                                // ZERO: path ends (:640-642); L: exact rerun
        sink((0x36u >> (2 * st)) & 3u);                                          // DEL -> D (2), INS -> I (1), DIAG -> M (3)
        // next state: gaps stay open while their "extended" bit is set (:648-653, :662-667): bit 0 for DEL, bit 1 for INS
        where = ((code >> st) & (st < FT_DIAG ? 1u : 0u)) ? st : FT_DIAG;
        const int mv_left = (st != FT_INS), mv_up = (st != FT_DEL);
        j -= mv_left; left_j -= mv_left;
        i -= mv_up; left_i -= mv_up;
    }
    if (off_band) return FAST_BAND;                                              // path left the stored band
    if (min(left_i, left_j) > 0 && st == FT_L) return FAST_LFLAG;                // long-gap candidate met in DIAG state
    out.query_offset = lim_i - left_i; out.ref_offset = lim_j - left_j; out.total = sink.count(); out.tflags = 0;
    return FAST_OK;
}

} // namespace gact

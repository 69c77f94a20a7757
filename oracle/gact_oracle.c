/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product (see gact_oracle.h).
 *
 * CPU restatement of the GACT alignment-extension path of yatisht/darwin.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/software).  Parity: PINNED against oracle/_ref (the compiled
 * reference) by tests/test_oracle_vs_ref.py and tests/golden/.
 *
 * The reference's AVX2 kernel reads two uninitialised vectors (vF_La, vF_La_ext,
 * Processor.cpp:259-260 used :405-408,:444); this restatement follows the
 * "patched" flavour in which they start as vF_L / vF_L_ext.  Only
 * trace bits 2048/4096 depend on it, and they are read only while a traceback is
 * in the long-insertion state (flag GACT_TILE_LONG_INS).
 *
 * Scoring precondition of the closed forms (STREAM, CLEAN): gap_open <= gap_extend <= 0
 * and long_gap_open <= long_gap_extend <= 0.  The STRIPED rule has no precondition.
 */
#include "gact_oracle.h"

#include <stdlib.h>
#include <string.h>

/* trace word bits, reference layout (Processor.h:21-34) */
#define T_ZERO      0
#define T_DEL       1
#define T_INS       2
#define T_DEL_L     4
#define T_INS_L     8
#define T_DIAG      16
#define B_DIAG_DEL  32
#define B_DEL       64
#define B_DIAG_INS  128
#define B_INS       256
#define B_DIAG_DEL_L 512
#define B_DEL_L     1024
#define B_DIAG_INS_L 2048
#define B_INS_L     4096
#define M_T         8160   /* TRACEBACK_T_MASK   */
#define M_F         7807   /* TRACEBACK_F_MASK   */
#define M_FL        2047   /* TRACEBACK_F_L_MASK */
#define NEG_INF16   ((int16_t)(-16384))   /* Processor.cpp:13 */

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* Processor.cpp:21-46 */
int gact_nt2int(char nt, int complement) {
    switch (nt) {
        case 'a': case 'A': return complement ? 3 : 0;
        case 'c': case 'C': return complement ? 2 : 1;
        case 'g': case 'G': return complement ? 1 : 2;
        case 't': case 'T': return complement ? 0 : 3;
        default: return 4;
    }
}

/* Processor.cpp:48-80 */
void gact_scoring_init(GactScoring* sc, const DarwinScoring* s) {
    int AA = s->sub_AA, AC = s->sub_AC, AG = s->sub_AG, AT = s->sub_AT, CC = s->sub_CC, CG = s->sub_CG,
        CT = s->sub_CT, GG = s->sub_GG, GT = s->sub_GT, TT = s->sub_TT, N = s->sub_N;
    int m[25] = { AA, AC, AG, AT, N,  AC, CC, CG, CT, N,  AG, CG, GG, GT, N,  AT, CT, GT, TT, N,  N, N, N, N, N };
    memcpy(sc->sub, m, sizeof(m));
    sc->go = s->gap_open; sc->ge = s->gap_extend; sc->lgo = s->long_gap_open; sc->lge = s->long_gap_extend;
    int t[11] = { AA, AC, AG, AT, CC, CG, CT, GG, GT, TT, N };
    memcpy(sc->tri, t, sizeof(t));
}

typedef struct TileSeq { uint8_t* q; uint8_t* r; int Q, R; } TileSeq;

/* sequence fetch with the request's flags: Processor.cpp:105-106 (query), :276-277 (reference) */
static void fetch_seqs(const char* dram, const DarwinTileReq* req, TileSeq* s) {
    int Q = req->query_size, R = req->ref_size;
    int rr = (req->align_fields >> 4) & 1, cr = (req->align_fields >> 3) & 1;
    int rq = (req->align_fields >> 2) & 1, cq = (req->align_fields >> 1) & 1;
    s->Q = Q; s->R = R;
    s->q = (uint8_t*)malloc((size_t)Q + 1); s->r = (uint8_t*)malloc((size_t)R + 1);
    for (int i = 0; i < Q; i++) {
        uint64_t a = rq ? req->query_bases_start_addr + (uint64_t)(Q - 1) - i : req->query_bases_start_addr + i;
        s->q[i] = (uint8_t)gact_nt2int(dram[a], cq);
    }
    for (int j = 0; j < R; j++) {
        uint64_t a = rr ? req->ref_bases_start_addr + (uint64_t)(R - 1) - j : req->ref_bases_start_addr + j;
        s->r[j] = (uint8_t)gact_nt2int(dram[a], cr);
    }
}

/* ------------------------------------------------------------------------------------------
 * Rule STRIPED: literal scalar emulation of DualAlignSIMD (Processor.cpp:164-566).
 * A "vector" is int16_t[16]; element `l` of the vector at segment index `t` is query row
 * l*segLen + t (Processor.cpp:96-111).  Output: trace words in natural [j*Q + i] order.
 * ------------------------------------------------------------------------------------------ */
typedef int16_t vec16[16];

static void v_shift(vec16 v, int16_t ins) {           /* _mm256_slli_si256_rpl(v,2) + insert at lane 0 */
    for (int l = 15; l > 0; l--) v[l] = v[l - 1];
    v[0] = ins;
}

static void tile_striped(const GactScoring* sc, const TileSeq* s, int start_end,
                         uint16_t* trace_out, int* score_out, int* end_query_out, int* end_ref_out) {
    const int Q = s->Q, R = s->R;
    const int segLen = (Q + 15) / 16;                                        /* :174 */
    const int16_t go = (int16_t)sc->go, ge = (int16_t)sc->ge, lgo = (int16_t)sc->lgo, lge = (int16_t)sc->lge;
    size_t vsz = (size_t)segLen * sizeof(vec16);
    vec16* prof = (vec16*)malloc(5 * vsz);                                   /* CreateVProfile :87-115 */
    for (int k = 0; k < 5; k++)
        for (int t = 0; t < segLen; t++)
            for (int l = 0; l < 16; l++) {
                int i = t + l * segLen;
                prof[k * segLen + t][l] = (int16_t)(i >= Q ? 0 : sc->sub[5 * k + s->q[i]]);
            }
    vec16* bufs = (vec16*)calloc(10, vsz);
    vec16 *pvHStore = bufs, *pvHLoad = bufs + segLen, *pvE = bufs + 2 * segLen, *pvEaStore = bufs + 3 * segLen,
          *pvEaLoad = bufs + 4 * segLen, *pvE_L = bufs + 5 * segLen, *pvE_LaStore = bufs + 6 * segLen,
          *pvE_LaLoad = bufs + 7 * segLen, *pvHT = bufs + 8 * segLen, *pvHMax = bufs + 9 * segLen;
    vec16* trace = (vec16*)calloc((size_t)segLen * R, sizeof(vec16));        /* result->trace :176 */
#define TR(j, t) trace[(size_t)(j) * segLen + (t)]
    for (int t = 0; t < segLen; t++)
        for (int l = 0; l < 16; l++) {                                        /* :229-239 */
            pvHStore[t][l] = 0; pvE[t][l] = go; pvEaStore[t][l] = go; pvE_L[t][l] = lgo; pvE_LaStore[t][l] = lgo;
            TR(0, t)[l] = B_DIAG_DEL + B_DIAG_DEL_L;
        }
    int16_t score = 0; int end_ref = 0, end_query = 0;
    vec16 vMaxH, vMaxHUnit;
    for (int l = 0; l < 16; l++) { vMaxH[l] = 0; vMaxHUnit[l] = 0; }

    int j;
    for (j = 0; j < R; j++) {                                                 /* :242 */
        vec16 vEF_opn, vF, vF_ext, vFa, vFa_ext, vH, vEF_L_opn, vF_L, vF_L_ext, vF_La, vF_La_ext, vHp;
        for (int l = 0; l < 16; l++) { vF[l] = go; vF_L[l] = lgo; vEF_opn[l] = 0; vEF_L_opn[l] = 0; vF_ext[l] = 0; vF_L_ext[l] = 0; }
        memcpy(vH, pvHStore[segLen - 1], sizeof(vec16));                      /* :272-273 */
        v_shift(vH, 0);
        const vec16* vP = prof + (size_t)s->r[j] * segLen;                    /* :276-277 */
        if (end_ref == j - 2) { vec16* tmp = pvHMax; pvHMax = pvHLoad; pvHLoad = pvHStore; pvHStore = tmp; }   /* :279-284 */
        else { vec16* tmp = pvHLoad; pvHLoad = pvHStore; pvHStore = tmp; }
        { vec16* tmp = pvEaLoad; pvEaLoad = pvEaStore; pvEaStore = tmp; }
        { vec16* tmp = pvE_LaLoad; pvE_LaLoad = pvE_LaStore; pvE_LaStore = tmp; }

        for (int t = 0; t < segLen; t++) {                                    /* main loop :293-380 */
            for (int l = 0; l < 16; l++) {
                int16_t E = pvE[t][l], EL = pvE_L[t][l];
                int16_t Hdag = (int16_t)(vH[l] + vP[t][l]); if (Hdag < 0) Hdag = 0;
                int16_t H = Hdag; if (E > H) H = E; if (vF[l] > H) H = vF[l]; if (EL > H) H = EL; if (vF_L[l] > H) H = vF_L[l];
                pvHStore[t][l] = H;
                int16_t T = (H == Hdag) ? (H == 0 ? T_ZERO : T_DIAG) : ((H == vF[l]) ? T_INS : T_DEL);    /* :309-317 */
                if (H == vF_L[l]) T = T_INS_L;                                /* :318-321 */
                if (H == EL) T = T_DEL_L;                                     /* :322-325 */
                pvHT[t][l] = T;
                TR(j, t)[l] = (int16_t)(T | TR(j, t)[l]);
                if (H > vMaxH[l]) vMaxH[l] = H;
                int16_t opn = (int16_t)(H + go), opnL = (int16_t)(H + lgo);   /* :332-333 */
                vEF_opn[l] = opn; vEF_L_opn[l] = opnL;
                int16_t E_ext = (int16_t)(E + ge); pvE[t][l] = opn > E_ext ? opn : E_ext;           /* :336-338 */
                int16_t EL_ext = (int16_t)(EL + lge); pvE_L[t][l] = opnL > EL_ext ? opnL : EL_ext; /* :339-341 */
                int16_t Ea_ext = (int16_t)(pvEaLoad[t][l] + ge), ELa_ext = (int16_t)(pvE_LaLoad[t][l] + lge);
                pvEaStore[t][l] = opn > Ea_ext ? opn : Ea_ext;                /* :343-350 */
                pvE_LaStore[t][l] = opnL > ELa_ext ? opnL : ELa_ext;
                if (j + 1 < R)                                                /* :352-359 */
                    TR(j + 1, t)[l] = (int16_t)((opn > Ea_ext ? B_DIAG_DEL : B_DEL) | (opnL > ELa_ext ? B_DIAG_DEL_L : B_DEL_L));
                vF_ext[l] = (int16_t)(vF[l] + ge); vF[l] = opn > vF_ext[l] ? opn : vF_ext[l];           /* :363-364 */
                vF_L_ext[l] = (int16_t)(vF_L[l] + lge); vF_L[l] = opnL > vF_L_ext[l] ? opnL : vF_L_ext[l];
                if (t + 1 < segLen)                                           /* :367-376 */
                    TR(j, t + 1)[l] = (int16_t)(TR(j, t + 1)[l] | (opn > vF_ext[l] ? B_DIAG_INS : B_INS) |
                                                (opnL > vF_L_ext[l] ? B_DIAG_INS_L : B_INS_L));
                vH[l] = pvHLoad[t][l];                                        /* :379 */
            }
        }

        /* lazy-F loop :383-497 (patched flavour: the long-gap twins start like the short ones) */
        memcpy(vFa_ext, vF_ext, sizeof(vec16)); memcpy(vFa, vF, sizeof(vec16));
        memcpy(vF_La_ext, vF_L_ext, sizeof(vec16)); memcpy(vF_La, vF_L, sizeof(vec16));
        int done = 0;
        for (int k = 0; k < 16 && !done; k++) {
            memcpy(vHp, pvHLoad[segLen - 1], sizeof(vec16)); v_shift(vHp, 0);             /* :386-387 */
            v_shift(vEF_opn, go); v_shift(vF_ext, NEG_INF16); v_shift(vF, go);
            v_shift(vFa_ext, NEG_INF16); v_shift(vFa, go);
            v_shift(vEF_L_opn, lgo); v_shift(vF_L_ext, NEG_INF16); v_shift(vF_L, lgo);
            v_shift(vF_La_ext, NEG_INF16); v_shift(vF_La, lgo);
            for (int t = 0; t < segLen; t++) {
                int any_f = 0, any_fl = 0;
                for (int l = 0; l < 16; l++) {
                    int16_t H = pvHStore[t][l];
                    if (vF[l] > H) H = vF[l];
                    if (vF_L[l] > H) H = vF_L[l];
                    pvHStore[t][l] = H;                                       /* :410-413 */
                    int16_t Hp = (int16_t)(vHp[l] + vP[t][l]); if (Hp < 0) Hp = 0; vHp[l] = Hp;   /* :422-423 */
                    int case1 = (H == Hp), case2 = (H == vF[l]), case3 = (H == vF_L[l]);
                    int16_t T = pvHT[t][l];
                    if (!case1 && case2) T = T_INS;                           /* :427,:430 */
                    if (!(case1 || case2) && case3) T = T_INS_L;              /* :428,:431 */
                    pvHT[t][l] = T;
                    int16_t w = (int16_t)((TR(j, t)[l] & M_T) | T);           /* :433-436 */
                    if (H > vMaxH[l]) vMaxH[l] = H;
                    w = (int16_t)((w & M_F) | (vEF_opn[l] > vFa_ext[l] ? B_DIAG_INS : B_INS));          /* :441-447 */
                    w = (int16_t)((w & M_FL) | (vEF_L_opn[l] > vF_La_ext[l] ? B_DIAG_INS_L : B_INS_L)); /* :448-449 */
                    TR(j, t)[l] = w;
                    vEF_opn[l] = (int16_t)(H + go); vF_ext[l] = (int16_t)(vF[l] + ge);    /* :453-456 */
                    vEF_L_opn[l] = (int16_t)(H + lgo); vF_L_ext[l] = (int16_t)(vF_L[l] + lge);
                    int16_t Ea_ext = (int16_t)(pvEaLoad[t][l] + ge), ELa_ext = (int16_t)(pvE_LaLoad[t][l] + lge);
                    pvEaStore[t][l] = vEF_opn[l] > Ea_ext ? vEF_opn[l] : Ea_ext;          /* :458-466 */
                    pvE_LaStore[t][l] = vEF_L_opn[l] > ELa_ext ? vEF_L_opn[l] : ELa_ext;
                    if (j + 1 < R)                                            /* :467-474 */
                        TR(j + 1, t)[l] = (int16_t)((vEF_opn[l] > Ea_ext ? B_DIAG_DEL : B_DEL) |
                                                    (vEF_L_opn[l] > ELa_ext ? B_DIAG_DEL_L : B_DEL_L));
                    if (vF_ext[l] >= vEF_opn[l]) any_f = 1;                   /* :476-483 */
                    if (vF_L_ext[l] >= vEF_L_opn[l]) any_fl = 1;
                }
                if (!any_f && !any_fl) { done = 1; break; }                   /* :486 goto end */
                for (int l = 0; l < 16; l++) {                                /* :488-495 */
                    vF[l] = vF_ext[l];
                    vFa_ext[l] = (int16_t)(vFa[l] + ge); vFa[l] = vEF_opn[l] > vFa_ext[l] ? vEF_opn[l] : vFa_ext[l];
                    vF_L[l] = vF_L_ext[l];
                    vF_La_ext[l] = (int16_t)(vF_La[l] + lge); vF_La[l] = vEF_L_opn[l] > vF_La_ext[l] ? vEF_L_opn[l] : vF_La_ext[l];
                    vHp[l] = pvHLoad[t][l];
                }
            }
        }
        {                                                                     /* :502-509 */
            int gt = 0; int16_t hm = vMaxH[0];
            for (int l = 0; l < 16; l++) { if (vMaxH[l] > vMaxHUnit[l]) gt = 1; if (vMaxH[l] > hm) hm = vMaxH[l]; }
            if (gt) { score = hm; for (int l = 0; l < 16; l++) vMaxHUnit[l] = score; end_ref = j; }
        }
    }

    if (start_end) {                                                          /* :514-517 */
        score = pvHStore[(Q - 1) % segLen][(Q - 1) / segLen];
        end_query = Q - 1; end_ref = R - 1;                                   /* :544-547 */
    } else {                                                                  /* :518-541 */
        if (end_ref == j - 1) { vec16* tmp = pvHMax; pvHMax = pvHStore; pvHStore = tmp; }
        else if (end_ref == j - 2) { vec16* tmp = pvHMax; pvHMax = pvHLoad; pvHLoad = tmp; }
        end_query = Q - 1;
        for (int t = 0; t < segLen; t++)
            for (int l = 0; l < 16; l++)
                if (pvHMax[t][l] == score) { int temp = t + l * segLen; if (temp < end_query) end_query = temp; }
    }
    *score_out = score; *end_query_out = end_query; *end_ref_out = end_ref;
    if (trace_out)
        for (int jj = 0; jj < R; jj++)
            for (int i = 0; i < Q; i++)
                trace_out[(size_t)jj * Q + i] = (uint16_t)TR(jj, i % segLen)[i / segLen];       /* :614 */
#undef TR
    free(trace); free(bufs); free(prof);
}

/* ------------------------------------------------------------------------------------------
 * Rule STREAM and rule CLEAN: one top-to-bottom pass per column.
 * STREAM reproduces the striped kernel's tie-breaking with O(1) extra state per vertical chain:
 *   f0/fl0  own-lane chains (restart at every row i with i % segLen == 0, Processor.cpp:265-269)
 *   fc/flc  best chain carried in from lanes above (lazy-F passes, :385-408), kf/kfl = lane distance
 *           of the farthest carried chain attaining it (later passes override earlier ones, :424-431)
 *   E'/EL'  the main loop's E, built from the not-yet-corrected H (:336-341)
 * ------------------------------------------------------------------------------------------ */
static void tile_scalar(const GactScoring* sc, const TileSeq* s, int start_end, int exact,
                        uint16_t* trace, int* score_out, int* end_query_out, int* end_ref_out) {
    const int Q = s->Q, R = s->R, go = sc->go, ge = sc->ge, lgo = sc->lgo, lge = sc->lge;
    const int segLen = (Q + 15) / 16;
    const int NINF = -(1 << 28);
    int* Hprev = (int*)calloc((size_t)Q + 1, sizeof(int));   /* true H of column j-1 */
    int* Hcur = (int*)calloc((size_t)Q + 1, sizeof(int));
    int* Ep = (int*)malloc(sizeof(int) * Q), *ELp = (int*)malloc(sizeof(int) * Q);   /* main-loop E', EL' */
    int* Ea = (int*)malloc(sizeof(int) * Q), *ELa = (int*)malloc(sizeof(int) * Q);   /* true E, E_L */
    uint8_t* Eo = (uint8_t*)malloc(Q), *ELo = (uint8_t*)malloc(Q);
    for (int i = 0; i < Q; i++) { Ep[i] = Ea[i] = go; ELp[i] = ELa[i] = lgo; Eo[i] = ELo[i] = 1; }
    int best = 0, best_j = 0, best_i = 0;
    for (int j = 0; j < R; j++) {
        int f0 = go, fl0 = lgo, fc = NINF, flc = NINF, kf = 0, kfl = 0;
        int F = go, FL = lgo, Fo = 1, FLo = 1;
        int colmax = -1, colmax_i = 0;
        for (int i = 0; i < Q; i++) {
            if (exact && i > 0 && (i % segLen) == 0) {
                if (fc >= f0) kf += 1; else { fc = f0; kf = 1; }
                if (flc >= fl0) kfl += 1; else { flc = fl0; kfl = 1; }
                f0 = go; fl0 = lgo;
            }
            int diag = (i > 0 && j > 0) ? Hprev[i - 1] : 0;
            int hd = imax(0, diag + sc->sub[5 * s->r[j] + s->q[i]]);
            int T, h;
            if (exact) {
                int hm = imax(imax(imax(hd, Ep[i]), imax(f0, ELp[i])), fl0);
                h = imax(hm, imax(fc, flc));
                int cs = (fc == h), cl = (flc == h);
                if (h == hd) T = (ELp[i] == h) ? T_DEL_L : (fl0 == h) ? T_INS_L : (h == 0 ? T_ZERO : T_DIAG);
                else if (cs || cl) T = (cs && (!cl || kf >= kfl)) ? T_INS : T_INS_L;
                else T = (ELp[i] == h) ? T_DEL_L : (fl0 == h) ? T_INS_L : (f0 == h) ? T_INS : T_DEL;
                Ep[i] = imax(hm + go, Ep[i] + ge); ELp[i] = imax(hm + lgo, ELp[i] + lge);
                f0 = imax(hm + go, f0 + ge); fl0 = imax(hm + lgo, fl0 + lge); fc += ge; flc += lge;
            } else {
                h = imax(imax(imax(hd, Ea[i]), imax(F, ELa[i])), FL);
                T = (ELa[i] == h) ? T_DEL_L : (FL == h) ? T_INS_L : (hd == h) ? (h == 0 ? T_ZERO : T_DIAG) : (F == h) ? T_INS : T_DEL;
            }
            if (trace)
                trace[(size_t)j * Q + i] = (uint16_t)(T | (Eo[i] ? B_DIAG_DEL : B_DEL) | (Fo ? B_DIAG_INS : B_INS) |
                                                      (ELo[i] ? B_DIAG_DEL_L : B_DEL_L) | (FLo ? B_DIAG_INS_L : B_INS_L));
            Eo[i] = (uint8_t)(h + go > Ea[i] + ge); Ea[i] = imax(h + go, Ea[i] + ge);
            ELo[i] = (uint8_t)(h + lgo > ELa[i] + lge); ELa[i] = imax(h + lgo, ELa[i] + lge);
            Fo = (h + go > F + ge); F = imax(h + go, F + ge);
            FLo = (h + lgo > FL + lge); FL = imax(h + lgo, FL + lge);
            Hcur[i] = h;
            if (h > colmax) { colmax = h; colmax_i = i; }
        }
        if (colmax > best) { best = colmax; best_j = j; best_i = colmax_i; }
        else if (j == 0) { best_i = colmax_i; }   /* all-zero tile: smallest zero row of column 0 */
        int* tmp = Hprev; Hprev = Hcur; Hcur = tmp;
    }
    if (start_end) { *score_out = Hprev[Q - 1]; *end_query_out = Q - 1; *end_ref_out = R - 1; }
    else { *score_out = best; *end_query_out = best_i; *end_ref_out = best_j; }
    free(Hprev); free(Hcur); free(Ep); free(ELp); free(Ea); free(ELa); free(Eo); free(ELo);
}

/* DualAlignSIMDTraceback + AddToTracebackPointers (Processor.cpp:568-716) on natural-order trace words. */
static void traceback(const uint16_t* trace, int Q, int R, int i, int j, int max_tb_steps,
                      DarwinTileRes* res, uint64_t* tb_words, int tb_cap, uint8_t* ops, int ops_cap,
                      uint32_t* flags, int* err) {
    int i_steps = 0, j_steps = 0, where = T_DIAG, total = 0, nwords = 0;
    uint64_t tb = 0;
    (void)R;
#define EMIT(code) do { \
        if (total % 32 == 0) { if (total > 0) { if (tb_words) { if (nwords < tb_cap) tb_words[nwords] = tb; else *err = 1; } nwords++; } tb = (uint64_t)(code); } \
        else tb = ((uint64_t)(code) << (2 * (total % 32))) + tb; \
        if (ops) { if (total < ops_cap) ops[total] = (uint8_t)(code); else *err = 1; } \
        total++; } while (0)
    while (i >= 0 && j >= 0) {
        uint16_t w = trace[(size_t)j * Q + i];
        if (i_steps == max_tb_steps || j_steps == max_tb_steps) break;
        if (where == T_DIAG) {
            if (w & T_DIAG) { EMIT(DARWIN_OP_M); i--; j--; i_steps++; j_steps++; }
            else if (w & T_DEL) where = T_DEL;
            else if (w & T_INS) where = T_INS;
            else if (w & T_DEL_L) { where = T_DEL_L; *flags |= GACT_TILE_LFLAG; }
            else if (w & T_INS_L) { where = T_INS_L; *flags |= GACT_TILE_LFLAG | GACT_TILE_LONG_INS; }
            else break;
        } else if (where == T_DEL) {
            EMIT(DARWIN_OP_D); j--; j_steps++; where = (w & B_DIAG_DEL) ? T_DIAG : T_DEL;
        } else if (where == T_INS) {
            EMIT(DARWIN_OP_I); i--; i_steps++; where = (w & B_DIAG_INS) ? T_DIAG : T_INS;
        } else if (where == T_INS_L) {
            EMIT(DARWIN_OP_I); i--; i_steps++; where = (w & B_DIAG_INS_L) ? T_DIAG : T_INS_L;   /* L_I % 4 == I */
        } else {
            EMIT(DARWIN_OP_D); j--; j_steps++; where = (w & B_DIAG_DEL_L) ? T_DIAG : T_DEL_L;   /* L_D % 4 == D */
        }
    }
    if (total > 0) { if (tb_words) { if (nwords < tb_cap) tb_words[nwords] = tb; else *err = 1; } nwords++; }
#undef EMIT
    res->query_offset = (uint16_t)i_steps; res->ref_offset = (uint16_t)j_steps; res->total_TB_pointers = (uint16_t)total;
}

/* One request of BatchAlignmentSIMD (Processor.cpp:722-761). */
int gact_tile(const GactScoring* sc, const char* dram, const DarwinTileReq* req, int do_traceback, int rule,
              DarwinTileRes* res, uint64_t* tb_words, int tb_words_cap, uint8_t* ops, int ops_cap, uint32_t* flags) {
    int Q = req->query_size, R = req->ref_size, se = req->align_fields & 1;
    uint32_t fl = 0; int err = 0;
    memset(res, 0, sizeof(*res));
    res->index = (uint8_t)req->index;
    if (Q == 0 || R == 0) {                                       /* :177-182: zeros; traceback loop never runs */
        /* with start_end the reference would start at i=-1/j=-1: the while loop exits immediately */
        if (flags) *flags = 0;
        return 0;
    }
    TileSeq s; fetch_seqs(dram, req, &s);
    uint16_t* trace = do_traceback ? (uint16_t*)malloc(sizeof(uint16_t) * (size_t)Q * R) : NULL;
    int score, eq, er;
    if (rule == GACT_RULE_STRIPED) tile_striped(sc, &s, se, trace, &score, &eq, &er);
    else tile_scalar(sc, &s, se, rule == GACT_RULE_STREAM, trace, &score, &eq, &er);
    res->score = score; res->ref_max_pos = (uint16_t)er; res->query_max_pos = (uint16_t)eq;
    if (do_traceback) {
        int i = se ? Q - 1 : eq, j = se ? R - 1 : er;             /* :593-598 */
        traceback(trace, Q, R, i, j, req->max_tb_steps, res, tb_words, tb_words_cap, ops, ops_cap, &fl, &err);
    }
    if (rule != GACT_RULE_CLEAN) fl &= ~GACT_TILE_LFLAG;
    if (flags) *flags = fl;
    free(trace); free(s.q); free(s.r);
    return err ? DARWIN_ERR_CAPACITY : 0;
}

int gact_tiles(const GactScoring* sc, const char* dram, int do_traceback, int rule, const DarwinTileReq* req, int n,
               DarwinTileRes* res, uint64_t* tb_words, int tb_words_per_req, uint32_t* flags) {
    for (int k = 0; k < n; k++) {
        int rc = gact_tile(sc, dram, &req[k], do_traceback, rule, &res[k],
                           tb_words ? tb_words + (size_t)k * tb_words_per_req : NULL, tb_words_per_req,
                           NULL, 0, flags ? &flags[k] : NULL);
        if (rc) return rc;
    }
    return 0;
}

/* AlignmentScore (extender.cpp:1161-1200); NtChar2Int = ntcoding.cpp:11-23 (toupper, non-ACGT -> N). */
int gact_alignment_score(const GactScoring* sc, const char* ref, const char* query, uint64_t n) {
    static const int mat_offset[4] = { 0, 1, 3, 6 };
    int score = 0, open = 0, sgp = 0, lgp = 0;
    for (uint64_t l = 0; l < n; l++) {
        char r = ref[l], q = query[l];
        if (r == '-' || q == '-') {
            sgp += open ? sc->ge : sc->go;
            lgp += open ? sc->lge : sc->lgo;
            open = 1;
        } else {
            int rn = gact_nt2int(r, 0), qn = gact_nt2int(q, 0);
            if (rn <= 3 && qn <= 3) {
                int idx = (rn > qn) ? qn * 4 + rn - mat_offset[qn] : rn * 4 + qn - mat_offset[rn];
                score += sc->tri[idx];
            } else score += sc->tri[10];
            score += (lgp < sgp) ? sgp : lgp;
            open = 0; sgp = 0; lgp = 0;
        }
    }
    return score;
}

static char comp_char(char c) {   /* main.cpp:83-113 (RevComp keeps case; anything else was rejected at load) */
    switch (c) {
        case 'a': return 't'; case 'A': return 'T'; case 'c': return 'g'; case 'C': return 'G';
        case 'g': return 'c'; case 'G': return 'C'; case 't': return 'a'; case 'T': return 'A';
        case 'n': return 'n'; default: return 'N';
    }
}

/* character of the strand-local read at offset k: forward read (rc=0) or rc_seq (rc=1, main.cpp:669-670) */
static char read_char_at(const char* dram, const DarwinAnchor* a, uint32_t k) {
    if (!a->strand) return dram[a->read_addr + k];
    if (k >= a->read_len) return 'N';                      /* rc_seq padding, main.cpp:116-118 */
    return comp_char(dram[a->read_addr + (a->read_len - 1 - k)]);
}

/* Gapped strings as extender.cpp:287-323 / :434-458 build them, from the op string. */
int gact_build_strings(const char* dram, const DarwinAnchor* a, const DarwinAlnRes* r, const uint8_t* ops,
                       char* ref_str, char* query_str) {
    /* Ops [0,n_left) are the left extension in left-to-right order: walking them backwards from the anchor
     * reproduces the reference's consumption; right ops walk forwards from anchor+1. */
    uint32_t cr = a->reference_pos - a->chr_start, cq = a->query_pos;
    for (int64_t k = (int64_t)r->n_left_ops - 1; k >= 0; k--) {
        uint8_t d = ops[k];
        ref_str[k] = (d == DARWIN_OP_I) ? '-' : dram[a->chr_start + cr];
        query_str[k] = (d == DARWIN_OP_D) ? '-' : read_char_at(dram, a, cq);
        if (d != DARWIN_OP_I && cr > 0) cr--;
        if (d != DARWIN_OP_D && cq > 0) cq--;
    }
    cr = a->reference_pos - a->chr_start + 1; cq = a->query_pos + 1;
    for (uint32_t k = r->n_left_ops; k < r->n_ops; k++) {
        uint8_t d = ops[k];
        ref_str[k] = (d == DARWIN_OP_I) ? '-' : dram[a->chr_start + cr];
        query_str[k] = (d == DARWIN_OP_D) ? '-' : read_char_at(dram, a, cq);
        if (d != DARWIN_OP_I && cr < a->ref_len) cr++;
        if (d != DARWIN_OP_D && cq < a->read_len) cq++;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Anchor state machine: extender_body::operator() for ONE anchor (extender.cpp:45-530 forward
 * strand, :557-1051 reverse strand) + makeForward/BackwardAlignment (:1067-1159).
 * ------------------------------------------------------------------------------------------ */
/* optional log of every tile request the state machine issues (debugging / tile-level parity of the extender) */
static DarwinTileReq* g_tile_log = 0; static int g_tile_log_cap = 0, g_tile_log_n = 0;
void gact_set_tile_log(DarwinTileReq* buf, int cap) { g_tile_log = buf; g_tile_log_cap = cap; g_tile_log_n = 0; }
int gact_tile_log_count(void) { return g_tile_log_n; }

typedef struct OpBuf { uint8_t* left; uint64_t nleft, capleft; uint8_t* right; uint64_t nright, capright; } OpBuf;

static void push_op(uint8_t** buf, uint64_t* n, uint64_t* cap, uint8_t d) {
    if (*n == *cap) { *cap = *cap ? *cap * 2 : 1024; *buf = (uint8_t*)realloc(*buf, *cap); }
    (*buf)[(*n)++] = d;
}

static int extend_one(const GactScoring* sc, const char* dram, const DarwinExtendParams* p, int rule,
                      const DarwinAnchor* a, const uint64_t* hit_pool, DarwinAlnRes* res,
                      uint8_t* ops_pool, uint64_t* used, uint64_t cap) {
    const int T = p->tile_size, O = p->tile_overlap, rc = a->strand;
    /* makeForwardAlignment / makeBackwardAlignment, extender.cpp:1083-1102 / :1130-1149 */
    uint32_t cr = a->reference_pos - a->chr_start, cq = a->query_pos;
    uint32_t rso = cr, reo = cr, qso = cq, qeo = cq;
    const uint64_t rsa = a->chr_start; const uint32_t RL = a->ref_len, QL = a->read_len;
    int large = 0, ldone = 0, rdone = 0, emit = 0;
    int64_t nl = a->left_hits_n, nr = a->right_hits_n;             /* vector sizes; .back() = [n-1] */
    const uint64_t* lh = hit_pool + a->left_hits_off; const uint64_t* rh = hit_pool + a->right_hits_off;
    OpBuf ob; memset(&ob, 0, sizeof(ob));
    int maxops = 4 * DARWIN_MAX_TILE + 64;
    uint8_t* tops = (uint8_t*)malloc((size_t)maxops);
    memset(res, 0, sizeof(*res));

    while (!(ldone && rdone)) {
        const int left = !ldone;
        int rt = T, qt = T;
        if (large) {                                              /* :61-78 / :136-153 */
            if ((left ? nl : nr) <= 0) { free(tops); free(ob.left); free(ob.right); return DARWIN_ERR_INVALID; }
            uint64_t ho = left ? lh[nl - 1] : rh[nr - 1];
            uint64_t h1 = rsa + cr, o1 = cq, h2 = ho >> 32, o2 = (ho << 32) >> 32;
            int wide = left ? ((h1 - h2) > (o1 - o2)) : ((h2 - h1) > (o2 - o1));   /* uint64 arithmetic */
            rt = wide ? 1984 : 960; qt = wide ? 960 : 1984;
            res->n_large_tiles++;
        }
        DarwinTileReq rq; memset(&rq, 0, sizeof(rq));
        if (left) {                                               /* :121-130 / :637-646 */
            rq.ref_size = (uint16_t)((uint64_t)cr + 1 < (uint64_t)rt ? cr + 1 : (uint32_t)rt);
            rq.query_size = (uint16_t)((uint64_t)cq + 1 < (uint64_t)qt ? cq + 1 : (uint32_t)qt);
            rq.ref_bases_start_addr = rsa + (cr >= (uint32_t)rt ? cr - rt + 1 : 0);
            uint32_t qoff = (cq >= (uint32_t)qt ? cq - qt + 1 : 0);
            rq.query_bases_start_addr = rc ? a->read_addr + QL - rq.query_size - qoff : a->read_addr + qoff;
            rq.align_fields = rc ? (DARWIN_REVERSE_QUERY | DARWIN_COMPLEMENT_QUERY | DARWIN_START_END) : DARWIN_START_END;
        } else {                                                  /* :197-206 / :712-721 */
            rq.ref_size = (uint16_t)((RL - cr) < (uint32_t)rt ? (RL - cr) : (uint32_t)rt);
            rq.query_size = (uint16_t)((QL - cq) < (uint32_t)qt ? (QL - cq) : (uint32_t)qt);
            rq.ref_bases_start_addr = rsa + cr;
            rq.query_bases_start_addr = rc ? a->read_addr + QL - rq.query_size - cq : a->read_addr + cq;
            rq.align_fields = rc ? (DARWIN_REVERSE_REF | DARWIN_COMPLEMENT_QUERY | DARWIN_START_END)
                                 : (DARWIN_REVERSE_REF | DARWIN_REVERSE_QUERY | DARWIN_START_END);
        }
        rq.max_tb_steps = (uint16_t)(2 * T);                      /* :127 */
        DarwinTileRes tr; uint32_t tfl = 0;
        if (g_tile_log) { if (g_tile_log_n < g_tile_log_cap) g_tile_log[g_tile_log_n] = rq; g_tile_log_n++; }
        int trc = gact_tile(sc, dram, &rq, 1, rule, &tr, NULL, 0, tops, maxops, &tfl);
        if (!trc && rule == GACT_RULE_CLEAN && (tfl & GACT_TILE_LFLAG)) {
            /* two-tier scheme of the product: a flagged clean tile is recomputed with the exact rule */
            res->flags |= DARWIN_ALN_EXACT_RERUN;
            trc = gact_tile(sc, dram, &rq, 1, GACT_RULE_STREAM, &tr, NULL, 0, tops, maxops, &tfl);
        }
        if (trc) { free(tops); free(ob.left); free(ob.right); return trc; }
        if (tfl & GACT_TILE_LONG_INS) res->flags |= DARWIN_ALN_LONG_INS_PATH;
        res->n_tiles++; res->cells += (uint64_t)rq.ref_size * rq.query_size;
        const int len = tr.total_TB_pointers;

        /* consumption, :258-331 / :406-466 (and rc twins) -- the `break` leaves only the 32-op loop */
        int crt = T, cqt = T;
        if (large && p->do_overlap == 0) { crt = rt; cqt = qt; }
        const int S = imin(crt, cqt) - O;
        int steps = 0;
        for (int w = 0; w < len; w += 32) {
            int np = imin(len - w, 32);
            for (int q = 0; q < np; q++) {
                uint8_t d = tops[w + q];
                if (left) {
                    push_op(&ob.left, &ob.nleft, &ob.capleft, d);
                    if (d == DARWIN_OP_M || d == DARWIN_OP_D) { if (cr > 0) cr--; else rso = 0; }
                    if (d == DARWIN_OP_M || d == DARWIN_OP_I) { if (cq > 0) cq--; else qso = 0; }
                } else {
                    push_op(&ob.right, &ob.nright, &ob.capright, d);
                    if (d == DARWIN_OP_M || d == DARWIN_OP_D) { if (cr < RL) cr++; }
                    if (d == DARWIN_OP_M || d == DARWIN_OP_I) { if (cq < QL) cq++; }
                }
                steps++;
                if (steps >= S && d == DARWIN_OP_M) break;
            }
        }
        if (g_tile_log && g_tile_log_n <= g_tile_log_cap) {      /* debugging: state after consumption */
            DarwinTileReq* L = &g_tile_log[g_tile_log_n - 1];
            L->score_threshold = (uint32_t)len; L->ref_bases_start_addr = cr; L->query_bases_start_addr = cq;
        }
        if (left) {
            while (nl > 0) {                                      /* :336-351 */
                uint64_t ho = lh[nl - 1], hit = ho >> 32, off = (ho << 32) >> 32;
                if (hit < rsa + cr && off < cq) break;
                nl--;
            }
            int stall = rc ? (len == 0 || rso == 0 || qso == 0)                 /* :867 */
                           : (len == 0 || nl == 0 || rso == 0 || qso == 0);     /* :353 */
            if (stall) {
                if (large || nl == 0 || rso == 0 || qso == 0) {                 /* :354 / :868 */
                    ldone = 1;
                    if (rso > 0) rso = cr + 1;
                    if (qso > 0) qso = cq + 1;
                    if ((cr + 1 < RL) && (cq + 1 < QL) && !rdone) { cr = reo + 1; cq = qeo + 1; }   /* :363-367 */
                    else { rdone = 1; if (rc) emit = 1; }          /* fwd strand drops it (:368-382), rc emits (:883-903) */
                } else large = 1;
            } else large = 0;
        } else {
            while (nr > 0) {                                      /* :472-488 */
                uint64_t ho = rh[nr - 1], hit = ho >> 32, off = (ho << 32) >> 32;
                if (hit > rsa + cr && off > cq) break;
                nr--;
            }
            if (len == 0 || cr == RL || cq == QL) {               /* :490 */
                if (large || nr == 0 || cr == RL || cq == QL) { reo = cr - 1; qeo = cq - 1; emit = 1; rdone = 1; }
                else large = 1;
            } else large = 0;
        }
    }
    free(tops);
    res->reference_start_offset = rso; res->reference_end_offset = reo;
    res->query_start_offset = qso; res->query_end_offset = qeo;
    if (emit) {
        res->flags |= DARWIN_ALN_EMITTED;
        uint64_t n = ob.nleft + ob.nright;
        res->n_ops = (uint32_t)n; res->n_left_ops = (uint32_t)ob.nleft; res->ops_offset = *used;
        if (*used + n > cap) res->flags |= DARWIN_ALN_OPS_OVERFLOW;
        else {
            uint8_t* o = ops_pool + *used;
            for (uint64_t k = 0; k < ob.nleft; k++) o[k] = ob.left[ob.nleft - 1 - k];   /* left output is prepended */
            memcpy(o + ob.nleft, ob.right, ob.nright);
            char* rs = (char*)malloc(n + 1), *qs = (char*)malloc(n + 1);
            gact_build_strings(dram, a, res, o, rs, qs);
            res->score = gact_alignment_score(sc, rs, qs, n);     /* :498 */
            free(rs); free(qs);
            *used += n;
        }
    }
    free(ob.left); free(ob.right);
    return 0;
}

int gact_extend(const GactScoring* sc, const char* dram, const DarwinExtendParams* p, int rule,
                const DarwinAnchor* anchors, int n, const uint64_t* hit_pool,
                DarwinAlnRes* res, uint8_t* ops_pool, uint64_t ops_pool_bytes) {
    uint64_t used = 0;
    for (int k = 0; k < n; k++) {
        int rc = extend_one(sc, dram, p, rule, &anchors[k], hit_pool, &res[k], ops_pool, &used, ops_pool_bytes);
        if (rc) return rc;
    }
    return 0;
}

/* ---- first-tile filter (filter.cpp) ------------------------------------------------------------------------------ */

/* Request construction of filter_body for one candidate: filter.cpp:44-71 (forward), :154-181 (reverse complement).
 * All arithmetic in the reference's own types (uint32_t hit/offset/chr_*, size_t read_len, int cfg.first_tile_size). */
static void filter_request(const DarwinFilterCand* c, int fts, DarwinTileReq* rq, uint32_t* rts_out, uint32_t* qts_out) {
    uint32_t hit = c->hit, offset = c->offset, chr_start = c->chr_start;
    uint32_t chr_end = chr_start + c->chr_len;                                                     /* :51 */
    size_t read_len = c->read_len;
    uint32_t ref_tile_start = (hit + fts < chr_end) ? hit : ((chr_end > (uint32_t)fts) ? chr_end - fts : 0);             /* :57 */
    uint32_t query_tile_start = (offset + fts < read_len) ? offset : ((read_len > (size_t)fts) ? (uint32_t)(read_len - fts) : 0);   /* :58 */
    uint32_t ref_tile_size = ((uint32_t)fts < (chr_end - chr_start)) ? (uint32_t)fts : (chr_end - chr_start);             /* :59 */
    uint32_t query_tile_size = ((size_t)fts < read_len) ? (uint32_t)fts : (uint32_t)read_len;                            /* :60 */
    memset(rq, 0, sizeof(*rq));
    rq->ref_size = (uint16_t)ref_tile_size; rq->query_size = (uint16_t)query_tile_size;
    rq->ref_bases_start_addr = ref_tile_start;                                                                           /* :66 */
    rq->query_bases_start_addr = c->strand ? c->read_addr + read_len - (query_tile_start + query_tile_size)              /* :176 */
                                           : c->read_addr + query_tile_start;                                            /* :67 */
    rq->max_tb_steps = (uint16_t)(2 * fts); rq->score_threshold = 0;                                                     /* :69-70 */
    rq->align_fields = c->strand ? (DARWIN_REVERSE_QUERY | DARWIN_COMPLEMENT_QUERY) : 0;                                 /* :72 / :181 */
    *rts_out = ref_tile_start; *qts_out = query_tile_start;
}

/* The tile part of filter_body::operator() for n candidates (filter.cpp:28-122, :131-223): score-only max-cell tiles
 * (do_traceback = 0), then the score test (:87) and the overlap test (:102-104). */
int gact_filter(const GactScoring* sc, const char* dram, const DarwinFilterParams* p, const DarwinFilterCand* cands, int n,
                DarwinFilterRes* res) {
    for (int k = 0; k < n; k++) {
        const DarwinFilterCand* c = &cands[k];
        DarwinTileReq rq; uint32_t rts, qts;
        filter_request(c, p->first_tile_size, &rq, &rts, &qts);
        rq.index = (uint16_t)(k & 63);                                                              /* :62 (c - b) */
        DarwinTileRes tr;
        int rc = gact_tile(sc, dram, &rq, 0, GACT_RULE_STREAM, &tr, NULL, 0, NULL, 0, NULL);
        if (rc) return rc;
        uint32_t chr_end = c->chr_start + c->chr_len;
        uint32_t ovl = c->offset + (chr_end - c->hit);                                              /* :102 */
        res[k].score = tr.score;
        res[k].reference_pos = rts + tr.ref_max_pos;                                                /* :109 */
        res[k].query_pos = qts + tr.query_max_pos;                                                  /* :110 */
        res[k].flags = ((uint32_t)tr.score >= (uint32_t)p->first_tile_score_threshold ? DARWIN_FILTER_SCORE_OK : 0u) |   /* :87: uint32 score vs int */
                       (ovl > (uint32_t)(p->min_overlap / 2) ? DARWIN_FILTER_OVERLAP_OK : 0u);      /* :104 */
    }
    return 0;
}

/* filter_body::slopeFilter (filter.cpp:227-289) on the locations of ONE strand of one batch.
 * in: read_num / score / reference_pos / query_pos per location (any order); order_out receives the indices of the
 * surviving locations in output order; returns their number.  The sort predicate is the reference's (:232-234):
 * read_num asc, score desc, reference_pos asc, query_pos asc; std::sort is not stable but the key is total up to exact
 * duplicates.  The slope test is evaluated in float like the reference (:266-271). */
typedef struct { int read_num, score; uint32_t rpos, qpos; int idx; } SlopeLoc;
static int slope_cmp(const void* a_, const void* b_) {
    const SlopeLoc* a = (const SlopeLoc*)a_; const SlopeLoc* b = (const SlopeLoc*)b_;
    if (a->read_num != b->read_num) return a->read_num < b->read_num ? -1 : 1;
    if (a->score != b->score) return a->score > b->score ? -1 : 1;
    if (a->rpos != b->rpos) return a->rpos < b->rpos ? -1 : 1;
    if (a->qpos != b->qpos) return a->qpos < b->qpos ? -1 : 1;
    return a->idx < b->idx ? -1 : (a->idx > b->idx);
}
int gact_slope_filter(const int* read_num, const int* score, const uint32_t* reference_pos, const uint32_t* query_pos, int n,
                      float slope_threshold, int* order_out) {
    SlopeLoc* v = (SlopeLoc*)malloc(sizeof(SlopeLoc) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; k++) { v[k].read_num = read_num[k]; v[k].score = score[k]; v[k].rpos = reference_pos[k]; v[k].qpos = query_pos[k]; v[k].idx = k; }
    qsort(v, (size_t)n, sizeof(SlopeLoc), slope_cmp);
    int kept = 0;
    for (int a = 0; a < n; a++) {
        if (v[a].read_num == -1) continue;                                                          /* :240-241 */
        order_out[kept++] = v[a].idx;                                                               /* :243 */
        for (int b = a + 1; b < n; b++) {
            if (v[b].read_num == -1) continue;
            if (v[b].read_num != v[a].read_num) break;                                              /* :250-251 */
            float r1 = (float)v[a].rpos, q1 = (float)v[a].qpos, r2 = (float)v[b].rpos, q2 = (float)v[b].qpos;
            float s = (r1 - r2) / (q1 - q2) - 1;                                                    /* :270 */
            if ((s < 0 ? -s : s) <= slope_threshold) v[b].read_num = -1;                            /* :271-274 */
        }
    }
    free(v);
    return kept;
}

// Shared device-side definitions of the B200-native GACT path.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/darwin_gpu.h"

namespace gact {

constexpr int kMaxTile   = DARWIN_MAX_TILE;   // 1984: largest tile edge (reference: extender.cpp:70-75)
constexpr int kSeqSmem   = 2048;              // bytes of shared memory per staged sequence
constexpr int kWarpsPerCta = 4;
constexpr int kNegInf    = -(1 << 28);

// ---- exact path geometry: lane owns KX consecutive query rows of a 32*KX-row strip ---------------
constexpr int KX     = 8;
constexpr int kStrip = 32 * KX;

// ---- traceback codes of the exact path (1 byte per cell) ------------------------------------------
// bits 0-2: T  (reference one-hot TRACEBACK_*_T, Processor.h:21-26), bits 3-6: open bits of E, F, E_L, F_L
enum : uint32_t { XT_ZERO = 0, XT_DEL = 1, XT_INS = 2, XT_DEL_L = 3, XT_INS_L = 4, XT_DIAG = 5 };
constexpr uint32_t XB_EOPEN = 8, XB_FOPEN = 16, XB_ELOPEN = 32, XB_FLOPEN = 64;

struct DevScoring {
    int sub[25];          // row = reference nt, col = query nt, 4 = N   (Processor.cpp:50-74)
    int go, ge, lgo, lge; // Processor.cpp:75-78
    int tri[11];          // cfg.gact_sub_mat order, for AlignmentScore (extender.cpp:1161-1200)
    int uniform;          // 1 when all matches score `match` and all mismatches `mismatch`
    int match, mismatch, subn;
};

// One tile as the kernels see it (== AlignmentInputFieldsDRAM, Darwin.bond:95-112)
struct TileJob {
    uint64_t ra, qa;      // arena offsets of the first reference / query base of the tile
    int R, Q;             // ref_size, query_size
    uint32_t flags;       // align_fields
    int max_tb;           // max_tb_steps
    int band_shift;       // single-strip fast path: the stored band is centred this many cells off the corner diagonal (0 = centred);
                          // a scheduling hint of the anchor walker -- it decides fast path vs exact rerun, never the result
};

struct TileOut {
    int score, ref_max_pos, query_max_pos;
    int ref_offset, query_offset, total;
    uint32_t tflags;      // bit0: traceback entered INS_L state, bit1: traceback entered a long-gap state
};

// vertical chain state handed from the lane above: 32 bytes
struct __align__(16) ChainRec {
    int hbot;             // true H of the row above, this column
    int F, FL;            // true vertical gaps entering the next row (open bits in misc)
    int f0, fl0;          // own-lane chains of the striped kernel
    int fc, flc;          // best carried chains
    int misc;             // kf | kfl << 8 | Fopen << 16 | FLopen << 17
};

// per-warp global scratch (exact path)
struct WarpScratch {
    uint8_t*  trace;      // kTraceBytes
    ChainRec* bound;      // kMaxTile records: chain state below the last row of the previous strip
};

constexpr size_t kTraceBytesPerWarp = (size_t)((kMaxTile + kStrip - 1) / kStrip) * (kMaxTile + 32) * 32 * KX;   // 4.13 MB worst case
__host__ __device__ inline size_t exact_trace_bytes(int Q, int R) {
    return (size_t)((Q + kStrip - 1) / kStrip) * (size_t)(R + 31) * 32 * KX;
}

// 4-bit packed arena: base k lives in nibble (k & 1) of byte k >> 1.  0..3 = ACGT, 4 = N.
__device__ __forceinline__ uint32_t arena_code(const uint8_t* __restrict__ arena, uint64_t a) {
    uint32_t b = __ldg(arena + (a >> 1));
    return (b >> ((a & 1) * 4)) & 0xF;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

} // namespace gact

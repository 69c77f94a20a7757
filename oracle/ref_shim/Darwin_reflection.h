// TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product.
// Stand-in for the header Microsoft Bond's `gbc` would generate from the
// reference's IDL (reference: software/Darwin.bond:36-141).  Only plain
// structs are needed: Bond contributes no arithmetic to the GACT path.
#pragma once
#include <vector>
#include <string>
#include <cassert>
#include <cstring>
#include <cstdint>
#include <cstdlib>

namespace Darwin {

enum Status { OK = 0, InvalidData = 1 };          // Darwin.bond:36-40

struct AlignmentScoringParams {                   // Darwin.bond:42-66
    int32_t sub_AA = 1, sub_AC = -1, sub_AG = -1, sub_AT = -1;
    int32_t sub_CC = 1, sub_CG = -1, sub_CT = -1;
    int32_t sub_GG = 1, sub_GT = -1;
    int32_t sub_TT = 1;
    int32_t sub_N = 0;
    int32_t gap_open = -1, gap_extend = -1;
    int32_t long_gap_open = -1, long_gap_extend = -1;
};
struct AlignmentScoringParamsResponse { Status status = OK; };

struct InitializeDRAMMessage {                    // Darwin.bond:73-78
    uint64_t start_addr = 0;
    uint16_t num_bytes = 0;
    std::vector<uint64_t> data;
};
struct InitializeDRAMMessageResponse { Status status = OK; };

struct AlignmentInputFieldsDRAM {                 // Darwin.bond:95-112
    uint8_t  align_fields = 0;
    uint16_t index = 0;
    uint64_t ref_bases_start_addr = 0;
    uint64_t query_bases_start_addr = 0;
    uint16_t ref_size = 0;
    uint16_t query_size = 0;
    uint16_t max_tb_steps = 512;
    uint32_t score_threshold = 0;
};

struct AlignmentResult {                          // Darwin.bond:114-129
    uint8_t  index = 0;
    uint32_t score = 0;
    uint16_t ref_offset = 0;
    uint16_t query_offset = 0;
    uint16_t ref_max_pos = 0;
    uint16_t query_max_pos = 0;
    uint16_t total_TB_pointers = 0;
    std::vector<uint64_t> TB_pointers;
    Status status = OK;
};

struct BatchAlignmentInputFieldsDRAM {            // Darwin.bond:131-135
    uint8_t do_traceback = 0;
    std::vector<AlignmentInputFieldsDRAM> requests;
};
struct BatchAlignmentResultDRAM {                 // Darwin.bond:137-141
    std::vector<AlignmentResult> results;
};

} // namespace Darwin

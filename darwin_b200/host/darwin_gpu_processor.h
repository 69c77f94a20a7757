// C++ host adapter: the reference's own Processor / extender signatures on top of the C-ABI
// (include/darwin_gpu.h).  A maintainer of yatisht/darwin adds this file + darwin_gpu_processor.cpp to the
// build, links libdarwin_gact.so and points the g_* table at these functions.
//
// Compiles against the reference's headers (software/graph.h, Processor.h and the Bond-generated
// Darwin_reflection.h); nothing here is needed by the CUDA library itself.
#pragma once
#include "graph.h"               // reference: software/graph.h
#include "darwin_gpu.h"          // include/darwin_gpu.h
#include "darwin_gpu_combiner.h" // cross-read batching of the GPU calls

namespace darwin_gpu_host {

// == InitializeProcessor_ptr (software/Processor.h:50): one GPU handle per token (token -> device round-robin).
// `fpgas` is read as the number of GPUs to use (params.cfg [FPGA] num_fpgas), `chip_ids` is ignored.  Every GPU gets
// up to DARWIN_GPU_LANES (default 4, at most threads / gpus) handles sharing one arena replica; token t is served by
// GPU t % gpus, lane (t / gpus) % lanes.
// Returns the number of processors created.  arena_bytes defaults to the reference's 4 GiB arena (DRAM.cpp:8).
size_t InitializeProcessor(int threads, int gpus, std::string chip_ids);
void   ShutdownProcessor();

// == the g_* table entries (software/Processor.h:51-55, defaults at Processor.cpp:1063-1069)
void InitializeScoringParameters(size_t token, Darwin::AlignmentScoringParams& request,
                                 Darwin::AlignmentScoringParamsResponse& response);
void InitializeReferenceMemory(size_t token, char* dram, Darwin::InitializeDRAMMessage& request,
                               Darwin::InitializeDRAMMessageResponse& response);
void InitializeReadMemory(size_t token, char* dram, Darwin::InitializeDRAMMessage& request,
                          Darwin::InitializeDRAMMessageResponse& response);
void BatchAlignmentSIMD(size_t token, char* dram, Darwin::BatchAlignmentInputFieldsDRAM& request,
                        Darwin::BatchAlignmentResultDRAM& result);

// == extender_body (software/graph.h:219-229, extender.cpp:9-1065): same input/output ports; all anchors of the
// batch go to the GPU in one darwin_gpu_extend call, ExtendAlignments (gapped strings, offsets, score) are rebuilt
// from the op strings.  Output order: forward-strand anchors in input order, then reverse-strand (the printer
// re-sorts by read and score, printer.cpp:18-21).
struct gpu_extender_body {
    void operator()(extender_input input, extender_node::output_ports_type& op);
};

// == seeder_body (software/graph.h:200-203, seeder.cpp:6-55): D-SOFT of the whole batch on the GPU.  Needs the seed
// position table: call BuildSeedIndex() once after the reference has been uploaded (it replaces
// `sa = new SeedPosTable(...)`, main.cpp:508, and the minimizer pass over the reference, main.cpp:323-341).
void BuildSeedIndex();
struct gpu_seeder_body {
    filter_input operator()(seeder_input input);
};

// == the seeder -> filter -> extender chain of main.cpp:590-624 as ONE node body (needs BuildSeedIndex()): wire the reader's
// gatekeeper output straight into it and its ports into the printer / ticketer like the extender's.
struct gpu_align_body {
    void operator()(seeder_input input, extender_node::output_ports_type& op);
};

// == the whole align graph of main.cpp:590-624 INCLUDING the output stage (printer_body::sam_printer, printer.cpp:7-98) for the
// reference-guided mode: reads in, SAM text out.  The alignments never become ExtendAlignments: print order and overlap
// suppression come from darwin_gpu_sam_select, the CIGAR from darwin_gpu_cigar -- straight from the op strings, no gapped
// strings are rebuilt (that rebuild was what bounded gpu_align_body + printer_body).  The text is byte-identical to what
// printer_body prints for the same batch (header included, once per process, under the reference's io_lock / done_header).
// De novo mode (cfg.do_overlap = 1) prints the gapped strings themselves and keeps using gpu_align_body + printer_body.
struct gpu_sam_body {
    std::string operator()(seeder_input input);
};

// == filter_body (software/graph.h:205-217, filter.cpp:8-225): same input/output tuples; the first tiles of the whole
// batch (both strands) go to the GPU in one darwin_gpu_filter call; the slope filter is the reference's own.
struct gpu_filter_body {
    extender_input operator()(filter_input input);
};

// install the functions above into the reference's table (g_InitializeScoringParameters, ...).
void InstallProcessorTable();

DarwinGpu* handle_for_token(size_t token);
GpuCombiner& combiner_for_token(size_t token);       // the per-GPU batcher all stage bodies go through
CombinerStats combiner_stats(size_t token);
CombinerStats combiner_stats_total();              // summed over all GPUs and lanes
void host_profile(double out_seconds[3]);          // gpu_align_body: request building / blocked in the combiner / rebuilding ExtendAlignments (thread-seconds since the last call)
const char* last_error();

} // namespace darwin_gpu_host

// Host side of the GPU D-SOFT path (darwin_gpu_seed_index / darwin_gpu_seed); included by darwin_gact.cu after the
// DarwinGpu handle is defined.  Kernels: dsoft.cuh.  CUB supplies the scans and the segmented sorts.
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include "dsoft.cuh"

#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return DARWIN_ERR_CUDA; } } while (0)

// Scoped device allocation of the seeding calls: stream-ordered (cudaMallocAsync on the handle's stream, the device's
// default pool keeps the memory between calls), so neither the allocation nor the release synchronises the device --
// cudaMalloc / cudaFree would stall the kernels of every other lane.
struct DevBuf {
    void* p = nullptr; cudaStream_t st = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) { st = stream; return cudaMallocAsync(&p, bytes ? bytes : 16, stream); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static int seed_const_from(const DarwinSeedParams* p, uint64_t reference_size, dsoft::SeedConst& sc, std::string& err) {
    if (p->seed_size < 4 || p->seed_size > 15) { err = "seed_size must be in [4,15] (seed_pos_table.cpp:51-52)"; return DARWIN_ERR_INVALID; }
    if (p->minimizer_window < 1 || p->minimizer_window > 32) { err = "minimizer_window must be in [1,32]"; return DARWIN_ERR_INVALID; }
    if (p->bin_size < 1 || p->max_stride < 1 || p->num_seeds < 0 || p->threshold < 0) { err = "bad D-SOFT parameters"; return DARWIN_ERR_INVALID; }
    sc.k = p->seed_size; sc.w = p->minimizer_window; sc.N = p->num_seeds; sc.threshold = p->threshold;
    sc.max_stride = p->max_stride; sc.overlap = p->do_overlap ? 1 : 0;
    sc.bin_size = (uint32_t)p->bin_size;
    sc.max_occ = (uint32_t)p->seed_occurence_multiple * (1u + (uint32_t)(reference_size >> (2 * p->seed_size)));   // seed_pos_table.cpp:57
    sc.sv_bins = sc.overlap ? 1u : (uint32_t)((1u << 12) / sc.bin_size);                                           // :397
    sc.kmask = (1u << (2 * p->seed_size)) - 1u;
    return DARWIN_OK;
}

static uint32_t min_end(uint32_t len, int k) { const uint32_t c = (~0x0fu & (len + 15u)) - (uint32_t)k; return c < 16u ? 16u : c; }

static void free_index(SeedIndex& ix) {
    if (ix.owner) { if (ix.d_buckets) cudaFree(ix.d_buckets); if (ix.d_positions) cudaFree(ix.d_positions); }
    ix = SeedIndex{};
}

static int seed_index_build(DarwinGpu* h, SeedIndex& ix, const DarwinSeedParams* p, const DarwinChrom* chroms, int n_chroms, uint64_t reference_size) {
    using namespace dsoft;
    if (n_chroms == 0) { h->err = "seed position table needs at least one chromosome"; return DARWIN_ERR_INVALID; }
    free_index(ix);
    int rc = seed_const_from(p, reference_size, ix.sc, h->err);
    if (rc) return rc;
    const SeedConst sc = ix.sc;
    ix.n_buckets = 1ull << (2 * sc.k);
    std::vector<MinJob> jobs(n_chroms);
    uint64_t list_cap = 0; uint32_t max_chunks = 1;
    for (int c = 0; c < n_chroms; c++) {
        if ((uint64_t)chroms[c].start + chroms[c].len_unpadded > h->arena_bytes) { h->err = "chromosome outside arena"; return DARWIN_ERR_INVALID; }
        jobs[c] = MinJob{chroms[c].start, chroms[c].len_unpadded, 0};
        const uint32_t e = min_end(chroms[c].len_unpadded, sc.k);
        list_cap += e;
        max_chunks = std::max(max_chunks, (e + kChunk - 1) / kChunk);
    }
    CKS(cudaMalloc(&ix.d_buckets, (ix.n_buckets + 1) * sizeof(uint32_t)));
    CKS(cudaMemsetAsync(ix.d_buckets, 0, (ix.n_buckets + 1) * sizeof(uint32_t), h->stream));
    DevBuf d_jobs, d_list, d_carry, d_cursor, d_fill, d_tmp;
    CKS(d_jobs.alloc(sizeof(MinJob) * n_chroms, h->stream)); CKS(d_list.alloc(sizeof(uint64_t) * list_cap, h->stream));
    CKS(d_carry.alloc(sizeof(int) * (size_t)n_chroms * max_chunks, h->stream)); CKS(d_cursor.alloc(sizeof(unsigned long long), h->stream));
    CKS(cudaMemcpyAsync(d_jobs.p, jobs.data(), sizeof(MinJob) * n_chroms, cudaMemcpyHostToDevice, h->stream));
    CKS(cudaMemsetAsync(d_cursor.p, 0, sizeof(unsigned long long), h->stream));
    if ((uint64_t)max_chunks * (uint64_t)n_chroms > 0x7FFFFFFFull) { h->err = "reference too fragmented: chromosomes x chunks exceeds 2^31 blocks"; return DARWIN_ERR_INVALID; }
    const unsigned grid = (unsigned)((uint64_t)max_chunks * (uint64_t)n_chroms);
    minimizer_kernel<2><<<grid, kMinThreads, 0, h->stream>>>(h->d_arena, sc, d_jobs.as<MinJob>(), nullptr, nullptr, d_carry.as<int>(), nullptr, nullptr, max_chunks);
    CKS(cudaGetLastError());
    // pass A left the position of the last change of every chunk; turn it into the run start carried INTO every chunk
    // (entries of chunks beyond a chromosome's end are never read)
    carry_kernel<<<(n_chroms + 127) / 128, 128, 0, h->stream>>>(d_carry.as<int>(), n_chroms, (int)max_chunks);
    CKS(cudaGetLastError());
    minimizer_kernel<1><<<grid, kMinThreads, 0, h->stream>>>(h->d_arena, sc, d_jobs.as<MinJob>(), d_list.as<uint64_t>(), nullptr, d_carry.as<int>(),
                                                            d_cursor.as<unsigned long long>(), ix.d_buckets, max_chunks);
    CKS(cudaGetLastError());
    h->stats.kernel_launches += 3;
    unsigned long long n_min = 0;
    CKS(cudaMemcpyAsync(&n_min, d_cursor.p, sizeof(n_min), cudaMemcpyDeviceToHost, h->stream));
    CKS(stream_wait(h, h->stream));
    if (n_min > 0xFFFFFFFFull) { h->err = "more than 2^32 minimizers"; return DARWIN_ERR_INVALID; }
    ix.n_positions = n_min;
    // buckets = prefix sums of the histogram (seed_pos_table.cpp:66-101); hist was counted at [m + 1]
    size_t tmp_bytes = 0;
    CKS(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, ix.d_buckets, ix.d_buckets, (int64_t)(ix.n_buckets + 1), h->stream));
    CKS(d_tmp.alloc(tmp_bytes, h->stream));
    CKS(cub::DeviceScan::InclusiveSum(d_tmp.p, tmp_bytes, ix.d_buckets, ix.d_buckets, (int64_t)(ix.n_buckets + 1), h->stream));
    CKS(cudaMalloc(&ix.d_positions, sizeof(uint32_t) * std::max<uint64_t>(n_min, 1)));
    CKS(d_fill.alloc(sizeof(uint32_t) * ix.n_buckets, h->stream));
    CKS(cudaMemsetAsync(d_fill.p, 0, sizeof(uint32_t) * ix.n_buckets, h->stream));
    if (n_min) scatter_kernel<<<(unsigned)((n_min + 255) / 256), 256, 0, h->stream>>>(d_list.as<uint64_t>(), n_min, ix.d_buckets, d_fill.as<uint32_t>(), ix.d_positions);
    CKS(cudaGetLastError());
    bucket_sort_kernel<<<(unsigned)((ix.n_buckets + 255) / 256), 256, 0, h->stream>>>(ix.d_buckets, ix.n_buckets, sc.max_occ, ix.d_positions);
    CKS(cudaGetLastError());
    h->stats.kernel_launches += 2;
    CKS(stream_wait(h, h->stream));
    // chromosome table for the resident pipeline: padded lengths = distance to the next chromosome (Index::chr_len)
    ix.chr_start.resize(n_chroms); ix.chr_len.resize(n_chroms);
    for (int c = 0; c < n_chroms; c++) {
        ix.chr_start[c] = chroms[c].start;
        const uint64_t next = (c + 1 < n_chroms) ? chroms[c + 1].start : reference_size;
        ix.chr_len[c] = (uint32_t)(next > chroms[c].start ? next - chroms[c].start : chroms[c].len_unpadded);
    }
    ix.ready = true;
    return DARWIN_OK;
}

template <class T>
static int exclusive_sum(DarwinGpu* h, const T* in, T* out, int64_t n) {
    size_t bytes = 0;
    CKS(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, h->stream));
    DevBuf tmp; CKS(tmp.alloc(bytes, h->stream));
    CKS(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, h->stream));
    return DARWIN_OK;                                // tmp is released in stream order
}

// 64-bit sum of a 32-bit array (synchronises the stream)
struct WidenU32 { __host__ __device__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; } };
static int sum_u64(DarwinGpu* h, const uint32_t* in, int64_t n, uint64_t* out) {
    thrust::transform_iterator<WidenU32, const uint32_t*, uint64_t> it(in, WidenU32());
    DevBuf d_out, tmp; size_t bytes = 0;
    CKS(d_out.alloc(sizeof(uint64_t), h->stream));
    CKS(cub::DeviceReduce::Sum(nullptr, bytes, it, d_out.as<uint64_t>(), n, h->stream));
    CKS(tmp.alloc(bytes, h->stream));
    CKS(cub::DeviceReduce::Sum(tmp.p, bytes, it, d_out.as<uint64_t>(), n, h->stream));
    CKS(cudaMemcpyAsync(out, d_out.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CKS(stream_wait(h, h->stream));
    return DARWIN_OK;
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, int n, uint32_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
constexpr size_t kSortSmemMax = 220 * 1024;     // shared memory one CTA of the segment sorts may use (227 KB per CTA on sm_100)
__global__ void widen_kernel(const uint32_t* __restrict__ src, uint32_t n, uint64_t* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// keep_pool != nullptr: the chained-hit pool stays on the device (handed to *keep_pool) and the anchors go to *anchor_vec
// (resized here) -- the resident pipeline (darwin_gpu_align_reads); otherwise everything is copied to the caller's buffers.
static int seed_query(DarwinGpu* h, const SeedIndex& ix, const DarwinSeedRead* reads, int n, uint32_t* anchor_begin,
                      DarwinSeedAnchor* anchors, uint64_t anchors_cap, uint64_t* n_anchors,
                      uint64_t* pool, uint64_t pool_cap, uint64_t* n_pool,
                      DevBuf* keep_pool = nullptr, std::vector<DarwinSeedAnchor>* anchor_vec = nullptr) {
    using namespace dsoft;
    const SeedConst sc = ix.sc;
    const int ns = 2 * n;
    std::vector<MinJob> jobs(ns);
    std::vector<uint32_t> seed_base(ns + 1);
    uint32_t slots = 0, max_cap = 1;
    for (int r = 0; r < n; r++) {
        if (reads[r].read_addr + reads[r].read_len > h->arena_bytes || reads[r].read_len == 0) { h->err = "read outside arena"; return DARWIN_ERR_INVALID; }
        const uint32_t e = min_end(reads[r].read_len, sc.k);
        const uint32_t head = (uint32_t)sc.N + 2u;
        const uint32_t cap = e <= head ? e : (sc.overlap ? head : head + (e - head) / (uint32_t)sc.max_stride + 1u);
        for (int s = 0; s < 2; s++) {
            seed_base[2 * r + s] = slots;
            jobs[2 * r + s] = MinJob{reads[r].read_addr, reads[r].read_len, slots};
            slots += cap;
        }
        max_cap = std::max(max_cap, cap);
    }
    seed_base[ns] = slots;
    // 32-bit offsets index the seed slots of one call: keep a call below ~1.5 G read bases (about 150 k reads of 10 kbp);
    // the hits get their own 64-bit check below; larger batches are the caller's to split
    uint64_t bases = 0, slots64 = 0;
    for (int r = 0; r < n; r++) { bases += reads[r].read_len; slots64 += 2ull * (seed_base[2 * r + 1] - seed_base[2 * r]); }
    if (bases > 1500000000ull || slots64 != slots) { h->err = "seeding batch too large (more than 1.5 G read bases in one call)"; return DARWIN_ERR_INVALID; }
    DevBuf d_jobs, d_base, d_seeds, d_nseeds, d_cnt, d_first, d_hoff, d_soff, d_max;
    CKS(d_jobs.alloc(sizeof(MinJob) * ns, h->stream)); CKS(d_base.alloc(sizeof(uint32_t) * (ns + 1), h->stream)); CKS(d_seeds.alloc(sizeof(uint64_t) * slots, h->stream));
    CKS(d_nseeds.alloc(sizeof(uint32_t) * ns, h->stream)); CKS(d_cnt.alloc(sizeof(uint32_t) * ((size_t)slots + 1), h->stream)); CKS(d_hoff.alloc(sizeof(uint32_t) * ((size_t)slots + 1), h->stream));
    CKS(d_first.alloc(sizeof(uint32_t) * ((size_t)slots + 1), h->stream));
    CKS(d_soff.alloc(sizeof(uint32_t) * (ns + 1), h->stream)); CKS(d_max.alloc(sizeof(uint32_t) * 2, h->stream));
    CKS(cudaMemcpyAsync(d_jobs.p, jobs.data(), sizeof(MinJob) * ns, cudaMemcpyHostToDevice, h->stream));
    CKS(cudaMemcpyAsync(d_base.p, seed_base.data(), sizeof(uint32_t) * (ns + 1), cudaMemcpyHostToDevice, h->stream));
    CKS(cudaMemsetAsync(d_cnt.p, 0, sizeof(uint32_t) * ((size_t)slots + 1), h->stream));
    CKS(cudaMemsetAsync(d_max.p, 0, sizeof(uint32_t) * 2, h->stream));
    CKS(cudaEventRecord(h->ev0, h->stream));
    minimizer_kernel<0><<<ns, kMinThreads, 0, h->stream>>>(h->d_arena, sc, d_jobs.as<MinJob>(), d_seeds.as<uint64_t>(), d_nseeds.as<uint32_t>(), nullptr, nullptr, nullptr, 0u);
    CKS(cudaGetLastError());
    const uint64_t grid_threads = (uint64_t)ns * max_cap;
    const unsigned gb = (unsigned)((grid_threads + 255) / 256);
    hit_count_kernel<<<gb, 256, 0, h->stream>>>(sc, ix.d_buckets, ix.d_positions, d_seeds.as<uint64_t>(), d_base.as<uint32_t>(), d_nseeds.as<uint32_t>(), ns, max_cap,
                                               d_cnt.as<uint32_t>(), d_first.as<uint32_t>());
    CKS(cudaGetLastError());
    int rc;
    // The hit offsets are 32-bit: the total (slots x up to max_occ hits each -- NOT bounded by the base count above) is
    // summed in 64 bits first, and a batch beyond 2^32 - 1 hits is refused instead of wrapping the prefix sums.
    uint64_t hits64 = 0;
    if ((rc = sum_u64(h, d_cnt.as<uint32_t>(), (int64_t)slots + 1, &hits64))) return rc;
    if (hits64 > 0xFFFFFFFFull) {
        h->err = "seeding batch too large: " + std::to_string(hits64) + " seed hits exceed 2^32 - 1; split the reads over several calls";
        return DARWIN_ERR_CAPACITY;
    }
    if ((rc = exclusive_sum(h, d_cnt.as<uint32_t>(), d_hoff.as<uint32_t>(), (int64_t)slots + 1))) return rc;
    gather_u32_kernel<<<(ns + 1 + 255) / 256, 256, 0, h->stream>>>(d_hoff.as<uint32_t>(), d_base.as<uint32_t>(), ns + 1, d_soff.as<uint32_t>());
    CKS(cudaGetLastError());
    seg_max_kernel<<<(ns + 255) / 256, 256, 0, h->stream>>>(d_soff.as<uint32_t>(), ns, d_max.as<uint32_t>());
    CKS(cudaGetLastError());
    const uint32_t n_hits = (uint32_t)hits64;
    uint32_t seg_max = 0;
    CKS(cudaMemcpyAsync(&seg_max, d_max.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CKS(stream_wait(h, h->stream));
    // One CTA sorts one strand's hits in shared memory (hit_sort_kernel) when the largest strand fits; a batch with a
    // larger strand (reads beyond ~100 kbp, or a very repetitive table) takes CUB's segmented sort instead.
    uint32_t seg_pad = 64; while (seg_pad < seg_max) seg_pad <<= 1;
    const size_t sort_smem = (size_t)seg_pad * 8 + (size_t)seg_max * 8;
    const bool smem_sort = !h->tune_cub_sort && sort_smem <= kSortSmemMax;
    DevBuf d_k0, d_k1, d_v0, d_v1, d_tmp, d_ncand, d_coff, d_ctmp;
    CKS(d_k1.alloc(sizeof(uint64_t) * n_hits, h->stream)); CKS(d_v1.alloc(sizeof(uint32_t) * n_hits, h->stream));
    CKS(d_ncand.alloc(sizeof(uint32_t) * (ns + 1), h->stream)); CKS(d_coff.alloc(sizeof(uint32_t) * (ns + 1), h->stream));
    CKS(cudaMemsetAsync(d_ncand.p, 0, sizeof(uint32_t) * (ns + 1), h->stream));
    if (smem_sort) {
        CKS(d_ctmp.alloc(sizeof(uint32_t) * n_hits, h->stream));
        const int threads = seg_pad <= 4096 ? 256 : seg_pad <= 8192 ? 512 : 1024;
        CKS(cudaFuncSetAttribute(hit_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemMax));
        hit_sort_kernel<<<ns, threads, sort_smem, h->stream>>>(sc, ix.d_positions, d_seeds.as<uint64_t>(), d_base.as<uint32_t>(), d_cnt.as<uint32_t>(), d_first.as<uint32_t>(),
                                                              d_hoff.as<uint32_t>(), d_soff.as<uint32_t>(), seg_max, seg_pad,
                                                              d_k1.as<uint64_t>(), d_v1.as<uint32_t>(), d_ncand.as<uint32_t>(), d_ctmp.as<uint32_t>());
        CKS(cudaGetLastError());
    } else {
        CKS(d_k0.alloc(sizeof(uint64_t) * n_hits, h->stream)); CKS(d_v0.alloc(sizeof(uint32_t) * n_hits, h->stream));
        hit_fill_kernel<<<gb, 256, 0, h->stream>>>(sc, ix.d_buckets, ix.d_positions, d_seeds.as<uint64_t>(), d_base.as<uint32_t>(), d_nseeds.as<uint32_t>(), ns, max_cap,
                                                  d_hoff.as<uint32_t>(), d_k0.as<uint64_t>(), d_v0.as<uint32_t>());
        CKS(cudaGetLastError());
        // std::stable_sort by bin_offset (seed_pos_table.cpp:338): ties (same seed) keep ascending hit order
        size_t tmp_bytes = 0;
        CKS(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tmp_bytes, d_k0.as<uint64_t>(), d_k1.as<uint64_t>(), d_v0.as<uint32_t>(), d_v1.as<uint32_t>(),
                                                      (int64_t)n_hits, (int64_t)ns, d_soff.as<uint32_t>(), d_soff.as<uint32_t>() + 1, h->stream));
        CKS(d_tmp.alloc(tmp_bytes, h->stream));
        CKS(cub::DeviceSegmentedSort::StableSortPairs(d_tmp.p, tmp_bytes, d_k0.as<uint64_t>(), d_k1.as<uint64_t>(), d_v0.as<uint32_t>(), d_v1.as<uint32_t>(),
                                                      (int64_t)n_hits, (int64_t)ns, d_soff.as<uint32_t>(), d_soff.as<uint32_t>() + 1, h->stream));
        candidate_kernel<false><<<(ns + 127) / 128, 128, 0, h->stream>>>(sc, d_k1.as<uint64_t>(), d_soff.as<uint32_t>(), ns, d_ncand.as<uint32_t>(), nullptr, nullptr);
        CKS(cudaGetLastError());
    }
    const uint64_t* keys = d_k1.as<uint64_t>(); const uint32_t* vals = d_v1.as<uint32_t>();
    if ((rc = exclusive_sum(h, d_ncand.as<uint32_t>(), d_coff.as<uint32_t>(), (int64_t)ns + 1))) return rc;
    CKS(cudaMemcpyAsync(anchor_begin, d_coff.p, sizeof(uint32_t) * (ns + 1), cudaMemcpyDeviceToHost, h->stream));
    CKS(stream_wait(h, h->stream));
    const uint32_t n_cands = anchor_begin[ns];
    *n_anchors = n_cands;
    h->stats.kernel_launches += 6;
    if (n_cands == 0) { *n_pool = 0; CKS(cudaEventRecord(h->ev1, h->stream)); CKS(stream_wait(h, h->stream)); CKS(cudaEventElapsedTime(&h->stats.last_kernel_ms, h->ev0, h->ev1)); return DARWIN_OK; }
    DevBuf d_cidx, d_cstr, d_wlo, d_wn, d_wn64, d_woff;
    CKS(d_cidx.alloc(sizeof(uint32_t) * n_cands, h->stream)); CKS(d_cstr.alloc(sizeof(uint32_t) * n_cands, h->stream)); CKS(d_wlo.alloc(sizeof(uint32_t) * n_cands, h->stream));
    CKS(d_wn.alloc(sizeof(uint32_t) * ((size_t)n_cands + 1), h->stream)); CKS(d_wn64.alloc(sizeof(uint64_t) * ((size_t)n_cands + 1), h->stream)); CKS(d_woff.alloc(sizeof(uint64_t) * ((size_t)n_cands + 1), h->stream));
    if (!smem_sort) {
        candidate_kernel<true><<<(ns + 127) / 128, 128, 0, h->stream>>>(sc, keys, d_soff.as<uint32_t>(), ns, nullptr, d_coff.as<uint32_t>(), d_cidx.as<uint32_t>());
        CKS(cudaGetLastError());
    }
    CKS(cudaMemsetAsync(d_wn.p, 0, sizeof(uint32_t) * ((size_t)n_cands + 1), h->stream));
    window_size_kernel<<<(n_cands + 127) / 128, 128, 0, h->stream>>>(sc, keys, d_soff.as<uint32_t>(), d_coff.as<uint32_t>(), ns, smem_sort ? nullptr : d_cidx.as<uint32_t>(), n_cands,
                                                                    d_cstr.as<uint32_t>(), d_wlo.as<uint32_t>(), d_wn.as<uint32_t>(),
                                                                    smem_sort ? d_ctmp.as<uint32_t>() : nullptr, d_cidx.as<uint32_t>(), d_max.as<uint32_t>() + 1);
    CKS(cudaGetLastError());
    widen_kernel<<<(n_cands + 1 + 255) / 256, 256, 0, h->stream>>>(d_wn.as<uint32_t>(), n_cands + 1, d_wn64.as<uint64_t>());
    CKS(cudaGetLastError());
    if ((rc = exclusive_sum(h, d_wn64.as<uint64_t>(), d_woff.as<uint64_t>(), (int64_t)n_cands + 1))) return rc;
    uint64_t n_win = 0; uint32_t win_max = 0;
    CKS(cudaMemcpyAsync(&n_win, d_woff.as<uint64_t>() + n_cands, sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CKS(cudaMemcpyAsync(&win_max, d_max.as<uint32_t>() + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CKS(stream_wait(h, h->stream));
    const uint64_t need_pool = n_win + n_cands;
    *n_pool = need_pool;
    if (keep_pool) { anchor_vec->resize(n_cands); anchors = anchor_vec->data(); }
    else if (n_cands > anchors_cap || need_pool > pool_cap) { h->err = "seed output capacity: need " + std::to_string(n_cands) + " anchors, " + std::to_string(need_pool) + " pool entries"; return DARWIN_ERR_CAPACITY; }
    DevBuf d_w0, d_w1, d_pool, d_tanc, d_anc, d_tmp2;
    CKS(d_pool.alloc(sizeof(uint64_t) * need_pool, h->stream));
    CKS(d_tanc.alloc(sizeof(DarwinSeedAnchor) * n_cands, h->stream)); CKS(d_anc.alloc(sizeof(DarwinSeedAnchor) * n_cands, h->stream));
    uint32_t win_pad = 64; while (win_pad < win_max) win_pad <<= 1;
    if (!h->tune_cub_sort && (size_t)win_pad * 8 <= kSortSmemMax) {
        // one CTA per candidate sorts its window in shared memory and chains it there (window_chain_kernel); one launch per
        // size class, so that the many small windows are not held to the CTA shape of the largest one
        CKS(cudaFuncSetAttribute(window_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemMax));
        const uint32_t cls_hi[3] = {1024u, 4096u, win_pad};
        const int cls_threads[3] = {128, 256, 1024};
        uint32_t n_lo = 0;
        for (int k = 0; k < 3; k++) {
            const uint32_t n_hi = std::min(cls_hi[k], win_pad);
            window_chain_kernel<<<n_cands, cls_threads[k], (size_t)n_hi * 8, h->stream>>>(
                keys, vals, d_cidx.as<uint32_t>(), d_wlo.as<uint32_t>(), d_wn.as<uint32_t>(), d_woff.as<uint64_t>(), n_lo, n_hi,
                d_pool.as<uint64_t>(), d_tanc.as<DarwinSeedAnchor>());
            CKS(cudaGetLastError());
            h->stats.kernel_launches++;
            if (n_hi >= win_pad) break;
            n_lo = n_hi;
        }
    } else {
        CKS(d_w0.alloc(sizeof(uint64_t) * n_win, h->stream)); CKS(d_w1.alloc(sizeof(uint64_t) * n_win, h->stream));
        window_copy_kernel<<<(unsigned)(((uint64_t)n_cands * 32 + 255) / 256), 256, 0, h->stream>>>(keys, vals, n_cands, d_wlo.as<uint32_t>(), d_wn.as<uint32_t>(), d_woff.as<uint64_t>(), d_w0.as<uint64_t>());
        CKS(cudaGetLastError());
        size_t tmp_bytes = 0;
        CKS(cub::DeviceSegmentedSort::SortKeys(nullptr, tmp_bytes, d_w0.as<uint64_t>(), d_w1.as<uint64_t>(), (int64_t)n_win, (int64_t)n_cands,
                                               d_woff.as<uint64_t>(), d_woff.as<uint64_t>() + 1, h->stream));
        CKS(d_tmp2.alloc(tmp_bytes, h->stream));
        CKS(cub::DeviceSegmentedSort::SortKeys(d_tmp2.p, tmp_bytes, d_w0.as<uint64_t>(), d_w1.as<uint64_t>(), (int64_t)n_win, (int64_t)n_cands,
                                               d_woff.as<uint64_t>(), d_woff.as<uint64_t>() + 1, h->stream));
        chain_kernel<<<(n_cands + 127) / 128, 128, 0, h->stream>>>(keys, vals, d_cidx.as<uint32_t>(), n_cands, d_wn.as<uint32_t>(), d_woff.as<uint64_t>(), d_w1.as<uint64_t>(),
                                                                  d_pool.as<uint64_t>(), d_tanc.as<DarwinSeedAnchor>());
        CKS(cudaGetLastError());
        h->stats.kernel_launches += 2;
    }
    order_kernel<<<(ns + 127) / 128, 128, 0, h->stream>>>(d_coff.as<uint32_t>(), ns, d_tanc.as<DarwinSeedAnchor>(), d_anc.as<DarwinSeedAnchor>());
    CKS(cudaGetLastError());
    CKS(cudaEventRecord(h->ev1, h->stream));
    h->stats.kernel_launches += 5;
    CKS(cudaMemcpyAsync(anchors, d_anc.p, sizeof(DarwinSeedAnchor) * n_cands, cudaMemcpyDeviceToHost, h->stream));
    if (keep_pool) { keep_pool->p = d_pool.p; keep_pool->st = h->stream; d_pool.p = nullptr; }
    else CKS(cudaMemcpyAsync(pool, d_pool.p, sizeof(uint64_t) * need_pool, cudaMemcpyDeviceToHost, h->stream));
    CKS(stream_wait(h, h->stream));
    CKS(cudaEventElapsedTime(&h->stats.last_kernel_ms, h->ev0, h->ev1));
    return DARWIN_OK;
}

// D-SOFT seeding on the GPU: minimizers of the reads, seed position table look-ups, bin counting,
// candidate anchors and their collinear chained hits -- SeedPosTable::DSOFT (software/seed_pos_table.cpp:252-553) and the
// minimizer iteration it rests on (software/seed_pos_table.h:280-372, ntcoding.h:35-67) -- with results identical to
// the reference's (CPU twin: oracle/dsoft_oracle.c, pinned to the compiled reference).
//
// The work is HBM-bound integer work (hash, table look-ups, sorting), so it is laid out as a short pipeline of flat
// kernels over the whole read batch with CUB's segmented sorts in between, not as one monolithic kernel:
//   minimizer_kernel   one CTA per read strand: 2-bit codes -> hash32 -> window minimum -> emission rule (block scans),
//                      keeps the seeds the reference visits (all up to index N+1, every max_stride-th afterwards)
//   hit_count / hit_fill   one thread per seed: bucket look-up, hits at or beyond the seed's read offset
//   (segmented stable sort by (bin, read offset); ties keep ascending hit order like std::stable_sort)
//   candidate_kernel   one thread per strand: the sequential bin-coverage scan (curr_count rule), count pass + fill pass
//   window_chain_kernel   one CTA per candidate: its SV window as (hit << 32 | offset) keys sorted in shared memory, the
//                      greedy collinear chains to the left and right of the anchor as min / max scans over the sorted window
//   (windows beyond shared memory: window_copy_kernel -> CUB segmented sort -> chain_kernel, one thread per candidate)
//   order_kernel       one thread per strand: final order (chained hits descending, hit_offset ascending)
#pragma once
#include <vector>
#include "gact_common.cuh"

namespace dsoft {

constexpr int kChunk = 2048;                 // positions per minimizer pass of a CTA
constexpr int kMinThreads = 256;

struct SeedConst {
    int k, w, N, threshold, max_stride, overlap;
    uint32_t bin_size, max_occ, sv_bins, kmask;
};

__device__ __forceinline__ uint32_t hash32(uint32_t key, uint32_t m) {        // ntcoding.h:56-67
    key = (~key + (key << 21)) & m;
    key = key ^ (key >> 24);
    key = ((key + (key << 3)) + (key << 8)) & m;
    key = key ^ (key >> 14);
    key = ((key + (key << 2)) + (key << 4)) & m;
    key = key ^ (key >> 28);
    key = (key + (key << 31)) & m;
    return key;
}

// 2-bit code of strand-local position q of a sequence (seed_pos_table.h:63-84: A 0, C 1, G 2, T 3, N and padding 0).
// strand 1 = reverse complement of the forward sequence; an N stays an N (main.cpp:83-113), i.e. codes as 0.
__device__ __forceinline__ uint32_t code_at(const uint8_t* __restrict__ arena, uint64_t addr, uint32_t len, int strand, uint32_t q) {
    if (q >= len) return 0;
    const uint32_t c = gact::arena_code(arena, strand ? addr + (len - 1 - q) : addr + q);
    if (c > 3) return 0;
    return strand ? 3 - c : c;
}

// Block-wide inclusive scans over kMinThreads threads, 8 items per thread handled by the caller.
__device__ __forceinline__ int block_scan_max(int v, int* warp_tot) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = max(v, t); }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    int pre = INT_MIN;
    for (int k = 0; k < wid; k++) pre = max(pre, warp_tot[k]);
    __syncthreads();
    return max(v, pre);
}
__device__ __forceinline__ int block_scan_sum(int v, int* warp_tot) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    int pre = 0;
    for (int k = 0; k < wid; k++) pre += warp_tot[k];
    __syncthreads();
    return v + pre;
}

// Minimizers of one sequence strand (iterate_minimizers_qw), kChunk positions at a time.
//   MODE 0  reads: one CTA per strand walks its chunks in order; seeds are numbered in emission order and only the ones
//           the reference visits are kept (seeds[out_base + slot] = (p << 32) | m).
//   MODE 1  reference chromosomes, pass B: one CTA per (chunk, chromosome); the run start carried into the chunk comes
//           from carry[]; every minimizer goes to seeds[] through an atomic cursor as (m << 32) | arena position and
//           is counted in hist[m + 1] (order is irrelevant for the table).
//   MODE 2  reference chromosomes, pass A: only the position of the last change inside the chunk (for carry[]).
struct MinJob { uint64_t addr; uint32_t len; uint32_t out_base; };   // out_base: first seed slot of this strand (MODE 0)

template <int MODE>
__global__ void __launch_bounds__(kMinThreads)
minimizer_kernel(const uint8_t* __restrict__ arena, const SeedConst sc, const MinJob* __restrict__ jobs,
                 uint64_t* __restrict__ seeds, uint32_t* __restrict__ n_seeds,
                 int* __restrict__ carry, unsigned long long* __restrict__ cursor, uint32_t* __restrict__ hist,
                 uint32_t chunks_per_job) {            // MODE 1/2: 1-D grid of jobs x chunks_per_job blocks (no 65535 limit on either)
    constexpr int IPT = kChunk / kMinThreads;                    // 8 positions per thread
    __shared__ uint8_t codes[kChunk + 64];
    __shared__ uint32_t hs[kChunk + 32];                         // hashes of positions base-w .. base+kChunk-1 (w <= 32)
    __shared__ uint32_t ms[kChunk + 1];                          // window minima, ms[0] = m of position base-1
    __shared__ int warp_tot[kMinThreads / 32];
    __shared__ int incl_s[kMinThreads];
    __shared__ int tot_s, last_s;
    const int k = sc.k, w = sc.w, lead = w - 1;
    const uint32_t jb = MODE == 0 ? blockIdx.x : blockIdx.x / chunks_per_job;
    const MinJob job = jobs[jb];
    const int strand = MODE == 0 ? (int)(blockIdx.x & 1) : 0;
    const uint32_t len = job.len;
    const uint32_t centinel = (~0x0fu & (len + 15u)) - (uint32_t)k;
    const uint32_t end = centinel < 16u ? 16u : centinel;       // the first batch always covers p = 0..15
    int carry_start = 0;                                         // run start carried into the chunk (last_p = 0 initially)
    int emitted = 0;                                             // minimizers emitted so far (MODE 0)
    const uint32_t n_chunks = (end + kChunk - 1) / kChunk;
    uint32_t chunk = MODE == 0 ? 0u : blockIdx.x % chunks_per_job;
    if (MODE != 0 && chunk >= n_chunks) return;
    if (MODE == 1) carry_start = carry[(size_t)jb * chunks_per_job + chunk];
    for (; chunk < n_chunks; chunk++) {
        const uint32_t base = chunk * kChunk;
        // codes of positions base-w .. base+kChunk+k-2
        for (int t = threadIdx.x; t < kChunk + w + k - 1; t += kMinThreads) {
            const int64_t q = (int64_t)base - w + t;
            codes[t] = (q >= 0) ? (uint8_t)code_at(arena, job.addr, len, strand, (uint32_t)q) : 0;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < kChunk + w; t += kMinThreads) {
            uint32_t seed = 0;
            for (int c = 0; c < k; c++) seed |= (uint32_t)codes[t + c] << (2 * c);
            hs[t] = hash32(seed & sc.kmask, sc.kmask);           // position base - w + t
        }
        __syncthreads();
        // window minima: ms[1 + j] = m(base + j); ms[0] = m(base - 1); positions before w-1 count as 0 (the initial last_m)
        for (int t = threadIdx.x; t < kChunk + 1; t += kMinThreads) {
            const int64_t p = (int64_t)base - 1 + t;
            uint32_t m = 0;
            if (p >= (int64_t)lead) {
                m = 0x7FFFFFFFu;                                 // Min_Window's start value, ntcoding.h:36
                for (int c = 0; c < w; c++) m = min(m, hs[t - 1 + w - c]);        // positions p-w+1 .. p
            }
            ms[t] = m;
        }
        __syncthreads();
        // per position: change flag and start of its run of equal minima (max-scan over the chunk)
        int rs[IPT]; uint32_t cmask = 0;
        int local_last = INT_MIN;
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const int idx = threadIdx.x * IPT + j;
            const uint32_t p = base + idx;
            const bool live = p < end && p >= (uint32_t)lead;
            if (live && ms[idx + 1] != ms[idx]) { cmask |= 1u << j; local_last = (int)p; }
            rs[j] = local_last;
        }
        const int incl = block_scan_max(local_last, warp_tot);
        incl_s[threadIdx.x] = incl;
        __syncthreads();
        if (MODE == 2) {
            if (threadIdx.x == kMinThreads - 1) carry[(size_t)jb * chunks_per_job + chunk] = incl;
            return;
        }
        const int excl = threadIdx.x ? incl_s[threadIdx.x - 1] : INT_MIN;
        int n_emit = 0; uint32_t emask = 0;
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const uint32_t p = base + threadIdx.x * IPT + j;
            const bool live = p < end && p >= (uint32_t)lead;
            int start = max(rs[j], excl);
            if (start == INT_MIN) start = carry_start;
            // emitted when the minimum changed or w positions after the last emission (seed_pos_table.h:312, :343)
            if (live && ((cmask >> j & 1u) || ((int)p - start) % w == 0)) { n_emit++; emask |= 1u << j; }
        }
        const int incl_sum = block_scan_sum(n_emit, warp_tot);
        int out_i = emitted + incl_sum - n_emit;
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            if (!(emask & (1u << j))) continue;
            const int idx = threadIdx.x * IPT + j;
            const uint32_t p = base + idx, m = ms[idx + 1];
            if (MODE == 0) {
                const int i = out_i++;
                // visited seeds: i <= N+1, then every max_stride-th (seed_pos_table.cpp:307-336); overlap mode stops at N+1
                int slot = -1;
                if (i <= sc.N + 1) slot = i;
                else if (!sc.overlap && (i - (sc.N + 1)) % sc.max_stride == 0) slot = sc.N + 1 + (i - (sc.N + 1)) / sc.max_stride;
                if (slot >= 0) seeds[job.out_base + slot] = ((uint64_t)p << 32) | m;
            } else {
                const unsigned long long at = atomicAdd(cursor, 1ull);
                seeds[at] = ((uint64_t)m << 32) | (uint32_t)(p + (uint32_t)job.addr);
                atomicAdd(hist + m + 1, 1u);
            }
        }
        if (MODE == 1) return;
        if (threadIdx.x == kMinThreads - 1) { tot_s = incl_sum; last_s = incl; }
        __syncthreads();
        emitted += tot_s;
        if (last_s != INT_MIN) carry_start = last_s;
        __syncthreads();
    }
    if (MODE == 0 && threadIdx.x == 0) {
        int visited = emitted;
        if (emitted > sc.N + 2) visited = sc.overlap ? sc.N + 2 : sc.N + 2 + (emitted - 1 - (sc.N + 1)) / sc.max_stride;
        n_seeds[blockIdx.x] = (uint32_t)visited;
    }
}

// run start carried into every chunk of a chromosome: position of the last change before it (0 = none yet)
__global__ void carry_kernel(int* __restrict__ carry, int n_chroms, int chunks_per_chrom) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chroms) return;
    int running = 0;
    for (int k = 0; k < chunks_per_chrom; k++) {
        const int last = carry[(size_t)c * chunks_per_chrom + k];
        carry[(size_t)c * chunks_per_chrom + k] = running;
        if (last != INT_MIN) running = last;
    }
}

// hits of one seed: bucket entries >= the seed's read offset (seed_pos_table.cpp:313-326); buckets are ascending
__global__ void hit_count_kernel(const SeedConst sc, const uint32_t* __restrict__ buckets, const uint32_t* __restrict__ positions,
                                 const uint64_t* __restrict__ seeds, const uint32_t* __restrict__ seed_base,
                                 const uint32_t* __restrict__ n_seeds, int n_strands, uint32_t slots_per_strand_max,
                                 uint32_t* __restrict__ cnt, uint32_t* __restrict__ first) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t strand = (uint32_t)(g / slots_per_strand_max), slot = (uint32_t)(g % slots_per_strand_max);
    if (strand >= (uint32_t)n_strands) return;
    const uint32_t cap = seed_base[strand + 1] - seed_base[strand];
    if (slot >= cap) return;
    uint32_t c = 0, f = 0;
    if (slot < n_seeds[strand]) {
        const uint64_t sd = seeds[seed_base[strand] + slot];
        const uint32_t offset = (uint32_t)(sd >> 32), m = (uint32_t)sd;
        const uint32_t s = buckets[m], e = buckets[m + 1];
        if (e - s <= sc.max_occ) {
            for (uint32_t j = s; j < e; j++) c += positions[j] >= offset;
            f = e - c;                                           // buckets are ascending: the hits are the suffix [f, e)
        }
    }
    cnt[seed_base[strand] + slot] = c;
    first[seed_base[strand] + slot] = f;
}

__global__ void hit_fill_kernel(const SeedConst sc, const uint32_t* __restrict__ buckets, const uint32_t* __restrict__ positions,
                                const uint64_t* __restrict__ seeds, const uint32_t* __restrict__ seed_base,
                                const uint32_t* __restrict__ n_seeds, int n_strands, uint32_t slots_per_strand_max,
                                const uint32_t* __restrict__ hit_off, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t strand = (uint32_t)(g / slots_per_strand_max), slot = (uint32_t)(g % slots_per_strand_max);
    if (strand >= (uint32_t)n_strands) return;
    if (slot >= n_seeds[strand] || slot >= seed_base[strand + 1] - seed_base[strand]) return;
    const uint64_t sd = seeds[seed_base[strand] + slot];
    const uint32_t offset = (uint32_t)(sd >> 32), m = (uint32_t)sd;
    const uint32_t s = buckets[m], e = buckets[m + 1];
    if (e - s > sc.max_occ) return;
    uint32_t at = hit_off[seed_base[strand] + slot];
    for (uint32_t j = s; j < e; j++) {
        const uint32_t hit = positions[j];
        if (hit >= offset) {
            keys[at] = ((uint64_t)((hit - offset) / sc.bin_size) << 32) | offset;
            vals[at] = hit;
            at++;
        }
    }
}

// ---- segment sorts in shared memory ---------------------------------------------------------------------------------
// The two sorts of D-SOFT work on short independent segments (the hits of one read strand: a few thousand; the SV window
// of one candidate: a few hundred), so each segment is sorted by ONE CTA in shared memory and touches HBM once on the way
// in and once on the way out.  Bitonic network on 64-bit keys; the stages whose partner distance is <= 32 stay inside
// the 64-element block a warp holds and need only a warp barrier (57 of the 78 stages at 4096 keys).
__device__ __forceinline__ uint32_t pow2_ceil(uint32_t n) { return n <= 64u ? 64u : 1u << (32 - __clz(n - 1)); }

__device__ __forceinline__ void cmpx(uint64_t* a, uint32_t t, uint32_t j, uint32_t k) {
    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
    const uint64_t x = a[i], y = a[p];
    if ((x > y) == ((i & k) == 0)) { a[i] = y; a[p] = x; }
}
// a[0, npad) ascending; npad a power of two >= 64; called by every thread of the CTA (blockDim a multiple of 32)
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, uint32_t npad) {
    const uint32_t half = npad >> 1;
    for (uint32_t k = 2; k <= npad; k <<= 1) {
        uint32_t j = k >> 1;
        for (; j > 32; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) cmpx(a, t, j, k);
            __syncthreads();
        }
        for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {      // a warp's 32 pairs = one 64-element block
            for (uint32_t jj = j; jj > 0; jj >>= 1) { cmpx(a, t, jj, k); __syncwarp(); }
        }
        __syncthreads();
    }
}

// inclusive block scan over blockDim (<= 1024) threads; warp_tot: 32 ints of shared memory
__device__ __forceinline__ int block_scan_sum_any(int v, int* warp_tot) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    int pre = 0;
    for (int k = 0; k < wid && k < nw; k++) pre += warp_tot[k];
    __syncthreads();
    return v + pre;
}

// Hits of one strand: fill (positions[first .. first + cnt) of every visited seed, in seed order = ascending read offset),
// STABLE sort by bin -- the input is already in (offset, hit) order, so this is the reference's std::stable_sort by
// (bin, offset) with ties in ascending hit order (seed_pos_table.cpp:338) -- and the sequential bin-coverage rule
// (:352-392), which restarts at every new bin and pushes at most once per bin: one thread per run of equal bins.
// Shared memory: key_s[npad_max] u64 (bin << 32 | index in fill order), off_s[cap], hit_s[cap].
// Output: keys (bin << 32 | offset), vals (hit) in sorted order; cand_tmp[lo ..): indices of the strand's candidates, ascending.
__global__ void hit_sort_kernel(const SeedConst sc, const uint32_t* __restrict__ positions, const uint64_t* __restrict__ seeds,
                                const uint32_t* __restrict__ seed_base, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ first,
                                const uint32_t* __restrict__ hit_off, const uint32_t* __restrict__ strand_hit_off, uint32_t cap, uint32_t npad_max,
                                uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ n_cand, uint32_t* __restrict__ cand_tmp) {
    extern __shared__ __align__(16) unsigned char dsoft_smem[];
    __shared__ int warp_tot[32];
    uint64_t* key_s = reinterpret_cast<uint64_t*>(dsoft_smem);
    uint32_t* off_s = reinterpret_cast<uint32_t*>(key_s + npad_max);
    uint32_t* hit_s = off_s + cap;
    const uint32_t s = blockIdx.x, lo = strand_hit_off[s], n = strand_hit_off[s + 1] - lo;
    if (n == 0) { if (threadIdx.x == 0) n_cand[s] = 0; return; }
    const uint32_t npad = pow2_ceil(n);
    for (uint32_t i = n + threadIdx.x; i < npad; i += blockDim.x) key_s[i] = ~0ull;
    const uint32_t base = seed_base[s], slots = seed_base[s + 1] - base;
    for (uint32_t slot = threadIdx.x; slot < slots; slot += blockDim.x) {
        const uint32_t c = cnt[base + slot];
        if (!c) continue;
        const uint32_t at = hit_off[base + slot] - lo, offset = (uint32_t)(seeds[base + slot] >> 32), p0 = first[base + slot];
        for (uint32_t j = 0; j < c; j++) {
            const uint32_t hit = positions[p0 + j];
            off_s[at + j] = offset; hit_s[at + j] = hit;
            key_s[at + j] = ((uint64_t)((hit - offset) / sc.bin_size) << 32) | (at + j);
        }
    }
    __syncthreads();
    bitonic_sort_smem(key_s, npad);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t k = key_s[i];
        const uint32_t idx = (uint32_t)k;
        const uint64_t out = (k & 0xFFFFFFFF00000000ull) | off_s[idx];
        keys[lo + i] = out; vals[lo + i] = hit_s[idx];
        key_s[i] = out;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) hit_s[i] = 0;       // now the push flags
    __syncthreads();
    const uint32_t ks = (uint32_t)sc.k, thr = (uint32_t)sc.threshold;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t k = key_s[i];
        const uint32_t bin = (uint32_t)(k >> 32);
        if (i && (uint32_t)(key_s[i - 1] >> 32) == bin) continue;             // not the head of its run
        uint32_t curr = ks, last_offset = (uint32_t)k;
        if (curr >= thr) { hit_s[i] = 1; continue; }
        for (uint32_t j = i + 1; j < n; j++) {
            const uint64_t kj = key_s[j];
            if ((uint32_t)(kj >> 32) != bin) break;
            const uint32_t offset = (uint32_t)kj;
            curr = ((offset - last_offset > ks) || curr == 0) ? curr + ks : curr + (offset - last_offset);
            if (curr >= thr) { hit_s[j] = 1; break; }
            last_offset = offset;
        }
    }
    __syncthreads();
    // ordered compaction: every thread owns a contiguous range
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x, a = min(n, threadIdx.x * per), b = min(n, a + per);
    int mine = 0;
    for (uint32_t i = a; i < b; i++) mine += (int)hit_s[i];
    const int incl = block_scan_sum_any(mine, warp_tot);
    uint32_t out = lo + (uint32_t)(incl - mine);
    for (uint32_t i = a; i < b; i++) if (hit_s[i]) cand_tmp[out++] = lo + i;
    if (threadIdx.x == blockDim.x - 1) n_cand[s] = (uint32_t)incl;
}

// largest segment of a prefix-sum array (sizes the shared memory of the segment sorts)
__global__ void seg_max_kernel(const uint32_t* __restrict__ off, int n_seg, uint32_t* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = s < n_seg ? off[s + 1] - off[s] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}

// candidate bins of one strand (seed_pos_table.cpp:352-392).  FILL = false: count only.
template <bool FILL>
__global__ void candidate_kernel(const SeedConst sc, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ strand_hit_off,
                                 int n_strands, uint32_t* __restrict__ n_cand, const uint32_t* __restrict__ cand_off,
                                 uint32_t* __restrict__ cand_hit_idx) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_strands) return;
    const uint32_t lo = strand_hit_off[s], hi = strand_hit_off[s + 1];
    uint32_t last_bin = 1u << 31, last_offset = 0, curr = 0, n = 0;
    const uint32_t ks = (uint32_t)sc.k, thr = (uint32_t)sc.threshold;
    uint32_t out = FILL ? cand_off[s] : 0;
    for (uint32_t i = lo; i < hi; i++) {
        const uint64_t key = keys[i];
        const uint32_t offset = (uint32_t)key, bin = (uint32_t)(key >> 32);
        bool push = false;
        if (bin == last_bin) {
            if (curr < thr) {
                curr = ((offset - last_offset > ks) || curr == 0) ? curr + ks : curr + (offset - last_offset);
                push = curr >= thr;
            }
        } else {
            last_bin = bin; curr = ks;
            push = curr >= thr;
        }
        if (push) { if (FILL) cand_hit_idx[out++] = i; n++; }
        last_offset = offset;
    }
    if (!FILL) n_cand[s] = n;
}

// SV window [ws, we) of every candidate inside its strand's sorted hits (seed_pos_table.cpp:403-428) -- sizes only
__global__ void window_size_kernel(const SeedConst sc, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ strand_hit_off,
                                   const uint32_t* __restrict__ cand_off, int n_strands, const uint32_t* __restrict__ cand_hit_idx,
                                   uint32_t n_cands, uint32_t* __restrict__ cand_strand, uint32_t* __restrict__ win_lo,
                                   uint32_t* __restrict__ win_n, const uint32_t* __restrict__ cand_tmp, uint32_t* __restrict__ cand_hit_out,
                                   uint32_t* __restrict__ win_max) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cands) return;
    // strand of this candidate: binary search in cand_off
    int a = 0, b = n_strands;
    while (b - a > 1) { const int mid = (a + b) >> 1; if (cand_off[mid] <= c) a = mid; else b = mid; }
    const uint32_t s = (uint32_t)a;
    cand_strand[c] = s;
    const uint32_t lo = strand_hit_off[s], hi = strand_hit_off[s + 1];
    uint32_t ci;
    if (cand_tmp) { ci = cand_tmp[lo + (c - cand_off[s])]; cand_hit_out[c] = ci; }      // hit_sort_kernel left the list per strand
    else ci = cand_hit_idx[c];
    const uint32_t cb = (uint32_t)(keys[ci] >> 32);
    const uint32_t bmin = cb >= sc.sv_bins ? cb - sc.sv_bins : 0u, bmax = cb + sc.sv_bins;   // bmin <= bin < bmax
    uint32_t x = lo, y = hi;
    while (x < y) { const uint32_t mid = (x + y) >> 1; if ((uint32_t)(keys[mid] >> 32) < bmin) x = mid + 1; else y = mid; }
    const uint32_t ws = x;
    y = hi;
    while (x < y) { const uint32_t mid = (x + y) >> 1; if ((uint32_t)(keys[mid] >> 32) < bmax) x = mid + 1; else y = mid; }
    win_lo[c] = ws; win_n[c] = x - ws;
    if (win_max) atomicMax(win_max, x - ws);
}

// Exclusive scan over the CTA's threads (blockDim a multiple of 32, <= 1024) for any associative op; *total = the fold over
// all threads (every thread gets it).  warp_tot: 32 words of shared memory.
template <class Op>
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t ident, uint32_t* warp_tot, uint32_t* total, Op op) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = op(t, inc); }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint32_t pre = ident, all = ident;
    for (int k = 0; k < nw; k++) { const uint32_t w = warp_tot[k]; if (k < wid) pre = op(pre, w); all = op(all, w); }
    __syncthreads();
    uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = ident;
    if (total) *total = all;
    return op(pre, ex);
}
struct OpMinU { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a < b ? a : b; } };
struct OpMaxU { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };
struct OpAddU { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; } };

// One side of the greedy collinear chain (seed_pos_table.cpp:430-497) on a window sorted by (hit, offset), as a scan.
// The reference walks away from the anchor and takes a hit v when hit(cur) >= hit(v) && offset(cur) >= offset(v) (left; <=
// on the right), then cur = v.  Walking a sorted window the hit condition always holds, so a hit is taken iff its offset is
// <= (>=) the offset of the last hit taken -- and since a hit that is NOT taken has an offset beyond that bound, the bound
// is simply the running minimum (maximum) of ALL offsets seen so far, anchor included: an exclusive min / max scan.
// Elements are a[first + DIR * p], p = 0 (next to the anchor) .. count - 1; the K hits taken go to out[K - 1 - r], r = their
// rank walking away from the anchor (left chain ascending, right chain descending, as the reference stores them).
template <int DIR>
__device__ __forceinline__ uint32_t chain_side(const uint64_t* a, uint32_t first, uint32_t count, uint32_t anchor_off,
                                               uint64_t* __restrict__ out, uint32_t* warp_tot) {
    const uint32_t seg = (count + blockDim.x - 1) / blockDim.x;
    const uint32_t p0 = min(count, threadIdx.x * seg), p1 = min(count, p0 + seg);
    uint32_t ext = DIR < 0 ? 0xFFFFFFFFu : 0u;
    for (uint32_t p = p0; p < p1; p++) {
        const uint32_t o = (uint32_t)a[DIR < 0 ? first - p : first + p];
        ext = DIR < 0 ? min(ext, o) : max(ext, o);
    }
    uint32_t bound = DIR < 0 ? block_scan_excl(ext, 0xFFFFFFFFu, warp_tot, nullptr, OpMinU())
                             : block_scan_excl(ext, 0u, warp_tot, nullptr, OpMaxU());
    bound = DIR < 0 ? min(bound, anchor_off) : max(bound, anchor_off);
    uint32_t cnt = 0, b = bound;
    for (uint32_t p = p0; p < p1; p++) {
        const uint32_t o = (uint32_t)a[DIR < 0 ? first - p : first + p];
        if (DIR < 0 ? o <= b : o >= b) { cnt++; b = o; }
    }
    uint32_t K = 0;
    uint32_t rank = block_scan_excl(cnt, 0u, warp_tot, &K, OpAddU());
    b = bound;
    for (uint32_t p = p0; p < p1; p++) {
        const uint64_t v = a[DIR < 0 ? first - p : first + p];
        const uint32_t o = (uint32_t)v;
        if (DIR < 0 ? o <= b : o >= b) { out[K - 1 - rank] = v; rank++; b = o; }
    }
    return K;
}

// SV window of one candidate, sorted by (hit, offset) in shared memory (seed_pos_table.cpp:403-428) AND chained there
// (:430-497): one CTA per candidate, straight from the strand's sorted hits, without a round trip of the sorted windows
// through HBM and with the chains as parallel scans (chain_side) instead of one thread per candidate walking its window
// twice per side.  A launch serves the candidates whose window
// size is in (n_lo, n_hi]: windows differ by two orders of magnitude, and a CTA sized for the largest one wastes the SM on
// the many small ones.  Output as chain_kernel: pool[win_off[c] + c ..): left chain ascending, then right chain descending.
__global__ void window_chain_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                    const uint32_t* __restrict__ cand_hit_idx, const uint32_t* __restrict__ win_lo,
                                    const uint32_t* __restrict__ win_n, const uint64_t* __restrict__ win_off, uint32_t n_lo, uint32_t n_hi,
                                    uint64_t* __restrict__ pool, DarwinSeedAnchor* __restrict__ tmp_anchor) {
    extern __shared__ __align__(16) unsigned char dsoft_smem[];
    __shared__ uint32_t warp_tot[32];
    uint64_t* a = reinterpret_cast<uint64_t*>(dsoft_smem);
    const uint32_t c = blockIdx.x, n = win_n[c], lo = win_lo[c];
    if (n > n_hi || (n <= n_lo && n_lo != 0)) return;             // another launch's size class (block-uniform)
    const uint32_t npad = pow2_ceil(n);
    for (uint32_t j = threadIdx.x; j < npad; j += blockDim.x)
        a[j] = j < n ? ((uint64_t)vals[lo + j] << 32) | (uint32_t)keys[lo + j] : ~0ull;
    __syncthreads();
    bitonic_sort_smem(a, npad);
    const uint32_t hi_idx = cand_hit_idx[c];
    const uint64_t anchor = ((uint64_t)vals[hi_idx] << 32) | (uint32_t)keys[hi_idx];
    uint32_t x = 0, y = n;                                          // position of the anchor ((hit, offset) pairs are unique)
    while (x < y) { const uint32_t mid = (x + y) >> 1; if (a[mid] < anchor) x = mid + 1; else y = mid; }
    const uint32_t ai = x;
    const uint64_t base = win_off[c] + c;                          // regions are win_n + 1 long
    uint64_t* out = pool + base;
    const uint32_t kl = chain_side<-1>(a, ai - 1, ai, (uint32_t)anchor, out, warp_tot);
    const uint32_t kr = chain_side<+1>(a, ai + 1, n > ai ? n - ai - 1 : 0u, (uint32_t)anchor, out + kl + 1, warp_tot);
    if (threadIdx.x == 0) {
        out[kl] = anchor; out[kl + 1 + kr] = anchor;
        DarwinSeedAnchor an;
        an.hit_offset = anchor; an.left_off = base; an.left_n = kl + 1; an.right_off = base + kl + 1; an.right_n = kr + 1;
        tmp_anchor[c] = an;
    }
}

// copy every candidate's window as (hit << 32 | offset) keys; one warp per candidate
__global__ void window_copy_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n_cands,
                                   const uint32_t* __restrict__ win_lo, const uint32_t* __restrict__ win_n,
                                   const uint64_t* __restrict__ win_off, uint64_t* __restrict__ wkeys) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_cands) return;
    const uint32_t lo = win_lo[c], n = win_n[c];
    const uint64_t at = win_off[c];
    for (uint32_t j = threadIdx.x & 31; j < n; j += 32)
        wkeys[at + j] = ((uint64_t)vals[lo + j] << 32) | (uint32_t)keys[lo + j];
}

// greedy collinear chains of one candidate on its window sorted by (hit, offset) (seed_pos_table.cpp:430-497).
// out region of candidate c: pool[c_off .. c_off + win_n + 1): left chain (ascending) first, right chain (descending) after it.
__global__ void chain_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ cand_hit_idx,
                             uint32_t n_cands, const uint32_t* __restrict__ win_n, const uint64_t* __restrict__ win_off,
                             const uint64_t* __restrict__ wkeys, uint64_t* __restrict__ pool, DarwinSeedAnchor* __restrict__ tmp_anchor) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cands) return;
    const uint32_t hi_idx = cand_hit_idx[c];
    const uint64_t anchor = ((uint64_t)vals[hi_idx] << 32) | (uint32_t)keys[hi_idx];
    const uint32_t n = win_n[c];
    const uint64_t* w = wkeys + win_off[c];
    uint64_t* out = pool + win_off[c] + c;                       // regions are win_n + 1 long
    // position of the anchor inside its sorted window ((hit, offset) pairs are unique)
    uint32_t x = 0, y = n;
    while (x < y) { const uint32_t mid = (x + y) >> 1; if (w[mid] < anchor) x = mid + 1; else y = mid; }
    const uint32_t ai = x;
    // left: walk from the anchor downwards; two passes (count, then write ascending)
    uint32_t kl = 1; uint64_t cur = anchor;
    for (uint32_t h = ai; h-- > 0;) {
        const uint64_t v = w[h];
        if ((uint32_t)(cur >> 32) >= (uint32_t)(v >> 32) && (uint32_t)cur >= (uint32_t)v) { kl++; cur = v; }
    }
    uint32_t pos = kl - 1; out[pos] = anchor; cur = anchor;
    for (uint32_t h = ai; h-- > 0;) {
        const uint64_t v = w[h];
        if ((uint32_t)(cur >> 32) >= (uint32_t)(v >> 32) && (uint32_t)cur >= (uint32_t)v) { out[--pos] = v; cur = v; }
    }
    // right: walk upwards; written descending
    uint32_t kr = 1; cur = anchor;
    for (uint32_t h = ai + 1; h < n; h++) {
        const uint64_t v = w[h];
        if ((uint32_t)(cur >> 32) <= (uint32_t)(v >> 32) && (uint32_t)cur <= (uint32_t)v) { kr++; cur = v; }
    }
    uint64_t* outr = out + kl;
    pos = kr - 1; outr[pos] = anchor; cur = anchor;
    for (uint32_t h = ai + 1; h < n; h++) {
        const uint64_t v = w[h];
        if ((uint32_t)(cur >> 32) <= (uint32_t)(v >> 32) && (uint32_t)cur <= (uint32_t)v) { outr[--pos] = v; cur = v; }
    }
    DarwinSeedAnchor a;
    a.hit_offset = anchor; a.left_off = win_off[c] + c; a.left_n = kl; a.right_off = win_off[c] + c + kl; a.right_n = kr;
    tmp_anchor[c] = a;
}

// final order of a strand's anchors: chained hits descending, hit_offset ascending (seed_pos_table.cpp:506-510)
__global__ void order_kernel(const uint32_t* __restrict__ cand_off, int n_strands, const DarwinSeedAnchor* __restrict__ tmp_anchor,
                             DarwinSeedAnchor* __restrict__ anchors) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_strands) return;
    const uint32_t lo = cand_off[s], hi = cand_off[s + 1];
    for (uint32_t i = lo; i < hi; i++) {                          // insertion sort into the output (a handful per strand)
        const DarwinSeedAnchor a = tmp_anchor[i];
        const uint32_t na = a.left_n + a.right_n;
        uint32_t j = i;
        while (j > lo) {
            const DarwinSeedAnchor b = anchors[j - 1];
            const uint32_t nb = b.left_n + b.right_n;
            if (nb > na || (nb == na && b.hit_offset < a.hit_offset)) break;
            anchors[j] = b; j--;
        }
        anchors[j] = a;
    }
}

// ---- index build helpers ------------------------------------------------------------------------------------------
// sort the positions of every bucket that DSOFT will ever read (non-empty, <= max_occ entries; seed_pos_table.cpp:146-150)
__global__ void bucket_sort_kernel(const uint32_t* __restrict__ buckets, uint64_t n_buckets, uint32_t max_occ, uint32_t* __restrict__ positions) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    const uint32_t lo = buckets[b], hi = buckets[b + 1];
    if (hi - lo < 2 || hi - lo > max_occ) return;
    for (uint32_t i = lo + 1; i < hi; i++) {
        const uint32_t v = positions[i];
        uint32_t j = i;
        while (j > lo && positions[j - 1] > v) { positions[j] = positions[j - 1]; j--; }
        positions[j] = v;
    }
}

// scatter the (m << 32 | position) list into the table through per-bucket cursors (order fixed by bucket_sort_kernel)
__global__ void scatter_kernel(const uint64_t* __restrict__ list, uint64_t n, const uint32_t* __restrict__ buckets,
                               uint32_t* __restrict__ fill, uint32_t* __restrict__ positions) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = (uint32_t)(list[i] >> 32);
    const uint32_t at = buckets[m] + atomicAdd(fill + m, 1u);
    positions[at] = (uint32_t)list[i];
}

} // namespace dsoft

// The seed position table of one device (SeedPosTable's seedBuckets / seedPositions, seed_pos_table.cpp:59-160).
struct SeedIndex {
    dsoft::SeedConst sc{};
    uint32_t* d_buckets = nullptr; uint64_t n_buckets = 0;
    uint32_t* d_positions = nullptr; uint64_t n_positions = 0;
    bool ready = false, owner = true;
    std::vector<uint32_t> chr_start, chr_len;       // host copy: Index::chr_coord / Index::chr_len (padded)
};

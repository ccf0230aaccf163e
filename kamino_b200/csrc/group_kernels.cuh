// group_kernels.cuh — K7 ("next" row 2 of SURVEY 8f): grouping of the block observations by anchor pair
// and per-group length selection on the device.
//
// Replaces the HashMap merge of merge_species_extraction (/root/reference/src/group_extraction.rs:335-357:
// observations of one (left, right) anchor pair end up in one Vec, in emission order) by a STABLE LSD
// radix sort of the observation indices by (left, right) — after it the observations of a pair are
// contiguous and still in emission order — and select_unique_best_supported_length
// (/root/reference/src/group_sorting.rs:90-143) by a segmented reduction over each run of equal keys:
// support of a block length = number of DISTINCT species observing it; the group survives iff exactly one
// length attains the maximum support, support >= min_needed and length >= 2k+1.
//
// The sort is a stable LSD radix sort with 8-bit digits (histogram / scan / stable scatter, 2048 elements per
// CTA): 2 x ceil(3k / 8) passes over a few million 3k-bit key halves.
#pragma once
#include "../../include/kamino_b200.h"
#include "kmn_device.cuh"

namespace kmn {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr uint32_t RS_TILE = RS_THREADS * RS_ROUNDS;   // elements per CTA
constexpr uint32_t RS_BITS = 8;
constexpr uint32_t RS_BINS = 1u << RS_BITS;            // 8-bit digits
constexpr int GROUP_MAX_LENS = 12;                     // distinct block lengths tracked per group on the device
static_assert(RS_BINS == RS_THREADS, "one thread per bin in the per-CTA tables");

// per-CTA digit histogram of the current pass: block_hist[bin * n_blocks + block]
__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ perm, uint32_t n, uint32_t shift,
                  uint32_t* __restrict__ block_hist) {
    __shared__ uint32_t s_h[RS_BINS];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_h[uint32_t(key[perm[i]] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    block_hist[threadIdx.x * gridDim.x + blockIdx.x] = s_h[threadIdx.x];
}

// stable scatter: element i goes to (scanned base of its digit and CTA) + (rank among the CTA's earlier
// elements with the same digit).  Inside a round the rank comes from a warp match (lanes with the same digit,
// lower lane first) plus the counts of the earlier warps.
__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ perm_in, uint32_t n, uint32_t shift,
                     const uint32_t* __restrict__ block_base, uint32_t* __restrict__ perm_out) {
    constexpr int WARPS = RS_THREADS / 32;
    __shared__ uint32_t s_base[RS_BINS];
    __shared__ uint16_t s_wcnt[WARPS][RS_BINS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    s_base[threadIdx.x] = block_base[threadIdx.x * gridDim.x + blockIdx.x];
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; r++) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) s_wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        const uint32_t p = valid ? perm_in[i] : 0u;
        const uint32_t d = valid ? (uint32_t(key[p] >> shift) & (RS_BINS - 1)) : RS_BINS;   // RS_BINS: padding lane
        const unsigned same = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(same & ((1u << lane) - 1u));
        if (valid && rank == 0) s_wcnt[warp][d] = uint16_t(__popc(same));
        __syncthreads();
        if (valid) {
            uint32_t off = 0;
            for (unsigned w = 0; w < warp; w++) off += s_wcnt[w][d];
            perm_out[s_base[d] + off + rank] = p;
        }
        __syncthreads();
        {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) t += s_wcnt[w][threadIdx.x];
            s_base[threadIdx.x] += t;
        }
        // the next round's zeroing is ordered after these reads by its barrier pair
        __syncthreads();
    }
}

// Observations straight from the anchor pairs kmn_pair_batch left on the device (kmn_group_batches): pair i of a
// batch is observation out0 + i; its species is found in the batch's per-species pair offsets.
__global__ void __launch_bounds__(256)
obs_from_pairs_kernel(const kmn_pair* __restrict__ pairs, uint32_t n_pairs, const uint32_t* __restrict__ sp_pairs_off,
                      uint32_t n_sp, uint32_t sid_base, uint32_t sid_stride, uint32_t out0, uint64_t* __restrict__ left,
                      uint64_t* __restrict__ right, uint32_t* __restrict__ sid, uint32_t* __restrict__ len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const kmn_pair p = pairs[i];
    uint32_t lo = 0, hi = n_sp;   // last species s with sp_pairs_off[s] <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sp_pairs_off[mid] <= i) lo = mid; else hi = mid;
    }
    left[out0 + i] = p.left;
    right[out0 + i] = p.right;
    sid[out0 + i] = sid_base + lo * sid_stride;
    len[out0 + i] = p.len;
}

__global__ void iota_kernel(uint32_t* __restrict__ perm, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = i;
}

// head[i] = 1 iff sorted position i starts a new (left, right) run
__global__ void group_heads_kernel(const uint64_t* __restrict__ left, const uint64_t* __restrict__ right,
                                   const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = 1;
    if (i > 0) {
        const uint32_t a = perm[i], b = perm[i - 1];
        h = (left[a] != left[b]) || (right[a] != right[b]);
    }
    head[i] = h;
}

// group_first[g] = sorted position of the g-th run (head_scan = exclusive scan of head)
__global__ void group_firsts_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ head_scan, uint32_t n,
                                    uint32_t* __restrict__ group_first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && head[i]) group_first[head_scan[i]] = i;
}

// one thread per group: select_unique_best_supported_length.  Within a group the observations are in
// emission order, i.e. species-major, so "distinct species of a length" is a run count.
__global__ void __launch_bounds__(128)
group_select_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ sid, const uint32_t* __restrict__ len,
                    const uint32_t* __restrict__ group_first, uint32_t n_groups, uint32_t n, uint32_t min_needed,
                    uint32_t min_len, kmn_group* __restrict__ groups) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t i0 = group_first[g], i1 = g + 1 < n_groups ? group_first[g + 1] : n;
    uint32_t lens[GROUP_MAX_LENS], cnt[GROUP_MAX_LENS], last[GROUP_MAX_LENS];
    int m = 0;
    bool host = false;
    for (uint32_t i = i0; i < i1 && !host; i++) {
        const uint32_t o = perm[i], l = len[o], s = sid[o];
        int slot = -1;
#pragma unroll
        for (int q = 0; q < GROUP_MAX_LENS; q++)
            if (q < m && lens[q] == l) slot = q;
        if (slot < 0) {
            if (m == GROUP_MAX_LENS) { host = true; break; }   // too many lengths: the host decides this group
#pragma unroll
            for (int q = 0; q < GROUP_MAX_LENS; q++)
                if (q == m) { lens[q] = l; cnt[q] = 1; last[q] = s; }
            m++;
            continue;
        }
#pragma unroll
        for (int q = 0; q < GROUP_MAX_LENS; q++)
            if (q == slot) {
                if (s < last[q]) host = true;            // not species-major: the host decides this group
                else if (s != last[q]) { cnt[q]++; last[q] = s; }
            }
    }
    uint32_t best_len = 0, best_c = 0;
    bool tie = false;
#pragma unroll
    for (int q = 0; q < GROUP_MAX_LENS; q++)
        if (q < m) {
            if (cnt[q] > best_c) { best_len = lens[q]; best_c = cnt[q]; tie = false; }
            else if (cnt[q] == best_c) tie = true;
        }
    kmn_group out;
    out.first = i0;
    out.count = i1 - i0;
    out.best_len = best_len;
    out.coverage = best_c;
    out.flags = host ? KMN_GROUP_HOST : ((!tie && best_c >= min_needed && best_len >= min_len) ? KMN_GROUP_VALID : 0u);
    groups[g] = out;
}

}  // namespace kmn

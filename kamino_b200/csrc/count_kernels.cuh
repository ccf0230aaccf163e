// count_kernels.cuh — pass 1 on the recoded stream: exact per-species Bloom semantics + CMS build.
//
// Replaces the hot loop of count_species_kmers (/root/reference/src/group_extraction.rs:385-391):
//     if have == k && !bloom.test_and_set(roll) { cmsa.increment(roll) }
//
// ---- Bloom (proba_filter.rs:59-74) -------------------------------------------------------------
// test_and_set is sequential per species: an occurrence is "new" iff at least one of its two probe
// bits was still clear, i.e. iff it is the FIRST occurrence (in file order) touching one of its two
// bits (i0 != i1 always because h2 is forced odd).  Order-free restatement:
//     new(p) <=> p is the file-order-first toucher of bit i0(p) or of bit i1(p).
// The dense traffic goes to per-species BITSETS (2^25 bits = 4 MiB each, like the reference's
// filter) that stay resident in the 126 MB L2 for a group of species:
//   pass A  touch : old = atomicOr(bits[bit]).  The temporally-first toucher of a bit sees it clear
//                   ("was_first").  A LATE toucher marks multi[bit], publishes its file-order index
//                   with atomicMax(first[bit], tag(p)) and appends itself to a compact late list.
//   pass B  settle: a was_first probe whose bit is not in multi[] is the only toucher -> first.
//                   A was_first probe of a contended bit is first iff its p is smaller than the
//                   smallest late p (first[bit]); if so it sets fw[bit] ("first toucher wins").
//   pass C  late  : (late list only) a late probe is first iff it holds first[bit] and !fw[bit].
// tag(p) = (species_epoch << 32) | ~p with a per-species epoch that only grows, so the 64-bit
// first[] array is never cleared and only contended bits (a few % of the probes on bacterial
// proteomes) ever touch it.  False positives AND their order dependence are reproduced exactly.
//
// The flags (one bit per codes[] position: this occurrence passed the filter) feed cms_kernels.cuh.
#pragma once
#include "kmn_device.cuh"

namespace kmn {

constexpr int COUNT_THREADS = 256;
constexpr int MAX_GROUP_SPECIES = 16;  // species per launch == number of Bloom slots

struct GroupBounds {                   // codes[] boundaries of the species of one launch
    uint32_t n;                        // species in the group (<= MAX_GROUP_SPECIES)
    uint32_t first_species;            // index of the first species inside the batch
    uint32_t epoch0;                   // epoch of the first species (monotone over the context's life)
};

struct BloomSlots {
    uint32_t* bits;                 // [slots][2^20] touched bits
    uint32_t* multi;                // [slots][2^20] bits touched more than once
    uint32_t* fw;                   // [slots][2^20] contended bits won by their temporally-first toucher
    unsigned long long* first;      // [slots][2^25] tags of the earliest LATE toucher
};
constexpr uint32_t BLOOM_WORDS_LOG2 = BLOOM_BITS_LOG2 - 5;

__device__ __forceinline__ unsigned long long bloom_tag(uint32_t epoch, uint32_t p) {
    return (static_cast<unsigned long long>(epoch) << 32) | (0xFFFFFFFFu - p);
}

// bit j set <=> position c0+j lies inside [cb, ce)
__device__ __forceinline__ uint32_t range_mask16(uint32_t c0, uint32_t cb, uint32_t ce) {
    const uint32_t lo = c0 >= cb ? 0u : min(cb - c0, 16u);
    const uint32_t hi = c0 + 16u <= ce ? 16u : (ce > c0 ? ce - c0 : 0u);
    if (hi <= lo) return 0u;
    return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}

// Rare path of pass A, kept out of line: a probe found its bit already set.
__device__ __noinline__ void bloom_mark_late(BloomSlots bs, uint32_t sl, uint32_t bit, unsigned long long tag) {
    atomicOr(bs.multi + (size_t(sl) << BLOOM_WORDS_LOG2) + (bit >> 5), 1u << (bit & 31u));
    atomicMax(bs.first + (size_t(sl) << BLOOM_BITS_LOG2) + bit, tag);
}

// Rare path of pass B, kept out of line: a was_first probe of a contended bit.
__device__ __noinline__ bool bloom_first_wins(BloomSlots bs, uint32_t sl, uint32_t bit, unsigned long long tag) {
    if (tag > bs.first[(size_t(sl) << BLOOM_BITS_LOG2) + bit]) {   // earlier than every late toucher
        atomicOr(bs.fw + (size_t(sl) << BLOOM_WORDS_LOG2) + (bit >> 5), 1u << (bit & 31u));
        return true;
    }
    return false;
}

// One lane's 16-position chunk of pass A.  UNIFORM: the whole chunk lies inside species slot `sl`
// and inside the launch's range (the common case) — no per-position range / slot bookkeeping.
// Two half chunks: 16 atomicOr in flight per thread before any result is consumed.
template <bool UNIFORM>
__device__ __forceinline__ void bloom_touch_chunk(const uint4& v, uint64_t roll, int have, int k, uint64_t k_mask,
                                                  uint32_t c0, uint32_t range, const uint32_t* s_b, uint32_t sl,
                                                  const GroupBounds& g, const BloomSlots& bs,
                                                  uint32_t& first_bits, uint32_t& late_bits) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t idx[8][2], old[8][2];
        uint32_t vmask = 0, slots = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int j = 8 * h + q;
            const uint32_t code = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            if (code == CODE_INVALID) {
                roll = 0;
                have = 0;
            } else {
                roll = ((roll << 3) | uint64_t(code)) & k_mask;
                if (have < k) have++;
                if (have == k && (UNIFORM || ((range >> j) & 1u))) vmask |= 1u << q;
            }
            bloom_indices(roll, idx[q][0], idx[q][1]);
            if (!UNIFORM) {
                if ((vmask >> q) & 1u)
                    while (c0 + j >= s_b[sl + 1]) sl++;
                slots |= sl << (4 * q);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t s = UNIFORM ? sl : ((slots >> (4 * q)) & 15u);
            uint32_t* words = bs.bits + (size_t(s) << BLOOM_WORDS_LOG2);
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                old[q][pr] = 0;
                if ((vmask >> q) & 1u) old[q][pr] = atomicOr(words + (idx[q][pr] >> 5), 1u << (idx[q][pr] & 31u));
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (!((vmask >> q) & 1u)) continue;
            const int j = 8 * h + q;
            const uint32_t s = UNIFORM ? sl : ((slots >> (4 * q)) & 15u);
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                if (old[q][pr] & (1u << (idx[q][pr] & 31u))) {   // late toucher of this bit
                    bloom_mark_late(bs, s, idx[q][pr], bloom_tag(g.epoch0 + s, c0 + j - s_b[s]));
                    late_bits |= 1u << (2 * j + pr);
                } else {
                    first_bits |= 1u << (2 * j + pr);
                }
            }
        }
    }
}

// slot of position c (c inside the launch's range)
__device__ __forceinline__ uint32_t slot_of(const uint32_t* s_b, uint32_t n, uint32_t c) {
    uint32_t sl = 0;
    while (sl + 1 < n && c >= s_b[sl + 1]) sl++;
    return sl;
}

// Pass A: touch.
__global__ void __launch_bounds__(COUNT_THREADS)
bloom_touch_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, GroupBounds g,
                   int k, uint64_t k_mask, BloomSlots bs, uint32_t* __restrict__ wf,
                   uint32_t* __restrict__ late_list, uint32_t* __restrict__ late_count) {
    __shared__ uint32_t s_b[MAX_GROUP_SPECIES + 1];
    if (threadIdx.x <= g.n) s_b[threadIdx.x] = sp_cstart[g.first_species + threadIdx.x];
    __syncthreads();
    const uint32_t cb = s_b[0], ce = s_b[g.n];
    const uint32_t warps_per_grid = gridDim.x * (COUNT_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (COUNT_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tile_lo = cb / WARP_TILE, tile_hi = (ce + WARP_TILE - 1) / WARP_TILE;
    for (uint32_t t = tile_lo + warp_id; t < tile_hi; t += warps_per_grid) {
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        uint4 v;
        uint64_t roll;
        int have;
        warp_tile_seed<true>(codes, t * WARP_TILE, k, k_mask, v, roll, have);
        const uint32_t range = range_mask16(c0, cb, ce);
        uint32_t first_bits = 0;  // bit 2j / 2j+1: probe 0 / 1 of position c0+j saw its bit clear
        uint32_t late_bits = 0;   // same indexing: the probe found its bit already set
        if (range) {
            const uint32_t sl0 = slot_of(s_b, g.n, max(c0, cb));
            if (range == 0xFFFFu && c0 + CHUNK <= s_b[sl0 + 1])
                bloom_touch_chunk<true>(v, roll, have, k, k_mask, c0, range, s_b, sl0, g, bs, first_bits, late_bits);
            else
                bloom_touch_chunk<false>(v, roll, have, k, k_mask, c0, range, s_b, sl0, g, bs, first_bits, late_bits);
            // chunks straddling a group boundary are shared by two stream-ordered launches
            wf[c0 / CHUNK] = range == 0xFFFFu ? first_bits : (wf[c0 / CHUNK] | first_bits);
        }
        // warp-aggregated append of the late probes: one global atomic per warp tile
        const uint32_t n_late = __popc(late_bits);
        uint32_t incl = n_late;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= unsigned(d)) incl += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            uint32_t base = 0;
            if (lane == 31) base = atomicAdd(late_count, total);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - n_late;
            while (late_bits) {
                const int b = __ffs(late_bits) - 1;
                late_bits &= late_bits - 1;
                late_list[base++] = (c0 + (b >> 1)) | (uint32_t(b & 1) << 31);
            }
        }
    }
}

// One lane's chunk of pass B: all multi[] loads of a half chunk are in flight together.
template <bool UNIFORM>
__device__ __forceinline__ uint32_t bloom_settle_chunk(const uint4& v, uint64_t roll, int have, int k, uint64_t k_mask,
                                                       uint32_t c0, uint32_t range, uint32_t first_bits,
                                                       const uint32_t* s_b, uint32_t sl, const GroupBounds& g,
                                                       const BloomSlots& bs, uint32_t* s_cnt) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t new_bits = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t idx[8][2], mw[8][2];
        uint32_t vmask = 0, slots = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int j = 8 * h + q;
            const uint32_t code = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            if (code == CODE_INVALID) {
                roll = 0;
                have = 0;
            } else {
                roll = ((roll << 3) | uint64_t(code)) & k_mask;
                if (have < k) have++;
                if (have == k && (UNIFORM || ((range >> j) & 1u))) vmask |= 1u << q;
            }
            bloom_indices(roll, idx[q][0], idx[q][1]);
            if (!UNIFORM) {
                if ((vmask >> q) & 1u)
                    while (c0 + j >= s_b[sl + 1]) sl++;
                slots |= sl << (4 * q);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int j = 8 * h + q;
            const uint32_t s = UNIFORM ? sl : ((slots >> (4 * q)) & 15u);
            const uint32_t* words = bs.multi + (size_t(s) << BLOOM_WORDS_LOG2);
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                mw[q][pr] = 0;
                if (((vmask >> q) & 1u) && ((first_bits >> (2 * j + pr)) & 1u)) mw[q][pr] = words[idx[q][pr] >> 5];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (!((vmask >> q) & 1u)) continue;
            const int j = 8 * h + q;
            const uint32_t s = UNIFORM ? sl : ((slots >> (4 * q)) & 15u);
            bool is_new = false;
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                if (!((first_bits >> (2 * j + pr)) & 1u)) continue;   // late probes: bloom_late_kernel
                if (!(mw[q][pr] & (1u << (idx[q][pr] & 31u))))
                    is_new = true;                                     // only toucher of this bit
                else if (bloom_first_wins(bs, s, idx[q][pr], bloom_tag(g.epoch0 + s, c0 + j - s_b[s])))
                    is_new = true;
            }
            if (is_new) {
                new_bits |= 1u << j;
                if (!UNIFORM) atomicAdd(&s_cnt[s], 1u);
            }
        }
    }
    if (UNIFORM && new_bits) atomicAdd(&s_cnt[sl], uint32_t(__popc(new_bits)));
    return new_bits;
}

// Pass B: settle every was_first probe.
__global__ void __launch_bounds__(COUNT_THREADS)
bloom_settle_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, GroupBounds g,
                    int k, uint64_t k_mask, BloomSlots bs, const uint32_t* __restrict__ wf,
                    uint16_t* __restrict__ flags, unsigned long long* __restrict__ new_per_species) {
    __shared__ uint32_t s_b[MAX_GROUP_SPECIES + 1];
    __shared__ uint32_t s_cnt[MAX_GROUP_SPECIES];
    if (threadIdx.x <= g.n) s_b[threadIdx.x] = sp_cstart[g.first_species + threadIdx.x];
    if (threadIdx.x < MAX_GROUP_SPECIES) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t cb = s_b[0], ce = s_b[g.n];
    const uint32_t warps_per_grid = gridDim.x * (COUNT_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (COUNT_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tile_lo = cb / WARP_TILE, tile_hi = (ce + WARP_TILE - 1) / WARP_TILE;
    for (uint32_t t = tile_lo + warp_id; t < tile_hi; t += warps_per_grid) {
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        uint4 v;
        uint64_t roll;
        int have;
        warp_tile_seed<true>(codes, t * WARP_TILE, k, k_mask, v, roll, have);
        const uint32_t range = range_mask16(c0, cb, ce);
        if (!range) continue;
        const uint32_t first_bits = __ldcs(wf + c0 / CHUNK);
        const uint32_t sl0 = slot_of(s_b, g.n, max(c0, cb));
        uint32_t new_bits;
        if (range == 0xFFFFu && c0 + CHUNK <= s_b[sl0 + 1])
            new_bits = bloom_settle_chunk<true>(v, roll, have, k, k_mask, c0, range, first_bits, s_b, sl0, g, bs, s_cnt);
        else
            new_bits = bloom_settle_chunk<false>(v, roll, have, k, k_mask, c0, range, first_bits, s_b, sl0, g, bs, s_cnt);
        if (new_bits) {
            uint16_t* f = flags + (c0 / CHUNK);
            *f = uint16_t(*f | new_bits);   // merge: a straddling chunk is shared by two launches
        }
    }
    __syncthreads();
    if (threadIdx.x < g.n && s_cnt[threadIdx.x])
        atomicAdd(new_per_species + g.first_species + threadIdx.x,
                  static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// Pass C: late probes only.
__global__ void __launch_bounds__(COUNT_THREADS)
bloom_late_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, GroupBounds g,
                  int k, BloomSlots bs, const uint32_t* __restrict__ late_list,
                  const uint32_t* __restrict__ late_count, uint32_t* __restrict__ flags32,
                  unsigned long long* __restrict__ new_per_species) {
    __shared__ uint32_t s_b[MAX_GROUP_SPECIES + 1];
    __shared__ uint32_t s_cnt[MAX_GROUP_SPECIES];
    if (threadIdx.x <= g.n) s_b[threadIdx.x] = sp_cstart[g.first_species + threadIdx.x];
    if (threadIdx.x < MAX_GROUP_SPECIES) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = *late_count;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint32_t v = late_list[e];
        const uint32_t c = v & 0x7FFFFFFFu, j = v >> 31;
        uint32_t sl = 0;
        while (c >= s_b[sl + 1]) sl++;
        uint64_t kmer = 0;
        for (int i = k - 1; i >= 0; i--) kmer = (kmer << 3) | uint64_t(codes[c - i]);
        uint32_t idx[2];
        bloom_indices(kmer, idx[0], idx[1]);
        const uint32_t bit = idx[j], w = bit >> 5, m = 1u << (bit & 31u);
        const unsigned long long tag = bloom_tag(g.epoch0 + sl, c - s_b[sl]);
        if (bs.first[(size_t(sl) << BLOOM_BITS_LOG2) + bit] == tag &&
            !(bs.fw[(size_t(sl) << BLOOM_WORDS_LOG2) + w] & m)) {
            const uint32_t fm = 1u << (c & 31u);
            const uint32_t old = atomicOr(flags32 + (c >> 5), fm);
            if (!(old & fm)) atomicAdd(&s_cnt[sl], 1u);   // not already new through its other probe
        }
    }
    __syncthreads();
    if (threadIdx.x < g.n && s_cnt[threadIdx.x])
        atomicAdd(new_per_species + g.first_species + threadIdx.x,
                  static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// Clears the bitsets of a group's slots and the late counter in ONE launch.
__global__ void __launch_bounds__(256)
bloom_clear_kernel(uint4* __restrict__ bits, uint4* __restrict__ multi, uint4* __restrict__ fw, uint32_t n_vec,
                   uint32_t* __restrict__ late_count) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += gridDim.x * blockDim.x) {
        bits[i] = z;
        multi[i] = z;
        fw[i] = z;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *late_count = 0;
}

}  // namespace kmn

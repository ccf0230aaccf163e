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
// ---- CMS (proba_filter.rs:146-168) -------------------------------------------------------------
// All updates are +1 and saturation is monotone, so accumulating in u32 and clamping to 65535 at
// snapshot time is bit-identical to saturating u16.  The four cell indices of every new k-mer are
// hashed ONCE (in settle / late) into four index streams; the 1 GiB accumulator table is then swept
// in L2-resident slices: one launch per (row, slice) streams 4 B per position and applies only the
// increments that land inside the slice, so the scattered red.add hits L2 instead of paying a 64 B
// DRAM read + 32 B write-back per 4-byte update.
#pragma once
#include "kmn_device.cuh"

namespace kmn {

constexpr int COUNT_THREADS = 256;
constexpr int MAX_GROUP_SPECIES = 16;  // species per launch == number of Bloom slots
constexpr uint32_t IDX_NONE = 0xFFFFFFFFu;   // index-stream sentinel: no increment for this position

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

struct IndexStreams { uint32_t* row[CMS_DEPTH]; };   // [n_codes] cell index per position, IDX_NONE = none

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

// slot (species index inside the group) of each of the 16 positions, 4 bits each
__device__ __forceinline__ uint64_t slots16(const uint32_t* s_b, uint32_t c0, uint32_t valid) {
    uint64_t packed = 0;
    uint32_t sl = 0;
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        if ((valid >> j) & 1u) {
            while (c0 + j >= s_b[sl + 1]) sl++;
            packed |= uint64_t(sl) << (4 * j);
        }
    }
    return packed;
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
        uint64_t kmer[CHUNK];
        uint32_t valid = load_chunk_kmers<true>(codes, t * WARP_TILE, k, k_mask, kmer);
        valid &= range_mask16(c0, cb, ce);
        const uint64_t slots = slots16(s_b, c0, valid);
        uint32_t first_bits = 0;  // bit 2j / 2j+1: probe 0 / 1 of position c0+j saw its bit clear
        uint32_t late_bits = 0;   // same indexing: the probe found its bit already set
#pragma unroll
        for (int h = 0; h < 2; h++) {   // two half chunks: 16 atomics in flight per thread
            uint32_t idx[8][2], old[8][2];
#pragma unroll
            for (int q = 0; q < 8; q++) bloom_indices(kmer[8 * h + q], idx[q][0], idx[q][1]);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int j = 8 * h + q;
                const size_t wbase = size_t((slots >> (4 * j)) & 15u) << BLOOM_WORDS_LOG2;
#pragma unroll
                for (int pr = 0; pr < 2; pr++) {
                    old[q][pr] = 0;
                    if ((valid >> j) & 1u)
                        old[q][pr] = atomicOr(bs.bits + wbase + (idx[q][pr] >> 5), 1u << (idx[q][pr] & 31u));
                }
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int j = 8 * h + q;
                if (!((valid >> j) & 1u)) continue;
                const uint32_t sl = uint32_t(slots >> (4 * j)) & 15u;
#pragma unroll
                for (int pr = 0; pr < 2; pr++) {
                    const uint32_t m = 1u << (idx[q][pr] & 31u);
                    if (old[q][pr] & m) {  // late toucher of this bit
                        atomicOr(bs.multi + (size_t(sl) << BLOOM_WORDS_LOG2) + (idx[q][pr] >> 5), m);
                        atomicMax(bs.first + (size_t(sl) << BLOOM_BITS_LOG2) + idx[q][pr],
                                  bloom_tag(g.epoch0 + sl, c0 + j - s_b[sl]));
                        late_bits |= 1u << (2 * j + pr);
                    } else {
                        first_bits |= 1u << (2 * j + pr);
                    }
                }
            }
        }
        if (c0 + CHUNK > cb && c0 < ce) {
            // chunks straddling a group boundary are shared by two stream-ordered launches
            const bool whole = c0 >= cb && c0 + CHUNK <= ce;
            wf[c0 / CHUNK] = whole ? first_bits : (wf[c0 / CHUNK] | first_bits);
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

// Pass B: settle every was_first probe; hash the four CMS cell indices of every new k-mer.
__global__ void __launch_bounds__(COUNT_THREADS)
bloom_settle_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, GroupBounds g,
                    int k, uint64_t k_mask, BloomSlots bs, const uint32_t* __restrict__ wf,
                    uint16_t* __restrict__ flags, IndexStreams ix,
                    unsigned long long* __restrict__ new_per_species) {
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
        uint64_t kmer[CHUNK];
        uint32_t valid = load_chunk_kmers<true>(codes, t * WARP_TILE, k, k_mask, kmer);
        const uint32_t in_range = range_mask16(c0, cb, ce);
        valid &= in_range;
        const uint64_t slots = slots16(s_b, c0, valid);
        const uint32_t first_bits = __ldcs(wf + c0 / CHUNK);
        uint32_t new_bits = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t idx[8][2], mw[8][2];
#pragma unroll
            for (int q = 0; q < 8; q++) bloom_indices(kmer[8 * h + q], idx[q][0], idx[q][1]);
#pragma unroll
            for (int q = 0; q < 8; q++) {          // all multi[] loads of the half chunk in flight
                const int j = 8 * h + q;
                const size_t wbase = size_t((slots >> (4 * j)) & 15u) << BLOOM_WORDS_LOG2;
#pragma unroll
                for (int pr = 0; pr < 2; pr++) {
                    mw[q][pr] = 0;
                    if (((valid >> j) & 1u) && ((first_bits >> (2 * j + pr)) & 1u))
                        mw[q][pr] = bs.multi[wbase + (idx[q][pr] >> 5)];
                }
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int j = 8 * h + q;
                if (!((valid >> j) & 1u)) continue;
                const uint32_t sl = uint32_t(slots >> (4 * j)) & 15u;
                bool is_new = false;
#pragma unroll
                for (int pr = 0; pr < 2; pr++) {
                    if (!((first_bits >> (2 * j + pr)) & 1u)) continue;  // late probes: bloom_late_kernel
                    const uint32_t m = 1u << (idx[q][pr] & 31u);
                    if (!(mw[q][pr] & m)) {
                        is_new = true;          // only toucher of this bit
                    } else {
                        const unsigned long long tag = bloom_tag(g.epoch0 + sl, c0 + j - s_b[sl]);
                        if (tag > bs.first[(size_t(sl) << BLOOM_BITS_LOG2) + idx[q][pr]]) {  // before every late toucher
                            atomicOr(bs.fw + (size_t(sl) << BLOOM_WORDS_LOG2) + (idx[q][pr] >> 5), m);
                            is_new = true;
                        }
                    }
                }
                if (is_new) new_bits |= 1u << j;
            }
        }
        // per-species counts (a chunk rarely spans species)
        if (new_bits) {
            uint32_t rest = new_bits;
            while (rest) {
                const int j = __ffs(rest) - 1;
                const uint32_t sl = uint32_t(slots >> (4 * j)) & 15u;
                uint32_t same = 0;
#pragma unroll
                for (int q = 0; q < CHUNK; q++)
                    if (((rest >> q) & 1u) && (uint32_t(slots >> (4 * q)) & 15u) == sl) same |= 1u << q;
                atomicAdd(&s_cnt[sl], uint32_t(__popc(same)));
                rest &= ~same;
            }
            uint16_t* f = flags + (c0 / CHUNK);
            *f = uint16_t(*f | new_bits);   // merge: a straddling chunk is shared by two launches
        }
        // CMS cell indices, hashed once, one stream per row
        if (in_range) {
            const bool whole = in_range == 0xFFFFu;
#pragma unroll
            for (int r = 0; r < int(CMS_DEPTH); r++) {
                const uint64_t seed = r == 0 ? SEED_CMS_0 : r == 1 ? SEED_CMS_1 : r == 2 ? SEED_CMS_2 : SEED_CMS_3;
                uint32_t v[CHUNK];
#pragma unroll
                for (int j = 0; j < CHUNK; j++)
                    v[j] = ((new_bits >> j) & 1u) ? cms_index(kmer[j], seed) : IDX_NONE;
                uint32_t* dst = ix.row[r] + c0;
                if (whole) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst);
                    d4[0] = make_uint4(v[0], v[1], v[2], v[3]);
                    d4[1] = make_uint4(v[4], v[5], v[6], v[7]);
                    d4[2] = make_uint4(v[8], v[9], v[10], v[11]);
                    d4[3] = make_uint4(v[12], v[13], v[14], v[15]);
                } else {
#pragma unroll
                    for (int j = 0; j < CHUNK; j++)
                        if ((in_range >> j) & 1u) dst[j] = v[j];
                }
            }
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
                  const uint32_t* __restrict__ late_count, uint32_t* __restrict__ flags32, IndexStreams ix,
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
            if (!(old & fm)) {   // not already new through its other probe
                atomicAdd(&s_cnt[sl], 1u);
                ix.row[0][c] = cms_index(kmer, SEED_CMS_0);
                ix.row[1][c] = cms_index(kmer, SEED_CMS_1);
                ix.row[2][c] = cms_index(kmer, SEED_CMS_2);
                ix.row[3][c] = cms_index(kmer, SEED_CMS_3);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < g.n && s_cnt[threadIdx.x])
        atomicAdd(new_per_species + g.first_species + threadIdx.x,
                  static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// K3: one launch per (row, slice): stream the row's index stream (evict-first) and apply the
// increments that fall inside the L2-resident slice [slice << shift, (slice+1) << shift).
__global__ void __launch_bounds__(COUNT_THREADS)
cms_sweep_kernel(const uint4* __restrict__ idx_row, uint32_t n_vec, uint32_t slice_shift, uint32_t slice,
                 uint32_t* __restrict__ acc_row) {
    for (uint32_t i = blockIdx.x * COUNT_THREADS + threadIdx.x; i < n_vec; i += gridDim.x * COUNT_THREADS) {
        const uint4 v = __ldcs(idx_row + i);
        if ((v.x >> slice_shift) == slice) atomicAdd(acc_row + v.x, 1u);
        if ((v.y >> slice_shift) == slice) atomicAdd(acc_row + v.y, 1u);
        if ((v.z >> slice_shift) == slice) atomicAdd(acc_row + v.z, 1u);
        if ((v.w >> slice_shift) == slice) atomicAdd(acc_row + v.w, 1u);
    }
}

// AtomicCountMinSketch::snapshot (proba_filter.rs:171-177) on u32 accumulators: clamp to u16.
__global__ void __launch_bounds__(256)
cms_finalize_kernel(const uint4* __restrict__ acc4, uint2* __restrict__ out2, uint64_t n_vec) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 a = acc4[i];
        const uint32_t x = min(a.x, 65535u), y = min(a.y, 65535u), z = min(a.z, 65535u), w = min(a.w, 65535u);
        out2[i] = make_uint2(x | (y << 16), z | (w << 16));
    }
}

}  // namespace kmn

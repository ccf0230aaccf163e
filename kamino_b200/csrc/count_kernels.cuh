// count_kernels.cuh — pass 1 on the recoded stream: exact per-species Bloom semantics + CMS build.
//
// Replaces the hot loop of count_species_kmers (/root/reference/src/group_extraction.rs:385-391):
//     if have == k && !bloom.test_and_set(roll) { cmsa.increment(roll) }
//
// Bloom (proba_filter.rs:59-74) is sequential per species: an occurrence is "new" iff at least one
// of its two probe bits was still clear, i.e. iff it is the FIRST occurrence (in file order) that
// touches one of its two bits.  Order-free restatement used here:
//     first[bit] = min file-order index p of the occurrences touching `bit`
//     new(p)   <=> first[i0(p)] == p  ||  first[i1(p)] == p
// (i0 != i1 always because h2 is forced odd.)  Phase A publishes candidates with a 64-bit atomicMax
// of tag = (species_epoch << 32) | ~p into a per-slot first[] array; the epoch makes stale entries of
// earlier species lose, so the array is never cleared.  Phase B re-derives the k-mers and tests.
//
// CMS (proba_filter.rs:146-168): all updates are +1 and saturation is monotone, so accumulating in
// u32 and clamping to 65535 at snapshot time (finalize kernel) is bit-identical to saturating u16.
#pragma once
#include "kmn_device.cuh"

namespace kmn {

constexpr int COUNT_THREADS = 256;
constexpr int MAX_GROUP_SPECIES = 16;  // species per launch == number of first[] slots

struct GroupBounds {                   // codes[] boundaries of the species of one launch
    uint32_t n;                        // species in the group (<= MAX_GROUP_SPECIES)
    uint32_t first_species;            // index of the first species inside the batch
    uint32_t epoch0;                   // epoch of the first species (monotone over the context's life)
};

__device__ __forceinline__ unsigned long long atomic_max_u64(unsigned long long* p, unsigned long long v) {
    return atomicMax(p, v);
}

// Phase A: publish first-touch candidates.
__global__ void __launch_bounds__(COUNT_THREADS)
bloom_mark_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, GroupBounds g,
                  int k, uint64_t k_mask, unsigned long long* __restrict__ first) {
    __shared__ uint32_t s_b[MAX_GROUP_SPECIES + 1];
    if (threadIdx.x <= g.n) s_b[threadIdx.x] = sp_cstart[g.first_species + threadIdx.x];
    __syncthreads();
    const uint32_t cb = s_b[0], ce = s_b[g.n];
    const uint32_t warps_per_grid = gridDim.x * (COUNT_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (COUNT_THREADS / 32) + (threadIdx.x >> 5);
    const uint32_t tile_lo = cb / WARP_TILE, tile_hi = (ce + WARP_TILE - 1) / WARP_TILE;
    for (uint32_t t = tile_lo + warp_id; t < tile_hi; t += warps_per_grid) {
        uint32_t sl = 0;  // local species slot of the current position
        for_each_kmer_in_warp_tile(codes, t * WARP_TILE, k, k_mask, [&](uint32_t c, uint64_t kmer) {
            if (c < cb || c >= ce) return;
            while (c >= s_b[sl + 1]) sl++;
            const uint32_t p = c - s_b[sl];
            const unsigned long long tag =
                (static_cast<unsigned long long>(g.epoch0 + sl) << 32) | (0xFFFFFFFFu - p);
            uint32_t i0, i1;
            bloom_indices(kmer, i0, i1);
            unsigned long long* slot = first + (size_t(sl) << BLOOM_BITS_LOG2);
            atomic_max_u64(slot + i0, tag);
            atomic_max_u64(slot + i1, tag);
        });
    }
}

// Phase B: new(p) <=> first[i0]==tag || first[i1]==tag.  Writes one flag bit per codes[] position
// (flags[c/16] bit c%16) and the per-species number of new k-mers.
__global__ void __launch_bounds__(COUNT_THREADS)
bloom_resolve_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, GroupBounds g,
                     int k, uint64_t k_mask, const unsigned long long* __restrict__ first,
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
        uint32_t sl = 0;
        uint32_t bits = 0;
        uint32_t cnt = 0, cnt_sl = 0;  // new k-mers of species cnt_sl seen by this thread
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        for_each_kmer_in_warp_tile(codes, t * WARP_TILE, k, k_mask, [&](uint32_t c, uint64_t kmer) {
            if (c < cb || c >= ce) return;
            while (c >= s_b[sl + 1]) sl++;
            const uint32_t p = c - s_b[sl];
            const unsigned long long tag =
                (static_cast<unsigned long long>(g.epoch0 + sl) << 32) | (0xFFFFFFFFu - p);
            uint32_t i0, i1;
            bloom_indices(kmer, i0, i1);
            const unsigned long long* slot = first + (size_t(sl) << BLOOM_BITS_LOG2);
            const bool is_new = (__ldg(slot + i0) == tag) | (__ldg(slot + i1) == tag);
            if (is_new) {
                bits |= 1u << (c - c0);
                if (sl != cnt_sl) {
                    if (cnt) atomicAdd(&s_cnt[cnt_sl], cnt);
                    cnt = 0;
                    cnt_sl = sl;
                }
                cnt++;
            }
        });
        if (cnt) atomicAdd(&s_cnt[cnt_sl], cnt);
        if (bits) {
            // chunks straddling a group boundary are touched by two stream-ordered launches:
            // merge, never overwrite (flags are zeroed when the batch is (re)counted)
            uint16_t* f = flags + (c0 / CHUNK);
            *f = uint16_t(*f | bits);
        }
    }
    __syncthreads();
    if (threadIdx.x < g.n && s_cnt[threadIdx.x])
        atomicAdd(new_per_species + g.first_species + threadIdx.x,
                  static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

// ---------------------------------------------------------------------------------------------
// Bloom, L2-resident formulation (default).  Same semantics as above, but the dense traffic goes to
// per-species BITSETS (2^25 bits = 4 MiB, like the reference's filter) that stay in the 126 MB L2:
//   pass A  touch : old = atomicOr(bits[bit]).  The temporally-first toucher of a bit sees it clear
//                   ("was_first").  A LATE toucher marks multi[bit], publishes its file-order index
//                   with atomicMax(first[bit], tag(p)) and appends itself to a compact late list.
//   pass B  settle: a probe whose bit is not in multi[] is the only toucher -> first.
//                   A was_first probe of a contended bit is first iff its p is smaller than the
//                   smallest late p (first[bit]); if so it sets fw[bit] ("first toucher wins").
//   pass C  late  : (late list only) a late probe is first iff it holds first[bit] and !fw[bit].
//   new(p) = first(probe 0) || first(probe 1), OR-ed into the flag bitmap.
// Only contended bits (a few % of the probes on bacterial proteomes) touch the 64-bit first[] array.
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
        uint32_t sl = 0;
        uint32_t first_bits = 0;  // bit 2j / 2j+1: probe 0 / 1 of position c0+j saw its bit clear
        uint32_t late_bits = 0;   // same indexing: the probe found its bit already set
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        for_each_kmer_in_warp_tile<true>(codes, t * WARP_TILE, k, k_mask, [&](uint32_t c, uint64_t kmer) {
            if (c < cb || c >= ce) return;
            while (c >= s_b[sl + 1]) sl++;
            uint32_t idx[2];
            bloom_indices(kmer, idx[0], idx[1]);
            const size_t wbase = size_t(sl) << BLOOM_WORDS_LOG2;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t w = idx[j] >> 5, m = 1u << (idx[j] & 31u);
                const uint32_t old = atomicOr(bs.bits + wbase + w, m);
                if (old & m) {  // late toucher of this bit
                    atomicOr(bs.multi + wbase + w, m);
                    atomicMax(bs.first + (size_t(sl) << BLOOM_BITS_LOG2) + idx[j],
                              bloom_tag(g.epoch0 + sl, c - s_b[sl]));
                    late_bits |= 1u << (2 * (c - c0) + j);
                } else {
                    first_bits |= 1u << (2 * (c - c0) + j);
                }
            }
        });
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
        uint32_t sl = 0, bits = 0, cnt = 0, cnt_sl = 0;
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        const uint32_t first_bits = __ldcs(wf + c0 / CHUNK);
        for_each_kmer_in_warp_tile<true>(codes, t * WARP_TILE, k, k_mask, [&](uint32_t c, uint64_t kmer) {
            if (c < cb || c >= ce) return;
            while (c >= s_b[sl + 1]) sl++;
            uint32_t idx[2];
            bloom_indices(kmer, idx[0], idx[1]);
            const size_t wbase = size_t(sl) << BLOOM_WORDS_LOG2;
            bool is_new = false;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t w = idx[j] >> 5, m = 1u << (idx[j] & 31u);
                const bool was_first = (first_bits >> (2 * (c - c0) + j)) & 1u;
                if (!was_first) continue;  // late probes are settled by bloom_late_kernel
                if (!(bs.multi[wbase + w] & m)) {
                    is_new = true;          // only toucher of this bit
                } else {
                    const unsigned long long tag = bloom_tag(g.epoch0 + sl, c - s_b[sl]);
                    if (tag > bs.first[(size_t(sl) << BLOOM_BITS_LOG2) + idx[j]]) {  // earlier than every late toucher
                        atomicOr(bs.fw + wbase + w, m);
                        is_new = true;
                    }
                }
            }
            if (is_new) {
                bits |= 1u << (c - c0);
                if (sl != cnt_sl) {
                    if (cnt) atomicAdd(&s_cnt[cnt_sl], cnt);
                    cnt = 0;
                    cnt_sl = sl;
                }
                cnt++;
            }
        });
        if (cnt) atomicAdd(&s_cnt[cnt_sl], cnt);
        if (bits) {
            uint16_t* f = flags + (c0 / CHUNK);
            *f = uint16_t(*f | bits);
        }
    }
    __syncthreads();
    if (threadIdx.x < g.n && s_cnt[threadIdx.x])
        atomicAdd(new_per_species + g.first_species + threadIdx.x,
                  static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

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

// K3: CMS increments for flagged positions (4 rows x mix64 + red.add.u32).
__global__ void __launch_bounds__(COUNT_THREADS)
cms_update_kernel(const uint8_t* __restrict__ codes, const uint16_t* __restrict__ flags, uint32_t c_begin,
                  uint32_t c_end, int k, uint64_t k_mask, uint32_t* __restrict__ acc) {
    const uint32_t warps_per_grid = gridDim.x * (COUNT_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (COUNT_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tile_lo = c_begin / WARP_TILE, tile_hi = (c_end + WARP_TILE - 1) / WARP_TILE;
    for (uint32_t t = tile_lo + warp_id; t < tile_hi; t += warps_per_grid) {
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        const uint32_t bits = flags[c0 / CHUNK];
        // all lanes must enter (warp shuffles inside), but work only where a flag is set
        for_each_kmer_in_warp_tile(codes, t * WARP_TILE, k, k_mask, [&](uint32_t c, uint64_t kmer) {
            if (!((bits >> (c - c0)) & 1u)) return;
            atomicAdd(acc + (0u << CMS_WIDTH_LOG2) + cms_index(kmer, SEED_CMS_0), 1u);
            atomicAdd(acc + (1u << CMS_WIDTH_LOG2) + cms_index(kmer, SEED_CMS_1), 1u);
            atomicAdd(acc + (2u << CMS_WIDTH_LOG2) + cms_index(kmer, SEED_CMS_2), 1u);
            atomicAdd(acc + (3u << CMS_WIDTH_LOG2) + cms_index(kmer, SEED_CMS_3), 1u);
        });
    }
}

// K3 (L2-sliced): one launch per (row, slice).  The 1 GiB accumulator table is swept in slices that
// fit the 126 MB L2; every launch re-streams the 1 B/residue codes + 1 bit/residue flags (evict-first)
// and applies only the increments of ONE row that land inside the slice, so the scattered red.add hits
// L2 instead of paying a 64 B DRAM read + 32 B write-back per 4-byte update.
__global__ void __launch_bounds__(COUNT_THREADS)
cms_sweep_kernel(const uint8_t* __restrict__ codes, const uint16_t* __restrict__ flags, uint32_t c_begin,
                 uint32_t c_end, int k, uint64_t k_mask, uint64_t row_seed, uint32_t slice_shift,
                 uint32_t slice, uint32_t* __restrict__ acc_row) {
    const uint32_t warps_per_grid = gridDim.x * (COUNT_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (COUNT_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tile_lo = c_begin / WARP_TILE, tile_hi = (c_end + WARP_TILE - 1) / WARP_TILE;
    for (uint32_t t = tile_lo + warp_id; t < tile_hi; t += warps_per_grid) {
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        const uint32_t bits = __ldcs(flags + c0 / CHUNK);
        for_each_kmer_in_warp_tile<true>(codes, t * WARP_TILE, k, k_mask, [&](uint32_t c, uint64_t kmer) {
            if (!((bits >> (c - c0)) & 1u)) return;
            const uint32_t idx = cms_index(kmer, row_seed);
            if ((idx >> slice_shift) == slice) atomicAdd(acc_row + idx, 1u);
        });
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

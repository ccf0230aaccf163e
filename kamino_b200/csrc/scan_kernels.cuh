// scan_kernels.cuh — pass 2: per-position occupancy + segment boundary rule.
//
// Replaces CountMinSketch::estimate (/root/reference/src/proba_filter.rs:103-112) evaluated at every
// k-mer position, the segmentation of extract_species_blocks (group_extraction.rs:455-498: split at
// residues that recode to 255, keep segments >= 2k+1) and the boundary rule of
// extract_from_valid_segment (group_extraction.rs:292-299).
#pragma once
#include "kmn_device.cuh"
#include "../../include/kamino_b200.h"

namespace kmn {

constexpr int SCAN_THREADS = 256;

// K5a: occ[c] = min over the 4 CMS rows for every valid k-mer END position c; 0 elsewhere.
// 4.8e8 scattered 2-byte gathers per 100 proteomes into a 512 MiB table run at the HBM random-sector rate
// (~60 Gop/s, 8 ms); against an L2-resident region the same loads run at 282 Gop/s
// (tools/microbench/l2_atomics.cu).  So the table is swept in 8 L2-sized slices (half a row = 64 MiB):
// sweep (row, half) streams the codes, hashes the row's index of every k-mer and folds the cells that
// fall into its half into the running minimum kept in occ[] (streamed with evict-first hints so that it
// does not displace the slice).
__global__ void __launch_bounds__(SCAN_THREADS)
occupancy_sweep_kernel(const uint8_t* __restrict__ codes, uint32_t n_tiles, int k, uint64_t k_mask,
                       const uint16_t* __restrict__ cms_row, uint64_t seed, uint32_t half, bool first,
                       uint16_t* __restrict__ occ) {
    const uint32_t warps_per_grid = gridDim.x * (SCAN_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (SCAN_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t fseed = fold30_keep(seed);
    for (uint32_t t = warp_id; t < n_tiles; t += warps_per_grid) {
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        uint64_t kmer[CHUNK];
        const uint32_t valid = load_chunk_kmers<true>(codes, t * WARP_TILE, k, k_mask, kmer);
        uint4* dst = reinterpret_cast<uint4*>(occ + c0);
        uint32_t o[CHUNK / 2];
        if (first) {
#pragma unroll
            for (int i = 0; i < CHUNK / 2; i++)
                o[i] = (((valid >> (2 * i)) & 1u) ? 0xFFFFu : 0u) | (((valid >> (2 * i + 1)) & 1u) ? 0xFFFF0000u : 0u);
        } else {
            const uint4 a = __ldcs(dst), b = __ldcs(dst + 1);
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
            o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
        }
        uint32_t v[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {   // all gathers of the chunk in flight together
            v[j] = 0xFFFFu;
            if ((valid >> j) & 1u) {
                const uint32_t idx = cms_index_folded(fold30_keep(kmer[j]), fseed);
                if ((idx >> (CMS_WIDTH_LOG2 - 1)) == half) v[j] = __ldg(cms_row + idx);
            }
        }
#pragma unroll
        for (int i = 0; i < CHUNK / 2; i++) {
            const uint32_t lo = min(o[i] & 0xFFFFu, v[2 * i]), hi = min(o[i] >> 16, v[2 * i + 1]);
            o[i] = lo | (hi << 16);
        }
        __stcs(dst, make_uint4(o[0], o[1], o[2], o[3]));
        __stcs(dst + 1, make_uint4(o[4], o[5], o[6], o[7]));
    }
}

// ---- narrow table: pass 2 in 4 sweeps instead of 8 ------------------------------------------------------------
// Pass 2 only ever compares occupancies with min_needed and with each other, and only when at least one side is
// shared (>= min_needed; group_extraction.rs:292-299: `previous.shared && previous.occupancy > current.occupancy`,
// `current.shared && previous.occupancy < current.occupancy`; the start ranking of :151-175 sees shared hits only).
// Any map that keeps the order of the values >= min_needed and sends everything below to something smaller
// gives the same hits.  When the largest cell of the finalized table is <= min_needed + 253 the map
// cell -> max(cell - (min_needed - 1), 0) fits a byte: the table shrinks to 256 MiB, ONE ROW (64 MiB) is an
// L2-sized slice, and the occupancy takes 4 sweeps (a position that falls to 0 — below min_needed — in one row skips
// the hash and the gather of the remaining rows: the sweeps run at the L2 gather rate, ~300 Gop/s).  The last sweep maps back (x ? x + min_needed - 1 : 0), so
// occ[] holds the exact estimate wherever it is >= min_needed and 0 elsewhere; the kernels downstream do not change.
// (kmn_fetch_occupancy, which also reports the estimates below min_needed, reruns the 8 wide sweeps.)
__global__ void __launch_bounds__(256)
table_max_kernel(const uint4* __restrict__ tab, size_t n_vec, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n_vec; i += size_t(gridDim.x) * blockDim.x) {
        const uint4 v = __ldcs(tab + i);
        m = __vmaxu2(__vmaxu2(m, v.x), __vmaxu2(__vmaxu2(v.y, v.z), v.w));
    }
    m = max(m & 0xFFFFu, m >> 16);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31u) == 0 && m) atomicMax(out, m);
}

// tab8[i] = min(max(tab[i] - base, 0), 255); 8 cells per thread
__global__ void __launch_bounds__(256)
narrow_table_kernel(const uint4* __restrict__ tab, uint2* __restrict__ tab8, size_t n_vec, uint32_t base) {
    const uint32_t b2 = base | (base << 16);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n_vec; i += size_t(gridDim.x) * blockDim.x) {
        const uint4 v = __ldcs(tab + i);
        const uint32_t a = __vminu2(__vsubus2(v.x, b2), 0x00FF00FFu), b = __vminu2(__vsubus2(v.y, b2), 0x00FF00FFu);
        const uint32_t c = __vminu2(__vsubus2(v.z, b2), 0x00FF00FFu), d = __vminu2(__vsubus2(v.w, b2), 0x00FF00FFu);
        tab8[i] = make_uint2(__byte_perm(a, b, 0x6420), __byte_perm(c, d, 0x6420));
    }
}

// One sweep per row over the narrow table (row8: 2^26 bytes).  `last`: map the running minimum back to estimates.
__global__ void __launch_bounds__(SCAN_THREADS)
occupancy_sweep8_kernel(const uint8_t* __restrict__ codes, uint32_t n_tiles, int k, uint64_t k_mask,
                        const uint8_t* __restrict__ row8, uint64_t seed, bool first, bool last, uint32_t base,
                        uint16_t* __restrict__ occ) {
    const uint32_t warps_per_grid = gridDim.x * (SCAN_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (SCAN_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t fseed = fold30_keep(seed);
    for (uint32_t t = warp_id; t < n_tiles; t += warps_per_grid) {
        const uint32_t c0 = t * WARP_TILE + lane * CHUNK;
        uint64_t kmer[CHUNK];
        const uint32_t valid = load_chunk_kmers<true>(codes, t * WARP_TILE, k, k_mask, kmer);
        uint4* dst = reinterpret_cast<uint4*>(occ + c0);
        uint32_t o[CHUNK / 2];
        if (first) {
#pragma unroll
            for (int i = 0; i < CHUNK / 2; i++)
                o[i] = (((valid >> (2 * i)) & 1u) ? 0xFFFFu : 0u) | (((valid >> (2 * i + 1)) & 1u) ? 0xFFFF0000u : 0u);
        } else {
            const uint4 a = __ldcs(dst), b = __ldcs(dst + 1);
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
            o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
        }
        uint32_t v[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {   // all gathers of the chunk in flight together
            v[j] = 0xFFFFu;
            // a running minimum of 0 (below min_needed after an earlier row) cannot change any more: no hash, no gather
            const uint32_t cur = (o[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
            if (((valid >> j) & 1u) && cur != 0u) v[j] = __ldg(row8 + cms_index_folded(fold30_keep(kmer[j]), fseed));
        }
#pragma unroll
        for (int i = 0; i < CHUNK / 2; i++) {
            uint32_t lo = min(o[i] & 0xFFFFu, v[2 * i]), hi = min(o[i] >> 16, v[2 * i + 1]);
            if (last) {
                lo = lo ? lo + base : 0u;
                hi = hi ? hi + base : 0u;
            }
            o[i] = lo | (hi << 16);
        }
        __stcs(dst, make_uint4(o[0], o[1], o[2], o[3]));
        __stcs(dst + 1, make_uint4(o[4], o[5], o[6], o[7]));
    }
}

// ---- list of reset positions (0xFF codes) of a batch: count / scatter around exclusive_scan_kernel
__global__ void __launch_bounds__(256)
invalid_count_kernel(const uint8_t* __restrict__ codes, uint32_t n_codes, uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t s_w[8];
    const uint32_t c0 = (blockIdx.x * 256u + threadIdx.x) * CHUNK;
    uint32_t n = 0;
    if (c0 < n_codes) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(codes + c0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++)
            n += (c0 + j < n_codes) && (((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) == CODE_INVALID);
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s_w[i];
        tile_cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
invalid_scatter_kernel(const uint8_t* __restrict__ codes, uint32_t n_codes, const uint32_t* __restrict__ tile_base,
                       uint32_t* __restrict__ inv_pos) {
    __shared__ uint32_t s_w[8];
    const uint32_t c0 = (blockIdx.x * 256u + threadIdx.x) * CHUNK;
    uint32_t mask = 0;
    if (c0 < n_codes) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(codes + c0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++)
            if ((c0 + j < n_codes) && (((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) == CODE_INVALID)) mask |= 1u << j;
    }
    const uint32_t n = __popc(mask);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= unsigned(d)) incl += y;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int i = 0; i < 8; i++) woff += (i < int(warp)) ? s_w[i] : 0u;
    uint32_t out = tile_base[blockIdx.x] + woff + incl - n;
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        inv_pos[out++] = c0 + j;
    }
}

__device__ __forceinline__ uint32_t upper_index_le(const uint32_t* __restrict__ a, uint32_t n, uint32_t x) {
    // largest i in [0, n) with a[i] <= x   (a non-decreasing, a[0] <= x guaranteed)
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo + 1) / 2;
        if (a[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// The k-mer that ENDS at position c (all k codes are known to be valid).
__device__ __forceinline__ uint64_t kmer_ending_at(const uint8_t* __restrict__ codes, uint32_t c, int k) {
    uint64_t r = 0;
    for (int i = k - 1; i >= 0; i--) r = (r << 3) | uint64_t(codes[c - i]);
    return r;
}

// K5b/K5c: one warp per valid segment (the codes strictly between two consecutive resets).
// EMIT == false: count start candidates / end anchors per segment.
// EMIT == true : write them, in position order, at the scanned offsets, as COMPACT hits (start inside the
//                segment, occupancy).  k-mer, record and segment offset of a hit are only materialised for
//                what is actually consumed: the anchor pairs (pair_segments_kernel) or kmn_fetch_hits
//                (expand_hits_kernel).
template <bool EMIT>
__global__ void __launch_bounds__(SCAN_THREADS)
segment_boundary_kernel(const uint16_t* __restrict__ occ, const uint32_t* __restrict__ inv_pos, uint32_t n_inv, int k,
                        uint32_t min_needed, uint32_t* __restrict__ seg_starts_cnt, uint32_t* __restrict__ seg_ends_cnt,
                        const uint32_t* __restrict__ seg_starts_off, const uint32_t* __restrict__ seg_ends_off,
                        uint2* __restrict__ starts, uint2* __restrict__ ends) {
    const uint32_t warps_per_grid = gridDim.x * (SCAN_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (SCAN_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (uint32_t j = warp_id; j < n_inv; j += warps_per_grid) {
        const uint32_t seg_lo = j == 0 ? 0u : inv_pos[j - 1] + 1;  // first code of the segment
        const uint32_t seg_hi = inv_pos[j];                        // one past its last code
        const uint32_t len = seg_hi - seg_lo;
        uint32_t ns = 0, ne = 0;
        uint32_t os = 0, oe = 0;
        if (len >= uint32_t(2 * k + 1)) {   // group_extraction.rs:424,462,484
            if (EMIT) {
                os = seg_starts_off[j];
                oe = seg_ends_off[j];
            }
            // hit h (h >= 1) ends at c = seg_lo + k - 1 + h; compare with hit h-1 (group_extraction.rs:292-299)
            for (uint32_t base = seg_lo + k; base < seg_hi; base += 32) {
                const uint32_t c = base + lane;
                bool is_s = false, is_e = false;
                uint32_t prev = 0, cur = 0;
                if (c < seg_hi) {
                    prev = occ[c - 1];
                    cur = occ[c];
                    is_s = prev >= min_needed && prev > cur;   // previous hit becomes a start candidate
                    is_e = cur >= min_needed && prev < cur;    // current hit becomes an end anchor
                }
                const unsigned bs = __ballot_sync(0xffffffffu, is_s), be = __ballot_sync(0xffffffffu, is_e);
                if (EMIT) {
                    if (is_s) starts[os + ns + __popc(bs & lt_mask)] = make_uint2((c - 1) - (k - 1) - seg_lo, prev);
                    if (is_e) ends[oe + ne + __popc(be & lt_mask)] = make_uint2(c - (k - 1) - seg_lo, cur);
                }
                ns += __popc(bs);
                ne += __popc(be);
            }
        }
        if (!EMIT && lane == 0) {
            seg_starts_cnt[j] = ns;
            seg_ends_cnt[j] = ne;
        }
    }
}

// record / segment coordinates of segment j (the segment starts at codes index seg_lo)
struct SegCoords { uint32_t rec, seg_start; };
__device__ __forceinline__ SegCoords segment_coords(uint32_t seg_lo, const uint32_t* __restrict__ rec_cstart, uint32_t n_rec,
                                                   const uint32_t* __restrict__ sp_cstart,
                                                   const uint64_t* __restrict__ sp_rec_off, uint32_t n_sp) {
    const uint32_t r = upper_index_le(rec_cstart, n_rec, seg_lo);
    const uint32_t s = upper_index_le(sp_cstart, n_sp, seg_lo);
    return SegCoords{r - uint32_t(sp_rec_off[s]), seg_lo - rec_cstart[r]};
}

// compact hits -> kmn_hit (kmn_fetch_hits only): one warp per segment
__global__ void __launch_bounds__(SCAN_THREADS)
expand_hits_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ inv_pos, uint32_t n_inv, int k,
                   const uint2* __restrict__ compact, const uint32_t* __restrict__ seg_off,
                   const uint32_t* __restrict__ rec_cstart, uint32_t n_rec, const uint32_t* __restrict__ sp_cstart,
                   const uint64_t* __restrict__ sp_rec_off, uint32_t n_sp, kmn_hit* __restrict__ out) {
    const uint32_t warps_per_grid = gridDim.x * (SCAN_THREADS / 32);
    const uint32_t warp_id = blockIdx.x * (SCAN_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t j = warp_id; j < n_inv; j += warps_per_grid) {
        const uint32_t h0 = seg_off[j], h1 = seg_off[j + 1];
        if (h1 == h0) continue;
        const uint32_t seg_lo = j == 0 ? 0u : inv_pos[j - 1] + 1;
        const SegCoords sc = segment_coords(seg_lo, rec_cstart, n_rec, sp_cstart, sp_rec_off, n_sp);
        for (uint32_t i = h0 + lane; i < h1; i += 32) {
            const uint2 c = compact[i];
            kmn_hit h;
            h.kmer = kmer_ending_at(codes, seg_lo + c.x + k - 1, k);
            h.rec = sc.rec;
            h.seg_start = sc.seg_start;
            h.start = c.x;
            h.occ = c.y;
            out[i] = h;
        }
    }
}

// hit-list offsets per species: first segment of species s is the first reset position >= sp_cstart[s]
__global__ void species_hit_offsets_kernel(const uint32_t* __restrict__ inv_pos, uint32_t n_inv,
                                           const uint32_t* __restrict__ sp_cstart, uint32_t n_sp,
                                           const uint32_t* __restrict__ seg_starts_off,
                                           const uint32_t* __restrict__ seg_ends_off,
                                           uint32_t* __restrict__ sp_starts_off, uint32_t* __restrict__ sp_ends_off) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_sp) return;
    const uint32_t x = sp_cstart[s];
    uint32_t lo = 0, hi = n_inv;  // first j with inv_pos[j] >= x
    while (lo < hi) {
        const uint32_t mid = (lo + hi) / 2;
        if (inv_pos[mid] >= x) hi = mid; else lo = mid + 1;
    }
    sp_starts_off[s] = seg_starts_off[lo];
    sp_ends_off[s] = seg_ends_off[lo];
}

// K5d ("next" row 1 of SURVEY 8f): anchor selection + bubble pairing on the device, one thread per segment.
// Restates select_best_starts_per_local_cluster (group_extraction.rs:151-175: a cluster spans
// [first.start, first.start + 10); better = higher occupancy, then smaller start) and extract_bubble_pairs
// (group_extraction.rs:178-242: best end anchor with left.end < right.start, right.start - left.end <=
// length_middle; better = higher occupancy, then smaller start) on the segment's start / end hit lists.
// ONE pass: a segment has at most as many pairs as start candidates, so its pairs are written, in left-anchor order,
// at seg_s_off[j] of a scratch array and counted (seg_p_cnt); pair_compact_kernel then moves them to their scanned
// offsets (a separate count pass cost as much as the emit pass: both walk the hit lists one thread per segment;
// 1.40 -> 1.08 ms per 100 proteomes).
__global__ void __launch_bounds__(256)
pair_segments_kernel(const uint2* __restrict__ starts, const uint2* __restrict__ ends,
                     const uint32_t* __restrict__ seg_s_off, const uint32_t* __restrict__ seg_e_off, uint32_t n_seg,
                     uint32_t k, uint32_t length_middle, uint32_t* __restrict__ seg_p_cnt, kmn_pair* __restrict__ pairs,
                     const uint8_t* __restrict__ codes, const uint32_t* __restrict__ inv_pos,
                     const uint32_t* __restrict__ rec_cstart, uint32_t n_rec, const uint32_t* __restrict__ sp_cstart,
                     const uint64_t* __restrict__ sp_rec_off, uint32_t n_sp) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_seg) return;
    const uint32_t s0 = seg_s_off[j], s1 = seg_s_off[j + 1], e0 = seg_e_off[j], e1 = seg_e_off[j + 1];
    uint32_t n = 0;
    if (s1 > s0 && e1 > e0) {
        const uint32_t out = s0;
        const uint32_t seg_lo = j == 0 ? 0u : inv_pos[j - 1] + 1;
        SegCoords sc{0, 0};
        bool have_sc = false;
        uint32_t first_end = e0;
        auto handle = [&](uint32_t lstart) {
            const uint32_t left_end = lstart + k;
            while (first_end < e1 && ends[first_end].x <= left_end) first_end++;
            uint32_t bo = 0, bs = 0;
            bool have = false;
            for (uint32_t q = first_end; q < e1; q++) {
                const uint2 r = ends[q];
                if (r.x <= left_end) continue;
                if (r.x - left_end > length_middle) break;
                if (!have || r.y > bo || (r.y == bo && r.x < bs)) { have = true; bo = r.y; bs = r.x; }
            }
            if (!have) return;
            if (!have_sc) {   // first pair of the segment
                sc = segment_coords(seg_lo, rec_cstart, n_rec, sp_cstart, sp_rec_off, n_sp);
                have_sc = true;
            }
            kmn_pair p;
            p.left = kmer_ending_at(codes, seg_lo + lstart + k - 1, int(k));
            p.right = kmer_ending_at(codes, seg_lo + bs + k - 1, int(k));
            p.rec = sc.rec;
            p.seg_start = sc.seg_start;
            p.left_start = lstart;
            p.len = bs + k - lstart;
            pairs[out + n] = p;
            n++;
        };
        const uint2 f = starts[s0];
        uint32_t cluster_at = f.x, best_start = f.x, best_occ = f.y;
        for (uint32_t i = s0 + 1; i < s1; i++) {
            const uint2 c = starts[i];
            if (c.x < cluster_at + 10u) {
                if (c.y > best_occ || (c.y == best_occ && c.x < best_start)) { best_start = c.x; best_occ = c.y; }
            } else {
                handle(best_start);
                cluster_at = c.x; best_start = c.x; best_occ = c.y;
            }
        }
        handle(best_start);
    }
    seg_p_cnt[j] = n;
}

// pairs of segment j: scratch[seg_s_off[j] ..] -> pairs[seg_p_off[j] ..]; one warp per segment, 16 pairs per step
__global__ void __launch_bounds__(256)
pair_compact_kernel(const kmn_pair* __restrict__ scratch, const uint32_t* __restrict__ seg_s_off,
                    const uint32_t* __restrict__ seg_p_off, uint32_t n_seg, kmn_pair* __restrict__ pairs) {
    static_assert(sizeof(kmn_pair) == 32, "two 16-byte vectors per pair");
    const uint32_t warps_per_grid = gridDim.x * 8u;
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t j = blockIdx.x * 8u + (threadIdx.x >> 5); j < n_seg; j += warps_per_grid) {
        const uint32_t o0 = seg_p_off[j], n = seg_p_off[j + 1] - o0;
        if (n == 0) continue;
        const uint4* src = reinterpret_cast<const uint4*>(scratch + seg_s_off[j]);
        uint4* dst = reinterpret_cast<uint4*>(pairs + o0);
        for (uint32_t i = lane; i < 2u * n; i += 32u) dst[i] = src[i];
    }
}

// Parity helper: occupancy per stripped residue of one species (separators removed).
__global__ void occupancy_gather_kernel(const uint16_t* __restrict__ occ, const uint32_t* __restrict__ rec_cstart,
                                        uint32_t r0, uint32_t r1, uint16_t* __restrict__ out) {
    // one CTA per record (grid-stride); out index = c - rec_cstart[r0] - (r - r0)
    for (uint32_t r = r0 + blockIdx.x; r < r1; r += gridDim.x) {
        const uint32_t lo = rec_cstart[r], hi = rec_cstart[r + 1] - 1;
        const uint32_t shift = rec_cstart[r0] + (r - r0);
        for (uint32_t c = lo + threadIdx.x; c < hi; c += blockDim.x) out[c - shift] = occ[c];
    }
}

}  // namespace kmn

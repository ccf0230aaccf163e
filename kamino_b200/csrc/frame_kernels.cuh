// frame_kernels.cuh — K0: raw seq() bytes -> whitespace-free recoded stream with record separators.
//
// Replaces, for a whole batch at once, the per-byte front of count_species_kmers
// (/root/reference/src/group_extraction.rs:375-384): skip ASCII whitespace, upper-case, recode
// (src/recode.rs:31-80) -> class 0..5 or 255.  Memory-bound streaming: 1 B read + 1 B written per
// raw byte, 16-byte vector loads, 256-entry recode LUT in shared memory.
#pragma once
#include "kmn_device.cuh"

namespace kmn {

constexpr int FRAME_THREADS = 256;
constexpr int FRAME_TILE = FRAME_THREADS * CHUNK;  // 4096 raw bytes per CTA
constexpr int FRAME_WARPS = FRAME_THREADS / 32;

struct RecodeLut { uint8_t v[256]; };  // passed by value as a kernel parameter (256 B)

__device__ __forceinline__ void load_lut(uint8_t* s_lut, const RecodeLut& lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut.v[i];
    __syncthreads();
}

// number of non-whitespace bytes among the 16 bytes of v
__device__ __forceinline__ uint32_t count_kept(const uint4& v, const uint8_t* s_lut) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t n = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) n += s_lut[(w[j >> 2] >> (8 * (j & 3))) & 0xFFu] != CODE_SKIP;
    return n;
}

// Pass A: kept-byte counts per CTA tile (4096 B), for the tiles [tile_begin, tile_begin + gridDim.x).
// raw is padded with spaces to a multiple of FRAME_TILE, so no tail guard is needed.
__global__ void __launch_bounds__(FRAME_THREADS)
frame_count_kernel(const uint8_t* __restrict__ raw, RecodeLut lut, uint32_t tile_begin, uint32_t* __restrict__ tile_cnt) {
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_w[FRAME_WARPS];
    load_lut(s_lut, lut);
    const uint32_t tile = tile_begin + blockIdx.x;
    const size_t base = size_t(tile) * FRAME_TILE + size_t(threadIdx.x) * CHUNK;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(raw + base));
    uint32_t n = count_kept(v, s_lut);
    n = __reduce_add_sync(0xffffffffu, n);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s_w[warp] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < FRAME_WARPS; i++) t += s_w[i];
        tile_cnt[tile] = t;
    }
}

// Single-CTA exclusive scan of n u32 values (up to a few 10^5 tiles / segments), 8 values per thread and
// iteration; total -> out[n].  With `carry` the scan continues a running total kept in device memory
// (chunked framing: *carry in, *carry + sum out).  blockIdx.x selects one of up to two independent scans.
struct ScanJob { const uint32_t* in; uint32_t* out; uint32_t n; uint32_t* carry; };
constexpr uint32_t SCAN_ITEMS = 8;

__global__ void __launch_bounds__(1024) exclusive_scan_kernel(ScanJob j0, ScanJob j1) {
    const ScanJob j = blockIdx.x == 0 ? j0 : j1;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = j.carry ? *j.carry : 0u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < j.n; base += 1024 * SCAN_ITEMS) {
        const uint32_t i0 = base + threadIdx.x * SCAN_ITEMS;
        uint32_t x[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < SCAN_ITEMS; q++) {
            x[q] = i0 + q < j.n ? j.in[i0 + q] : 0u;
            sum += x[q];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= unsigned(d)) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= unsigned(d)) wi += y;
            }
            s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        uint32_t run = carry + s_warp[warp] + incl - sum;
#pragma unroll
        for (uint32_t q = 0; q < SCAN_ITEMS; q++) {
            if (i0 + q < j.n) j.out[i0 + q] = run;
            run += x[q];
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        j.out[j.n] = s_carry;
        if (j.carry) *j.carry = s_carry;
    }
}

// Two-level scan for long arrays (segments, hits, observations: 10^5 .. 10^7 values): per-CTA sums of
// SCAN_TILE values, a single-CTA scan of those sums (exclusive_scan_kernel), then per-CTA scans offset by them.
constexpr uint32_t SCAN_TILE = 1024 * SCAN_ITEMS;

__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t s_warp[32];
    const uint32_t i0 = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t q = 0; q < SCAN_ITEMS; q++) sum += i0 + q < n ? in[i0 + q] : 0u;
    sum = __reduce_add_sync(0xffffffffu, sum);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t t = __reduce_add_sync(0xffffffffu, s_warp[threadIdx.x]);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}

// out[i] = tile_off[tile] + exclusive prefix inside the tile; the last CTA also writes out[n] = total
__global__ void __launch_bounds__(1024) scan_tiles_kernel(const uint32_t* __restrict__ in, uint32_t n,
                                                          const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ out) {
    __shared__ uint32_t s_warp[32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t i0 = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t x[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (uint32_t q = 0; q < SCAN_ITEMS; q++) {
        x[q] = i0 + q < n ? in[i0 + q] : 0u;
        sum += x[q];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= unsigned(d)) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= unsigned(d)) wi += y;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = tile_off[blockIdx.x] + s_warp[warp] + incl - sum;
#pragma unroll
    for (uint32_t q = 0; q < SCAN_ITEMS; q++) {
        if (i0 + q < n) out[i0 + q] = run;
        run += x[q];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) out[n] = run;   // padding items are zero
}

// index of the record that contains raw byte offset x: largest r with rec_off[r] <= x, skipping
// empty records (rec_off[r] == rec_off[r+1]); search window [lo, hi].
__device__ __forceinline__ uint32_t record_of(const uint64_t* __restrict__ rec_off, uint32_t lo,
                                              uint32_t hi, uint64_t x) {
    // find the largest r in [lo, hi] with rec_off[r] <= x
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo + 1) / 2;
        if (rec_off[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Pass B: scatter recoded residues and frame the records.  Output index of a kept raw byte i =
//   (#kept bytes before i) + (index of the record containing i)      [one separator per earlier record]
// The thread whose 16-byte chunk contains raw offset rec_off[b] also owns record boundary b (0..n_rec):
//   rec_cstart[b] = (#kept bytes before rec_off[b]) + b, and record b-1's separator (0xFF) sits right before it.
// raw is padded with at least one space, so the chunk that contains offset n_raw exists.
// The codes of a tile (residues + separators) form one contiguous range of codes[]: they are staged in
// shared memory at the same 16-byte alignment and written out as 16-byte vectors (bytes only at the two
// ragged ends); a tile with thousands of empty records falls back to direct byte stores.
// tile_rec[t] = record_of(rec_off, 0, n_rec, t * FRAME_TILE) for the tiles [tile_begin, tile_begin + n]: one
// thread per tile, so the ~19 dependent loads of each search overlap across all tiles instead of heading
// every CTA of frame_scatter_kernel.
__global__ void frame_tile_records_kernel(const uint64_t* __restrict__ rec_off, uint32_t n_rec, uint32_t tile_begin, uint32_t n,
                                          uint32_t* __restrict__ tile_rec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = tile_begin + i;
    tile_rec[t] = record_of(rec_off, 0, n_rec, uint64_t(t) * FRAME_TILE);
}

constexpr uint32_t FRAME_MAX_BOUNDS = 2048;                                  // record boundaries staged per tile
constexpr uint32_t FRAME_STAGE = FRAME_TILE + FRAME_MAX_BOUNDS + 32;

__global__ void __launch_bounds__(FRAME_THREADS)
frame_scatter_kernel(const uint8_t* __restrict__ raw, uint64_t n_raw, RecodeLut lut,
                     const uint32_t* __restrict__ tile_base, const uint64_t* __restrict__ rec_off,
                     uint32_t n_rec, uint8_t* __restrict__ codes, uint32_t* __restrict__ rec_cstart,
                     uint32_t tile_begin, const uint32_t* __restrict__ tile_rec) {
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_w[FRAME_WARPS];
    __shared__ uint32_t s_rlo, s_rhi, s_blo, s_first, s_last;
    __shared__ __align__(16) uint8_t s_stage[FRAME_STAGE];
    load_lut(s_lut, lut);
    const uint32_t tile = tile_begin + blockIdx.x;
    const uint64_t tile0 = uint64_t(tile) * FRAME_TILE;
    if (threadIdx.x == 0) {
        const uint32_t rlo = tile_rec[tile];
        uint32_t blo = rlo;   // empty records that end exactly at tile0: their boundaries belong to this tile too
        while (blo > 0 && rec_off[blo - 1] == tile0) blo--;
        s_rlo = rlo;
        s_blo = blo;
        s_first = 0xFFFFFFFFu;
        s_last = 0;
    }
    if (threadIdx.x == 32) s_rhi = tile_rec[tile + 1];   // record of the next tile's first byte: bounds every search below
    const uint64_t i0 = tile0 + uint64_t(threadIdx.x) * CHUNK;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(raw + i0));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t kept = 0;   // bit j: byte j of the chunk is a residue (not whitespace, inside the data)
    uint8_t c[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        c[j] = s_lut[(w[j >> 2] >> (8 * (j & 3))) & 0xFFu];
        if (c[j] != CODE_SKIP && i0 + j < n_raw) kept |= 1u << j;
    }
    const uint32_t n = __popc(kept);
    // exclusive scan of n over the CTA
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
#pragma unroll
    for (int i = 0; i < FRAME_WARPS; i++) woff += (i < int(warp)) ? s_w[i] : 0u;
    const uint32_t tb = tile_base[tile];
    uint32_t out = tb + woff + incl - n;   // #kept bytes before i0
    // staging origin: no code of this tile lies before q0; smem index = position - g0, g0 16-byte aligned
    const bool staged = s_rhi - s_blo < FRAME_MAX_BOUNDS;   // CTA-uniform
    const uint32_t q0 = tb + max(s_blo, 1u) - 1u;
    const uint32_t g0 = q0 & ~15u;
    uint32_t my_first = 0xFFFFFFFFu, my_last = 0;
    auto put = [&](uint32_t pos, uint8_t val) {
        if (staged) {
            s_stage[pos - g0] = val;
            my_first = min(my_first, pos);
            my_last = max(my_last, pos + 1);
        } else {
            codes[pos] = val;
        }
    };
    if (i0 <= n_raw) {
        // r = last boundary at or before i0 (rec_off[0] == 0, so it exists)
        uint32_t r = record_of(rec_off, s_rlo, s_rhi, i0);
        // record boundaries inside [i0, i0 + 16)
        {
            uint32_t b = r;
            if (rec_off[b] < i0) b++;
            else while (b > 0 && rec_off[b - 1] == i0) b--;   // empty records share the offset
            for (; b <= n_rec; b++) {
                const uint64_t x = rec_off[b];
                if (x >= i0 + CHUNK) break;
                const uint32_t cs = out + __popc(kept & ((1u << uint32_t(x - i0)) - 1u)) + b;
                rec_cstart[b] = cs;
                if (b > 0) put(cs - 1, CODE_INVALID);
            }
        }
        if (n) {
            if (r >= n_rec) r = n_rec - 1;
            uint64_t r_end = rec_off[r + 1];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if ((kept >> j) & 1u) {
                    const uint64_t i = i0 + j;
                    while (i >= r_end) { r++; r_end = rec_off[r + 1]; }
                    put(out + r, c[j]);
                    out++;
                }
            }
        }
    }
    if (!staged) return;
    // the tile's codes are one contiguous range [first, last): copy it out, vectors where whole
    my_first = __reduce_min_sync(0xffffffffu, my_first);
    my_last = __reduce_max_sync(0xffffffffu, my_last);
    if (lane == 0 && my_last) {
        atomicMin(&s_first, my_first);
        atomicMax(&s_last, my_last);
    }
    __syncthreads();
    const uint32_t first = s_first, last = s_last;
    if (last <= first) return;
    const uint32_t v_lo = (first - g0 + 15u) & ~15u, v_hi = (last - g0) & ~15u;   // whole vectors in smem coordinates
    if (v_hi > v_lo) {
        for (uint32_t x = v_lo + threadIdx.x * 16u; x < v_hi; x += FRAME_THREADS * 16u)
            *reinterpret_cast<uint4*>(codes + g0 + x) = *reinterpret_cast<const uint4*>(s_stage + x);
        for (uint32_t x = first - g0 + threadIdx.x; x < v_lo; x += FRAME_THREADS) codes[g0 + x] = s_stage[x];
        for (uint32_t x = v_hi + threadIdx.x; x < last - g0; x += FRAME_THREADS) codes[g0 + x] = s_stage[x];
    } else {
        for (uint32_t x = first - g0 + threadIdx.x; x < last - g0; x += FRAME_THREADS) codes[g0 + x] = s_stage[x];
    }
}

// species -> codes[] offsets for the species boundaries [s_begin, s_end]: sp_cstart[s] = rec_cstart[sp_rec_off[s]]
__global__ void species_offsets_kernel(const uint32_t* __restrict__ rec_cstart,
                                       const uint64_t* __restrict__ sp_rec_off, uint32_t s_begin, uint32_t s_end,
                                       uint32_t* __restrict__ sp_cstart) {
    const uint32_t s = s_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= s_end) sp_cstart[s] = rec_cstart[sp_rec_off[s]];
}

// Per-species source tiles of `tile_wt` warp tiles for the species [s0, s1) (count_kernels.cuh,
// cms_kernels.cuh): sp_tile0[s] = first tile of species s (indices local to this range, sp_tile0[s0] = 0),
// tile_sp[t] = species of tile t, *n_tiles = total.  One CTA; no host round trip.
// running (optional): [1] accumulates the tiles of the species ranges of one batch, [0] receives the value it had
// before this range (= global index of the range's first tile).
struct TilesJob { uint32_t tile_wt; uint32_t* sp_tile0; uint32_t* tile_sp; uint32_t* n_tiles; uint32_t* running; };

__global__ void __launch_bounds__(1024)
species_tiles_kernel(const uint32_t* __restrict__ sp_cstart, uint32_t s0, uint32_t s1, TilesJob j0, TilesJob j1) {
    // blockIdx.x selects one of two independent tile tables (Bloom partition / count-min partition): one launch
    const TilesJob job = blockIdx.x == 0 ? j0 : j1;
    const uint32_t tile_wt = job.tile_wt;
    uint32_t* __restrict__ sp_tile0 = job.sp_tile0;
    uint32_t* __restrict__ tile_sp = job.tile_sp;
    uint32_t* __restrict__ n_tiles = job.n_tiles;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = s0; base < s1; base += 1024) {
        const uint32_t s = base + threadIdx.x;
        uint32_t x = 0;
        if (s < s1) {
            const uint32_t cb = sp_cstart[s], ce = sp_cstart[s + 1];
            if (ce > cb) x = ((ce + WARP_TILE - 1) / WARP_TILE - cb / WARP_TILE + tile_wt - 1) / tile_wt;
        }
        uint32_t incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= unsigned(d)) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= unsigned(d)) wi += y;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t first = s_carry + s_warp[warp] + incl - x;
        if (s < s1) {
            sp_tile0[s] = first;
            for (uint32_t t = 0; t < x; t++) tile_sp[first + t] = s;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = first + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sp_tile0[s1] = s_carry;
        *n_tiles = s_carry;
        if (job.running) {
            const uint32_t before = job.running[1];
            job.running[0] = before;
            job.running[1] = before + s_carry;
        }
    }
}

}  // namespace kmn

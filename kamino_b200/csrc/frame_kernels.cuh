// frame_kernels.cuh — K0: raw seq() bytes -> whitespace-free recoded stream with record separators.
//
// Replaces, for a whole batch at once, the per-byte front of count_species_kmers
// (/root/reference/src/group_extraction.rs:375-384): skip ASCII whitespace, upper-case, recode
// (src/recode.rs:31-80) -> class 0..5 or 255.  Memory-bound streaming: 1 B read + 1 B written per
// raw byte in ONE pass (decoupled look-back), 16-byte vector loads, 256-entry recode LUT in shared memory.
#pragma once
#include "kmn_device.cuh"

namespace kmn {

constexpr int FRAME_THREADS = 256;
constexpr int FRAME_TILE = FRAME_THREADS * CHUNK;  // 4096 bytes per CTA of the FASTA kernels (fasta_kernels.cuh)
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

// ---- K0 in one pass ------------------------------------------------------------------------------------
// Output index of a kept raw byte i = (#kept bytes before i) + (index of the record containing i)   [one separator
// per earlier record];  rec_cstart[b] = (#kept bytes before rec_off[b]) + b, and record b-1's separator (0xFF) sits
// right before it.  One CTA per tile of 16 KiB raw bytes (4 chunks of 16 bytes per thread, a warp's four loads are
// 512 contiguous bytes each): raw is read ONCE; the running number of kept bytes travels from tile to tile by a
// decoupled look-back over `state` (tiles are taken in ticket order, so a tile only ever waits for tiles that are
// already running; the chain continues across the launches of one call because finished tiles keep their
// inclusive prefix).  The record boundaries that fall into the tile are staged in shared memory once (offsets
// relative to the tile), every chunk finds its record there, and the tile that contains raw offset rec_off[b]
// owns boundary b.  The codes of a tile (residues + separators) form one contiguous range of codes[]: they are
// staged in shared memory at the same 16-byte alignment (words swizzled so that the byte stores of a warp, 16 bytes
// apart, spread over the banks) and leave as 16-byte vectors (bytes only at the two ragged ends); a tile with
// thousands of empty records reads its boundaries from global memory and stores bytes directly.
constexpr int FR_CHUNKS = 4;
constexpr int FR_TILE = FRAME_THREADS * FR_CHUNKS * CHUNK;                 // 16384 raw bytes per CTA
constexpr int FR_TILE_CHUNKS = FR_TILE / CHUNK;
constexpr uint32_t FR_MAX_BOUNDS = 2048;                                    // record boundaries staged per tile
constexpr uint32_t FR_STAGE = FR_TILE + FR_MAX_BOUNDS + 32;
constexpr unsigned long long FR_AGGREGATE = 1ull << 32, FR_INCLUSIVE = 2ull << 32;

// tile_rec[t] = record_of(rec_off, 0, n_rec, t * FR_TILE); tile_blo[t] = first boundary that may belong to tile t
// (empty records that end exactly at the tile's first byte share its offset).  One thread per tile, so the ~19
// dependent loads of each search overlap across all tiles instead of heading every CTA of frame_kernel.
__global__ void frame_tile_records_kernel(const uint64_t* __restrict__ rec_off, uint32_t n_rec, uint32_t tile_begin, uint32_t n,
                                          uint32_t* __restrict__ tile_rec, uint32_t* __restrict__ tile_blo) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = tile_begin + i;
    const uint64_t tile0 = uint64_t(t) * FR_TILE;
    const uint32_t r = record_of(rec_off, 0, n_rec, tile0);
    uint32_t blo = r;
    while (blo > 0 && rec_off[blo - 1] == tile0) blo--;
    tile_rec[t] = r;
    tile_blo[t] = blo;
}

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// byte x of the staging buffer: words are XOR-swizzled inside their 16-byte vector by the 128-byte row
__device__ __forceinline__ uint32_t fr_swz(uint32_t x) { return x ^ (((x >> 7) & 3u) << 2); }

#ifndef KMN_FRAME_CTAS
#define KMN_FRAME_CTAS 3
#endif
__global__ void __launch_bounds__(FRAME_THREADS, KMN_FRAME_CTAS)
frame_kernel(const uint8_t* __restrict__ raw, uint64_t n_raw, RecodeLut lut, const uint64_t* __restrict__ rec_off,
             uint32_t n_rec, uint8_t* __restrict__ codes, uint32_t* __restrict__ rec_cstart, uint32_t tile_begin,
             const uint32_t* __restrict__ tile_rec, const uint32_t* __restrict__ tile_blo,
             unsigned long long* __restrict__ state, uint32_t* __restrict__ ticket) {
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_w[FRAME_WARPS];
    __shared__ uint32_t s_tile, s_tb, s_first, s_last;
    __shared__ uint16_t s_cpre[FR_TILE_CHUNKS], s_kept[FR_TILE_CHUNKS];
    __shared__ int32_t s_boff[FR_MAX_BOUNDS];
    __shared__ __align__(16) uint8_t s_stage[FR_STAGE];
    if (threadIdx.x == 0) {
        s_tile = tile_begin + atomicAdd(ticket, 1u);
        s_first = 0xFFFFFFFFu;
        s_last = 0;
    }
    load_lut(s_lut, lut);   // (barrier inside)
    const uint32_t tile = s_tile;
    const uint64_t tile0 = uint64_t(tile) * FR_TILE;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // ---- load + recode: chunk ci = warp * 128 + i * 32 + lane
    uint4 v[FR_CHUNKS];
#pragma unroll
    for (int i = 0; i < FR_CHUNKS; i++)
        v[i] = __ldcs(reinterpret_cast<const uint4*>(raw + tile0) + (warp * (FR_CHUNKS * 32) + i * 32 + lane));
    const uint32_t rhi = tile_rec[tile + 1], blo = tile_blo[tile];
    // boundaries blo .. rhi + 1 (the last one is a sentinel at or beyond the tile's end)
    const uint32_t nb = rhi + 2u - blo;
    const bool staged = nb <= FR_MAX_BOUNDS;   // CTA-uniform
    auto boff_global = [&](uint32_t j) -> int32_t {
        const uint32_t b = blo + j;
        if (b > n_rec) return FR_TILE + 1;
        const uint64_t x = rec_off[b];
        return x < tile0 ? -1 : (x - tile0 > uint64_t(FR_TILE) ? FR_TILE + 1 : int32_t(x - tile0));
    };
    if (staged)
        for (uint32_t j = threadIdx.x; j < nb; j += FRAME_THREADS) s_boff[j] = boff_global(j);
    auto boff = [&](uint32_t j) -> int32_t { return staged ? s_boff[j] : boff_global(j); };
    uint32_t cw[FR_CHUNKS][4], kept[FR_CHUNKS], n_kept[FR_CHUNKS];
#pragma unroll
    for (int i = 0; i < FR_CHUNKS; i++) {
        const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        const uint64_t i0 = tile0 + uint64_t(warp * (FR_CHUNKS * 32) + i * 32 + lane) * CHUNK;
        uint32_t k = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t c = s_lut[(w[q] >> (8 * j)) & 0xFFu];
                o |= c << (8 * j);
                if (c != CODE_SKIP) k |= 1u << (4 * q + j);
            }
            cw[i][q] = o;
        }
        if (i0 + CHUNK > n_raw) k &= i0 < n_raw ? (1u << uint32_t(n_raw - i0)) - 1u : 0u;   // nothing beyond the data
        kept[i] = k;
        n_kept[i] = __popc(k);
    }
    // ---- exclusive prefix of the kept bytes inside the tile (order: warp, i, lane)
    uint32_t p01 = n_kept[0] | (n_kept[1] << 16), p23 = n_kept[2] | (n_kept[3] << 16);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y01 = __shfl_up_sync(0xffffffffu, p01, d), y23 = __shfl_up_sync(0xffffffffu, p23, d);
        if (lane >= unsigned(d)) { p01 += y01; p23 += y23; }
    }
    const uint32_t t01 = __shfl_sync(0xffffffffu, p01, 31), t23 = __shfl_sync(0xffffffffu, p23, 31);
    const uint32_t tot[4] = {t01 & 0xFFFFu, t01 >> 16, t23 & 0xFFFFu, t23 >> 16};
    uint32_t excl[FR_CHUNKS];
    excl[0] = (p01 & 0xFFFFu) - n_kept[0];
    excl[1] = tot[0] + (p01 >> 16) - n_kept[1];
    excl[2] = tot[0] + tot[1] + (p23 & 0xFFFFu) - n_kept[2];
    excl[3] = tot[0] + tot[1] + tot[2] + (p23 >> 16) - n_kept[3];
    if (lane == 0) s_w[warp] = tot[0] + tot[1] + tot[2] + tot[3];
    __syncthreads();
    uint32_t woff = 0, agg = 0;
#pragma unroll
    for (int i = 0; i < FRAME_WARPS; i++) {
        woff += (i < int(warp)) ? s_w[i] : 0u;
        agg += s_w[i];
    }
#pragma unroll
    for (int i = 0; i < FR_CHUNKS; i++) {
        excl[i] += woff;
        const uint32_t ci = warp * (FR_CHUNKS * 32) + i * 32 + lane;
        s_cpre[ci] = uint16_t(excl[i]);
        s_kept[ci] = uint16_t(kept[i]);
    }
    // ---- decoupled look-back (warp 0): kept bytes before this tile
    if (warp == 0) {
        if (lane == 0) st_state(state + tile, (tile == 0 ? FR_INCLUSIVE : FR_AGGREGATE) | agg);
        uint32_t before = 0;
        if (tile > 0) {
            int64_t look = int64_t(tile) - 1;
            for (;;) {
                const int64_t idx = look - int64_t(lane);
                unsigned long long sv = idx >= 0 ? ld_state(state + idx) : FR_INCLUSIVE;   // before tile 0: nothing
                const uint32_t invalid = __ballot_sync(0xffffffffu, (sv >> 32) == 0ull);
                const uint32_t incl = __ballot_sync(0xffffffffu, (sv >> 32) == 2ull);
                const int f = incl ? __ffs(incl) - 1 : 32;             // nearest predecessor with an inclusive prefix
                const uint32_t need = f >= 31 ? 0xffffffffu : (2u << f) - 1u;   // lanes 0 .. f
                if (invalid & need) continue;                           // a predecessor has not published yet: look again
                before += __reduce_add_sync(0xffffffffu, ((need >> lane) & 1u) ? uint32_t(sv) : 0u);
                if (f < 32) break;
                look -= 32;
            }
            if (lane == 0) st_state(state + tile, FR_INCLUSIVE | (before + agg));
        }
        if (lane == 0) s_tb = before;
    }
    __syncthreads();
    const uint32_t tb = s_tb;
    // staging origin: no code of this tile lies before q0; smem index = position - g0, g0 16-byte aligned
    const uint32_t q0 = tb + max(blo, 1u) - 1u;
    const uint32_t g0 = q0 & ~15u;
    uint32_t my_first = 0xFFFFFFFFu, my_last = 0;
    auto put = [&](uint32_t pos, uint32_t val) {
        if (staged) s_stage[fr_swz(pos - g0)] = uint8_t(val);
        else codes[pos] = uint8_t(val);
    };
    // ---- residues
#pragma unroll
    for (int i = 0; i < FR_CHUNKS; i++) {
        if (!kept[i]) continue;
        const int32_t x0 = int32_t((warp * (FR_CHUNKS * 32) + i * 32 + lane) * CHUNK);
        // j = last boundary at or before x0 (boundary 0 of the stage list is never after the tile's first byte)
        uint32_t lo = 0, hi = nb - 1;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo + 1) / 2;
            if (boff(mid) <= x0) lo = mid; else hi = mid - 1;
        }
        uint32_t j = lo;
        uint32_t out = tb + excl[i] + blo + j;       // position of the chunk's first kept byte
        my_first = min(my_first, out);
        int32_t next = boff(j + 1);
        if (next >= x0 + CHUNK) {   // the whole chunk lies in one record
#pragma unroll
            for (int q = 0; q < 16; q++)
                if ((kept[i] >> q) & 1u) put(out++, (cw[i][q >> 2] >> (8 * (q & 3))) & 0xFFu);
        } else {
#pragma unroll
            for (int q = 0; q < 16; q++)
                if ((kept[i] >> q) & 1u) {
                    while (x0 + q >= next) { j++; out++; next = boff(j + 1); }
                    put(out++, (cw[i][q >> 2] >> (8 * (q & 3))) & 0xFFu);
                }
        }
        my_last = max(my_last, out);
    }
    // ---- record boundaries that fall into this tile: rec_cstart and the previous record's separator
    for (uint32_t j = threadIdx.x; j < nb; j += FRAME_THREADS) {
        const uint32_t b = blo + j;
        if (b > n_rec) break;
        const int32_t x = boff(j);
        if (x < 0 || x >= FR_TILE) continue;
        const uint32_t ci = uint32_t(x) / CHUNK;
        const uint32_t cs = tb + s_cpre[ci] + __popc(uint32_t(s_kept[ci]) & ((1u << (uint32_t(x) % CHUNK)) - 1u)) + b;
        rec_cstart[b] = cs;
        if (b > 0) {
            put(cs - 1, CODE_INVALID);
            my_first = min(my_first, cs - 1);
            my_last = max(my_last, cs);
        }
    }
    if (!staged) return;
    // ---- the tile's codes are one contiguous range [first, last): copy it out, vectors where whole
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
        for (uint32_t x = v_lo + threadIdx.x * 16u; x < v_hi; x += FRAME_THREADS * 16u) {
            uint4 o = *reinterpret_cast<const uint4*>(s_stage + x);
            const uint32_t sw = (x >> 7) & 3u;
            if (sw & 1u) { uint32_t t = o.x; o.x = o.y; o.y = t; t = o.z; o.z = o.w; o.w = t; }
            if (sw & 2u) { uint32_t t = o.x; o.x = o.z; o.z = t; t = o.y; o.y = o.w; o.w = t; }
            *reinterpret_cast<uint4*>(codes + g0 + x) = o;
        }
        for (uint32_t x = first - g0 + threadIdx.x; x < v_lo; x += FRAME_THREADS) codes[g0 + x] = s_stage[fr_swz(x)];
        for (uint32_t x = v_hi + threadIdx.x; x < last - g0; x += FRAME_THREADS) codes[g0 + x] = s_stage[fr_swz(x)];
    } else {
        for (uint32_t x = first - g0 + threadIdx.x; x < last - g0; x += FRAME_THREADS) codes[g0 + x] = s_stage[fr_swz(x)];
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

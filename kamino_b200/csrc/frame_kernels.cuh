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
// right before it.  One CTA per tile of 16 KiB raw bytes, one thread per 64 contiguous bytes: raw is read ONCE; the
// running number of kept bytes travels from tile to tile by a decoupled look-back over `state` (tiles are taken in
// ticket order, so a tile only ever waits for tiles that are already running; the chain continues across the
// launches of one call because finished tiles keep their inclusive prefix).  The record boundaries that fall into
// the tile are staged in shared memory once (offsets relative to the tile); a thread finds its first boundary there
// and then walks its 64 bytes and its boundaries together; the thread whose bytes contain raw offset rec_off[b]
// owns boundary b.  The codes of a tile (residues + separators) form one contiguous range of codes[]: they are
// assembled in a zeroed shared-memory image of that range — a 16-byte chunk without an inner boundary is placed
// run by run of kept bytes, each run as five funnel-shifted words OR-ed in (`red.shared.or`), never byte by byte —
// and leave as 16-byte vectors (bytes only at the two ragged ends).  Words of the image are XOR-swizzled so that
// the threads of a warp, 64 bytes apart, hit different banks.  A tile with thousands of empty records reads its
// boundaries from global memory and stores bytes directly.
constexpr int FR_SEG = 64;                                                  // raw bytes per thread (4 chunks of 16)
constexpr int FR_TILE = FRAME_THREADS * FR_SEG;                             // 16384 raw bytes per CTA
constexpr uint32_t FR_MAX_BOUNDS = 2048;                                    // record boundaries staged per tile
constexpr uint32_t FR_SLACK = 16;                                           // image bytes before the tile's first code
constexpr uint32_t FR_STAGE = FR_TILE + FR_MAX_BOUNDS + 64;                 // + alignment, slack and the 5-word overhang
constexpr unsigned long long FR_AGGREGATE = 1ull << 32, FR_INCLUSIVE = 2ull << 32;
static_assert(FR_STAGE % 64 == 0, "the swizzle permutes words inside 64-byte groups");

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
__device__ __forceinline__ void reds_or(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// word w of the image lives at word w ^ ((w >> 5) & 15): a permutation inside every 16-word group
__device__ __forceinline__ uint32_t fr_swz(uint32_t w) { return w ^ ((w >> 5) & 15u); }
// the low 4 bits of n -> a byte mask
__device__ __forceinline__ uint32_t nib_to_bytes(uint32_t n) { return (((n & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu; }

// OR the chunk bytes selected by the bit mask pm (c[0..3] = the 16 codes) into the image so that chunk byte i lands
// at image byte off0 + i
__device__ __forceinline__ void fr_place(uint32_t a_stage, const uint32_t (&c)[4], uint32_t pm, uint32_t off0) {
    const uint32_t m0 = c[0] & nib_to_bytes(pm), m1 = c[1] & nib_to_bytes(pm >> 4), m2 = c[2] & nib_to_bytes(pm >> 8),
                   m3 = c[3] & nib_to_bytes(pm >> 12);
    const uint32_t sh = (off0 & 3u) * 8u, w0 = off0 >> 2;
    reds_or(a_stage + (fr_swz(w0) << 2), m0 << sh);
    reds_or(a_stage + (fr_swz(w0 + 1) << 2), __funnelshift_l(m0, m1, sh));
    reds_or(a_stage + (fr_swz(w0 + 2) << 2), __funnelshift_l(m1, m2, sh));
    reds_or(a_stage + (fr_swz(w0 + 3) << 2), __funnelshift_l(m2, m3, sh));
    reds_or(a_stage + (fr_swz(w0 + 4) << 2), __funnelshift_l(m3, 0u, sh));
}

#ifndef KMN_FRAME_CTAS
#define KMN_FRAME_CTAS 6   // measured: 3 / 4 / 5 / 6 CTAs per SM -> 0.240 / 0.216 / 0.207 / 0.200 ms per 100 proteomes
#endif
__global__ void __launch_bounds__(FRAME_THREADS, KMN_FRAME_CTAS)
frame_kernel(const uint8_t* __restrict__ raw, uint64_t n_raw, RecodeLut lut, const uint64_t* __restrict__ rec_off,
             uint32_t n_rec, uint8_t* __restrict__ codes, uint32_t* __restrict__ rec_cstart, uint32_t tile_begin,
             const uint32_t* __restrict__ tile_rec, const uint32_t* __restrict__ tile_blo,
             unsigned long long* __restrict__ state, uint32_t* __restrict__ ticket) {
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_w[FRAME_WARPS];
    __shared__ uint32_t s_tile, s_tb, s_first, s_last;
    __shared__ int32_t s_boff[FR_MAX_BOUNDS + 1];
    __shared__ __align__(16) uint32_t s_stage[FR_STAGE / 4];
    if (threadIdx.x == 0) {
        s_tile = tile_begin + atomicAdd(ticket, 1u);
        s_first = 0xFFFFFFFFu;
        s_last = 0;
    }
    for (uint32_t i = threadIdx.x; i < FR_STAGE / 16; i += FRAME_THREADS) reinterpret_cast<uint4*>(s_stage)[i] = make_uint4(0u, 0u, 0u, 0u);
    load_lut(s_lut, lut);   // (barrier inside)
    const uint32_t tile = s_tile;
    const uint64_t tile0 = uint64_t(tile) * FR_TILE;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int32_t x_base = int32_t(threadIdx.x) * FR_SEG;
    // ---- load + recode this thread's 64 bytes
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __ldg(reinterpret_cast<const uint4*>(raw + tile0 + x_base) + i);
    const uint32_t rhi = tile_rec[tile + 1], blo = tile_blo[tile];
    // boundaries blo .. min(rhi + 1, n_rec), then a sentinel
    const uint32_t nb = min(rhi + 1u, n_rec) + 1u - blo;
    const bool staged = nb <= FR_MAX_BOUNDS;   // CTA-uniform
    auto boff_global = [&](uint32_t j) -> int32_t {   // offset of boundary blo + j relative to the tile, clamped
        if (j >= nb) return 0x7FFFFFFF;
        const uint64_t x = rec_off[blo + j];
        return x < tile0 ? -1 : (x - tile0 > uint64_t(FR_TILE) ? FR_TILE + 1 : int32_t(x - tile0));
    };
    if (staged)
        for (uint32_t j = threadIdx.x; j <= nb; j += FRAME_THREADS) s_boff[j] = boff_global(j);
    auto boff = [&](uint32_t j) -> int32_t { return staged ? s_boff[j] : boff_global(j); };
    uint32_t cw[4][4], kept[4], nk = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        uint32_t k = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t o = uint32_t(s_lut[w[q] & 0xFFu]) | (uint32_t(s_lut[(w[q] >> 8) & 0xFFu]) << 8) |
                               (uint32_t(s_lut[(w[q] >> 16) & 0xFFu]) << 16) | (uint32_t(s_lut[w[q] >> 24]) << 24);
            cw[i][q] = o;
            // a byte is whitespace iff it is CODE_SKIP = 0xFE: bit 7 set and bit 0 clear (classes are < 0x80, a reset is 0xFF)
            const uint32_t skip = (o >> 7) & ~o & 0x01010101u;
            k |= (((skip * 0x01020408u) >> 24) ^ 0xFu) << (4 * q);
        }
        const uint64_t i0 = tile0 + uint64_t(x_base + 16 * i);
        if (i0 + CHUNK > n_raw) k &= i0 < n_raw ? (1u << uint32_t(n_raw - i0)) - 1u : 0u;   // nothing beyond the data
        kept[i] = k;
        nk += __popc(k);
    }
    // ---- exclusive prefix of the kept bytes inside the tile
    uint32_t incl = nk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= unsigned(d)) incl += y;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, agg = 0;
#pragma unroll
    for (int i = 0; i < FRAME_WARPS; i++) {
        woff += (i < int(warp)) ? s_w[i] : 0u;
        agg += s_w[i];
    }
    const uint32_t excl = woff + incl - nk;
    // ---- decoupled look-back (warp 0): kept bytes before this tile
    if (warp == 0) {
        if (lane == 0) st_state(state + tile, (tile == 0 ? FR_INCLUSIVE : FR_AGGREGATE) | agg);
        uint32_t before = 0;
        if (tile > 0) {
            int64_t look = int64_t(tile) - 1;
            for (;;) {
                const int64_t idx = look - int64_t(lane);
                const unsigned long long sv = idx >= 0 ? ld_state(state + idx) : FR_INCLUSIVE;   // before tile 0: nothing
                const uint32_t invalid = __ballot_sync(0xffffffffu, (sv >> 32) == 0ull);
                const uint32_t inclusive = __ballot_sync(0xffffffffu, (sv >> 32) == 2ull);
                const int f = inclusive ? __ffs(inclusive) - 1 : 32;   // nearest predecessor with an inclusive prefix
                const uint32_t need = f >= 31 ? 0xffffffffu : (2u << f) - 1u;   // lanes 0 .. f
                if (invalid & need) continue;                            // a predecessor has not published yet: look again
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
    // image origin: no code of this tile lies before q0; image byte = position - g0; g0 is 16-byte aligned and leaves
    // FR_SLACK bytes in front (a run of kept bytes behind dropped ones is placed from up to 15 bytes before its chunk)
    const uint32_t q0 = tb + max(blo, 1u) - 1u;
    const uint32_t g0 = (q0 & ~15u) - FR_SLACK;
    const uint32_t a_stage = uint32_t(__cvta_generic_to_shared(s_stage));
    uint32_t my_first = 0xFFFFFFFFu, my_last = 0;
    auto put = [&](uint32_t pos, uint32_t val) {   // one byte
        if (staged) {
            const uint32_t off = pos - g0;
            reds_or(a_stage + (fr_swz(off >> 2) << 2), val << (8u * (off & 3u)));
        } else {
            codes[pos] = uint8_t(val);
        }
        my_first = min(my_first, pos);
        my_last = max(my_last, pos + 1u);
    };
    // ---- this thread's first boundary: the first one at or after x_base
    uint32_t jn;
    {
        uint32_t lo = 0, hi = nb;   // boff(nb) is the sentinel
        while (lo < hi) {
            const uint32_t mid = (lo + hi) / 2;
            if (boff(mid) < x_base) lo = mid + 1; else hi = mid;
        }
        jn = lo;
    }
    uint32_t kb = tb + excl;   // kept bytes before the next byte
    int32_t next = boff(jn);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int32_t x0 = x_base + 16 * i;
        const uint32_t k = kept[i];
        if (staged) {
            // The boundaries inside the chunk cut it into segments; every segment lies in one record (blo + jn - 1 while
            // boundary jn is the next one) and is placed run by run of kept bytes.  Chunk byte q lands at
            // kb + (#kept chunk bytes before q) + record.  Usually one turn; two in the chunk where a record ends.
            uint32_t lo = 0;
            for (;;) {
                const uint32_t hi = next < x0 + CHUNK ? uint32_t(next - x0) : uint32_t(CHUNK);
                uint32_t rem = k & ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                const uint32_t out = kb + blo + jn - 1u;
                while (rem) {   // one run of kept bytes per turn (one turn, or two around a line break)
                    const uint32_t a = __ffs(rem) - 1u, len = __ffs(~(rem >> a)) - 1u;
                    const uint32_t pm = ((1u << len) - 1u) << a;
                    const uint32_t pos = out + __popc(k & ((1u << a) - 1u));   // of chunk byte a
                    fr_place(a_stage, cw[i], pm, pos - g0 - a);
                    my_first = min(my_first, pos);
                    my_last = max(my_last, pos + len);
                    rem &= ~pm;
                }
                if (hi == uint32_t(CHUNK)) break;
                // boundary blo + jn sits at chunk byte hi: its record starts with the next kept byte
                const uint32_t b = blo + jn, cs = kb + __popc(k & ((1u << hi) - 1u)) + b;
                rec_cstart[b] = cs;
                if (b > 0) put(cs - 1u, CODE_INVALID);
                jn++;
                next = boff(jn);
                lo = hi;
            }
            kb += __popc(k);
        } else {
#pragma unroll 1
            for (int q = 0; q < CHUNK; q++) {
                while (next <= x0 + q) {
                    const uint32_t b = blo + jn, cs = kb + b;
                    rec_cstart[b] = cs;
                    if (b > 0) put(cs - 1u, CODE_INVALID);
                    jn++;
                    next = boff(jn);
                }
                if ((k >> q) & 1u) {
                    const uint32_t wq = q < 4 ? cw[i][0] : q < 8 ? cw[i][1] : q < 12 ? cw[i][2] : cw[i][3];   // (no indexed register array)
                    put(kb + blo + jn - 1u, (wq >> (8 * (q & 3))) & 0xFFu);
                    kb++;
                }
            }
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
    const uint8_t* s_bytes = reinterpret_cast<const uint8_t*>(s_stage);
    auto image_byte = [&](uint32_t x) -> uint8_t { return s_bytes[(fr_swz(x >> 2) << 2) | (x & 3u)]; };
    const uint32_t v_lo = (first - g0 + 15u) & ~15u, v_hi = (last - g0) & ~15u;   // whole vectors in image coordinates
    if (v_hi > v_lo) {
        for (uint32_t x = v_lo + threadIdx.x * 16u; x < v_hi; x += FRAME_THREADS * 16u) {
            // logical words x/4 .. x/4+3 share (w >> 5): they sit in vector (x/16) ^ (sw >> 2), permuted by sw & 3
            const uint32_t sw = (x >> 7) & 15u;
            uint4 o = *reinterpret_cast<const uint4*>(s_bytes + ((((x >> 4) ^ (sw >> 2))) << 4));
            if (sw & 1u) { uint32_t t = o.x; o.x = o.y; o.y = t; t = o.z; o.z = o.w; o.w = t; }
            if (sw & 2u) { uint32_t t = o.x; o.x = o.z; o.z = t; t = o.y; o.y = o.w; o.w = t; }
            *reinterpret_cast<uint4*>(codes + uint32_t(g0 + x)) = o;   // g0 may be "negative" (tile 0): wrap in 32 bits
        }
        for (uint32_t x = first - g0 + threadIdx.x; x < v_lo; x += FRAME_THREADS) codes[uint32_t(g0 + x)] = image_byte(x);
        for (uint32_t x = v_hi + threadIdx.x; x < last - g0; x += FRAME_THREADS) codes[uint32_t(g0 + x)] = image_byte(x);
    } else {
        for (uint32_t x = first - g0 + threadIdx.x; x < last - g0; x += FRAME_THREADS) codes[uint32_t(g0 + x)] = image_byte(x);
    }
}

// codes[] behind the last separator: 0xFF up to the end of the buffer (consumers read whole warp tiles and scan
// the whole capacity).  Everything before *n_codes is written by frame_kernel, so the buffer needs no memset.
__global__ void __launch_bounds__(256) codes_tail_kernel(uint8_t* __restrict__ codes, const uint32_t* __restrict__ n_codes, size_t cap) {
    const size_t n = *n_codes, v0 = min((n + 15) & ~size_t(15), cap);
    const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = size_t(gridDim.x) * blockDim.x;
    if (n + tid < v0) codes[n + tid] = CODE_INVALID;   // at most 15 bytes
    for (size_t i = v0 / 16 + tid; i < cap / 16; i += nth) reinterpret_cast<uint4*>(codes)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
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

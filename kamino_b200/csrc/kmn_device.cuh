// kmn_device.cuh — device-side building blocks shared by the sm_100a kernels.
//
// HBM layout of a staged batch:
//   codes[n_codes]   one byte per whitespace-stripped residue: class 0..5 (recode.rs:40-80) or
//                    0xFF for a residue that resets the roll; ONE extra 0xFF separator after the
//                    last residue of every record (group_extraction.rs:373-374 resets roll/have
//                    per record, so a separator is exactly a reset).  Padded with 0xFF to a
//                    multiple of 512 bytes so a warp always loads 32 aligned 16-byte chunks.
//   rec_cstart[n_rec+1]  index in codes[] of the first residue of each record
//                    (rec_cstart[r+1]-1 is record r's separator).
//   sp_cstart[n_sp+1]    index in codes[] where each species starts.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kmn {

constexpr uint32_t BLOOM_BITS_LOG2 = 25;                 // proba_filter.rs:13
constexpr uint32_t BLOOM_MASK = (1u << BLOOM_BITS_LOG2) - 1;
constexpr uint32_t CMS_WIDTH_LOG2 = 26;                  // proba_filter.rs:17
constexpr uint32_t CMS_MASK = (1u << CMS_WIDTH_LOG2) - 1;
constexpr uint32_t CMS_DEPTH = 4;                        // proba_filter.rs:19
constexpr uint64_t CMS_CELLS = uint64_t(CMS_DEPTH) << CMS_WIDTH_LOG2;

// proba_filter.rs:21-28
constexpr uint64_t SEED_BLOOM_1 = 0x9e3779b97f4a7c15ull;
constexpr uint64_t SEED_BLOOM_2 = 0xbf58476d1ce4e5b9ull;
constexpr uint64_t SEED_CMS_0 = 0x94d049bb133111ebull;
constexpr uint64_t SEED_CMS_1 = 0x3c79ac492ba7b653ull;
constexpr uint64_t SEED_CMS_2 = 0x1c69b3f74ac4ae35ull;
constexpr uint64_t SEED_CMS_3 = 0x6a09e667f3bcc909ull;

constexpr uint8_t CODE_INVALID = 0xFF;  // resets roll/have
constexpr uint8_t CODE_SKIP = 0xFE;     // ASCII whitespace: skipped without a reset (LUT only)

constexpr int CHUNK = 16;               // residues per thread (one 16-byte vector load)
constexpr int WARP_TILE = 32 * CHUNK;   // residues per warp iteration

// SplitMix64 finalizer (proba_filter.rs:30-37)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// The same function with fewer issue slots on the device.  mix64 is always applied to key ^ seed, and its first
// step is linear over XOR:  (k ^ s) ^ ((k ^ s) >> 30) == (k ^ (k >> 30)) ^ (s ^ (s >> 30)).  A kernel that hashes
// one k-mer with several seeds therefore folds the k-mer once (fold30) and pays two LOP3 per seed for the step.
// The 64 x 64 -> 64 bit multiplies are spelled as one wide multiply-add plus two 32-bit ones (ptxas otherwise
// emits a fourth instruction for the cross-term addition).
__host__ __device__ constexpr uint64_t fold30(uint64_t x) { return x ^ (x >> 30); }
// fold30 whose result the compiler must keep (it otherwise re-derives the fold from the raw key next to every use)
__device__ __forceinline__ uint64_t fold30_keep(uint64_t x) {
    uint64_t f = x ^ (x >> 30);
    asm volatile("" : "+l"(f));
    return f;
}
// (lo, hi) * c mod 2^64 in three issue slots: wide lo*clo, then both cross terms chained into the high word
__device__ __forceinline__ void mul64c(uint32_t& lo, uint32_t& hi, uint64_t c) {
    const uint32_t clo = uint32_t(c), chi = uint32_t(c >> 32);
    uint32_t plo, phi;
    asm("{\n\t.reg .u64 p;\n\t"
        "mul.wide.u32 p, %2, %4;\n\t"
        "mov.b64 {%0, %1}, p;\n\t"
        "mad.lo.u32 %1, %3, %4, %1;\n\t"
        "mad.lo.u32 %1, %2, %5, %1;\n\t}"
        : "=&r"(plo), "=&r"(phi) : "r"(lo), "r"(hi), "r"(clo), "r"(chi));
    lo = plo;
    hi = phi;
}
// mix64(key ^ seed) from folded_key = fold30(key) and folded_seed = fold30(seed)
__device__ __forceinline__ uint64_t mix64_folded(uint64_t folded_key, uint64_t folded_seed) {
    const uint64_t x0 = folded_key ^ folded_seed;
    uint32_t lo = uint32_t(x0), hi = uint32_t(x0 >> 32);
    mul64c(lo, hi, 0xbf58476d1ce4e5b9ull);
    lo ^= __funnelshift_r(lo, hi, 27);   // x ^= x >> 27
    hi ^= hi >> 27;
    mul64c(lo, hi, 0x94d049bb133111ebull);
    lo ^= __funnelshift_r(lo, hi, 31);   // x ^ (x >> 31)
    hi ^= hi >> 31;
    return (uint64_t(hi) << 32) | lo;
}
// low 32 bits of the same (all the sketches need: indices are at most 26 bits wide)
__device__ __forceinline__ uint32_t mix64_folded_lo(uint64_t folded_key, uint64_t folded_seed) {
    const uint64_t x0 = folded_key ^ folded_seed;
    uint32_t lo = uint32_t(x0), hi = uint32_t(x0 >> 32);
    mul64c(lo, hi, 0xbf58476d1ce4e5b9ull);
    lo ^= __funnelshift_r(lo, hi, 27);
    hi ^= hi >> 27;
    mul64c(lo, hi, 0x94d049bb133111ebull);
    return lo ^ __funnelshift_r(lo, hi, 31);
}

// Bloom probe indices (proba_filter.rs:60-64): idx_i = (h1 + i*h2) & mask, h2 forced odd.
__device__ __forceinline__ void bloom_indices(uint64_t kmer, uint32_t& i0, uint32_t& i1) {
    const uint64_t fk = fold30_keep(kmer);
    const uint32_t h1 = mix64_folded_lo(fk, fold30(SEED_BLOOM_1));        // only the low 25 bits of h1 and h1 + h2 are used
    const uint32_t h2 = mix64_folded_lo(fk, fold30(SEED_BLOOM_2)) | 1u;
    i0 = h1 & BLOOM_MASK;
    i1 = (h1 + h2) & BLOOM_MASK;
}

// CMS cell index of row r (proba_filter.rs:107,148); flat offset = r * 2^26 + idx.
__device__ __forceinline__ uint32_t cms_index(uint64_t kmer, uint64_t seed) {
    return mix64_folded_lo(fold30(kmer), fold30(seed)) & CMS_MASK;
}
// the same from a k-mer folded once for all rows
__device__ __forceinline__ uint32_t cms_index_folded(uint64_t folded_kmer, uint64_t folded_seed) {
    return mix64_folded_lo(folded_kmer, folded_seed) & CMS_MASK;
}

// Pack one 16-code chunk: symbol j (memory order) lands at bits [3*(15-j), 3*(15-j)+3) of P, so
// P equals the reference's roll after shifting in the 16 symbols.  last_inv = index of the last
// 0xFF in the chunk, -1 if none.
__device__ __forceinline__ void pack_chunk(const uint4& v, uint64_t& P, int& last_inv) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint64_t p = 0;
    int li = -1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t x = w[i];
        // bytes b0..b3 (little endian) -> 12 bits, b0 most significant
        uint32_t t = ((x & 7u) << 9) | (((x >> 8) & 7u) << 6) | (((x >> 16) & 7u) << 3) | ((x >> 24) & 7u);
        p = (p << 12) | t;
        uint32_t inv = x & 0x80808080u;  // only 0xFF has bit 7 set
        if (inv) li = 4 * i + ((31 - __clz(inv)) >> 3);
    }
    P = p;
    last_inv = li;
}

// Warp-cooperative load of one warp tile (32 lanes x 16 codes).  Every lane owns codes [c0, c0+16),
// c0 = tile_base + 16*lane, and receives the roll / have state of the reference's loop
// (group_extraction.rs:373-388) as it stands just BEFORE c0: the (k-1)-symbol history comes from the
// two previous lanes by warp shuffle (lanes 0/1 fetch the two chunks before the tile).
// ALL 32 lanes must call this (shuffles inside).
// STREAM: evict-first loads, so re-streamed codes do not displace an L2-resident table slice.
template <bool STREAM>
__device__ __forceinline__ void warp_tile_seed(const uint8_t* __restrict__ codes, uint32_t tile_base, int k,
                                               uint64_t k_mask, uint4& v, uint64_t& roll, int& have) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t c0 = tile_base + lane * CHUNK;
    v = STREAM ? __ldcs(reinterpret_cast<const uint4*>(codes + c0))
               : __ldg(reinterpret_cast<const uint4*>(codes + c0));
    uint64_t P;
    int li;
    pack_chunk(v, P, li);
    // history chunks for lanes 0 and 1: chunk (tile-2) on lane 0, chunk (tile-1) on lane 1
    uint64_t E = 0;
    int eli = 15;  // "before the buffer" behaves like a reset
    if (lane < 2 && tile_base >= (2 - lane) * CHUNK) {
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(codes + tile_base - (2 - lane) * CHUNK));
        pack_chunk(e, E, eli);
    }
    uint64_t P1 = __shfl_up_sync(0xffffffffu, P, 1), P2 = __shfl_up_sync(0xffffffffu, P, 2);
    int li1 = __shfl_up_sync(0xffffffffu, li, 1), li2 = __shfl_up_sync(0xffffffffu, li, 2);
    const uint64_t E0 = __shfl_sync(0xffffffffu, E, 0), E1 = __shfl_sync(0xffffffffu, E, 1);
    const int el0 = __shfl_sync(0xffffffffu, eli, 0), el1 = __shfl_sync(0xffffffffu, eli, 1);
    if (lane == 0) { P1 = E1; li1 = el1; P2 = E0; li2 = el0; }
    if (lane == 1) { P2 = E1; li2 = el1; }
    roll = (((P2 & 0x7FFFull) << 48) | P1) & k_mask;
    have = li1 >= 0 ? 15 - li1 : (li2 >= 0 ? 31 - li2 : 32);
    if (have > k) have = k;
}

// Calls f(c, kmer) for every position c of the lane's chunk whose k-mer is valid (have == k, no
// reset inside the window).
template <bool STREAM = false, typename F>
__device__ __forceinline__ void for_each_kmer_in_warp_tile(const uint8_t* __restrict__ codes,
                                                           uint32_t tile_base, int k, uint64_t k_mask,
                                                           F&& f) {
    uint4 v;
    uint64_t roll;
    int have;
    warp_tile_seed<STREAM>(codes, tile_base, k, k_mask, v, roll, have);
    const uint32_t c0 = tile_base + (threadIdx.x & 31u) * CHUNK;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        const uint32_t code = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if (code == CODE_INVALID) {
            roll = 0;
            have = 0;
        } else {
            roll = ((roll << 3) | uint64_t(code)) & k_mask;
            if (have < k) have++;
            if (have == k) f(c0 + j, roll);
        }
    }
}

// Pack one 16-code chunk like pack_chunk; inv = bit j set <=> code j is 0xFF (a reset).
__device__ __forceinline__ void pack_chunk_mask(const uint4& v, uint64_t& P, uint32_t& inv) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint64_t p = 0;
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t x = w[i];
        const uint32_t t = ((x & 7u) << 9) | (((x >> 8) & 7u) << 6) | (((x >> 16) & 7u) << 3) | ((x >> 24) & 7u);
        p = (p << 12) | t;
        // bit 7 of the four bytes -> bits 24..27 of the product (the other partial products stay below bit 24)
        m |= ((((x >> 7) & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
    }
    P = p;
    inv = m;
}

// Register-resident k-mers of one warp tile: the lane ends up with the k-mers ENDING at its 16 positions
// (kmer[j], garbage where invalid) and returns the validity mask (bit j: position c0+j carries a valid k-mer,
// i.e. the reference's loop has have == k there, group_extraction.rs:373-388).  Lets the caller issue all of a
// chunk's memory operations before consuming any result.
// No rolling: the packed symbols of the lane's chunk and of the two chunks before it (previous lanes by shuffle;
// lanes 0 / 1 load the two chunks before the tile) form one bit string, and the k-mer ending at symbol j is the
// window of that string at bit offset 3 (15 - j) — two funnel shifts and the mask, 4 instructions per position
// instead of the ~12 of the roll / have recurrence.  A window is valid iff none of its k symbols is a reset:
// the reset bits of the three chunks are OR-ed over windows of k by log2(k) shift steps, once per chunk.
// "Before the buffer" counts as a reset.  k <= 21 (3 k <= 63 bits; 20 symbols of history <= the 32 available).
// ALL 32 lanes must call this (shuffles inside).  STREAM: evict-first loads.
template <bool STREAM = false>
__device__ __forceinline__ uint32_t load_chunk_kmers(const uint8_t* __restrict__ codes, uint32_t tile_base,
                                                     int k, uint64_t k_mask, uint64_t (&kmer)[CHUNK]) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t c0 = tile_base + lane * CHUNK;
    const uint4 v = STREAM ? __ldcs(reinterpret_cast<const uint4*>(codes + c0))
                           : __ldg(reinterpret_cast<const uint4*>(codes + c0));
    uint64_t P;
    uint32_t I;
    pack_chunk_mask(v, P, I);
    // history chunks for lanes 0 and 1: chunk (tile - 2) on lane 0, chunk (tile - 1) on lane 1
    uint64_t E = 0;
    uint32_t EI = 0x8000u;   // a chunk before the buffer: its last symbol is a reset
    if (lane < 2 && tile_base >= (2 - lane) * CHUNK) {
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(codes + tile_base - (2 - lane) * CHUNK));
        pack_chunk_mask(e, E, EI);
    }
    uint64_t P1 = __shfl_up_sync(0xffffffffu, P, 1);
    uint32_t P2 = __shfl_up_sync(0xffffffffu, uint32_t(P), 2);   // 5 symbols of history from the chunk before the previous one
    uint32_t I1 = __shfl_up_sync(0xffffffffu, I, 1), I2 = __shfl_up_sync(0xffffffffu, I, 2);
    const uint64_t E1 = __shfl_sync(0xffffffffu, E, 1);
    const uint32_t E0 = __shfl_sync(0xffffffffu, uint32_t(E), 0);
    const uint32_t EI0 = __shfl_sync(0xffffffffu, EI, 0), EI1 = __shfl_sync(0xffffffffu, EI, 1);
    if (lane == 0) { P1 = E1; I1 = EI1; P2 = E0; I2 = EI0; }
    if (lane == 1) { P2 = uint32_t(E1); I2 = EI1; }
    // the symbol string, low bits first: P (48 bits), P1 (48 bits), low 32 bits of the chunk before
    const uint32_t S[4] = {uint32_t(P), uint32_t(P >> 32) | (uint32_t(P1) << 16), uint32_t(P1 >> 16), P2};
    const uint32_t mlo = uint32_t(k_mask), mhi = uint32_t(k_mask >> 32);
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
        const int o = 3 * (15 - j), w = o >> 5, sh = o & 31;
        const uint32_t lo = sh ? __funnelshift_r(S[w], S[w + 1], sh) : S[w];
        const uint32_t hi = sh ? __funnelshift_r(S[w + 1], S[w + 2], sh) : S[w + 1];
        kmer[j] = (uint64_t(hi & mhi) << 32) | (lo & mlo);
    }
    // resets OR-ed over windows of k symbols (bit t of J: symbol t of the 48, oldest first)
    uint64_t acc = uint64_t(I2) | (uint64_t(I1) << 16) | (uint64_t(I) << 32);
    int w = 1;
    while (2 * w <= k) {
        acc |= acc << w;
        w *= 2;
    }
    if (k > w) acc |= acc << (k - w);
    return ~uint32_t(acc >> 32) & 0xFFFFu;
}

// bit j set <=> position c0+j lies inside [cb, ce)
__device__ __forceinline__ uint32_t range_mask16(uint32_t c0, uint32_t cb, uint32_t ce) {
    const uint32_t lo = c0 >= cb ? 0u : min(cb - c0, 16u);
    const uint32_t hi = c0 + 16u <= ce ? 16u : (ce > c0 ? ce - c0 : 0u);
    if (hi <= lo) return 0u;
    return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}

// Lost map of the Bloom stage (count_kernels.cuh): bit 2j+pr of a chunk's word <=> probe pr of position j
// lost.  Returns bit j set <=> BOTH probes of position j lost, i.e. the occurrence does not pass the filter.
__device__ __forceinline__ uint32_t both_lost_mask16(uint32_t lw) {
    uint32_t x = lw & (lw >> 1) & 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

}  // namespace kmn

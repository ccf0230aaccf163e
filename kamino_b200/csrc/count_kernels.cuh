// count_kernels.cuh — pass 1 on the recoded stream: exact per-species Bloom semantics.
//
// Replaces the filter half of the hot loop of count_species_kmers
// (/root/reference/src/group_extraction.rs:385-391):
//     if have == k && !bloom.test_and_set(roll) { cmsa.increment(roll) }
// and produces the LOST MAP (two bits per codes[] position: probe 0 / probe 1 lost), from which
// cms_kernels.cuh derives "this occurrence passed the filter" and the count-min increments.
//
// ---- exact semantics, order-free (proba_filter.rs:59-74) ---------------------------------------
// test_and_set is sequential per species: an occurrence passes iff at least one of its two probe
// bits was still clear, i.e. iff it is the FIRST occurrence (in file order) touching one of its two
// bits (i0 != i1 always because h2 is forced odd).  Call a probe LOST when an earlier position of
// the species touches the same bit:
//     passes(p) <=> not (probe 0 of p lost and probe 1 of p lost).
// False positives AND their dependence on file order are reproduced exactly.
//
// ---- machine mapping ----------------------------------------------------------------------------
// Measured on B200 (tools/microbench/l2_atomics.cu): scattered atomicOr-with-result on an L2-resident
// bitset tops out at 128 Gop/s chip-wide (the first CUDA path of this repo: 2 x 120 M probes -> a
// 1.9 ms floor per 100 proteomes, ~5 ms achieved), shared-memory atomics run at ~900 Gop/s.  So the
// filter bits never live in global memory:
//
//   bloom_partition_kernel  one CTA per source tile (8 Ki positions of ONE species): hashes every
//                           k-mer once and splits its two probes over the 128 buckets of 2^18 filter
//                           bits; entry = (bit inside the bucket, position inside the tile).  Each
//                           (tile, bucket) run has a FIXED slot in global memory, so there are no
//                           global atomics and the runs of a bucket are ordered by tile = by file
//                           position.  A position whose k-mer equals the previous position's (a
//                           homopolymer run) can never pass and is marked without probing.
//   bloom_bucket_kernel     one CTA per (species, bucket): the bucket's 2^18 bits in shared memory,
//                           runs consumed in file order, 8 tiles per step (one per warp).  A probe whose bit was set
//                           by an earlier step is lost.  Inside a step: every probe sets its bit with a
//                           shared-memory atomicOr; a probe that finds the bit set by a probe of the
//                           same step joins a small conflict list (and marks a 2048-bit mini filter);
//                           the probes that set a marked bit join it too; inside the list the smallest
//                           file position wins.  A lost probe sets its bit (2 per position) in the lost map.
//   bloom_slow_kernel       same step logic on a 4 MiB global bitset, one CTA per species, straight from
//                           codes[]: only for species whose probes overflow a run slot (massively
//                           repeated k-mers) — exact for every input, fast for natural ones.
// An occurrence passes iff it is a valid k-mer and not both of its probes are in the lost map; the
// count-min partition (cms_kernels.cuh) reads the map directly.  The map is written with atomicOr only,
// so redoing a species on the slow path is idempotent.
#pragma once
#include "kmn_device.cuh"

namespace kmn {

constexpr int BP_THREADS = 512;
constexpr uint32_t BP_TILE_WT = BP_THREADS / 32;          // warp tiles per source tile (one per warp)
constexpr uint32_t BP_TILE = BP_TILE_WT * WARP_TILE;       // 8192 positions
constexpr uint32_t BP_POS_BITS = 13;
constexpr uint32_t BB_LOG2 = 18;                           // filter bits per bucket (32 KiB of shared memory)
constexpr uint32_t BB_BUCKETS = 1u << (BLOOM_BITS_LOG2 - BB_LOG2);   // 128
constexpr uint32_t BB_WORDS = 1u << (BB_LOG2 - 5);
constexpr uint32_t BB_MASK = (1u << BB_LOG2) - 1;
constexpr uint32_t RUN_CAP = 208;                          // slots per (tile, bucket): mean 128, sigma 11
constexpr uint32_t RUN_Q = (RUN_CAP + 31) / 32;            // entries per lane when a warp reads one run
#ifndef KMN_BB_THREADS
#define KMN_BB_THREADS 256
#endif
#ifndef KMN_BB_CTAS
#define KMN_BB_CTAS 4
#endif
constexpr int BB_THREADS = KMN_BB_THREADS;
constexpr uint32_t BB_STEP_RUNS = BB_THREADS / 32;         // runs (source tiles) per step: one per warp
constexpr uint32_t BB_LIST_CAP = 2048;                     // conflict list of a step; overflow -> slow path (exact)
constexpr uint32_t ENT_BIT_SHIFT = BP_POS_BITS + 1;         // entry = bit_in_bucket << 14 | pos_in_tile << 1 | probe
constexpr uint32_t MF_WORDS = 256;                         // 8192-bit mini filter of a step's conflicting bits
constexpr uint32_t MT_SLOTS = 256;                         // hashed min table of a step's conflict list
constexpr size_t BB_SMEM = size_t(BB_LIST_CAP) * sizeof(unsigned long long) + size_t(BB_WORDS) * sizeof(uint32_t);
static_assert(BP_TILE == (1u << BP_POS_BITS), "entry layout");
static_assert(BB_LOG2 + ENT_BIT_SHIFT <= 32, "entry layout");
static_assert(BB_BUCKETS % BP_TILE_WT == 0 && BB_BUCKETS <= BP_THREADS, "bucket ownership");
// staging rows are cap + 1 words apart: an odd stride spreads the stores of a warp (random buckets, similar ranks)
// over the shared-memory banks
constexpr size_t bp_smem_for(uint32_t cap) {
    return size_t(BB_BUCKETS) * (cap + 1) * sizeof(uint32_t) + BB_BUCKETS * sizeof(uint32_t);
}

struct BloomRuns {
    uint32_t* ent;              // [n_tiles][BB_BUCKETS][cap] bit_in_bucket << 14 | pos_in_tile << 1 | probe
    uint16_t* cnt;              // per species: [BB_BUCKETS][tiles of the species] (a bucket's counts are contiguous)
    uint32_t cap;               // RUN_CAP (smaller only under the test knob)
    const uint32_t* tile_sp;    // [n_tiles] species (index inside the batch) of every source tile
    const uint32_t* sp_tile0;   // [n_sp + 1] first source tile of every species (species_tiles_kernel)
    const uint32_t* n_tiles;    // device scalar: source tiles of this launch (the grid is an upper bound)
    uint32_t* slow;             // [n_sp] != 0: a run slot / conflict list overflowed, bloom_slow_kernel owns the species
    uint32_t* lost2;            // lost map: bit 2j+pr of word c/16 <=> probe pr of position c+j lost
};

// Positions whose k-mer equals the k-mer of the previous position (both valid): the earlier one set
// both filter bits, so these can never pass (proba_filter.rs:59-74) and are not probed at all.
// All 32 lanes must call this.
__device__ __forceinline__ uint32_t dup_mask16(const uint64_t (&kmer)[CHUNK], uint32_t valid) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t prev_k = __shfl_up_sync(0xffffffffu, kmer[CHUNK - 1], 1);
    const uint32_t prev_v = __shfl_up_sync(0xffffffffu, valid >> (CHUNK - 1), 1) & 1u;
    uint32_t dup = 0;
    if (lane > 0 && prev_v && (valid & 1u) && kmer[0] == prev_k) dup |= 1u;   // lane 0: not checked across warp tiles
#pragma unroll
    for (int j = 1; j < CHUNK; j++)
        if (((valid >> (j - 1)) & 3u) == 3u && kmer[j] == kmer[j - 1]) dup |= 1u << j;
    return dup;
}

// bit j of m -> value v in the 2-bit field j of a chunk's lost-map word
__device__ __forceinline__ uint32_t spread2(uint32_t m, uint32_t v) {
    uint32_t x = m & 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x * v;
}

// first codes[] index of source tile t of a species that starts at cb (tiles sit on the warp-tile grid)
__device__ __forceinline__ uint32_t tile_c0(uint32_t cb, uint32_t t) {
    return (cb / WARP_TILE + t * BP_TILE_WT) * WARP_TILE;
}

__global__ void __launch_bounds__(BP_THREADS, 2)
bloom_partition_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, int k,
                       uint64_t k_mask, BloomRuns R) {
    extern __shared__ __align__(16) uint32_t bp_smem[];
    const uint32_t stride = R.cap + 1;
    uint32_t* stage = bp_smem;                               // [BB_BUCKETS][cap + 1]
    uint32_t* fill = bp_smem + size_t(BB_BUCKETS) * stride;  // [BB_BUCKETS]
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tg = blockIdx.x;
    if (tg >= *R.n_tiles) return;
    const uint32_t s = R.tile_sp[tg];
    const uint32_t cb = sp_cstart[s], ce = sp_cstart[s + 1];
    const uint32_t tg0 = R.sp_tile0[s], nt = R.sp_tile0[s + 1] - tg0;
    const uint32_t base = tile_c0(cb, tg - tg0);
    if (threadIdx.x < BB_BUCKETS) fill[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t c0 = base + warp * WARP_TILE + lane * CHUNK;
    if (base + warp * WARP_TILE < ce) {   // warp-uniform
        uint64_t kmer[CHUNK];
        const uint32_t valid_all = load_chunk_kmers<true>(codes, base + warp * WARP_TILE, k, k_mask, kmer);
        const uint32_t valid = valid_all & range_mask16(c0, cb, ce);
        const uint32_t dup = dup_mask16(kmer, valid_all) & valid;
        if (dup) atomicOr(R.lost2 + c0 / CHUNK, spread2(dup, 3u));
        const uint32_t todo = valid & ~dup;
        bool ovf = false;
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {
            if (!((todo >> j) & 1u)) continue;
            uint32_t idx[2];
            bloom_indices(kmer[j], idx[0], idx[1]);
            const uint32_t pos = warp * WARP_TILE + lane * CHUNK + j;
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                const uint32_t b = idx[pr] >> BB_LOG2;
                const uint32_t rank = atomicAdd(&fill[b], 1u);
                if (rank < R.cap) stage[b * stride + rank] = ((idx[pr] & BB_MASK) << ENT_BIT_SHIFT) | (pos << 1) | uint32_t(pr);
                else ovf = true;
            }
        }
        if (ovf) R.slow[s] = 1u;
    }
    __syncthreads();
    if (threadIdx.x < BB_BUCKETS)
        R.cnt[size_t(tg0) * BB_BUCKETS + size_t(threadIdx.x) * nt + (tg - tg0)] = uint16_t(min(fill[threadIdx.x], R.cap));
    // warp w copies the runs of buckets 8w .. 8w+7 to their fixed slots
#pragma unroll 1
    for (uint32_t q = 0; q < BB_BUCKETS / BP_TILE_WT; q++) {
        const uint32_t b = warp * (BB_BUCKETS / BP_TILE_WT) + q;
        const uint32_t n = min(fill[b], R.cap);
        uint32_t* dst = R.ent + (size_t(tg) * BB_BUCKETS + b) * R.cap;
        for (uint32_t i = lane; i < n; i += 32) dst[i] = stage[b * stride + i];
    }
}

// ---- one step of exact first-touch resolution on a set of probes -----------------------------------
// Shared by bloom_bucket_kernel (bitset in shared memory) and bloom_slow_kernel (bitset in global
// memory, GLOBAL_BITS: loads go past L1 because the bits are set by L2 atomics).  Every thread owns up
// to NE probes of the step: bit[e] (index into the bitset) and p[e] (file order inside the step; unique
// among the probes of one bit), bit e of `live` set.  `lost` receives the mask of the probes that are
// NOT the file-order-first toucher of their bit.  Returns false (CTA-uniform, nothing decided) when the
// conflict list is too small.  All threads of the CTA must call it (barriers inside); `par` alternates
// 0/1 between consecutive steps.
struct StepShared {
    unsigned long long* list;    // [list_cap] (bit << 32) | p
    uint32_t list_cap;
    uint32_t* mf;                // [2][MF_WORDS] mini filter: bits that saw a same-step race
    unsigned long long* mt;      // [2][MT_SLOTS] min key per hash slot of the conflict list
    uint32_t* n_list;            // [2]
};
constexpr size_t STEP_STATIC_SMEM = 2 * MF_WORDS * 4 + 2 * MT_SLOTS * 8 + 8;

__device__ __forceinline__ uint32_t mt_slot(uint32_t bit) { return (bit ^ (bit >> 8) ^ (bit >> 16)) & (MT_SLOTS - 1); }

// NQ_GUARD: probes e >= nq are known dead for the whole WARP (warp-uniform skip of unrolled slots).
template <int NE, bool GLOBAL_BITS>
__device__ __forceinline__ bool bloom_step(uint32_t* bitset, const StepShared& S, uint32_t par, const uint32_t (&bit)[NE],
                                           const uint32_t (&p)[NE], uint32_t live, int nq, uint32_t& lost) {
    // phase 1: bits set by earlier steps (= by smaller file positions)
    uint32_t pre = 0;
#pragma unroll
    for (int e = 0; e < NE; e++) {
        if (e >= nq) break;
        if (!((live >> e) & 1u)) continue;
        const uint32_t w = GLOBAL_BITS ? __ldcg(bitset + (bit[e] >> 5)) : bitset[bit[e] >> 5];
        if (w & (1u << (bit[e] & 31u))) pre |= 1u << e;
    }
    __syncthreads();
    // the other parity's scratch was last read in the previous step
    for (uint32_t i = threadIdx.x; i < MF_WORDS; i += blockDim.x) S.mf[(par ^ 1u) * MF_WORDS + i] = 0;
    for (uint32_t i = threadIdx.x; i < MT_SLOTS; i += blockDim.x) S.mt[(par ^ 1u) * MT_SLOTS + i] = ~0ull;
    if (threadIdx.x == 0) S.n_list[par ^ 1u] = 0;
    // phase 2: every remaining probe sets its bit; the loser of a same-step race joins the conflict list
    uint32_t won = 0, member = 0;
    uint32_t* mf = S.mf + par * MF_WORDS;
    unsigned long long* mt = S.mt + par * MT_SLOTS;
#pragma unroll
    for (int e = 0; e < NE; e++) {
        if (e >= nq) break;
        if (!((live >> e) & 1u) || ((pre >> e) & 1u)) continue;
        const uint32_t m = 1u << (bit[e] & 31u);
        const uint32_t old = atomicOr(&bitset[bit[e] >> 5], m);
        if (old & m) {
            member |= 1u << e;
            const unsigned long long key = (static_cast<unsigned long long>(bit[e]) << 32) | p[e];
            const uint32_t slot = atomicAdd(&S.n_list[par], 1u);
            if (slot < S.list_cap) S.list[slot] = key;
            atomicMin(&mt[mt_slot(bit[e])], key);
            atomicOr(&mf[(bit[e] >> 5) & (MF_WORDS - 1)], m);
        } else {
            won |= 1u << e;
        }
    }
    __syncthreads();
    lost = pre;
    if (S.n_list[par] == 0) return true;   // CTA-uniform: no same-step conflicts
    // phase 3a: the probes that set a marked bit join the list
#pragma unroll
    for (int e = 0; e < NE; e++) {
        if (e >= nq) break;
        if (!((won >> e) & 1u)) continue;
        if (mf[(bit[e] >> 5) & (MF_WORDS - 1)] & (1u << (bit[e] & 31u))) {
            member |= 1u << e;
            const unsigned long long key = (static_cast<unsigned long long>(bit[e]) << 32) | p[e];
            const uint32_t slot = atomicAdd(&S.n_list[par], 1u);
            if (slot < S.list_cap) S.list[slot] = key;
            atomicMin(&mt[mt_slot(bit[e])], key);
        }
    }
    __syncthreads();
    // phase 3b: inside the list the smallest file position of a bit wins.  The hashed min table answers
    // directly when its slot holds this probe's bit; a slot taken by a smaller bit falls back to the list.
    const uint32_t n = S.n_list[par];
    if (n > S.list_cap) return false;
    if (member) {
#pragma unroll
        for (int e = 0; e < NE; e++) {
            if (!((member >> e) & 1u)) continue;
            const unsigned long long lo = static_cast<unsigned long long>(bit[e]) << 32, me = lo | p[e];
            const unsigned long long best = mt[mt_slot(bit[e])];
            bool l;
            if ((best >> 32) == bit[e]) {
                l = best != me;
            } else {
                l = false;
                for (uint32_t i = 0; i < n; i++) {
                    const unsigned long long o = S.list[i];
                    l |= (o >= lo) && (o < me);   // same bit, smaller position
                }
            }
            if (l) lost |= 1u << e;
        }
    }
    return true;
}

// ---- bloom_bucket_kernel: the same step, hand-scheduled for shared memory ---------------------------
// (32-bit shared addresses and explicit ld/atom.shared keep the per-probe instruction count low: this
// kernel is bound by instruction issue, not by memory.)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t atoms_or_u32(uint32_t a, uint32_t v) {
    uint32_t o;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
    return o;
}

// The bucket kernel's conflict table in 32 bits (a 64-bit shared-memory atomicMin is a CAS loop; ncu showed the
// conflict path, run by one or two lanes in almost every step, taking a quarter of the kernel's instructions):
// slot = low 8 bits of the bit index, key = remaining bit-index bits << 18 | order inside the step.  The minimum
// of a slot is the smallest bit that maps to it and, for that bit, its first toucher in file order.
constexpr uint32_t MT32_ORDER_BITS = 18;   // order = warp << 14 | position inside the tile << 1 | probe
static_assert((BB_THREADS / 32) <= (1 << (MT32_ORDER_BITS - ENT_BIT_SHIFT)), "order field");
static_assert(BB_LOG2 - 8 + MT32_ORDER_BITS <= 32, "key layout");
__device__ __forceinline__ uint32_t mt32_slot(uint32_t bit) { return bit & (MT_SLOTS - 1); }
__device__ __forceinline__ uint32_t mt32_key(uint32_t bit, uint32_t order) { return ((bit >> 8) << MT32_ORDER_BITS) | order; }

// rare: a probe of this step found its bit set by another probe of the same step
__device__ __noinline__ void bucket_conflict_join(unsigned long long* list, uint32_t* mt, uint32_t* mf,
                                                  uint32_t* n_list, uint32_t bit, uint32_t p, bool mark) {
    const uint32_t slot = atomicAdd(n_list, 1u);
    if (slot < BB_LIST_CAP) list[slot] = (static_cast<unsigned long long>(bit) << 32) | p;
    atomicMin(&mt[mt32_slot(bit)], mt32_key(bit, p));
    if (mark) atomicOr(&mf[(bit >> 5) & (MF_WORDS - 1)], 1u << (bit & 31u));
}

// rare: is a member of the conflict list beaten by a smaller file position of the same bit?
__device__ __noinline__ bool bucket_conflict_lost(const unsigned long long* list, const uint32_t* mt, uint32_t n,
                                                  uint32_t bit, uint32_t p) {
    const uint32_t best = mt[mt32_slot(bit)];
    if ((best >> MT32_ORDER_BITS) == (bit >> 8)) return best != mt32_key(bit, p);
    const unsigned long long lo = static_cast<unsigned long long>(bit) << 32, me = lo | p;
    bool l = false;   // the slot belongs to a smaller bit: scan the list
    for (uint32_t i = 0; i < n; i++) {
        const unsigned long long o = list[i];
        l |= (o >= lo) && (o < me);
    }
    return l;
}

__global__ void __launch_bounds__(BB_THREADS, KMN_BB_CTAS)
bloom_bucket_kernel(const uint32_t* __restrict__ sp_cstart, uint32_t s0, BloomRuns R) {
    extern __shared__ __align__(16) unsigned long long bb_smem[];
    unsigned long long* s_list = bb_smem;                                        // BB_LIST_CAP
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(bb_smem + BB_LIST_CAP);       // BB_WORDS
    __shared__ uint32_t s_mt[2 * MT_SLOTS];
    __shared__ uint32_t s_mf[2 * MF_WORDS];
    __shared__ uint32_t s_n[2];
    const uint32_t s = s0 + blockIdx.x / BB_BUCKETS, beta = blockIdx.x % BB_BUCKETS;
    if (R.slow[s]) return;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tg0 = R.sp_tile0[s], n_tiles = R.sp_tile0[s + 1] - tg0;
    if (n_tiles == 0) return;
    const uint32_t cb = sp_cstart[s];
    for (uint32_t i = threadIdx.x; i < BB_WORDS; i += BB_THREADS) s_bits[i] = 0;
    for (uint32_t i = threadIdx.x; i < 2 * MF_WORDS; i += BB_THREADS) s_mf[i] = 0;
    for (uint32_t i = threadIdx.x; i < 2 * MT_SLOTS; i += BB_THREADS) s_mt[i] = ~0u;
    if (threadIdx.x < 2) s_n[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t a_bits = smem_addr(s_bits);
    // this bucket's run lengths are contiguous; warp w owns run t0 + w of every step
    const uint16_t* __restrict__ cnt = R.cnt + size_t(tg0) * BB_BUCKETS + size_t(beta) * n_tiles;
    const uint32_t* __restrict__ ent = R.ent + (size_t(tg0) * BB_BUCKETS + beta) * R.cap + lane;
    const size_t run_stride = size_t(BB_BUCKETS) * R.cap;
    auto run_len = [&](uint32_t t) -> uint32_t { return t < n_tiles ? uint32_t(__ldg(cnt + t)) : 0u; };
    auto lane_live = [&](uint32_t n) -> uint32_t {   // bit q set <=> lane + 32 q < n
        return n > lane ? (2u << ((n - lane - 1u) >> 5)) - 1u : 0u;
    };
    // software pipeline: entries one step ahead, run lengths two steps ahead
    uint32_t raw[RUN_Q];
    uint32_t n_next = run_len(warp);
    {
        const uint32_t lv = lane_live(n_next);
        const uint32_t* e = ent + size_t(warp) * run_stride;
#pragma unroll
        for (int q = 0; q < int(RUN_Q); q++)
            if ((lv >> q) & 1u) raw[q] = __ldcs(e + 32 * q);
    }
    uint32_t n_next2 = run_len(BB_STEP_RUNS + warp);
    uint32_t par = 0;
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += BB_STEP_RUNS, par ^= 1u) {
        const uint32_t t = t0 + warp;
        const uint32_t n = n_next;
        const uint32_t live = lane_live(n);
        uint32_t cur[RUN_Q];
#pragma unroll
        for (int q = 0; q < int(RUN_Q); q++) cur[q] = raw[q];
        n_next = n_next2;
        {
            const uint32_t lv = lane_live(n_next);
            const uint32_t* e = ent + size_t(t + BB_STEP_RUNS) * run_stride;
#pragma unroll
            for (int q = 0; q < int(RUN_Q); q++)
                if ((lv >> q) & 1u) raw[q] = __ldcs(e + 32 * q);
        }
        n_next2 = run_len(t + 2 * BB_STEP_RUNS);
        uint32_t* mt = s_mt + par * MT_SLOTS;
        uint32_t* mf = s_mf + par * MF_WORDS;
        const uint32_t a_mf = smem_addr(mf);
        // phase 1: bits set by earlier steps (= by smaller file positions)
        uint32_t pre = 0;
#pragma unroll
        for (int q = 0; q < int(RUN_Q); q++) {
            if (32u * q >= n) break;   // warp-uniform
            if ((live >> q) & 1u) {
                const uint32_t bit = cur[q] >> ENT_BIT_SHIFT;
                pre |= ((lds_u32(a_bits + ((bit >> 5) << 2)) >> (bit & 31u)) & 1u) << q;
            }
        }
        __syncthreads();
        // the other parity's scratch was last read in the previous step
        for (uint32_t i = threadIdx.x; i < MF_WORDS; i += BB_THREADS) s_mf[(par ^ 1u) * MF_WORDS + i] = 0;
        for (uint32_t i = threadIdx.x; i < MT_SLOTS; i += BB_THREADS) s_mt[(par ^ 1u) * MT_SLOTS + i] = ~0u;
        if (threadIdx.x == 0) s_n[par ^ 1u] = 0;
        // phase 2: every remaining probe sets its bit; the loser of a same-step race joins the conflict list
        uint32_t won = 0, member = 0;
        const uint32_t todo = live & ~pre;
#pragma unroll
        for (int q = 0; q < int(RUN_Q); q++) {
            if (32u * q >= n) break;
            if ((todo >> q) & 1u) {
                const uint32_t bit = cur[q] >> ENT_BIT_SHIFT, m = 1u << (bit & 31u);
                if (atoms_or_u32(a_bits + ((bit >> 5) << 2), m) & m) {
                    member |= 1u << q;
                    bucket_conflict_join(s_list, mt, mf, &s_n[par], bit, (warp << ENT_BIT_SHIFT) | (cur[q] & ((1u << ENT_BIT_SHIFT) - 1)), true);
                } else {
                    won |= 1u << q;
                }
            }
        }
        __syncthreads();
        uint32_t lost = pre;
        if (s_n[par] != 0) {   // CTA-uniform: same-step conflicts exist
            // phase 3a: the probes that set a marked bit join the list
#pragma unroll
            for (int q = 0; q < int(RUN_Q); q++) {
                if (32u * q >= n) break;
                if ((won >> q) & 1u) {
                    const uint32_t bit = cur[q] >> ENT_BIT_SHIFT;
                    if ((lds_u32(a_mf + (((bit >> 5) & (MF_WORDS - 1)) << 2)) >> (bit & 31u)) & 1u) {
                        member |= 1u << q;
                        bucket_conflict_join(s_list, mt, mf, &s_n[par], bit, (warp << ENT_BIT_SHIFT) | (cur[q] & ((1u << ENT_BIT_SHIFT) - 1)), false);
                    }
                }
            }
            __syncthreads();
            // phase 3b: inside the list the smallest file position of a bit wins
            const uint32_t nl = s_n[par];
            if (nl > BB_LIST_CAP) {   // conflict list too small: bloom_slow_kernel redoes the species
                if (threadIdx.x == 0) R.slow[s] = 1u;
                return;
            }
            while (member) {
                const int q = __ffs(member) - 1;
                member &= member - 1;
                uint32_t c = 0;
#pragma unroll
                for (int i = 0; i < int(RUN_Q); i++) c = i == q ? cur[i] : c;
                if (bucket_conflict_lost(s_list, mt, nl, c >> ENT_BIT_SHIFT, (warp << ENT_BIT_SHIFT) | (c & ((1u << ENT_BIT_SHIFT) - 1))))
                    lost |= 1u << q;
            }
        }
        // lost probes -> lost map (atomicOr: idempotent).  About one probe in ten is lost: the loop runs over the set
        // bits (once or twice per warp) instead of over all RUN_Q entries with three lanes in 32 active each time.
        if (lost) {
            const uint32_t c_tile = tile_c0(cb, t);
            while (lost) {
                const int q = __ffs(lost) - 1;
                lost &= lost - 1;
                uint32_t e = 0;
#pragma unroll
                for (int i = 0; i < int(RUN_Q); i++) e = i == q ? cur[i] : e;
                const uint32_t c = c_tile + ((e >> 1) & (BP_TILE - 1));
                atomicOr(R.lost2 + c / CHUNK, 1u << (2u * (c % CHUNK) + (e & 1u)));
            }
        }
    }
}

// Exact path for species with a run-slot overflow: one CTA per species, bitset in global memory
// (4 MiB per CTA, L2-resident), one source tile per step straight from codes[].
constexpr int BS_THREADS = BP_THREADS;
constexpr uint32_t BS_NE = 2 * CHUNK;                       // probes per thread and step
constexpr uint32_t BS_LIST_CAP = BP_TILE * 2;
constexpr size_t BS_SMEM = size_t(BS_LIST_CAP) * sizeof(unsigned long long);
constexpr uint32_t BS_MAX_CTAS = 32;                        // 4 MiB of scratch bits each

__global__ void __launch_bounds__(BS_THREADS, 1)
bloom_slow_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, uint32_t s0, uint32_t s1,
                  int k, uint64_t k_mask, BloomRuns R, uint32_t* __restrict__ scratch_bits) {
    extern __shared__ __align__(16) unsigned long long bs_list[];   // BS_LIST_CAP
    __shared__ unsigned long long s_mt[2 * MT_SLOTS];
    __shared__ uint32_t s_mf[2 * MF_WORDS];
    __shared__ uint32_t s_n[2];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t* bitset = scratch_bits + (size_t(blockIdx.x) << (BLOOM_BITS_LOG2 - 5));
    const StepShared S{bs_list, BS_LIST_CAP, s_mf, s_mt, s_n};
    for (uint32_t s = s0 + blockIdx.x; s < s1; s += gridDim.x) {
        if (!R.slow[s]) continue;   // CTA-uniform
        const uint32_t cb = sp_cstart[s], ce = sp_cstart[s + 1];
        const uint32_t n_tiles = R.sp_tile0[s + 1] - R.sp_tile0[s];
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < (1u << (BLOOM_BITS_LOG2 - 5)); i += BS_THREADS) bitset[i] = 0;
        for (uint32_t i = threadIdx.x; i < 2 * MF_WORDS; i += BS_THREADS) s_mf[i] = 0;
        for (uint32_t i = threadIdx.x; i < 2 * MT_SLOTS; i += BS_THREADS) s_mt[i] = ~0ull;
        if (threadIdx.x < 2) s_n[threadIdx.x] = 0;
        __threadfence();
        __syncthreads();
        uint32_t par = 0;
        for (uint32_t t = 0; t < n_tiles; t++, par ^= 1u) {
            const uint32_t base = tile_c0(cb, t);
            const uint32_t c0 = base + warp * WARP_TILE + lane * CHUNK;
            uint32_t bit[BS_NE], p[BS_NE], live = 0;
#pragma unroll
            for (int e = 0; e < int(BS_NE); e++) bit[e] = p[e] = 0;
            if (base + warp * WARP_TILE < ce) {   // warp-uniform
                uint64_t kmer[CHUNK];
                const uint32_t valid_all = load_chunk_kmers<true>(codes, base + warp * WARP_TILE, k, k_mask, kmer);
                const uint32_t valid = valid_all & range_mask16(c0, cb, ce);
                const uint32_t todo = valid & ~dup_mask16(kmer, valid_all);   // duplicates: marked by bloom_partition_kernel
#pragma unroll
                for (int j = 0; j < CHUNK; j++) {
                    if (!((todo >> j) & 1u)) continue;
                    bloom_indices(kmer[j], bit[2 * j], bit[2 * j + 1]);
                    p[2 * j] = p[2 * j + 1] = warp * WARP_TILE + lane * CHUNK + j;   // the two probes of a position never share a bit
                    live |= 3u << (2 * j);
                }
            }
            uint32_t lost = 0;
            bloom_step<int(BS_NE), true>(bitset, S, par, bit, p, live, int(BS_NE), lost);   // the list holds every probe: no overflow
            if (lost) atomicOr(R.lost2 + c0 / CHUNK, lost);   // probe e = 2j + pr is bit 2j + pr of the chunk's word
            __threadfence();   // the bits live in global memory: order this step's atomics before the next step's loads
        }
    }
}

}  // namespace kmn

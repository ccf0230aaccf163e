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
//                           k-mer once and splits its two probes over the 256 buckets of 2^17 filter
//                           bits; entry = (bit inside the bucket, position inside the tile).  Each
//                           (tile, bucket) run has a FIXED slot in global memory (96 entries, mean 64), so
//                           there are no global atomics and the runs of a bucket are ordered by tile = by
//                           file position; the few probes that do not fit their slot go to the tile's
//                           overflow list.  A position whose k-mer equals the previous position's (a
//                           homopolymer run) can never pass and is marked without probing.
//   bloom_bucket_kernel     one WARP per (species, bucket): the bucket's 2^17 bits in the warp's own 16 KiB of
//                           shared memory (14 warps per SM, one persistent CTA), runs consumed in file order,
//                           one run (= one source tile) per step, no CTA barrier anywhere.  A probe whose bit
//                           was set by an earlier step is lost.  Inside a step every probe sets its bit with a
//                           shared-memory atomicOr; two probes of one step on the same bit (1.6 % of the steps)
//                           are settled by the warp itself: the smallest file position wins.  A lost probe
//                           sets its bit (2 per position) in the lost map.
//   bloom_slow_kernel       same step logic on a 4 MiB global bitset, one CTA per species, straight from
//                           codes[]: only for species whose probes overflow a run slot (massively
//                           repeated k-mers) — exact for every input, fast for natural ones.
// An occurrence passes iff it is a valid k-mer and not both of its probes are in the lost map; the
// count-min partition (cms_kernels.cuh) reads the map directly.  The map is written with atomicOr only,
// so redoing a species on the slow path is idempotent.
#pragma once
#include <type_traits>
#include "kmn_device.cuh"

namespace kmn {

#ifndef KMN_BP_ILP
#define KMN_BP_ILP 1
#endif
constexpr int BP_THREADS = 512;
constexpr uint32_t BP_TILE_WT = BP_THREADS / 32;          // warp tiles per source tile (one per warp)
constexpr uint32_t BP_TILE = BP_TILE_WT * WARP_TILE;       // 8192 positions
constexpr uint32_t BP_POS_BITS = 13;
#ifndef KMN_BB_LOG2
#define KMN_BB_LOG2 17
#endif
constexpr uint32_t BB_LOG2 = KMN_BB_LOG2;                  // filter bits per bucket (16 KiB of shared memory per warp)
constexpr uint32_t BB_BUCKETS = 1u << (BLOOM_BITS_LOG2 - BB_LOG2);   // 256
constexpr uint32_t BB_WORDS = 1u << (BB_LOG2 - 5);
constexpr uint32_t BB_MASK = (1u << BB_LOG2) - 1;
// slots per (tile, bucket): the probes of a tile land in a bucket's run as Binomial(16 Ki, 1 / buckets): mean 64,
// sigma 8 at 256 buckets.  4 sigma of head room; what does not fit (3e-5 of the runs) goes to the tile's overflow list.
#ifndef KMN_RUN_CAP
#define KMN_RUN_CAP (BB_LOG2 == 17 ? 96 : BB_LOG2 == 16 ? 52 : 176)
#endif
constexpr uint32_t RUN_CAP = KMN_RUN_CAP;
constexpr uint32_t RUN_Q = (RUN_CAP + 31) / 32;            // entries per lane when a warp reads one run
constexpr uint32_t OVF_CAP = 64;                           // overflow entries per source tile; more -> slow path (exact)
constexpr uint32_t OVF_Q = OVF_CAP / 32;
constexpr uint32_t CNT_OVF = 0x8000u;                      // flag in a run's count: part of the run sits in the overflow list
constexpr uint32_t BW_NQ = RUN_Q + OVF_Q;                  // probes per lane and step, overflow entries included
constexpr uint32_t BW_WARPS = (227u * 1024u - 2048u) / (BB_WORDS * 4u);   // 14 warps, each with its own bucket
constexpr int BW_THREADS = int(BW_WARPS) * 32;
constexpr size_t BW_SMEM = size_t(BW_WARPS) * BB_WORDS * sizeof(uint32_t);
#ifndef KMN_BW_DEPTH
#define KMN_BW_DEPTH 4
#endif
constexpr uint32_t BW_DEPTH = KMN_BW_DEPTH;                           // runs in flight per warp (global-load latency is hidden by depth, not by occupancy)
constexpr uint32_t BB_LIST_CAP = 2048;                     // (bloom_slow_kernel) conflict list of a step
constexpr uint32_t ENT_BIT_SHIFT = BP_POS_BITS + 1;         // entry = bit_in_bucket << 14 | pos_in_tile << 1 | probe
constexpr uint32_t ENT_ORDER_MASK = (1u << ENT_BIT_SHIFT) - 1;
constexpr uint32_t MF_WORDS = 256;                         // (bloom_slow_kernel) 8192-bit mini filter of a step's conflicting bits
constexpr uint32_t MT_SLOTS = 256;                         // (bloom_slow_kernel) hashed min table of a step's conflict list
static_assert(BP_TILE == (1u << BP_POS_BITS), "entry layout");
static_assert(BB_LOG2 + ENT_BIT_SHIFT <= 32, "entry layout");
static_assert(BB_BUCKETS % BP_TILE_WT == 0 && BB_BUCKETS <= BP_THREADS, "bucket ownership");
static_assert(RUN_CAP < CNT_OVF && BW_WARPS >= 1 && BW_THREADS <= 1024 && 32 % BW_DEPTH == 0, "bucket kernel shape");
constexpr size_t bp_smem_for(uint32_t cap) {
    return size_t(BB_BUCKETS) * cap * sizeof(uint32_t) + BB_BUCKETS * sizeof(uint32_t);
}

struct BloomRuns {
    uint32_t* ent;              // [n_tiles][BB_BUCKETS][cap] bit_in_bucket << 14 | pos_in_tile << 1 | probe
    uint16_t* cnt;              // per species: [BB_BUCKETS][tiles of the species] (a bucket's counts are contiguous); | CNT_OVF
    uint32_t cap;               // RUN_CAP (smaller only under the test knob)
    unsigned long long* ovf;    // [n_tiles][OVF_CAP] (bucket << 32) | entry: probes that did not fit their run slot
    uint32_t* ovf_n;            // [n_tiles] entries in the tile's overflow list
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

// The staging buffer IS the tile's slot array ([bucket][cap], the layout of R.ent): when a tile is ranked, one
// elected thread hands the whole buffer to the bulk-copy engine (cp.async.bulk.global.shared::cta, SASS UBLKCP) and
// the CTA — persistent, two per SM — moves on to load and hash its next tile while the copy drains; it only waits
// for the copy to have READ the buffer right before it ranks into it again.  No thread copies entries out.
// Position of rank r inside a run: r ^ (bucket & 31) when the slot size is a multiple of 32 (bank = function of
// the bucket: the stores of a warp — random buckets, similar ranks — would otherwise pile up on a few banks); the
// bucket kernel reads a run in whatever order it lies.
__device__ __forceinline__ void bulk_store_group(void* dst_global, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t run_swizzle(uint32_t cap, uint32_t bucket) { return (cap & 31u) ? 0u : (bucket & 31u); }

// CAP: the run slot size as a compile-time constant (RUN_CAP: the slot address and the swizzle then cost 6 instead
// of 11 instructions per probe); 0: R.cap at run time (test knob KMN_BLOOM_RUN_CAP).
template <uint32_t CAP>
__global__ void __launch_bounds__(BP_THREADS, 2)
bloom_partition_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ sp_cstart, int k,
                       uint64_t k_mask, BloomRuns R) {
    extern __shared__ __align__(128) uint32_t bp_smem[];
    const uint32_t cap = CAP ? CAP : R.cap;
    const uint32_t stride = cap;
    uint32_t* stage = bp_smem;                               // [BB_BUCKETS][cap]
    uint32_t* fill = bp_smem + size_t(BB_BUCKETS) * stride;  // [BB_BUCKETS]
    __shared__ uint32_t s_ovf_n;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = *R.n_tiles;
    const uint32_t stage_bytes = BB_BUCKETS * stride * uint32_t(sizeof(uint32_t));
    for (uint32_t tg = blockIdx.x; tg < n_tiles; tg += gridDim.x) {
        const uint32_t s = R.tile_sp[tg];
        const uint32_t cb = sp_cstart[s], ce = sp_cstart[s + 1];
        const uint32_t tg0 = R.sp_tile0[s], nt = R.sp_tile0[s + 1] - tg0;
        const uint32_t base = tile_c0(cb, tg - tg0);
        const uint32_t c0 = base + warp * WARP_TILE + lane * CHUNK;
        const bool active = base + warp * WARP_TILE < ce;   // warp-uniform
        uint64_t kmer[CHUNK];
        uint32_t todo = 0;
        if (active) {
            const uint32_t valid_all = load_chunk_kmers<true>(codes, base + warp * WARP_TILE, k, k_mask, kmer);
            const uint32_t valid = valid_all & range_mask16(c0, cb, ce);
            const uint32_t dup = dup_mask16(kmer, valid_all) & valid;
            if (dup) atomicOr(R.lost2 + c0 / CHUNK, spread2(dup, 3u));
            todo = valid & ~dup;
        }
        // the previous tile's copy has read the staging buffer; its counters were consumed before the copy was issued
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            s_ovf_n = 0;
        }
        if (threadIdx.x < BB_BUCKETS) fill[threadIdx.x] = 0;
        __syncthreads();
        if (active) {
            // explicit shared-memory addresses; with a compile-time slot size the slot address is
            // (stage + b * 4 CAP + 4 rank) ^ 4 (b & 31): the base is 128-byte aligned and a slot a multiple of 128 bytes,
            // so the swizzle only touches bits the sum leaves alone, and everything but one LEA and one LOP3 is
            // computed while the rank atomic is in flight
            const uint32_t a_fill = static_cast<uint32_t>(__cvta_generic_to_shared(fill));
            const uint32_t a_stage = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
            auto emit = [&](uint32_t i, uint32_t pos, uint32_t pr) {
                const uint32_t b = i >> BB_LOG2;
                uint32_t rank;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(rank) : "r"(a_fill + 4u * b) : "memory");
                const uint32_t entry = ((i & BB_MASK) << ENT_BIT_SHIFT) | (pos << 1) | pr;
                if (rank < cap) {
                    uint32_t a;
                    if (CAP && CAP % 32u == 0) a = (a_stage + b * (4u * CAP) + 4u * rank) ^ ((i >> (BB_LOG2 - 2)) & 0x7Cu);
                    else a = a_stage + 4u * (b * stride + (rank ^ run_swizzle(cap, b)));
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(entry) : "memory");
                } else {   // rare: the run slot is full
                    const uint32_t o = atomicAdd(&s_ovf_n, 1u);
                    if (o < OVF_CAP) R.ovf[size_t(tg) * OVF_CAP + o] = (static_cast<unsigned long long>(b) << 32) | entry;
                    else R.slow[s] = 1u;
                }
            };
#if KMN_BP_ILP > 1
#pragma unroll
            for (int j = 0; j < CHUNK; j += KMN_BP_ILP) {
                uint32_t idx[KMN_BP_ILP][2];
#pragma unroll
                for (int u = 0; u < KMN_BP_ILP; u++) bloom_indices(kmer[j + u], idx[u][0], idx[u][1]);
#pragma unroll
                for (int u = 0; u < KMN_BP_ILP; u++) asm volatile("" : "+r"(idx[u][0]), "+r"(idx[u][1]));
#pragma unroll
                for (int u = 0; u < KMN_BP_ILP; u++) {
                    if (!((todo >> (j + u)) & 1u)) continue;
                    const uint32_t pos = warp * WARP_TILE + lane * CHUNK + j + u;
                    emit(idx[u][0], pos, 0u);
                    emit(idx[u][1], pos, 1u);
                }
            }
#else
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                if (!((todo >> j) & 1u)) continue;
                uint32_t idx[2];
                bloom_indices(kmer[j], idx[0], idx[1]);
                const uint32_t pos = warp * WARP_TILE + lane * CHUNK + j;
                emit(idx[0], pos, 0u);
                emit(idx[1], pos, 1u);
            }
#endif
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's staging stores, before the bulk copy reads them
        __syncthreads();
        if (threadIdx.x < BB_BUCKETS) {
            const uint32_t f = fill[threadIdx.x];
            R.cnt[size_t(tg0) * BB_BUCKETS + size_t(threadIdx.x) * nt + (tg - tg0)] = uint16_t(f > cap ? (cap | CNT_OVF) : f);
        }
        if (threadIdx.x == 0) {
            R.ovf_n[tg] = min(s_ovf_n, OVF_CAP);
            // the whole slot array of the tile (dead positions included) in pieces of at most 32 KiB
            uint8_t* dst = reinterpret_cast<uint8_t*>(R.ent + size_t(tg) * BB_BUCKETS * cap);
            const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
            for (uint32_t o = 0; o < stage_bytes; o += 32768u) bulk_store_group(dst + o, src + o, min(stage_bytes - o, 32768u));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

// ---- bloom_bucket_kernel: one warp per (species, bucket) --------------------------------------------
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

__device__ __forceinline__ void red_or_if(uint32_t* p, uint32_t v, bool pred) {   // predicated, no branch
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.global.or.b32 [%0], %1;\n\t}" ::"l"(p), "r"(v), "r"(uint32_t(pred)) : "memory");
}

// Rare (1.6 % of the steps): probes of ONE step share a bit.  `conf` marks this lane's probes that found their bit
// set by another probe of the step; all probes of that bit (the one that set it included) are compared and every
// one but the first in file order is lost.  All 32 lanes call this.
template <int NQ>
__device__ __forceinline__ uint32_t warp_step_conflicts(const uint32_t (&cur)[NQ], uint32_t todo, uint32_t conf) {
    uint32_t lost = 0;
    for (;;) {
        const uint32_t open = __ballot_sync(0xffffffffu, conf != 0);
        if (!open) break;
        uint32_t mine = 0;
        if (conf) {
            const int q = __ffs(conf) - 1;
#pragma unroll
            for (int i = 0; i < NQ; i++) mine = i == q ? (cur[i] >> ENT_BIT_SHIFT) : mine;
        }
        const uint32_t bit = __shfl_sync(0xffffffffu, mine, __ffs(open) - 1);
        uint32_t match = 0, best = 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < NQ; q++)
            if (((todo >> q) & 1u) && (cur[q] >> ENT_BIT_SHIFT) == bit) {
                match |= 1u << q;
                best = min(best, cur[q] & ENT_ORDER_MASK);
            }
        const uint32_t first = __reduce_min_sync(0xffffffffu, best);
#pragma unroll
        for (int q = 0; q < NQ; q++)
            if (((match >> q) & 1u) && (cur[q] & ENT_ORDER_MASK) != first) lost |= 1u << q;
        conf &= ~match;
    }
    return lost;
}

// One step on a run whose tail sits in the tile's overflow list (3e-5 of the runs; every run under the
// KMN_BLOOM_RUN_CAP test knob): the run's entries are read again here, the overflow entries of this bucket join
// them, same three phases with plain branches.  `lost2_tile` = lost-map word of the tile's first position.
__device__ __noinline__ void bucket_step_overflow(uint32_t a_bits, const uint32_t* __restrict__ run, uint32_t n,
                                                  const unsigned long long* __restrict__ ovf, uint32_t n_ovf,
                                                  uint32_t beta, uint32_t* __restrict__ lost2_tile, uint32_t cap) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t rlane = lane ^ run_swizzle(cap, beta);
    uint32_t cur[BW_NQ], live = 0;
#pragma unroll
    for (int q = 0; q < int(RUN_Q); q++) {
        cur[q] = 0;
        if (rlane + 32u * q < n) {
            cur[q] = __ldcs(run + 32 * q);
            live |= 1u << q;
        }
    }
#pragma unroll
    for (int x = 0; x < int(OVF_Q); x++) {
        cur[RUN_Q + x] = 0;
        const uint32_t j = lane + 32u * uint32_t(x);
        if (j < n_ovf) {
            const unsigned long long e = ovf[j];
            if (uint32_t(e >> 32) == beta) {
                cur[RUN_Q + x] = uint32_t(e);
                live |= 1u << (RUN_Q + x);
            }
        }
    }
    uint32_t pre = 0;
#pragma unroll
    for (int q = 0; q < int(BW_NQ); q++)
        if ((live >> q) & 1u) {
            const uint32_t bit = cur[q] >> ENT_BIT_SHIFT;
            pre |= ((lds_u32(a_bits + ((bit >> 5) << 2)) >> (bit & 31u)) & 1u) << q;
        }
    __syncwarp();
    uint32_t conf = 0;
    const uint32_t todo = live & ~pre;
#pragma unroll
    for (int q = 0; q < int(BW_NQ); q++)
        if ((todo >> q) & 1u) {
            const uint32_t bit = cur[q] >> ENT_BIT_SHIFT, m = 1u << (bit & 31u);
            if (atoms_or_u32(a_bits + ((bit >> 5) << 2), m) & m) conf |= 1u << q;
        }
    __syncwarp();
    uint32_t lost = pre;
    if (__any_sync(0xffffffffu, conf != 0)) lost |= warp_step_conflicts<int(BW_NQ)>(cur, todo, conf);
#pragma unroll
    for (int q = 0; q < int(BW_NQ); q++)
        if ((lost >> q) & 1u) atomicOr(lost2_tile + ((cur[q] >> 5) & (BP_TILE / CHUNK - 1)), 1u << (cur[q] & 31u));
}

// One persistent CTA per SM; every warp owns BB_WORDS of shared memory and takes (species, bucket) items on its
// own.  With 14 warps per SM there is little occupancy to hide global-load latency behind: a warp keeps the
// entries of BW_DEPTH runs in flight in registers, and the run lengths of the next 32 .. 64 tiles in two registers
// (lane l holds tile blk + l; a shuffle hands a length to the warp).  The step itself is branch-free: a dead lane
// slot ORs a zero mask into some word of the warp's own bitset, a lost probe's lost-map update is a predicated RED;
// the only branches are warp-uniform (third lane slot in use? overflow run? conflict inside the step?).
// Entry e = bit << 14 | pos << 1 | probe: its bitset word is (e >> 19), its bit (e >> 14) & 31; its lost-map word
// (2 bits per position, 16 positions per word) is tile_word + ((e >> 5) & 511), its lost-map bit e & 31.
__global__ void __launch_bounds__(BW_THREADS, 1)
bloom_bucket_kernel(const uint32_t* __restrict__ sp_cstart, uint32_t s0, uint32_t n_items, BloomRuns R) {
    extern __shared__ __align__(16) uint32_t bw_smem[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t* bits = bw_smem + size_t(warp) * BB_WORDS;
    const uint32_t a_bits = smem_addr(bits);
    const size_t run_stride = size_t(BB_BUCKETS) * R.cap;
    for (uint32_t item = blockIdx.x * BW_WARPS + warp; item < n_items; item += gridDim.x * BW_WARPS) {
        const uint32_t s = s0 + item / BB_BUCKETS, beta = item % BB_BUCKETS;
        if (R.slow[s]) continue;   // warp-uniform
        const uint32_t tg0 = R.sp_tile0[s], n_tiles = R.sp_tile0[s + 1] - tg0;
        if (n_tiles == 0) continue;
        const uint32_t cb = sp_cstart[s];
        __syncwarp();
        for (uint32_t i = lane; i < BB_WORDS / 4; i += 32) reinterpret_cast<uint4*>(bits)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        // this bucket's run lengths are contiguous over the tiles
        const uint16_t* __restrict__ cnt = R.cnt + size_t(tg0) * BB_BUCKETS + size_t(beta) * n_tiles;
        const uint32_t* __restrict__ ent0 = R.ent + (size_t(tg0) * BB_BUCKETS + beta) * R.cap + lane;
        const uint32_t rlane = lane ^ run_swizzle(R.cap, beta);   // this lane's slot positions hold ranks rlane + 32 q
        auto cnt_block = [&](uint32_t t) -> uint32_t { return t + lane < n_tiles ? uint32_t(__ldg(cnt + t + lane)) : 0u; };
        uint32_t blk = 0, c_cur = cnt_block(0), c_nxt = cnt_block(32);
        auto run_len = [&](uint32_t t) -> uint32_t {   // t in [blk, blk + 64); 0 past the last tile
            return __shfl_sync(0xffffffffu, t - blk < 32u ? c_cur : c_nxt, int(t & 31u));
        };
        uint32_t raw[BW_DEPTH][RUN_Q], n_raw[BW_DEPTH];
        const uint32_t* ent[BW_DEPTH];
        auto load_run = [&](const uint32_t* e, uint32_t n, uint32_t (&dst)[RUN_Q]) {
            const uint32_t nn = n & ~CNT_OVF;
#pragma unroll
            for (int q = 0; q < int(RUN_Q); q++)
                if (rlane + 32u * q < nn) dst[q] = __ldcs(e + 32 * q);
        };
#pragma unroll
        for (int i = 0; i < int(BW_DEPTH); i++) {
#pragma unroll
            for (int q = 0; q < int(RUN_Q); q++) raw[i][q] = 0;
            n_raw[i] = run_len(uint32_t(i));
            ent[i] = ent0 + size_t(i) * run_stride;
            load_run(ent[i], n_raw[i], raw[i]);
        }
        uint32_t* lost2_tile = R.lost2 + tile_c0(cb, 0) / CHUNK;
        // one step: the run of tile t sits in (e_raw, n_run, run); PREFETCH: the run BW_DEPTH tiles ahead takes its place
        auto step = [&](uint32_t t, uint32_t (&e_raw)[RUN_Q], uint32_t& n_run, const uint32_t*& run, auto prefetch) {
            const uint32_t nr = n_run;
            uint32_t cur[RUN_Q];
#pragma unroll
            for (int q = 0; q < int(RUN_Q); q++) cur[q] = e_raw[q];
            const uint32_t* this_run = run;
            if (decltype(prefetch)::value) {
                n_run = run_len(t + BW_DEPTH);
                run += size_t(BW_DEPTH) * run_stride;
                load_run(run, n_run, e_raw);
            }
            if (nr & CNT_OVF) {   // rare: the rest of this run sits in the tile's overflow list
                const uint32_t tile = tg0 + t;
                bucket_step_overflow(a_bits, this_run, nr & ~CNT_OVF, R.ovf + size_t(tile) * OVF_CAP, min(R.ovf_n[tile], OVF_CAP), beta,
                                     lost2_tile, R.cap);
            } else {
                const int nq = int((nr + 31u) >> 5);   // warp-uniform: lane slots in use
                uint32_t a[RUN_Q], m[RUN_Q];
                bool pre[RUN_Q];
                // phase 1: bits set by earlier steps (= by smaller file positions)
#pragma unroll
                for (int q = 0; q < int(RUN_Q); q++) {
                    if (q >= 2 && q >= nq) break;   // warp-uniform
                    a[q] = a_bits + ((cur[q] >> (ENT_BIT_SHIFT + 3)) & (BB_WORDS * 4u - 4u));
                    m[q] = rlane + 32u * q < nr ? 1u << ((cur[q] >> ENT_BIT_SHIFT) & 31u) : 0u;
                    pre[q] = (lds_u32(a[q]) & m[q]) != 0u;
                }
                // phase 2: every remaining probe sets its bit; finding it set now means a probe of this very step did it
                // (no __syncwarp between the phases: the step is branch-free, the warp converged, and a warp's
                // shared-memory accesses execute in program order)
                bool conf[RUN_Q], any_conf = false;
#pragma unroll
                for (int q = 0; q < int(RUN_Q); q++) {
                    if (q >= 2 && q >= nq) break;
                    const uint32_t m2 = pre[q] ? 0u : m[q];
                    conf[q] = (atoms_or_u32(a[q], m2) & m2) != 0u;
                    any_conf |= conf[q];
                }
                uint32_t extra = 0;
                if (__any_sync(0xffffffffu, any_conf)) {   // rare
                    uint32_t todo = 0, cm = 0;
#pragma unroll
                    for (int q = 0; q < int(RUN_Q); q++) {
                        if (q >= 2 && q >= nq) break;
                        if (m[q] && !pre[q]) todo |= 1u << q;
                        if (conf[q]) cm |= 1u << q;
                    }
                    extra = warp_step_conflicts<int(RUN_Q)>(cur, todo, cm);
                }
                // lost probes -> lost map (idempotent OR); about one probe in ten
#pragma unroll
                for (int q = 0; q < int(RUN_Q); q++) {
                    if (q >= 2 && q >= nq) break;
                    red_or_if(lost2_tile + ((cur[q] >> 5) & (BP_TILE / CHUNK - 1)), 1u << (cur[q] & 31u),
                              pre[q] || ((extra >> q) & 1u));
                }
            }
            lost2_tile += BP_TILE / CHUNK;
        };
        // full groups of BW_DEPTH steps: no exit inside the unrolled body, so each of its copies keeps its own registers
        // (an exit test per step made ptxas rotate the prefetched entries through one register set at the end of every
        // step, i.e. wait for the loads it had just issued)
        const uint32_t n_full = n_tiles - n_tiles % BW_DEPTH;
        for (uint32_t t0 = 0; t0 < n_full; t0 += BW_DEPTH) {
            if (t0 == blk + 32u) {
                blk += 32u;
                c_cur = c_nxt;
                c_nxt = cnt_block(blk + 32u);
            }
#pragma unroll
            for (int i = 0; i < int(BW_DEPTH); i++) step(t0 + uint32_t(i), raw[i], n_raw[i], ent[i], std::true_type{});
        }
#pragma unroll
        for (int i = 0; i < int(BW_DEPTH) - 1; i++)
            if (n_full + uint32_t(i) < n_tiles) step(n_full + uint32_t(i), raw[i], n_raw[i], ent[i], std::false_type{});
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

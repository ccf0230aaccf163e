// cms_kernels.cuh — count-min build of pass 1 (AtomicCountMinSketch::increment / snapshot,
// /root/reference/src/proba_filter.rs:146-177) as partition -> shared-memory histogram -> merge.
//
// Measured on B200 (tools/microbench/l2_atomics.cu): scattered RED.add into an L2-resident slice
// runs at ~190 Gop/s chip-wide, into HBM at ~23-34 Gop/s, while shared-memory atomics run at
// ~1000 Gop/s.  So the 4 x (new k-mers) increments never touch global memory as atomics:
//
//   cms_partition_kernel  each CTA takes a tile of 16 Ki positions of one species, derives "passed the
//                         Bloom filter" from the lost map (count_kernels.cuh), keeps the k-mers in
//                         registers for the four rows and splits the cell indices of one row at a time
//                         over the row's 1024 buckets of 2^16 cells: rank = shared atomicAdd on the
//                         bucket's fill counter, the low 16 index bits go to the bucket's 32-entry slot of a
//                         64 KiB staging buffer in shared memory (31 entries + their count), and the whole
//                         buffer leaves for FIXED (bucket, tile) slots in global memory as four TMA tensor
//                         stores issued by one thread while the next row is being ranked in the second
//                         buffer: no reservation atomics, no copy-out instructions.  Also counts the passes
//                         per species.
//   cms_apply_kernel      one CTA per bucket: a 128 KiB shared-memory histogram (two u16 counters per
//                         word) of the bucket's slots (contiguous over the tiles), then
//                         cell = min(cell + count, 65535)  on the u16 table by the bucket's only owner —
//                         no global atomics, no u32 accumulators, no separate snapshot pass.
//
// The table itself is the reference's saturating u16 table at all times: every update is +1 and
// saturation is monotone, so adding a batch's exact per-cell count with a clamp is bit-identical to
// the CAS loop of proba_filter.rs:151-166, and partial tables of several GPUs combine as
// min(sum_of_clamped, 65535) == min(sum, 65535).
//
// Rare paths, all exact: a slot that overflows inside a tile spills single entries to a global list
// (cms_spill_kernel, CAS loop on the u16 cell); a shared counter that would pass 65535 inside one
// batch makes the CTA redo its bucket with u32 counters in two halves; a spill list
// that runs out of space raises a flag that disables apply/spill and enables cms_direct_kernel,
// which rebuilds the batch's increments with global CAS loops straight from codes[] + the lost map.
#pragma once
#include "kmn_device.cuh"
#include <cuda.h>   // CUtensorMap (the driver entry point is resolved at run time, nothing links against libcuda)

namespace kmn {

#ifndef KMN_PART_ILP
#define KMN_PART_ILP 2   // positions whose hashes cms_partition_kernel computes together
#endif
constexpr uint32_t CMS_BUCKET_LOG2 = 16;                                             // cells per bucket
constexpr uint32_t CMS_BUCKETS_PER_ROW = 1u << (CMS_WIDTH_LOG2 - CMS_BUCKET_LOG2);    // 1024
constexpr uint32_t CMS_BUCKETS = CMS_DEPTH * CMS_BUCKETS_PER_ROW;                     // 4096
constexpr int PART_THREADS = 1024;                   // == CMS_BUCKETS_PER_ROW: thread t closes the slot of bucket t
constexpr int PART_TILE = PART_THREADS * CHUNK;      // 16384 positions per CTA tile
constexpr int PART_WTILES = PART_TILE / WARP_TILE;   // one warp tile per warp
constexpr int PART_SLOT = 32;                        // u16 per (tile, bucket) slot: 31 entries (mean 15) + their count
constexpr int PART_SLOT_CAP = PART_SLOT - 1;
constexpr int PART_BOX_BUCKETS = 256;                // buckets per TMA store (box dimensions are limited to 256)
constexpr size_t PART_STAGE_BYTES = size_t(CMS_BUCKETS_PER_ROW) * PART_SLOT * sizeof(uint16_t);   // 64 KiB per row
constexpr size_t PART_SMEM = 2 * PART_STAGE_BYTES + 2 * size_t(CMS_BUCKETS_PER_ROW) * sizeof(uint32_t) + 16;
constexpr int APPLY_THREADS = 1024;
constexpr uint32_t APPLY_WORDS = 1u << (CMS_BUCKET_LOG2 - 1);   // 32768 u32 words = 2^16 u16 counters
static_assert(PART_THREADS == int(CMS_BUCKETS_PER_ROW), "one closing thread per bucket");
static_assert(PART_SLOT * sizeof(uint16_t) == 64, "a slot is four 16-byte vectors");

// Partition output: FIXED slots.  ent[bucket][tile][32]: the low 16 index bits of the entries of (bucket, tile)
// in arrival order, entry 31 = their count.  A bucket's slots are contiguous over the tiles, so the apply pass
// streams them; a tile's 1024 slots of one row leave shared memory as four TMA tensor stores (no reservation
// atomics, no per-entry copy-out instructions).
// Inside a slot of 32 u16 the i-th entry (i < 31) sits at position i ^ slot_origin(bucket) and the count at
// position 31 ^ slot_origin(bucket).  Without the permutation the staging stores of a warp (random buckets, similar
// ranks) pile up on a handful of shared-memory banks (ncu: 80 % of the kernel's shared-memory wavefronts were
// conflict replays); with it the bank is a function of the bucket.  The apply pass owns one bucket per CTA, so the
// origin is a CTA constant there: position p is live iff (p ^ origin) < count.
__host__ __device__ __forceinline__ uint32_t slot_origin(uint32_t bucket) { return bucket & 31u; }

struct CmsSlots {
    uint16_t* ent;             // [CMS_BUCKETS][tiles_cap][PART_SLOT]
    uint32_t tiles_cap;        // slot columns allocated per bucket
    const uint32_t* running;   // device: [0] first tile index of the current species range, [1] tiles written so far
    uint32_t* spill;           // [spill_cap] flat cell offsets (row << 26 | idx) of slot overflows
    uint32_t* state;           // [0] fallback flag, [1] spill entries
    uint32_t spill_cap;
};

__host__ __device__ __forceinline__ uint64_t cms_seed(int r) {
    return r == 0 ? SEED_CMS_0 : r == 1 ? SEED_CMS_1 : r == 2 ? SEED_CMS_2 : SEED_CMS_3;
}

// saturating +1 on one u16 cell of the global table (the CAS loop of proba_filter.rs:151-166)
__device__ __forceinline__ void cms_cell_increment(uint16_t* __restrict__ tab, uint32_t cell) {
    uint32_t* w = reinterpret_cast<uint32_t*>(tab) + (cell >> 1);
    const uint32_t sh = (cell & 1u) * 16u;
    uint32_t old = *w;
    for (;;) {
        if (((old >> sh) & 0xFFFFu) == 0xFFFFu) return;
        const uint32_t seen = atomicCAS(w, old, old + (1u << sh));
        if (seen == old) return;
        old = seen;
    }
}

// Source tiles of the partition: 16 Ki positions of ONE species on the warp-tile grid (so the per-species
// counts of passing occurrences fall out of a block reduction).
struct CmsTiles {
    const uint32_t* tile_sp;    // [n_tiles] species of every tile
    const uint32_t* sp_tile0;   // [n_sp + 1] first tile of every species (species_tiles_kernel)
    const uint32_t* n_tiles;    // device scalar
};

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- TMA (bulk tensor) store of one staged row: shared memory -> the fixed slots in global memory -----------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"(smem), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One CTA per SM, persistent over the tiles.  Per tile the k-mers stay in registers for the four rows; per row
// every passing k-mer's cell index is ranked inside its bucket by a shared-memory atomicAdd and its low 16 bits
// land in the bucket's slot of the row's staging buffer; thread b then closes slot b (writes the count) and one
// elected thread hands the 64 KiB buffer to the TMA unit.  Two staging buffers: the store of row r drains while
// row r+1 is being ranked.  Two CTA barriers per row.
__global__ void __launch_bounds__(PART_THREADS, 1)
cms_partition_kernel(const __grid_constant__ CUtensorMap slots_map, const uint8_t* __restrict__ codes,
                     const uint32_t* __restrict__ sp_cstart, CmsTiles T, const uint32_t* __restrict__ lost2, int k,
                     uint64_t k_mask, CmsSlots L, uint32_t slot_limit, unsigned long long* __restrict__ new_per_species) {
    extern __shared__ __align__(1024) uint8_t part_smem[];   // TMA sources must be 128-byte aligned
    uint16_t* stage = reinterpret_cast<uint16_t*>(part_smem);                          // [2][1024][32]
    uint32_t* fill = reinterpret_cast<uint32_t*>(part_smem + 2 * PART_STAGE_BYTES);    // [2][1024]
    uint32_t& s_new = fill[2 * CMS_BUCKETS_PER_ROW];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        if (static_cast<uint32_t>(__cvta_generic_to_shared(part_smem)) & 127u) __trap();
        s_new = 0;
    }
    fill[threadIdx.x] = 0;
    fill[CMS_BUCKETS_PER_ROW + threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n_tiles = *T.n_tiles;
    const uint32_t tile0 = L.running[0];
    uint32_t buf = 0;
    for (uint32_t tg = blockIdx.x; tg < n_tiles; tg += gridDim.x) {
        const uint32_t s = T.tile_sp[tg];
        const uint32_t cb = sp_cstart[s], ce = sp_cstart[s + 1];
        const uint32_t wt = cb / WARP_TILE + (tg - T.sp_tile0[s]) * PART_WTILES + warp;
        const uint32_t c0 = wt * WARP_TILE + lane * CHUNK;
        uint64_t kmer[CHUNK];
        uint32_t new_bits = 0;
        if (uint64_t(wt) * WARP_TILE < ce) {   // warp-uniform
            const uint32_t valid = load_chunk_kmers<true>(codes, wt * WARP_TILE, k, k_mask, kmer) & range_mask16(c0, cb, ce);
            if (valid) new_bits = valid & ~both_lost_mask16(__ldcs(lost2 + c0 / CHUNK));
        }
        const uint32_t wn = __reduce_add_sync(0xffffffffu, uint32_t(__popc(new_bits)));
        if (lane == 0 && wn) atomicAdd(&s_new, wn);
#pragma unroll
        for (int j = 0; j < CHUNK; j++) kmer[j] = fold30_keep(kmer[j]);   // the registers keep the folded k-mers for the four rows
#pragma unroll 1
        for (int r = 0; r < int(CMS_DEPTH); r++, buf ^= 1u) {
            uint16_t* st = stage + size_t(buf) * (PART_STAGE_BYTES / 2);
            uint32_t* fl = fill + buf * CMS_BUCKETS_PER_ROW;
            const uint64_t seed = fold30_keep(cms_seed(r));   // (kept: ptxas otherwise re-folds the seed next to every use)
            // (Two branch-free forms of this loop were measured slower: all 16 hashes hoisted ahead of predicated atomics
            // and stores, 64 registers with spills: 1.27 ms; straight-line entries with a dummy bucket for dead lanes
            // and min(rank, cap) stores, 64 registers: 1.03 ms; the branchy form, 56 registers: 0.92 ms then, 0.84 ms with
            // the hashes of two positions computed together and the addressing below.)
            // Entry index inside the staging buffer = b * 32 + ((b & 31) ^ rank) = Tq ^ rank with Tq one LOP3 of two
            // shifts of the index (idx < 2^26, so idx >> 16 is the bucket itself); explicit shared-memory addresses so
            // that neither the counter nor the store address is re-derived from generic pointers per entry.
            const uint32_t a_fl = smem_addr_u32(fl), a_st = smem_addr_u32(st);
            auto emit = [&](uint32_t idx) {
                const uint32_t b = idx >> CMS_BUCKET_LOG2;
                uint32_t Tq;   // ((idx >> 11) & 0x7FE0) | (b & 31) as one LOP3: (x & c) | (b & ~c), c = 0x7FE0 (b < 1024)
                asm("lop3.b32 %0, %1, %2, 0x7FE0, 0xE4;" : "=r"(Tq) : "r"(idx >> (CMS_BUCKET_LOG2 - 5)), "r"(b));
                uint32_t rank;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(rank) : "r"(a_fl + 4u * b) : "memory");
                if (rank < slot_limit) {
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_st + 2u * (Tq ^ rank)), "h"(uint16_t(idx)) : "memory");
                } else {   // slot full: rare single-entry spill
                    const uint32_t sp = atomicAdd(L.state + 1, 1u);
                    if (sp < L.spill_cap) L.spill[sp] = (uint32_t(r) << CMS_WIDTH_LOG2) | idx;
                    else L.state[0] = 1u;
                }
            };
#if KMN_PART_ILP > 1
            // the hashes of KMN_PART_ILP neighbouring positions are computed together (independent dependency chains),
            // dead positions included; only the rank / store part is conditional
#pragma unroll
            for (int j = 0; j < CHUNK; j += KMN_PART_ILP) {
                uint32_t idx[KMN_PART_ILP];
#pragma unroll
                for (int u = 0; u < KMN_PART_ILP; u++) idx[u] = cms_index_folded(kmer[j + u], seed);
#pragma unroll
                for (int u = 0; u < KMN_PART_ILP; u++) asm volatile("" : "+r"(idx[u]));
#pragma unroll
                for (int u = 0; u < KMN_PART_ILP; u++)
                    if ((new_bits >> (j + u)) & 1u) emit(idx[u]);
            }
#else
#pragma unroll
            for (int j = 0; j < CHUNK; j++) {
                if (!((new_bits >> j) & 1u)) continue;
                emit(cms_index_folded(kmer[j], seed));
            }
#endif
            __syncthreads();
            // thread b closes slot b; the other buffer's counters (read one row ago) are cleared for the next row
            st[threadIdx.x * PART_SLOT + (slot_origin(threadIdx.x) ^ PART_SLOT_CAP)] = uint16_t(min(fl[threadIdx.x], slot_limit));
            fill[(buf ^ 1u) * CMS_BUCKETS_PER_ROW + threadIdx.x] = 0;
            fence_proxy_async_smem();
            if (threadIdx.x == 0) {
                if (r == 0 && s_new) {   // this tile's adds precede the barrier above, the next tile's follow the one below
                    atomicAdd(new_per_species + s, static_cast<unsigned long long>(s_new));
                    s_new = 0;
                }
                tma_wait_read_all();     // the previous row's store has left the other buffer
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(st));
#pragma unroll
                for (int q = 0; q < int(CMS_BUCKETS_PER_ROW) / PART_BOX_BUCKETS; q++)
                    tma_store_3d(&slots_map, a + q * (PART_BOX_BUCKETS * PART_SLOT * 2), 0, int(tile0 + tg),
                                 r * int(CMS_BUCKETS_PER_ROW) + q * PART_BOX_BUCKETS);
                tma_commit();
            }
        }
    }
    if (threadIdx.x == 0) tma_wait_all();
}

__device__ __forceinline__ uint32_t sat_add_u16x2(uint32_t g, uint32_t d) {
    const uint32_t lo = min((g & 0xFFFFu) + (d & 0xFFFFu), 0xFFFFu);
    const uint32_t hi = min((g >> 16) + (d >> 16), 0xFFFFu);
    return lo | (hi << 16);
}

// ---- mbarrier / bulk-copy primitives (TMA loads of the apply pass) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the same with an L2 evict-first hint (a stream that is read once and should not displace what else sits in L2)
__device__ __forceinline__ void bulk_load_evict_first(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// One CTA per bucket: histogram of the bucket's slots in 128 KiB of shared memory (two u16 counters per word),
// then the saturating merge into the table.  The bucket's slot array is contiguous (ent[bucket][tile][32]); it is
// streamed through a four-stage ring of 16 KiB shared-memory buffers by bulk async copies (one elected thread
// issues them, mbarriers signal "landed" / "consumed"), so no thread ever waits on a global load.  Thread t reads
// the t-th 16-byte vector of a stage (4 vectors per slot); the slot's count (position 31 ^ slot_origin)
// reaches the slot's four lanes by shuffle.  The histogram atomics do not return a value; a counter that
// wrapped or carried into its neighbour is detected afterwards (the histogram's total falls short of the number of
// entries) and the bucket is then redone with u32 counters.
#ifndef KMN_APPLY_VPT
#define KMN_APPLY_VPT 2
#endif
#ifndef KMN_APPLY_STAGES
#define KMN_APPLY_STAGES 3
#endif
constexpr int APPLY_STAGES = KMN_APPLY_STAGES;
constexpr uint32_t APPLY_VPT = KMN_APPLY_VPT;                              // vectors per thread and stage
constexpr uint32_t APPLY_STAGE_VEC = APPLY_THREADS * APPLY_VPT;
constexpr uint32_t APPLY_STAGE_BYTES = APPLY_STAGE_VEC * 16;               // 16 KiB per vector and thread
constexpr size_t APPLY_SMEM = size_t(APPLY_WORDS) * sizeof(uint32_t) + size_t(APPLY_STAGES) * APPLY_STAGE_BYTES + 256;   // + barriers (64 B) + one dummy word per lane

// Chunks carry a sequence number g that keeps counting over repeated passes (the exact redo streams the bucket
// again): chunk g uses stage g % STAGES for the (g / STAGES)-th time, which fixes the barrier parities.
struct ApplyRing {
    uint32_t ring, full, empty;   // shared-memory addresses: stage buffers, "landed" barriers, "consumed" barriers
    const uint4* src;
    uint32_t n_vec, n_chunks;
    uint32_t origin;              // slot_origin of the CTA's bucket
    bool evict_first;             // merging into a live table: the prefetched cells should survive the slot stream in L2
    uint32_t issued;              // sequence number of the next chunk to issue (elected thread)
    uint32_t done;                // sequence number of the first chunk of the current pass
};

__device__ __forceinline__ void apply_issue(ApplyRing& R) {   // elected thread only
    const uint32_t g = R.issued++, st = g % APPLY_STAGES, use = g / APPLY_STAGES;
    if (use > 0) mbar_wait(R.empty + 8 * st, (use - 1) & 1u);   // every thread has read the stage's previous content
    const uint32_t v0 = (g - R.done) * APPLY_STAGE_VEC;
    const uint32_t bytes = min(R.n_vec - v0, APPLY_STAGE_VEC) * 16u;
    mbar_expect_tx(R.full + 8 * st, bytes);
    if (R.evict_first) bulk_load_evict_first(R.ring + st * APPLY_STAGE_BYTES, R.src + v0, bytes, R.full + 8 * st);
    else bulk_load(R.ring + st * APPLY_STAGE_BYTES, R.src + v0, bytes, R.full + 8 * st);
}

// bump(c) for every live entry c of the bucket; returns this thread's number of live entries
template <typename Bump>
__device__ __forceinline__ uint32_t apply_stream(ApplyRing& R, Bump&& bump) {
    const unsigned lane = threadIdx.x & 31u;
    // rank of the entry at this lane's first position (its vector holds slot positions 8 (lane & 3) ..+7; position p
    // holds rank p ^ origin, and origin < 32, so position first_pos + j holds rank (first ^ j)) and
    // where the count sits: vector cnt_pos >> 3 of the slot, word cnt_word, its high half iff cnt_pos is odd
    const uint32_t first = (8u * (lane & 3u)) ^ R.origin;
    const uint32_t cnt_pos = R.origin ^ PART_SLOT_CAP;
    const int cnt_lane = int((lane & ~3u) | (cnt_pos >> 3));
    const uint32_t cnt_word = (cnt_pos & 7u) >> 1;
    uint32_t mine = 0;
    if (threadIdx.x == 0)
        for (uint32_t c = 0; c < min(R.n_chunks, uint32_t(APPLY_STAGES) - 1u); c++) apply_issue(R);
    for (uint32_t c = 0; c < R.n_chunks; c++) {
        if (threadIdx.x == 0 && c + APPLY_STAGES - 1 < R.n_chunks) apply_issue(R);
        const uint32_t g = R.done + c, st = g % APPLY_STAGES;
        mbar_wait(R.full + 8 * st, (g / APPLY_STAGES) & 1u);
        uint4 v[APPLY_VPT];
#pragma unroll
        for (uint32_t u = 0; u < APPLY_VPT; u++) {
            const uint32_t i = c * APPLY_STAGE_VEC + u * APPLY_THREADS + threadIdx.x;
            v[u] = make_uint4(0, 0, 0, 0);
            if (i < R.n_vec) {
                const uint32_t a = R.ring + st * APPLY_STAGE_BYTES + (u * APPLY_THREADS + threadIdx.x) * 16u;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "r"(a) : "memory");
            }
        }
        mbar_arrive(R.empty + 8 * st);
#pragma unroll
        for (uint32_t u = 0; u < APPLY_VPT; u++) {
            const uint32_t cw = cnt_word == 0 ? v[u].x : cnt_word == 1 ? v[u].y : cnt_word == 2 ? v[u].z : v[u].w;
            const uint32_t cnt = __shfl_sync(0xffffffffu, (cnt_pos & 1u) ? cw >> 16 : cw & 0xFFFFu, cnt_lane);
            if ((lane & 3u) == 0) mine += cnt;   // one lane per slot books the slot's entries
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const bool live = (first ^ uint32_t(j)) < cnt;   // rank 31 (the count's position) is never live: cnt <= 31
                bump((w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu, live);
            }
        }
    }
    R.done += R.n_chunks;
    return mine;
}

__global__ void __launch_bounds__(APPLY_THREADS, 1)
cms_apply_kernel(CmsSlots L, uint16_t* __restrict__ tab, int force_exact, int fresh) {
    // fresh: the table is logically zero (kmn_reset_table does not touch the 512 MiB; the first batch after it
    // overwrites every cell here instead of read-modify-writing it)
    extern __shared__ __align__(128) uint32_t s_cnt[];   // APPLY_WORDS words, then the ring, then the barriers
    __shared__ uint32_t s_tot[2];
    const uint32_t B = blockIdx.x;
    const bool fallback = L.state[0] != 0;              // cms_direct_kernel owns this batch
    const uint32_t n_vec = fallback ? 0u : L.running[1] * 4u;
    if (n_vec == 0) {
        if (fresh) {
            uint4* z4 = reinterpret_cast<uint4*>(tab + (size_t(B) << CMS_BUCKET_LOG2));
            for (uint32_t i = threadIdx.x; i < APPLY_WORDS / 4; i += APPLY_THREADS) z4[i] = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    if (!fresh)   // the merge at the end reads the bucket's 128 KiB of cells: one 128-byte line per thread, into L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tab + (size_t(B) << CMS_BUCKET_LOG2) + size_t(threadIdx.x) * 64) : "memory");
    ApplyRing R;
    R.ring = smem_u32(s_cnt + APPLY_WORDS);
    R.full = R.ring + APPLY_STAGES * APPLY_STAGE_BYTES;
    R.empty = R.full + 8 * APPLY_STAGES;
    R.src = reinterpret_cast<const uint4*>(L.ent) + size_t(B) * L.tiles_cap * (PART_SLOT / 8);
    R.n_vec = n_vec;
    R.n_chunks = (n_vec + APPLY_STAGE_VEC - 1) / APPLY_STAGE_VEC;
    R.origin = slot_origin(B & (CMS_BUCKETS_PER_ROW - 1));
    R.evict_first = !fresh;   // (measured: 0.728 -> 0.711 ms when merging, 0.650 -> 0.665 ms on a fresh table)
    R.issued = 0;
    R.done = 0;
    if (threadIdx.x == 0) {
        for (int st = 0; st < APPLY_STAGES; st++) {
            mbar_init(R.full + 8 * st, 1);
            mbar_init(R.empty + 8 * st, APPLY_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_tot[0] = s_tot[1] = 0;
    }
    uint4* s4 = reinterpret_cast<uint4*>(s_cnt);
    for (uint32_t i = threadIdx.x; i < APPLY_WORDS / 4; i += APPLY_THREADS) s4[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // No branch around the histogram atomic (half of the slot positions are dead): a dead position adds to the
    // lane's own dummy word instead (bank = lane; the dummy words collect garbage, nothing reads them).  ptxas turns
    // a predicated red.shared back into BSSY / BRA / ATOMS / BSYNC, and the branch per entry cost 11 % of the kernel
    // (0.76 -> 0.675 ms; its top stall reason was branch resolving).  Counter word c >> 1, low or high half by c & 1.
    const uint32_t a_cnt = smem_u32(s_cnt);
    const uint32_t off_dummy = R.full + 128u + 4u * (threadIdx.x & 31u) - a_cnt;
    const uint32_t mine = apply_stream(R, [&](uint32_t c, bool live) {
        const uint32_t off = live ? ((c << 1) & 0x3FFFCu) : off_dummy;
        const uint32_t val = (c & 1u) ? 0x10000u : 1u;
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a_cnt + off), "r"(val) : "memory");
    });
    {   // entries streamed by the CTA
        const uint32_t w = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31u) == 0 && w) atomicAdd(&s_tot[0], w);
    }
    __syncthreads();
    uint16_t* __restrict__ cells = tab + (size_t(B) << CMS_BUCKET_LOG2);   // B = row * 1024 + bucket
    {   // total of the histogram: falls short of the entries iff some counter wrapped or carried
        uint32_t sum = 0;
        for (uint32_t i = threadIdx.x; i < APPLY_WORDS / 4; i += APPLY_THREADS) {
            const uint4 d = s4[i];
            sum += (d.x & 0xFFFFu) + (d.x >> 16) + (d.y & 0xFFFFu) + (d.y >> 16) + (d.z & 0xFFFFu) + (d.z >> 16) + (d.w & 0xFFFFu) + (d.w >> 16);
        }
        sum = __reduce_add_sync(0xffffffffu, sum);
        if ((threadIdx.x & 31u) == 0 && sum) atomicAdd(&s_tot[1], sum);
    }
    __syncthreads();
    const bool exact = force_exact != 0 || s_tot[0] != s_tot[1];   // CTA-uniform
    if (!exact) {
        uint4* g4 = reinterpret_cast<uint4*>(cells);
        if (fresh) {   // no counter passed 65535: the counters are the cells
            for (uint32_t i = threadIdx.x; i < APPLY_WORDS / 4; i += APPLY_THREADS) g4[i] = s4[i];
            return;
        }
        // all of the thread's cell vectors in flight together (they were prefetched into L2 when the CTA started)
        constexpr uint32_t NV = APPLY_WORDS / 4 / APPLY_THREADS;   // 8
        uint4 g[NV];
#pragma unroll
        for (uint32_t u = 0; u < NV; u++) g[u] = g4[u * APPLY_THREADS + threadIdx.x];
#pragma unroll
        for (uint32_t u = 0; u < NV; u++) {
            const uint4 d = s4[u * APPLY_THREADS + threadIdx.x];
            if (!(d.x | d.y | d.z | d.w)) continue;
            g[u].x = sat_add_u16x2(g[u].x, d.x);
            g[u].y = sat_add_u16x2(g[u].y, d.y);
            g[u].z = sat_add_u16x2(g[u].z, d.z);
            g[u].w = sat_add_u16x2(g[u].w, d.w);
            g4[u * APPLY_THREADS + threadIdx.x] = g[u];
        }
        return;
    }
    // exact redo: u32 counters, the bucket's two halves one after the other
    for (uint32_t half = 0; half < 2; half++) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < APPLY_WORDS; i += APPLY_THREADS) s_cnt[i] = 0;
        __syncthreads();
        apply_stream(R, [&](uint32_t c, bool live) {
            if (live && (c >> 15) == half) atomicAdd(&s_cnt[c & 0x7FFFu], 1u);
        });
        __syncthreads();
        uint32_t* g32 = reinterpret_cast<uint32_t*>(cells + half * APPLY_WORDS);
        for (uint32_t i = threadIdx.x; i < APPLY_WORDS / 2; i += APPLY_THREADS) {
            const uint32_t d0 = s_cnt[2 * i], d1 = s_cnt[2 * i + 1];
            if (!(d0 | d1) && !fresh) continue;
            const uint32_t g = fresh ? 0u : g32[i];
            const uint32_t lo = min((g & 0xFFFFu) + min(d0, 0xFFFFu), 0xFFFFu);
            const uint32_t hi = min((g >> 16) + min(d1, 0xFFFFu), 0xFFFFu);
            g32[i] = lo | (hi << 16);
        }
    }
}

// slot overflows of the partition (after cms_apply_kernel in stream order)
__global__ void __launch_bounds__(256)
cms_spill_kernel(CmsSlots L, uint16_t* __restrict__ tab) {
    if (L.state[0]) return;
    const uint32_t n = min(L.state[1], L.spill_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        cms_cell_increment(tab, L.spill[i]);
}

// Exact slow path straight from codes[] + the lost map: runs only when the partition raised the fallback
// flag (or when forced, for the parity tests).  The per-species counts were already produced by the partition.
__global__ void __launch_bounds__(256)
cms_direct_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ lost2, uint32_t n_wtiles,
                  int k, uint64_t k_mask, uint16_t* __restrict__ tab, const uint32_t* __restrict__ state) {
    if (!state[0]) return;
    const uint32_t warps_per_grid = gridDim.x * (blockDim.x / 32);
    const uint32_t warp_id = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t t = warp_id; t < n_wtiles; t += warps_per_grid) {
        uint64_t kmer[CHUNK];
        const uint32_t valid = load_chunk_kmers<true>(codes, t * WARP_TILE, k, k_mask, kmer);
        const uint32_t new_bits = valid & ~both_lost_mask16(lost2[(size_t(t) * WARP_TILE) / CHUNK + lane]);
#pragma unroll
        for (int j = 0; j < CHUNK; j++) {
            if (!((new_bits >> j) & 1u)) continue;
#pragma unroll
            for (int r = 0; r < int(CMS_DEPTH); r++)
                cms_cell_increment(tab, (uint32_t(r) << CMS_WIDTH_LOG2) | cms_index(kmer[j], cms_seed(r)));
        }
    }
}

// Row sums + a position-sensitive 64-bit hash of the u16 table (sum of cell * mix64(cell index), wrapping):
// lets ranks / runs compare 512 MiB tables through 40 bytes (kmn_table_checksum).  While nothing saturates,
// row_sum[r] == number of AtomicCountMinSketch::increment calls (proba_filter.rs:146-168).
__global__ void __launch_bounds__(256)
table_checksum_kernel(const uint4* __restrict__ tab, unsigned long long* __restrict__ out /* [4] row sums, [4] hash */) {
    constexpr uint64_t VEC_PER_ROW = (uint64_t(1) << CMS_WIDTH_LOG2) / 8;
    unsigned long long sum[4] = {0, 0, 0, 0}, hash = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < CMS_CELLS / 8; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 v = __ldcs(tab + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        unsigned long long s = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t lo = w[j] & 0xFFFFu, hi = w[j] >> 16;
            s += lo + hi;
            if (lo) hash += lo * mix64(i * 8 + 2 * j + 0x51ull);
            if (hi) hash += hi * mix64(i * 8 + 2 * j + 1 + 0x51ull);
        }
        const int r = int(i / VEC_PER_ROW);
        sum[0] += r == 0 ? s : 0; sum[1] += r == 1 ? s : 0; sum[2] += r == 2 ? s : 0; sum[3] += r == 3 ? s : 0;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        for (int o = 16; o; o >>= 1) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
    }
    for (int o = 16; o; o >>= 1) hash += __shfl_xor_sync(0xffffffffu, hash, o);
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (sum[r]) atomicAdd(out + r, sum[r]);
        atomicAdd(out + 4, hash);
    }
}

// ---- multi-GPU -----------------------------------------------------------------------------------
// NCCL variant: widen the u16 table to u32 (summable with ncclSum), then clamp the sum back.
__global__ void __launch_bounds__(256)
cms_widen_kernel(const uint2* __restrict__ tab2, uint4* __restrict__ wide4, uint64_t n_vec) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint2 v = tab2[i];
        wide4[i] = make_uint4(v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16);
    }
}
__global__ void __launch_bounds__(256)
cms_narrow_kernel(const uint4* __restrict__ wide4, uint2* __restrict__ tab2, uint64_t n_vec) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 a = wide4[i];
        tab2[i] = make_uint2(min(a.x, 65535u) | (min(a.y, 65535u) << 16), min(a.z, 65535u) | (min(a.w, 65535u) << 16));
    }
}

// Fused exchange over NVLink peer memory: this rank sums its 1/world cell range of every rank's u16
// table (saturating), and stores the result into EVERY rank's exchanged table — reduce-scatter +
// snapshot + all-gather in one launch per rank (see kmn_exchange_reduce).
constexpr int MAX_PEERS = 16;
struct PeerTables {
    const uint4* part[MAX_PEERS];   // every rank's own table (peer-mapped; part[rank] is local)
    uint4* full[MAX_PEERS];         // every rank's exchanged table
    uint32_t* flags[MAX_PEERS];     // every rank's signal words: [0..16) "table of rank q complete", [16..32) "rank q done"
    int world, rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Bounded spin on a signal word: returns false when `limit_ns` passed (or the host raised the abort word)
// before *p reached `epoch`.  A rank that died, threw in pass 1 or never launched must not hang the other
// GPUs inside a cooperative kernel: the waiter gives up, records who it waited for and the kernel drains.
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct ExchangeGuard {
    uint32_t* err;                 // device word: 0 = fine, else 1 + rank waited for (| 0x100 epilogue, 0x200 host abort)
    const uint32_t* abort_flag;    // DEVICE word the host raises with a tiny copy on a side stream (kmn_exchange_abort).
                                   // (A host-mapped word here cost 0.6-2 ms per exchange: ~1200 spinning CTAs polling
                                   // it over PCIe delayed the very flags they were waiting for.)
    uint64_t limit_ns;
};
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// SYS: the word is written by another GPU (system scope); otherwise by a CTA of this GPU
template <bool SYS>
__device__ __forceinline__ bool wait_epoch(const uint32_t* p, uint32_t epoch, const ExchangeGuard& g, uint32_t who) {
    if ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) >= epoch) return true;
    const uint64_t t0 = global_ns();
    for (uint32_t it = 1;; it++) {
        if ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) >= epoch) return true;
        if (!SYS) __nanosleep(64);   // ~1200 CTAs poll the local go word: keep them off the memory system
        if ((it & 255u) == 0) {
            if (ld_acquire_gpu(g.abort_flag)) { atomicCAS(g.err, 0u, 0x200u | (1u + who)); return false; }
            if (global_ns() - t0 > g.limit_ns) { atomicCAS(g.err, 0u, 1u + who); return false; }
        }
    }
}
// prologue shared by both exchange kernels: CTA 0 publishes "my table is complete" to every peer, waits for
// theirs and releases the other CTAs; returns false (CTA-uniform) when the exchange must be skipped.
__device__ __forceinline__ bool exchange_prologue(uint32_t* const* flags, int world, int rank, uint32_t epoch,
                                                  uint32_t* local_sync, const ExchangeGuard& g) {
    if (blockIdx.x == 0) {
        if (int(threadIdx.x) < world) {
            const int q = (rank + int(threadIdx.x)) % world;
            st_release_sys(flags[q] + rank, epoch);
            wait_epoch<true>(flags[rank] + q, epoch, g, uint32_t(q));
        }
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(local_sync, epoch);
    } else {
        if (threadIdx.x == 0) {
            ExchangeGuard g2 = g;
            g2.limit_ns = 2 * g.limit_ns;   // CTA 0 always releases, at the latest when its own waits give up
            wait_epoch<false>(local_sync, epoch, g2, uint32_t(rank));
        }
        __syncthreads();
    }
    return ld_acquire_gpu(g.err) == 0;
}
// epilogue: the last CTA of this rank publishes "done" and waits for every peer's "done"
__device__ __forceinline__ void exchange_epilogue(uint32_t* const* flags, int world, int rank, uint32_t epoch,
                                                  uint32_t* local_sync, const ExchangeGuard& g, bool moved) {
    __threadfence_system();   // this thread's peer stores are visible system-wide
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(local_sync + 1, 1u);
        if (ticket == gridDim.x - 1) {   // last CTA of this rank
            local_sync[1] = 0;
            __threadfence_system();
            if (!moved) return;          // a failed prologue: nothing was read or written, peers time out on their own
            for (int p = 0; p < world; p++) st_release_sys(flags[(rank + p) % world] + MAX_PEERS + rank, epoch);
            for (int p = 0; p < world; p++) {
                ExchangeGuard g2 = g;
                g2.err = g.err;
                if (!wait_epoch<true>(flags[rank] + MAX_PEERS + p, epoch, g2, 0x100u | uint32_t(p))) break;
            }
        }
    }
}

// WORLD > 0: compile-time rank count; U vectors per thread and iteration, so that WORLD x U 16-byte loads
// (all peers) are in flight per thread: NVLink round trips are ~2-3 us and one load at a time leaves the
// links idle.  WORLD == 0: generic loop for other rank counts.
//
// The ranks synchronise through signal words in peer memory instead of host barriers (cooperative launch:
// every CTA is resident).  Prologue: CTA 0 publishes "my table is complete" (epoch) to every peer and waits
// for theirs, then releases the other CTAs.  Epilogue: the last CTA to finish publishes "done" and waits
// for every peer's "done": when the kernel ends, nobody still reads this rank's table and every peer has
// written its slice into this rank's exchanged table.
template <int WORLD, int U>
__global__ void __launch_bounds__(256)
exchange_reduce_kernel(PeerTables pt, uint64_t vec_begin, uint64_t vec_end, uint32_t epoch, uint32_t* local_sync,
                       ExchangeGuard guard) {
    constexpr int W = WORLD > 0 ? WORLD : 1;
    const int world = pt.world;
    const bool go = exchange_prologue(pt.flags, world, pt.rank, epoch, local_sync, guard);
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i0 = vec_begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; go && i0 < vec_end; i0 += stride * U) {
        if (WORLD > 0) {
            uint4 a[U][W];
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int p = 0; p < W; p++)
                    if (i0 + u * stride < vec_end) a[u][p] = __ldcs(pt.part[(pt.rank + p) % W] + i0 + u * stride);   // staggered across ranks
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (i0 + u * stride >= vec_end) break;
                uint4 s = a[u][0];
#pragma unroll
                for (int p = 1; p < W; p++) {
                    s.x = sat_add_u16x2(s.x, a[u][p].x); s.y = sat_add_u16x2(s.y, a[u][p].y);
                    s.z = sat_add_u16x2(s.z, a[u][p].z); s.w = sat_add_u16x2(s.w, a[u][p].w);
                }
#pragma unroll
                for (int p = 0; p < W; p++) pt.full[(pt.rank + p) % W][i0 + u * stride] = s;
            }
        } else {
            for (int u = 0; u < U; u++) {
                const uint64_t i = i0 + u * stride;
                if (i >= vec_end) break;
                uint4 s = pt.part[pt.rank][i];
#pragma unroll 1
                for (int p = 1; p < world; p++) {
                    const uint4 a = pt.part[(pt.rank + p) % world][i];
                    s.x = sat_add_u16x2(s.x, a.x); s.y = sat_add_u16x2(s.y, a.y);
                    s.z = sat_add_u16x2(s.z, a.z); s.w = sat_add_u16x2(s.w, a.w);
                }
                for (int p = 0; p < world; p++) pt.full[(pt.rank + p) % world][i] = s;
            }
        }
    }
    exchange_epilogue(pt.flags, world, pt.rank, epoch, local_sync, guard, go);
}

// ---- byte-plane exchange ---------------------------------------------------------------------------
// The exchange is NVLink-bound (0.94 GB per link direction and GPU at 8 ranks with u16 cells), and almost
// every cell of a partial table — and most cells of the final one — fits a byte.  So the ranks exchange
// byte planes: lo = cell & 0xFF (always), hi = cell >> 8 only for the 16-cell blocks whose bit is set in a
// bitmap (1 bit per block).  planes_split_kernel builds the planes of this rank's table,
// exchange_planes_kernel reduces this rank's slice from every peer's planes and scatters the planes of the
// result to every peer, planes_join_kernel rebuilds the u16 table.  Exact: nothing is approximated, a block
// with a large cell simply also moves its hi bytes.
constexpr uint32_t PLANE_BLOCK = 16;                                  // cells per bitmap bit
constexpr uint64_t PLANE_BLOCKS = CMS_CELLS / PLANE_BLOCK;            // 2^24
constexpr uint64_t PLANE_BM_WORDS = PLANE_BLOCKS / 32;

struct PeerPlanes {
    const uint4* p_lo[MAX_PEERS];      // partial-table planes of every rank (peer-mapped)
    const uint4* p_hi[MAX_PEERS];
    const uint32_t* p_bm[MAX_PEERS];
    uint4* f_lo[MAX_PEERS];            // planes of the exchanged table in every rank
    uint4* f_hi[MAX_PEERS];
    uint32_t* f_bm[MAX_PEERS];
    uint32_t* flags[MAX_PEERS];
    int world, rank;
};

// 16 u16 cells (two uint4) -> 16 lo bytes, 16 hi bytes
__device__ __forceinline__ void split_block(const uint4& a, const uint4& b, uint4& lo, uint4& hi) {
    lo = make_uint4(__byte_perm(a.x, a.y, 0x6420), __byte_perm(a.z, a.w, 0x6420), __byte_perm(b.x, b.y, 0x6420), __byte_perm(b.z, b.w, 0x6420));
    hi = make_uint4(__byte_perm(a.x, a.y, 0x7531), __byte_perm(a.z, a.w, 0x7531), __byte_perm(b.x, b.y, 0x7531), __byte_perm(b.z, b.w, 0x7531));
}
__device__ __forceinline__ void join_block(const uint4& lo, const uint4& hi, uint4& a, uint4& b) {
    a = make_uint4(__byte_perm(lo.x, hi.x, 0x5140), __byte_perm(lo.x, hi.x, 0x7362), __byte_perm(lo.y, hi.y, 0x5140), __byte_perm(lo.y, hi.y, 0x7362));
    b = make_uint4(__byte_perm(lo.z, hi.z, 0x5140), __byte_perm(lo.z, hi.z, 0x7362), __byte_perm(lo.w, hi.w, 0x5140), __byte_perm(lo.w, hi.w, 0x7362));
}

__global__ void __launch_bounds__(256)
planes_split_kernel(const uint4* __restrict__ tab, uint4* __restrict__ lo, uint4* __restrict__ hi, uint32_t* __restrict__ bm) {
    const unsigned lane = threadIdx.x & 31u;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < PLANE_BLOCKS; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 a = __ldcs(tab + 2 * i), b = __ldcs(tab + 2 * i + 1);
        uint4 l, h;
        split_block(a, b, l, h);
        const bool any_hi = (h.x | h.y | h.z | h.w) != 0;
        lo[i] = l;
        if (any_hi) hi[i] = h;
        const unsigned m = __ballot_sync(0xffffffffu, any_hi);   // i is warp-aligned: 32 consecutive blocks per warp
        if (lane == 0) bm[i >> 5] = m;
    }
}

__global__ void __launch_bounds__(256)
planes_join_kernel(const uint4* __restrict__ lo, const uint4* __restrict__ hi, const uint32_t* __restrict__ bm, uint4* __restrict__ tab) {
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < PLANE_BLOCKS; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 l = __ldcs(lo + i);
        uint4 h = make_uint4(0, 0, 0, 0);
        if ((bm[i >> 5] >> (i & 31u)) & 1u) h = __ldcs(hi + i);
        uint4 a, b;
        join_block(l, h, a, b);
        tab[2 * i] = a;
        tab[2 * i + 1] = b;
    }
}

template <int WORLD>
__global__ void __launch_bounds__(256)
exchange_planes_kernel(PeerPlanes pt, uint64_t blk_begin, uint64_t blk_end, uint32_t epoch, uint32_t* local_sync,
                       ExchangeGuard guard) {
    constexpr int W = WORLD;
    // every rank's planes are complete before anybody loads them
    const bool go = exchange_prologue(pt.flags, W, pt.rank, epoch, local_sync, guard);
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    // blk_begin, blk_end and stride are multiples of 32: a warp always owns 32 consecutive blocks (one bitmap word)
    for (uint64_t i = blk_begin + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; go && i < blk_end; i += stride) {
        uint4 l[W];
        uint32_t bits = 0;
#pragma unroll
        for (int p = 0; p < W; p++) {
            const int q = (pt.rank + p) % W;   // staggered across ranks
            l[p] = __ldcs(pt.p_lo[q] + i);
            bits |= ((__ldcs(pt.p_bm[q] + (i >> 5)) >> lane) & 1u) << p;
        }
        uint4 sa = make_uint4(0, 0, 0, 0), sb = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int p = 0; p < W; p++) {
            uint4 h = make_uint4(0, 0, 0, 0);
            if ((bits >> p) & 1u) h = __ldcs(pt.p_hi[(pt.rank + p) % W] + i);   // rare for partial tables
            uint4 a, b;
            join_block(l[p], h, a, b);
            sa.x = sat_add_u16x2(sa.x, a.x); sa.y = sat_add_u16x2(sa.y, a.y); sa.z = sat_add_u16x2(sa.z, a.z); sa.w = sat_add_u16x2(sa.w, a.w);
            sb.x = sat_add_u16x2(sb.x, b.x); sb.y = sat_add_u16x2(sb.y, b.y); sb.z = sat_add_u16x2(sb.z, b.z); sb.w = sat_add_u16x2(sb.w, b.w);
        }
        uint4 ol, oh;
        split_block(sa, sb, ol, oh);
        const bool any_hi = (oh.x | oh.y | oh.z | oh.w) != 0;
        const unsigned m = __ballot_sync(0xffffffffu, any_hi);
#pragma unroll
        for (int p = 0; p < W; p++) {
            const int q = (pt.rank + p) % W;
            pt.f_lo[q][i] = ol;
            if (any_hi) pt.f_hi[q][i] = oh;
            if (lane == 0) pt.f_bm[q][i >> 5] = m;
        }
    }
    exchange_epilogue(pt.flags, W, pt.rank, epoch, local_sync, guard, go);
}

}  // namespace kmn

// pair_kernels.cuh — K6: bit-sliced pairwise valid/mismatch counts for --nj.
//
// Replaces the byte-compare loop of lg_f81_distance (/root/reference/src/phylo.rs:84-95) for every
// pair i<j of compute_distance_matrix (phylo.rs:114-141):
//     valid    = #columns with a<20 && b<20
//     mismatch = #valid columns with a != b
// Codes 0..20 are sliced into 5 bit planes + a validity plane per 32-column word, so one pair-word
// costs 7 LOP3 + 2 POPC.  Integer only — no tensor cores (not a floating-point contraction).
#pragma once
#include <cuda_pipeline.h>

#include "kmn_device.cuh"

namespace kmn {

constexpr int PAIR_PLANES = 8;     // 5 code planes + valid + 2 pad -> 32 B per (row, word)
constexpr int PAIR_TILE = 64;      // pairs tile is 64 x 64 rows
constexpr int PAIR_WB = 8;         // 32-column words per smem stage
constexpr int PAIR_ROW_STRIDE = PAIR_WB * PAIR_PLANES + 4;  // u32; +4 keeps LDS.128 conflict-free
constexpr int PAIR_THREADS = 256;
constexpr size_t PAIR_COUNT_SMEM = size_t(4) * PAIR_TILE * PAIR_ROW_STRIDE * sizeof(uint32_t);   // two stages x two operands

// planes[(i * W + w) * 8 + p]; columns >= L and rows >= n are "invalid" (valid plane 0).
__global__ void __launch_bounds__(256)
pair_pack_kernel(const uint8_t* __restrict__ mapped, uint32_t n, uint64_t L, uint32_t n_pad, uint32_t W, uint32_t row0,
                 uint32_t* __restrict__ planes) {
    const uint64_t idx = uint64_t(row0) * W + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;   // rows below row0 are not needed
    if (idx >= uint64_t(n_pad) * W) return;
    const uint32_t i = uint32_t(idx / W), w = uint32_t(idx % W);
    uint32_t p[6] = {0, 0, 0, 0, 0, 0};
    if (i < n) {
        const uint8_t* row = mapped + uint64_t(i) * L;
        const uint64_t col0 = uint64_t(w) * 32;
#pragma unroll 4
        for (int b = 0; b < 32; b++) {
            const uint64_t col = col0 + b;
            const uint32_t code = col < L ? row[col] : 20u;
            const uint32_t ok = code < 20u;
            p[0] |= (code & 1u) << b;
            p[1] |= ((code >> 1) & 1u) << b;
            p[2] |= ((code >> 2) & 1u) << b;
            p[3] |= ((code >> 3) & 1u) << b;
            p[4] |= ((code >> 4) & 1u) << b;
            p[5] |= ok << b;
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(planes + idx * PAIR_PLANES);
    dst[0] = make_uint4(p[0], p[1], p[2], p[3]);
    dst[1] = make_uint4(p[4], p[5], 0u, 0u);
}

// One CTA per 64x64 tile (ti <= tj) of the pair matrix; each thread owns a 4x4 micro-tile.  The operand
// stages (64 rows x 8 words x 8 planes each) are double-buffered with cp.async so that the next stage's
// global loads fly while the current one is being counted.
__device__ __forceinline__ void pair_stage_load(uint32_t* sA, uint32_t* sB, const uint32_t* __restrict__ planes, uint32_t ti,
                                                uint32_t tj, uint32_t W, uint32_t w0) {
    for (uint32_t e = threadIdx.x; e < PAIR_TILE * PAIR_WB * 2; e += PAIR_THREADS) {
        const uint32_t row = e / (PAIR_WB * 2), q = e % (PAIR_WB * 2);  // q: (word, half)
        const uint32_t w = w0 + (q >> 1);
        uint32_t* da = &sA[row * PAIR_ROW_STRIDE + q * 4];
        uint32_t* db = &sB[row * PAIR_ROW_STRIDE + q * 4];
        if (w < W) {
            __pipeline_memcpy_async(da, reinterpret_cast<const uint4*>(planes + (uint64_t(ti * PAIR_TILE + row) * W + w) * PAIR_PLANES) + (q & 1), 16);
            __pipeline_memcpy_async(db, reinterpret_cast<const uint4*>(planes + (uint64_t(tj * PAIR_TILE + row) * W + w) * PAIR_PLANES) + (q & 1), 16);
        } else {
            *reinterpret_cast<uint4*>(da) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(db) = make_uint4(0, 0, 0, 0);
        }
    }
}

__global__ void __launch_bounds__(PAIR_THREADS, 3)
pair_count_kernel(const uint32_t* __restrict__ planes, uint32_t n, uint32_t W, uint32_t n_tiles, uint32_t tile_row0,
                  uint32_t* __restrict__ valid, uint32_t* __restrict__ mismatch) {
    extern __shared__ __align__(16) uint32_t pair_smem[];   // [2 buffers][A, B][PAIR_TILE * PAIR_ROW_STRIDE]
    constexpr uint32_t OPER = PAIR_TILE * PAIR_ROW_STRIDE;
    // linear block index -> (ti, tj), ti <= tj; the launch covers tile rows tile_row0.. (a rank's row block
    // when the matrix is sharded), and the outputs hold rows tile_row0 * PAIR_TILE.. only
    uint32_t rem = blockIdx.x, ti = tile_row0;
    while (rem >= n_tiles - ti) { rem -= n_tiles - ti; ti++; }
    const uint32_t tj = ti + rem;
    const uint32_t tx = threadIdx.x & 15u, ty = threadIdx.x >> 4;
    uint32_t cv[4][4], cm[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) { cv[a][b] = 0; cm[a][b] = 0; }

    pair_stage_load(pair_smem, pair_smem + OPER, planes, ti, tj, W, 0);
    __pipeline_commit();
    uint32_t buf = 0;
    for (uint32_t w0 = 0; w0 < W; w0 += PAIR_WB, buf ^= 1u) {
        const uint32_t* sA = pair_smem + buf * 2 * OPER;
        const uint32_t* sB = sA + OPER;
        if (w0 + PAIR_WB < W) {   // CTA-uniform
            uint32_t* nA = pair_smem + (buf ^ 1u) * 2 * OPER;
            pair_stage_load(nA, nA + OPER, planes, ti, tj, W, w0 + PAIR_WB);
            __pipeline_commit();
            __pipeline_wait_prior(1);
        } else {
            __pipeline_wait_prior(0);
        }
        __syncthreads();
#pragma unroll 2
        for (int w = 0; w < PAIR_WB; w++) {
            uint4 a03[4], b03[4];
            uint2 a45[4], b45[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t* pa = &sA[(ty * 4 + r) * PAIR_ROW_STRIDE + w * PAIR_PLANES];
                const uint32_t* pb = &sB[(r * 16 + tx) * PAIR_ROW_STRIDE + w * PAIR_PLANES];
                a03[r] = *reinterpret_cast<const uint4*>(pa);
                a45[r] = *reinterpret_cast<const uint2*>(pa + 4);
                b03[r] = *reinterpret_cast<const uint4*>(pb);
                b45[r] = *reinterpret_cast<const uint2*>(pb + 4);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t d = a03[a].x ^ b03[b].x;
                    d |= a03[a].y ^ b03[b].y;
                    d |= a03[a].z ^ b03[b].z;
                    d |= a03[a].w ^ b03[b].w;
                    d |= a45[a].x ^ b45[b].x;
                    const uint32_t v = a45[a].y & b45[b].y;
                    cv[a][b] += __popc(v);
                    cm[a][b] += __popc(d & v);
                }
        }
        __syncthreads();   // everybody is done with this buffer before the stage after next overwrites it
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t i = ti * PAIR_TILE + ty * 4 + a;
            const uint32_t j = tj * PAIR_TILE + b * 16 + tx;
            if (i < n && j < n && i < j) {
                const uint64_t o = uint64_t(i - tile_row0 * PAIR_TILE) * n + j;
                valid[o] = cv[a][b];
                mismatch[o] = cm[a][b];
            }
        }
}

}  // namespace kmn

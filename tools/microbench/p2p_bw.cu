// p2p_bw.cu — what SM-issued peer accesses over NVLink / NVSwitch achieve on this box (2 GPUs, one process):
// 16-byte loads from the peer, 16-byte stores to the peer, and both at once (the exchange kernel's pattern),
// for several grid sizes and per-thread unroll depths.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U>
__global__ void read_kernel(const uint4* __restrict__ src, uint4* __restrict__ sink, size_t n) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) v[u] = __ldcs(src + i + u * stride); else v[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; u++) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = acc;
}
template <int U>
__global__ void write_kernel(uint4* __restrict__ dst, size_t n) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride * U)
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) dst[i + u * stride] = v;
}
template <int U>
__global__ void rw_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {   // read peer, write peer
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) dst[i + u * stride] = v[u];
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    const size_t bytes = size_t(512) << 20, n = bytes / 16;
    uint4 *a0, *b0, *a1, *b1;
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&a0, bytes)); CK(cudaMalloc(&b0, bytes)); CK(cudaMemset(a0, 1, bytes));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&a1, bytes)); CK(cudaMalloc(&b1, bytes)); CK(cudaMemset(a1, 2, bytes));
    cudaStream_t s[2]; cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaStreamCreate(&s[d])); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d])); }
    auto run = [&](const char* what, int grid, int u, bool both, auto launch) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            for (int d = 0; d < (both ? 2 : 1); d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], s[d])); launch(d, grid); CK(cudaEventRecord(e1[d], s[d])); }
            float worst = 0;
            for (int d = 0; d < (both ? 2 : 1); d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(s[d])); float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); worst = ms > worst ? ms : worst; }
            best = worst < best ? worst : best;
        }
        printf("{\"what\": \"%s\", \"gpus_active\": %d, \"grid\": %d, \"unroll\": %d, \"ms\": %.3f, \"GBps_per_direction\": %.1f}\n", what, both ? 2 : 1, grid, u, best,
               bytes / (best * 1e-3) / 1e9);
    };
    for (int grid : {148 * 2, 148 * 8, 148 * 16}) {
        run("peer read (GPU0 reads GPU1)", grid, 1, false, [&](int d, int g) { read_kernel<1><<<g, 256, 0, s[d]>>>(d ? a0 : a1, d ? b1 : b0, n); });
        run("peer read", grid, 4, false, [&](int d, int g) { read_kernel<4><<<g, 256, 0, s[d]>>>(d ? a0 : a1, d ? b1 : b0, n); });
        run("peer write (GPU0 writes GPU1)", grid, 1, false, [&](int d, int g) { write_kernel<1><<<g, 256, 0, s[d]>>>(d ? b0 : b1, n); });
        run("peer read, both GPUs at once", grid, 4, true, [&](int d, int g) { read_kernel<4><<<g, 256, 0, s[d]>>>(d ? a0 : a1, d ? b1 : b0, n); });
        run("peer write, both GPUs at once", grid, 1, true, [&](int d, int g) { write_kernel<1><<<g, 256, 0, s[d]>>>(d ? b0 : b1, n); });
        run("peer read + peer write in one kernel, both GPUs", grid, 1, true, [&](int d, int g) { rw_kernel<1><<<g, 256, 0, s[d]>>>(d ? a0 : a1, d ? b0 : b1, n); });
        run("peer read + peer write in one kernel, both GPUs", grid, 4, true, [&](int d, int g) { rw_kernel<4><<<g, 256, 0, s[d]>>>(d ? a0 : a1, d ? b0 : b1, n); });
    }
    // local copy for scale
    CK(cudaSetDevice(0));
    run("local read + local write (GPU0)", 148 * 8, 4, false, [&](int d, int g) { rw_kernel<4><<<g, 256, 0, s[d]>>>(a0, b0, n); });
    return 0;
}

// l2_atomics.cu — B200 microbenchmark: random-address RED / ATOM / LDG throughput against region size
// (L2-resident vs HBM), and shared-memory atomics.  Sets the floor for the Bloom / CMS kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_atomics l2_atomics.cu && ./l2_atomics
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; return x ^ (x >> 31);
}

template <int OP, int U>   // 0 RED.add.u32, 1 ATOM.or.b32 (result used), 2 LDG.u32, 3 RED.or.b32
__global__ void __launch_bounds__(256) rnd_kernel(uint32_t* tab, uint32_t mask, uint64_t n_per_thread, uint32_t* sink) {
    const uint64_t tid = blockIdx.x * 256ull + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += U) {
        uint32_t idx[U];
#pragma unroll
        for (int u = 0; u < U; u++) idx[u] = uint32_t(mix64((tid * n_per_thread + i + u) ^ 0x9e3779b97f4a7c15ull)) & mask;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (OP == 0) atomicAdd(tab + idx[u], 1u);
            if (OP == 1) acc += atomicOr(tab + idx[u], 1u << ((idx[u] >> 7) & 31));
            if (OP == 2) acc += __ldg(tab + idx[u]);
            if (OP == 3) atomicOr(tab + idx[u], 1u << (idx[u] & 31));
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// stream of precomputed indices (uint4 loads, evict-first) -> RED
__global__ void __launch_bounds__(256) stream_red_kernel(const uint4* idx, uint64_t n_vec, uint32_t* tab) {
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n_vec; i += gridDim.x * 256ull) {
        const uint4 v = __ldcs(idx + i);
        atomicAdd(tab + v.x, 1u); atomicAdd(tab + v.y, 1u); atomicAdd(tab + v.z, 1u); atomicAdd(tab + v.w, 1u);
    }
}
__global__ void fill_idx(uint32_t* idx, uint64_t n, uint32_t mask) {
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) idx[i] = uint32_t(mix64(i ^ 0x1234567ull)) & mask;
}

template <int OP>  // 0 atomicAdd, 1 atomicOr with result, 2 plain LDS
__global__ void __launch_bounds__(1024) smem_kernel(uint32_t words_mask, uint64_t n_per_thread, uint32_t* sink) {
    extern __shared__ uint32_t s[];
    for (uint32_t i = threadIdx.x; i <= words_mask; i += 1024) s[i] = 0;
    __syncthreads();
    const uint64_t tid = blockIdx.x * 1024ull + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t i = 0; i < n_per_thread; i += 8) {
        uint32_t idx[8];
#pragma unroll
        for (int u = 0; u < 8; u++) idx[u] = uint32_t(mix64((tid * n_per_thread + i + u) ^ 0x9e3779b97f4a7c15ull));
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0) atomicAdd(s + (idx[u] & words_mask), 1u);
            if (OP == 1) acc += atomicOr(s + ((idx[u] >> 5) & words_mask), 1u << (idx[u] & 31));
            if (OP == 2) acc += s[idx[u] & words_mask];
        }
    }
    __syncthreads();
    if (acc == 0x12345678u) *sink = acc + s[threadIdx.x];
}

template <typename F> float time_ms(F&& f, int reps = 3) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d L2 %d MiB persistingL2Max %d MiB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.persistingL2CacheMaxSize >> 20);
    uint32_t *tab, *sink; CK(cudaMalloc(&tab, 1ull << 30)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(tab, 0, 1ull << 30));
    const int grid = p.multiProcessorCount * 8;
    const uint64_t per_thread = 256;   // ops per thread
    const double n_ops = double(grid) * 256 * per_thread;
    const char* names[4] = {"RED.add.u32", "ATOM.or (ret)", "LDG.u32", "RED.or.b32"};
    for (int mib : {4, 16, 32, 48, 64, 96, 128, 256, 1024}) {
        const uint32_t mask = uint32_t((uint64_t(mib) << 20) / 4 - 1);
        if (((uint64_t(mib) << 20) / 4) & mask) {   // not a power of two: use modulo-free by masking to next lower pow2 + offset skip
        }
        uint32_t m2 = 1; while (m2 * 2ull * 4 <= (uint64_t(mib) << 20)) m2 *= 2; m2 -= 1;
        float t0 = time_ms([&] { rnd_kernel<0, 16><<<grid, 256>>>(tab, m2, per_thread, sink); });
        float t1 = time_ms([&] { rnd_kernel<1, 16><<<grid, 256>>>(tab, m2, per_thread, sink); });
        float t2 = time_ms([&] { rnd_kernel<2, 16><<<grid, 256>>>(tab, m2, per_thread, sink); });
        float t3 = time_ms([&] { rnd_kernel<3, 16><<<grid, 256>>>(tab, m2, per_thread, sink); });
        printf("region %4d MiB (pow2 words %u): %s %.1f Gop/s | %s %.1f | %s %.1f | %s %.1f\n", mib, m2 + 1, names[0], n_ops / t0 / 1e6,
               names[1], n_ops / t1 / 1e6, names[2], n_ops / t2 / 1e6, names[3], n_ops / t3 / 1e6);
    }
    // unroll sensitivity at 64 MiB
    {
        const uint32_t m2 = (64u << 20) / 4 - 1;
        float a = time_ms([&] { rnd_kernel<0, 4><<<grid, 256>>>(tab, m2, per_thread, sink); });
        float b = time_ms([&] { rnd_kernel<0, 8><<<grid, 256>>>(tab, m2, per_thread, sink); });
        float c = time_ms([&] { rnd_kernel<1, 4><<<grid, 256>>>(tab, m2, per_thread, sink); });
        float d = time_ms([&] { rnd_kernel<1, 8><<<grid, 256>>>(tab, m2, per_thread, sink); });
        printf("64 MiB unroll: RED U4 %.1f U8 %.1f | ATOM U4 %.1f U8 %.1f Gop/s\n", n_ops / a / 1e6, n_ops / b / 1e6, n_ops / c / 1e6, n_ops / d / 1e6);
        for (int g : {1, 2, 4, 16}) {
            const int gr = p.multiProcessorCount * g;
            const double n2 = double(gr) * 256 * per_thread;
            float e = time_ms([&] { rnd_kernel<0, 16><<<gr, 256>>>(tab, m2, per_thread, sink); });
            float f = time_ms([&] { rnd_kernel<1, 16><<<gr, 256>>>(tab, m2, per_thread, sink); });
            printf("64 MiB ctas/SM %2d: RED %.1f ATOM %.1f Gop/s\n", g, n2 / e / 1e6, n2 / f / 1e6);
        }
    }
    // index stream -> RED (the cms sweep shape), 64 MiB slice and full 256 MiB row
    {
        const uint64_t n = 1ull << 27;
        uint32_t* idx; CK(cudaMalloc(&idx, n * 4));
        for (int mib : {64, 256}) {
            const uint32_t m2 = uint32_t((uint64_t(mib) << 20) / 4 - 1);
            fill_idx<<<grid, 256>>>(idx, n, m2);
            float t = time_ms([&] { stream_red_kernel<<<grid, 256>>>((const uint4*)idx, n / 4, tab); });
            printf("stream->RED %d MiB: %.1f Gop/s (%.3f ms for %llu)\n", mib, n / t / 1e6, t, (unsigned long long)n);
        }
        cudaFree(idx);
    }
    // shared memory
    for (int kib : {32, 128}) {
        const uint32_t wm = uint32_t(kib * 1024 / 4 - 1);
        CK(cudaFuncSetAttribute(smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kib * 1024));
        CK(cudaFuncSetAttribute(smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kib * 1024));
        CK(cudaFuncSetAttribute(smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kib * 1024));
        const int g = p.multiProcessorCount;
        const double n2 = double(g) * 1024 * 1024;
        float a = time_ms([&] { smem_kernel<0><<<g, 1024, kib * 1024>>>(wm, 1024, sink); });
        float b = time_ms([&] { smem_kernel<1><<<g, 1024, kib * 1024>>>(wm, 1024, sink); });
        float c = time_ms([&] { smem_kernel<2><<<g, 1024, kib * 1024>>>(wm, 1024, sink); });
        printf("smem %3d KiB: atomicAdd %.1f Gop/s | atomicOr(ret) %.1f | LDS %.1f\n", kib, n2 / a / 1e6, n2 / b / 1e6, n2 / c / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}

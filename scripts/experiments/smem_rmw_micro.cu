// Micro-benchmark: random read-modify-write on a 200 KB shared-memory accumulator array, as apply_sparse_kernel's walk does it.
//   mode 0: red.shared.add.u32 (what the kernel does)      mode 1: ld.shared + add + st.shared (warp-private slice, no atomic)
//   mode 2: like 0 but conflict-free addresses (lane-distinct banks)   mode 3: like 1, conflict-free
// 512 threads per CTA, one CTA per SM, indices from a per-lane LCG (no global loads in the loop).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NACC = 51200;
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, int iters) {
    extern __shared__ uint32_t acc[];
    for (int i = threadIdx.x; i < NACC; i += 512) acc[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const uint32_t base = uint32_t(__cvta_generic_to_shared(acc));
    const int slice = NACC / 16;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            s = s * 1664525u + 1013904223u;
            uint32_t idx;
            if (MODE == 0) idx = (s >> 8) % NACC;
            else if (MODE == 1) idx = warp * slice + (s >> 8) % slice;
            else if (MODE == 2) idx = (((s >> 8) % (NACC / 32)) * 32) + lane;
            else idx = warp * slice + (((s >> 8) % (slice / 32)) * 32) + lane;
            const uint32_t a = base + idx * 4, v = (s & 7u) + 1u;
            if (MODE == 0 || MODE == 2) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
            else {
                uint32_t x;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a) : "memory");
                x += v;
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
            }
        }
    }
    __syncthreads();
    uint32_t t = 0;
    for (int i = threadIdx.x; i < NACC; i += 512) t += acc[i];
    atomicAdd(out, t);
}
template <int MODE> void run(const char *name) {
    uint32_t *d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, NACC * 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    k<MODE><<<sms, 512, NACC * 4>>>(d, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, 512, NACC * 4>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = double(sms) * 512 * 16 * iters;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-34s %8.3f ms  %7.2f G updates/s  %.2f lane-updates/clk/SM (at %d MHz)  err=%s\n", name, ms, ops / ms / 1e6, ops / sms / (ms * 1e-3 * clk * 1e3), clk / 1000,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("red.shared random");
    run<1>("ld+add+st warp-slice random");
    run<2>("red.shared conflict-free");
    run<3>("ld+add+st warp-slice conflict-free");
    return 0;
}

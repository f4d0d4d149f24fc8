// skm_bench.cu — micro-benchmark kernel behind the fp32-FMA peak that the sparse scoring roofline is quoted against
// ("min(HBM, fp32-FMA)": MEASURED_PEAKS.json has HBM and bf16 only).  Not on the product path.
#include "skm_common.cuh"

namespace skm {

// 16 independent FMA chains per thread (the FMA pipe has 4-cycle latency and issues one warp instruction per cycle
// and SM sub-partition: 8+ warps x 16 chains keep it full), `iters` rounds of 16 FMAs each.
__global__ void __launch_bounds__(256) bench_fma_f32_kernel(int64_t iters, float seed, float *__restrict__ out) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = seed + float(threadIdx.x + j);
    const float m = 1.000001f, c = 0.5f;
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

}  // namespace skm

extern "C" int skm_bench_fma_f32(int64_t iters, int blocks, float *d_out, double *flops_out, skm_stream_t stream) {
    using namespace skm;
    if (iters <= 0 || blocks <= 0 || !d_out) { set_error("skm_bench_fma_f32: bad arguments"); return SKM_ERR_INVALID; }
    bench_fma_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, d_out);
    SKM_LAUNCH_CHECK("bench_fma_f32_kernel");
    if (flops_out) *flops_out = 2.0 * 16.0 * double(iters) * 256.0 * double(blocks);
    return SKM_OK;
}

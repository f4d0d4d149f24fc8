// skm_learn.cu — kernel (c): learn-mode reduction of per-annotation k-mer counts.
//
// Replaces Library.generate_kmer_counts / filter_and_construct /
// _process_annotation_counts (learn.smk:306-326, 359-408): per-sequence count
// lists over the whole basis, summed list-by-list into annotation rows.
// Here the per-sequence counts are never materialised: sequences arrive grouped
// by annotation (d_order), a CTA histograms every window of a run of
// same-annotation sequences into ONE shared-memory row and adds that row to
// M[a, :] once per run (64-bit integer adds of the non-zero entries only —
// integer, hence order-independent and bit-reproducible).  Unannotated
// sequences go to an extra "rest" row so that Totals (learn.smk:380, all
// sequences) is the column sum of all rows.
// Algorithmic bytes: R + 8(N+1) + 4N (annotation ids) + 8 (A+1) K.
#include "skm_common.cuh"

namespace skm {

template <typename CodeT, int NW>
__global__ void __launch_bounds__(256) learn_dense_kernel(const uint8_t *__restrict__ res, int64_t nres,
                                                          const int64_t *__restrict__ off, int64_t nseq,
                                                          const uint8_t *__restrict__ lut, int nsym, int k,
                                                          const int32_t *__restrict__ col_of_code, int K,
                                                          const int32_t *__restrict__ ann_id,
                                                          const int64_t *__restrict__ order, int n_ann, int64_t chunk,
                                                          unsigned long long *__restrict__ M /* [n_ann+1, K] */) {
    extern __shared__ __align__(16) uint32_t s_row[];   // K counters
    __shared__ uint8_t s_lut[256];
    __shared__ unsigned int s_next;
    __shared__ int64_t s_run_end;
    s_lut[threadIdx.x] = lut[threadIdx.x];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_row[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nchunks = (nseq + chunk - 1) / chunk;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int64_t i1 = min(nseq, (ch + 1) * chunk);
        int64_t i = ch * chunk;
        while (i < i1) {
            const int64_t s_first = order ? __ldg(order + i) : i;
            int a = __ldg(ann_id + s_first);
            if (a < 0 || a >= n_ann) a = n_ann;
            // end of the run of equal annotation ids inside this chunk
            if (threadIdx.x == 0) { s_run_end = i1; s_next = 0; }
            __syncthreads();
            for (int64_t j = i + 1 + threadIdx.x; j < i1; j += blockDim.x) {
                const int64_t sj = order ? __ldg(order + j) : j;
                int aj = __ldg(ann_id + sj);
                if (aj < 0 || aj >= n_ann) aj = n_ann;
                if (aj != a) { atomicMin(reinterpret_cast<unsigned long long *>(&s_run_end), (unsigned long long)j); break; }
            }
            __syncthreads();
            const int64_t j1 = s_run_end;
            for (;;) {
                unsigned int t = 0;
                if (lane == 0) t = atomicAdd(&s_next, 1u);
                t = __shfl_sync(FULL, t, 0);
                if (i + t >= j1) break;
                const int64_t s = order ? __ldg(order + i + t) : (i + t);
                const int64_t b = __ldg(off + s), e = __ldg(off + s + 1);
                warp_scan_sequence<CodeT, NW>(res, nres, b, e, s_lut, nsym, k, [&](int64_t, CodeT code, bool ok) {
                    if (ok) {
                        const int32_t col = col_of_code ? __ldg(col_of_code + code) : int32_t(code);
                        if (col >= 0) atomicAdd(&s_row[col], 1u);
                    }
                });
            }
            __syncthreads();
            unsigned long long *dst = M + (size_t)a * K;
            for (int c = threadIdx.x; c < K; c += blockDim.x) {
                const uint32_t v = s_row[c];
                if (v) { atomicAdd(dst + c, (unsigned long long)v); s_row[c] = 0; }
            }
            __syncthreads();
            i = j1;
        }
    }
}

__global__ void learn_totals_kernel(const int64_t *__restrict__ M, const int64_t *__restrict__ rest, int64_t n_ann,
                                    int64_t K, int64_t *__restrict__ totals) {
    for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < K; c += int64_t(gridDim.x) * blockDim.x) {
        int64_t t = rest[c];
        for (int64_t a = 0; a < n_ann; ++a) t += M[a * K + c];
        totals[c] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) row_norm2_kernel(const T *__restrict__ X, int64_t rows, int64_t cols,
                                                        double *__restrict__ out) {
    // one warp per row; EXACT integer accumulation (128 bits per lane, so that sums of squares of 64-bit counts cannot
    // wrap), one conversion to double at the end: below 2^53 — every real matrix — the result is the exact integer,
    // the same value skm_csc_build's integer sums and sklearn's float64 accumulation give.
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const T *x = X + r * cols;
        unsigned __int128 acc = 0;
        for (int64_t c = lane; c < cols; c += 32) {
            const int64_t v = (int64_t)x[c];
            const unsigned long long a = (unsigned long long)(v < 0 ? -v : v);
            acc += (unsigned __int128)a * a;
        }
        unsigned long long lo = (unsigned long long)acc, hi = (unsigned long long)(acc >> 64);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long lo2 = __shfl_xor_sync(FULL, lo, o), hi2 = __shfl_xor_sync(FULL, hi, o);
            const unsigned long long s = lo + lo2;
            hi += hi2 + (s < lo ? 1ull : 0ull);
            lo = s;
        }
        if (lane == 0) out[r] = hi ? (double)hi * 18446744073709551616.0 + (double)lo : (double)lo;
    }
}

}  // namespace skm

extern "C" {

int skm_learn_dense(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq,
                    const uint8_t *d_lut, int nsym, int k, const int32_t *d_col_of_code, int64_t S, int64_t K,
                    const int32_t *d_ann_id, const int64_t *d_order, int64_t n_ann, int64_t *d_M, int64_t *d_totals,
                    skm_stream_t stream) {
    using namespace skm;
    int rc = check_common(d_residues, nres, d_offsets, nseq, d_lut, nsym, k);
    if (rc) return rc;
    unsigned __int128 S128;
    code_space(nsym, k, &S128);
    if (S128 > (unsigned __int128)SKM_DENSE_MAX_SPACE || (int64_t)S128 != S) { set_error("skm_learn_dense: S must equal nsym^k and be <= 2^27"); return SKM_ERR_UNSUPPORTED; }
    if (!d_col_of_code && K != S) { set_error("skm_learn_dense: identity basis needs K == S"); return SKM_ERR_INVALID; }
    if (K < 0 || n_ann < 0 || n_ann >= (1ll << 31) - 1) { set_error("skm_learn_dense: bad K / n_ann"); return SKM_ERR_INVALID; }
    if (size_t(K) * 4 > 200 * 1024) { set_error("skm_learn_dense: K=%lld does not fit a shared-memory row; use the sparse path", (long long)K); return SKM_ERR_UNSUPPORTED; }
    if (nseq > 0 && !d_ann_id) { set_error("skm_learn_dense: d_ann_id is NULL"); return SKM_ERR_INVALID; }
    if (K == 0) return SKM_OK;
    if (!d_totals || !d_M) { set_error("skm_learn_dense: NULL output"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_M, 0, (size_t)(n_ann + 1) * K * 8, st));
    if (nseq > 0 && nres > 0) {
        const size_t smem = ((size_t)K * 4 + 15) / 16 * 16;
        const int per_sm = smem <= 48 * 1024 ? 4 : (smem <= 100 * 1024 ? 2 : 1);
        const int64_t max_ctas = int64_t(sm_count()) * per_sm;
        int64_t chunk = (nseq + max_ctas * 4 - 1) / (max_ctas * 4);
        if (chunk < 16) chunk = 16;
        const int64_t nchunks = (nseq + chunk - 1) / chunk;
        const int grid = (int)std::min<int64_t>(nchunks, max_ctas);
        const int nw = neighbour_words(k);
        SKM_DISPATCH_NW(nw, {
            auto kern = learn_dense_kernel<uint32_t, NW>;
            SKM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 256, smem, st>>>(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, d_col_of_code, (int)K,
                                          d_ann_id, d_order, (int)n_ann, chunk,
                                          reinterpret_cast<unsigned long long *>(d_M));
        });
        SKM_LAUNCH_CHECK("learn_dense_kernel");
    }
    const int g2 = (int)std::min<int64_t>((K + 255) / 256, int64_t(sm_count()) * 4);
    learn_totals_kernel<<<g2, 256, 0, st>>>(d_M, d_M + (size_t)n_ann * K, n_ann, K, d_totals);
    SKM_LAUNCH_CHECK("learn_totals_kernel");
    return SKM_OK;
}

int skm_row_norm2_i32(const int32_t *d_X, int64_t rows, int64_t cols, double *d_out, skm_stream_t stream) {
    using namespace skm;
    if (rows < 0 || cols < 0 || (rows > 0 && (!d_out || (cols > 0 && !d_X)))) { set_error("skm_row_norm2_i32: bad arguments"); return SKM_ERR_INVALID; }
    if (rows == 0) return SKM_OK;
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, int64_t(sm_count()) * 8);
    row_norm2_kernel<int32_t><<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, rows, cols, d_out);
    SKM_LAUNCH_CHECK("row_norm2_kernel<i32>");
    return SKM_OK;
}

int skm_row_norm2_i64(const int64_t *d_X, int64_t rows, int64_t cols, double *d_out, skm_stream_t stream) {
    using namespace skm;
    if (rows < 0 || cols < 0 || (rows > 0 && (!d_out || (cols > 0 && !d_X)))) { set_error("skm_row_norm2_i64: bad arguments"); return SKM_ERR_INVALID; }
    if (rows == 0) return SKM_OK;
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, int64_t(sm_count()) * 8);
    row_norm2_kernel<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, rows, cols, d_out);
    SKM_LAUNCH_CHECK("row_norm2_kernel<i64>");
    return SKM_OK;
}

}  // extern "C"

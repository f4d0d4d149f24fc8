// skm_apply.cu — kernel (d): apply-mode cosine scoring + top-2 (exact integer dots).
//
// Replaces sklearn.metrics.pairwise.cosine_similarity(M, Q).T followed by
// np.argsort(-S)[:, :2] (apply.smk:278-335, learn.smk:811-849).  Counts are
// integers, so the dot products are computed EXACTLY (int64) and only the final
// division by the row norms is floating point (float64, like the reference):
//   score[q,a] = dot(q, m_a) / (sqrt(qnorm2[q]) * sqrt(mnorm2[a])), 0 if a norm is 0.
// This file holds the exact integer-core path (any K, any magnitude) and the
// top-2 epilogue; skm_apply_tc.cu holds the tcgen05 tensor-core path used when
// the operands fit its exactness envelope.
#include "skm_common.cuh"

namespace skm {

constexpr int AQ = 32;   // queries per CTA tile
constexpr int AA = 64;   // annotations per CTA tile
constexpr int AK = 32;   // k-mers per step

__global__ void __launch_bounds__(256) apply_dots_kernel(const int32_t *__restrict__ Q, int64_t nq, int64_t K,
                                                         const int64_t *__restrict__ M, int64_t n_ann,
                                                         int64_t *__restrict__ dots /* [nq, n_ann] */) {
    __shared__ int32_t s_q[AQ][AK + 1];
    __shared__ int64_t s_m[AA][AK + 1];
    const int64_t q0 = int64_t(blockIdx.y) * AQ, a0 = int64_t(blockIdx.x) * AA;
    const int tq = threadIdx.x >> 3;          // 0..31 : query row of this thread
    const int ta = threadIdx.x & 7;           // 0..7  : annotation column group (8 columns each, stride 8)
    int64_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t k0 = 0; k0 < K; k0 += AK) {
        for (int i = threadIdx.x; i < AQ * AK; i += 256) {
            const int r = i / AK, c = i % AK;
            const int64_t q = q0 + r, kk = k0 + c;
            s_q[r][c] = (q < nq && kk < K) ? __ldg(Q + q * K + kk) : 0;
        }
        for (int i = threadIdx.x; i < AA * AK; i += 256) {
            const int r = i / AK, c = i % AK;
            const int64_t a = a0 + r, kk = k0 + c;
            s_m[r][c] = (a < n_ann && kk < K) ? __ldg(M + a * K + kk) : 0;
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < AK; ++c) {
            const int64_t qv = s_q[tq][c];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += qv * s_m[ta + 8 * j][c];
        }
        __syncthreads();
    }
    const int64_t q = q0 + tq;
    if (q < nq) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t a = a0 + ta + 8 * j;
            if (a < n_ann) dots[q * n_ann + a] = acc[j];
        }
    }
}

struct Top2 {
    double s1, s2;
    int i1, i2;
};
// order: higher score first, ties -> lower index (np.argsort(-S) on exact ties)
__device__ __forceinline__ bool better(double s, int i, double t, int j) { return s > t || (s == t && i < j); }
__device__ __forceinline__ void top2_push(Top2 &t, double s, int i) {
    if (i < 0) return;
    if (t.i1 < 0 || better(s, i, t.s1, t.i1)) { t.s2 = t.s1; t.i2 = t.i1; t.s1 = s; t.i1 = i; }
    else if (t.i2 < 0 || better(s, i, t.s2, t.i2)) { t.s2 = s; t.i2 = i; }
}

// one warp per query over its row of exact dots
__global__ void __launch_bounds__(256) apply_top2_kernel(const int64_t *__restrict__ dots, int64_t nq, int64_t n_ann,
                                                         const double *__restrict__ qn2, const double *__restrict__ mn2,
                                                         int32_t *__restrict__ top1, int32_t *__restrict__ top2,
                                                         double *__restrict__ sc1, double *__restrict__ sc2,
                                                         double *__restrict__ full) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t q = warp; q < nq; q += nwarps) {
        const double qn = sqrt(qn2[q]);
        Top2 t{0.0, 0.0, -1, -1};
        for (int64_t a = lane; a < n_ann; a += 32) {
            const double mn = sqrt(__ldg(mn2 + a));
            const double den = qn * mn;
            const double s = (den > 0.0) ? double(dots[q * n_ann + a]) / den : 0.0;
            if (full) full[q * n_ann + a] = s;
            top2_push(t, s, int(a));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double os1 = __shfl_xor_sync(FULL, t.s1, o), os2 = __shfl_xor_sync(FULL, t.s2, o);
            const int oi1 = __shfl_xor_sync(FULL, t.i1, o), oi2 = __shfl_xor_sync(FULL, t.i2, o);
            top2_push(t, os1, oi1);
            top2_push(t, os2, oi2);
        }
        if (lane == 0) {
            top1[q] = t.i1; sc1[q] = (t.i1 >= 0) ? t.s1 : 0.0;
            top2[q] = t.i2; sc2[q] = (t.i2 >= 0) ? t.s2 : nan("");
        }
    }
}

// 2-way merge of per-shard top-2 lists (annotation-sharded apply): one thread per query
__global__ void __launch_bounds__(256) top2_merge_kernel(const int64_t *__restrict__ idx, const double *__restrict__ score,
                                                         int64_t n_shards, int64_t nq, int32_t *__restrict__ top1,
                                                         int32_t *__restrict__ top2, double *__restrict__ sc1,
                                                         double *__restrict__ sc2) {
    for (int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < nq; q += int64_t(gridDim.x) * blockDim.x) {
        Top2 t{0.0, 0.0, -1, -1};
        for (int64_t w = 0; w < n_shards; ++w)
            for (int j = 0; j < 2; ++j) {
                const int64_t i = idx[(w * 2 + j) * nq + q];
                if (i >= 0) top2_push(t, score[(w * 2 + j) * nq + q], int(i));
            }
        top1[q] = t.i1; sc1[q] = (t.i1 >= 0) ? t.s1 : 0.0;
        top2[q] = t.i2; sc2[q] = (t.i2 >= 0) ? t.s2 : nan("");
    }
}

}  // namespace skm

extern "C" {

int skm_top2_merge(const int64_t *d_idx, const double *d_score, int64_t n_shards, int64_t nq, int32_t *d_top1,
                   int32_t *d_top2, double *d_score1, double *d_score2, skm_stream_t stream) {
    using namespace skm;
    if (n_shards < 0 || nq < 0) { set_error("skm_top2_merge: negative size"); return SKM_ERR_INVALID; }
    if (nq == 0) return SKM_OK;
    if (!d_top1 || !d_top2 || !d_score1 || !d_score2 || (n_shards > 0 && (!d_idx || !d_score))) { set_error("skm_top2_merge: NULL argument"); return SKM_ERR_INVALID; }
    const int grid = (int)std::min<int64_t>((nq + 255) / 256, int64_t(sm_count()) * 8);
    top2_merge_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_idx, d_score, n_shards, nq, d_top1, d_top2, d_score1, d_score2);
    SKM_LAUNCH_CHECK("top2_merge_kernel");
    return SKM_OK;
}

size_t skm_apply_dense_workspace(int64_t nq, int64_t n_ann, int64_t K) {
    (void)K;
    if (nq <= 0 || n_ann <= 0) return 256;
    return (size_t)nq * (size_t)n_ann * 8 + 256;
}

int skm_apply_dense(const int32_t *d_Q, int64_t nq, int64_t K, const int64_t *d_M, int64_t n_ann,
                    const double *d_qnorm2, const double *d_mnorm2, int32_t *d_top1, int32_t *d_top2,
                    double *d_score1, double *d_score2, double *d_scores_full, void *workspace,
                    size_t workspace_bytes, skm_stream_t stream) {
    using namespace skm;
    if (nq < 0 || K < 0 || n_ann < 0 || n_ann > 0x7FFFFFFF) { set_error("skm_apply_dense: bad sizes"); return SKM_ERR_INVALID; }
    if (nq == 0) return SKM_OK;
    if (!d_top1 || !d_top2 || !d_score1 || !d_score2 || !d_qnorm2 || (n_ann > 0 && !d_mnorm2)) { set_error("skm_apply_dense: NULL argument"); return SKM_ERR_INVALID; }
    if (K > 0 && n_ann > 0 && (!d_Q || !d_M)) { set_error("skm_apply_dense: NULL matrix"); return SKM_ERR_INVALID; }
    const size_t need = skm_apply_dense_workspace(nq, n_ann, K);
    if (!workspace || workspace_bytes < need) { set_error("skm_apply_dense: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t *dots = reinterpret_cast<int64_t *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    if (n_ann > 0) {
        if (K == 0) {
            SKM_CUDA_TRY(cudaMemsetAsync(dots, 0, (size_t)nq * n_ann * 8, st));
        } else {
            dim3 grid((unsigned)((n_ann + AA - 1) / AA), (unsigned)((nq + AQ - 1) / AQ));
            if (grid.y > 65535) { set_error("skm_apply_dense: more than %d queries per call; chunk the queries", 65535 * AQ); return SKM_ERR_UNSUPPORTED; }
            apply_dots_kernel<<<grid, 256, 0, st>>>(d_Q, nq, K, d_M, n_ann, dots);
            SKM_LAUNCH_CHECK("apply_dots_kernel");
        }
    }
    const int g2 = (int)std::min<int64_t>((nq + 7) / 8, int64_t(sm_count()) * 8);
    apply_top2_kernel<<<g2, 256, 0, st>>>(dots, nq, n_ann, d_qnorm2, d_mnorm2, d_top1, d_top2, d_score1, d_score2, d_scores_full);
    SKM_LAUNCH_CHECK("apply_top2_kernel");
    return SKM_OK;
}

}  // extern "C"

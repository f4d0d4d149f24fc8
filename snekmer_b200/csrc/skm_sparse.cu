// skm_sparse.cu — sort-based (atomic-free) sparse paths for large k-mer bases.
//
//  skm_learn_sparse   kernel (c) for bases that do not fit a shared-memory row
//                     (learn.smk:306-326,359-408 at K ~ 1e6): every valid window of an
//                     annotated sequence becomes the 64-bit key  ann * S + code;  the keys of a
//                     shard are radix-sorted and run-length encoded into a sorted COO
//                     (key, count) list = the annotation x k-mer count matrix.  No atomics,
//                     bit-reproducible.  Unannotated sequences only feed the Totals table
//                     (skm_basis_accumulate gives it).
//  skm_coo_merge      sum of COO lists with equal keys (sort by key + reduce by key):
//                     Merge.merge_dataframes (learn.smk:467-494) for sparse matrices, and the
//                     fan-in after an all-gather of per-GPU lists.
//  skm_csc_build      annotation-major COO -> k-mer-major CSC with cosine weights
//                     w = M[a, c] / ||m_a||  (float32) for the SpMM.
//  skm_apply_sparse   kernel (d) as SpMM: per query (CSR row of k-mer codes + counts) the
//                     weighted columns of the CSC are accumulated in a shared-memory score
//                     vector (one warp per query, lanes own distinct annotations: no atomics),
//                     followed by the top-2 scan (apply.smk:278-335).
//  window_keys_kernel is also the key generator of skm_count_csr (skm_count.cu).
#include <cub/cub.cuh>

#include "skm_common.cuh"
#include "skm_tile.cuh"

namespace skm {

constexpr int SP_SEG = ts_seg_cap(52);
constexpr int SP_SYM_BYTES = ts_sym_bytes(SP_SEG);

// keys[p] for every residue position p of [off[0], off[nseq]) (p = position of the window's LAST residue):
//   MODE 0: column (col_of_code) or code, 32-bit, all-ones when invalid / filtered
//   MODE 1: ann_id[row] * S + code, 64-bit, all-ones when invalid or the sequence is not annotated
template <int MODE, typename KeyT>
__global__ void __launch_bounds__(TS_THREADS) window_keys_kernel(const uint8_t *__restrict__ res, int64_t nres,
                                                                 const int64_t *__restrict__ off, int64_t nseq,
                                                                 const uint8_t *__restrict__ lut, uint32_t nsym, int k,
                                                                 uint32_t pow_k1, const int32_t *__restrict__ col_of_code,
                                                                 const int32_t *__restrict__ ann_id, uint64_t S,
                                                                 KeyT *__restrict__ keys) {
    extern __shared__ __align__(128) uint8_t s_sym[];
    __shared__ uint8_t s_lut[256];
    __shared__ int64_t s_ctl[4];
    ts_lut_init(s_lut, lut);
    __syncthreads();
    int64_t ann_row = -2;          // row whose annotation is cached
    uint64_t ann_base = ~0ull;
    ts_range_scan_rows<uint32_t>(res, nres, off, nseq, s_lut, s_sym, SP_SEG, k, nsym, pow_k1, /*origin=*/0, s_ctl,
                                 [&](int64_t rel_end, int64_t row, uint32_t code, bool ok) {
                                     KeyT key = KeyT(~KeyT(0));
                                     if (ok) {
                                         if (MODE == 0) {
                                             if (col_of_code) { const int32_t c = __ldg(col_of_code + code); if (c >= 0) key = KeyT(c); }
                                             else key = KeyT(code);
                                         } else {
                                             if (row != ann_row) {
                                                 ann_row = row;
                                                 const int32_t a = __ldg(ann_id + row);
                                                 ann_base = (a >= 0) ? uint64_t(a) * S : ~0ull;
                                             }
                                             if (ann_base != ~0ull) key = KeyT(ann_base + code);
                                         }
                                     }
                                     keys[rel_end] = key;
                                 });
}

__global__ void coo_finish_kernel(const uint64_t *__restrict__ uniq, const int64_t *__restrict__ num_runs, int64_t *__restrict__ nnz) {
    const int64_t n = *num_runs;
    *nnz = (n > 0 && uniq[n - 1] == ~0ull) ? n - 1 : n;
}

// ---- CSC build -------------------------------------------------------------------------------
// ||m_a||^2 as exact integers (order-independent, hence reproducible); requires sum v^2 < 2^64
__global__ void __launch_bounds__(256) coo_row_norm2_kernel(const uint64_t *__restrict__ keys, const int64_t *__restrict__ vals,
                                                            int64_t nnz, uint64_t S, unsigned long long *__restrict__ norm2) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += int64_t(gridDim.x) * blockDim.x) {
        const unsigned long long v = (unsigned long long)vals[i];
        atomicAdd(norm2 + keys[i] / S, v * v);
    }
}
__global__ void norm2_to_double_kernel(const unsigned long long *__restrict__ in, int64_t n, double *__restrict__ out) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) out[i] = double(in[i]);
}
__global__ void __launch_bounds__(256) csc_keys_kernel(const uint64_t *__restrict__ keys, int64_t nnz, uint64_t S,
                                                       uint64_t n_ann, uint64_t *__restrict__ out, int64_t *__restrict__ perm) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += int64_t(gridDim.x) * blockDim.x) {
        const uint64_t a = keys[i] / S, c = keys[i] - a * S;
        out[i] = c * n_ann + a;
        perm[i] = i;
    }
}
__global__ void __launch_bounds__(256) csc_emit_kernel(const uint64_t *__restrict__ skeys, const int64_t *__restrict__ perm,
                                                       const int64_t *__restrict__ vals, const unsigned long long *__restrict__ norm2,
                                                       int64_t nnz, uint64_t n_ann, int32_t *__restrict__ rows,
                                                       float *__restrict__ w, unsigned long long *__restrict__ colcount) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += int64_t(gridDim.x) * blockDim.x) {
        const uint64_t c = skeys[i] / n_ann, a = skeys[i] - c * n_ann;
        rows[i] = int32_t(a);
        const double n2 = double(norm2[a]);
        w[i] = n2 > 0.0 ? float(double(vals[perm[i]]) / sqrt(n2)) : 0.0f;
        atomicAdd(colcount + c + 1, 1ull);
    }
}

// ---- SpMM + top-2 ------------------------------------------------------------------------------
struct Top2f {
    float s1, s2;
    int i1, i2;
};
__device__ __forceinline__ bool better_f(float s, int i, float t, int j) { return s > t || (s == t && i < j); }
__device__ __forceinline__ void top2f_push(Top2f &t, float s, int i) {
    if (i < 0) return;
    if (t.i1 < 0 || better_f(s, i, t.s1, t.i1)) { t.s2 = t.s1; t.i2 = t.i1; t.s1 = s; t.i1 = i; }
    else if (t.i2 < 0 || better_f(s, i, t.s2, t.i2)) { t.s2 = s; t.i2 = i; }
}

// one warp per query; acc = n_ann floats of shared memory per warp
__global__ void __launch_bounds__(256) apply_sparse_kernel(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ qcols,
                                                           const int32_t *__restrict__ qvals, int64_t nq,
                                                           const int64_t *__restrict__ colptr, const int32_t *__restrict__ rows,
                                                           const float *__restrict__ w, int n_ann,
                                                           int32_t *__restrict__ top1, int32_t *__restrict__ top2,
                                                           double *__restrict__ sc1, double *__restrict__ sc2,
                                                           double *__restrict__ qnorm2_out) {
    extern __shared__ float s_acc[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float *acc = s_acc + size_t(wid) * n_ann;
    for (int a = lane; a < n_ann; a += 32) acc[a] = 0.0f;
    __syncwarp();
    for (int64_t q = int64_t(blockIdx.x) * nw + wid; q < nq; q += int64_t(gridDim.x) * nw) {
        const int64_t e0 = __ldg(rowptr + q), e1 = __ldg(rowptr + q + 1);
        double n2 = 0.0;
        for (int64_t e = e0; e < e1; ++e) {
            const uint32_t c = __ldg(qcols + e);
            const float cnt = float(__ldg(qvals + e));
            n2 += double(cnt) * double(cnt);
            const int64_t p0 = __ldg(colptr + c), p1 = __ldg(colptr + c + 1);
            for (int64_t p = p0 + lane; p < p1; p += 32) acc[__ldg(rows + p)] += cnt * __ldg(w + p);   // distinct annotations per column
            __syncwarp();
        }
        const float inv = n2 > 0.0 ? float(1.0 / sqrt(n2)) : 0.0f;
        Top2f t{0.f, 0.f, -1, -1};
        for (int a = lane; a < n_ann; a += 32) {
            top2f_push(t, acc[a] * inv, a);
            acc[a] = 0.0f;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os1 = __shfl_xor_sync(FULL, t.s1, o), os2 = __shfl_xor_sync(FULL, t.s2, o);
            const int oi1 = __shfl_xor_sync(FULL, t.i1, o), oi2 = __shfl_xor_sync(FULL, t.i2, o);
            top2f_push(t, os1, oi1);
            top2f_push(t, os2, oi2);
        }
        if (lane == 0) {
            top1[q] = t.i1; sc1[q] = (t.i1 >= 0) ? double(t.s1) : 0.0;
            top2[q] = t.i2; sc2[q] = (t.i2 >= 0) ? double(t.s2) : nan("");
            if (qnorm2_out) qnorm2_out[q] = n2;
        }
        __syncwarp();
    }
}

static size_t al(size_t x) { return (x + 255) & ~size_t(255); }
static int bits_for(unsigned __int128 n) { int b = 1; while (b < 64 && (((unsigned __int128)1) << b) < n) ++b; return b; }

}  // namespace skm

extern "C" {

// window keys for skm_count_csr (declared in skm_count.cu)
int skm_window_keys_u32(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq,
                        const uint8_t *d_lut, int nsym, int k, const int32_t *d_col_of_code, uint32_t *d_keys,
                        cudaStream_t st) {
    using namespace skm;
    uint32_t pow_k1 = 1;
    for (int i = 0; i + 1 < k; ++i) pow_k1 *= (uint32_t)nsym;
    int64_t grid = int64_t(sm_count()) * 8;
    const int64_t max_grid = (nres + SP_SEG - 1) / SP_SEG;
    if (grid > max_grid) grid = max_grid < 1 ? 1 : max_grid;
    window_keys_kernel<0, uint32_t><<<(unsigned)grid, TS_THREADS, SP_SYM_BYTES, st>>>(d_residues, nres, d_offsets, nseq, d_lut, (uint32_t)nsym, k,
                                                                                      pow_k1, d_col_of_code, nullptr, 0, d_keys);
    SKM_LAUNCH_CHECK("window_keys_kernel<0>");
    return SKM_OK;
}

size_t skm_learn_sparse_workspace(int64_t nres) {
    using namespace skm;
    if (nres <= 0) return 256;
    size_t t_sort = 0, t_rle = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, t_sort, (const uint64_t *)nullptr, (uint64_t *)nullptr, nres, 0, 64);
    cub::DeviceRunLengthEncode::Encode(nullptr, t_rle, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t *)nullptr,
                                       (int64_t *)nullptr, (int)std::min<int64_t>(nres, (1ll << 31) - 1));
    return 2 * al(size_t(nres) * 8) + al(std::max(t_sort, t_rle)) + 1024;
}

int skm_learn_sparse(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq,
                     const uint8_t *d_lut, int nsym, int k, const int32_t *d_ann_id, int64_t n_ann,
                     uint64_t *d_keys_out, int64_t *d_vals_out, int64_t *d_nnz, void *workspace,
                     size_t workspace_bytes, skm_stream_t stream) {
    using namespace skm;
    int rc = check_common(d_residues, nres, d_offsets, nseq, d_lut, nsym, k);
    if (rc) return rc;
    if (!ts_supported(nsym, k)) { set_error("skm_learn_sparse: nsym=%d k=%d outside the kernel envelope", nsym, k); return SKM_ERR_UNSUPPORTED; }
    unsigned __int128 S128;
    code_space(nsym, k, &S128);
    if (S128 >= ((unsigned __int128)1 << 32)) { set_error("skm_learn_sparse: code space must be < 2^32"); return SKM_ERR_UNSUPPORTED; }
    if (n_ann < 0 || (unsigned __int128)(n_ann + 1) * S128 >= (((unsigned __int128)1) << 63)) { set_error("skm_learn_sparse: n_ann * S does not fit 63 bits"); return SKM_ERR_UNSUPPORTED; }
    if (!d_nnz) { set_error("skm_learn_sparse: d_nnz is NULL"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_nnz, 0, 8, st));
    if (nseq == 0 || nres == 0 || n_ann == 0) return SKM_OK;
    if (nres >= (1ll << 31)) { set_error("skm_learn_sparse: more than 2^31 residues per call; split the shard"); return SKM_ERR_UNSUPPORTED; }
    if (!d_ann_id || !d_keys_out || !d_vals_out) { set_error("skm_learn_sparse: NULL argument"); return SKM_ERR_INVALID; }
    const size_t need = skm_learn_sparse_workspace(nres);
    if (!workspace || workspace_bytes < need) { set_error("skm_learn_sparse: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    const size_t seg = al(size_t(nres) * 8);
    uint64_t *keys_a = (uint64_t *)p, *keys_b = (uint64_t *)(p + seg);
    void *temp = p + 2 * seg;
    size_t temp_bytes = workspace_bytes - size_t((char *)temp - (char *)workspace);
    const uint64_t S = (uint64_t)S128;
    uint32_t pow_k1 = 1;
    for (int i = 0; i + 1 < k; ++i) pow_k1 *= (uint32_t)nsym;
    int64_t grid = int64_t(sm_count()) * 8;
    const int64_t max_grid = (nres + SP_SEG - 1) / SP_SEG;
    if (grid > max_grid) grid = max_grid < 1 ? 1 : max_grid;
    // positions outside [off[0], off[nseq]) are not written by the kernel
    SKM_CUDA_TRY(cudaMemsetAsync(keys_a, 0xFF, size_t(nres) * 8, st));
    window_keys_kernel<1, uint64_t><<<(unsigned)grid, TS_THREADS, SP_SYM_BYTES, st>>>(d_residues, nres, d_offsets, nseq, d_lut, (uint32_t)nsym, k,
                                                                                      pow_k1, nullptr, d_ann_id, S, keys_a);
    SKM_LAUNCH_CHECK("window_keys_kernel<1>");
    // all-ones (invalid) sorts last because every bit up to 63 takes part: sort on [0, 64) only when needed
    const int end_bit = 64;
    (void)bits_for;
    SKM_CUDA_TRY(cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys_a, keys_b, nres, 0, end_bit, st));
    int64_t *num_runs = reinterpret_cast<int64_t *>(keys_a);     // keys_a is free after the sort
    temp_bytes = workspace_bytes - size_t((char *)temp - (char *)workspace);
    SKM_CUDA_TRY(cub::DeviceRunLengthEncode::Encode(temp, temp_bytes, keys_b, d_keys_out, d_vals_out, num_runs, (int)nres, st));
    coo_finish_kernel<<<1, 1, 0, st>>>(d_keys_out, num_runs, d_nnz);
    SKM_LAUNCH_CHECK("coo_finish_kernel");
    return SKM_OK;
}

size_t skm_coo_merge_workspace(int64_t n) {
    using namespace skm;
    if (n <= 0) return 256;
    size_t t_sort = 0, t_red = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int64_t *)nullptr,
                                    (int64_t *)nullptr, n, 0, 64);
    cub::DeviceReduce::ReduceByKey(nullptr, t_red, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int64_t *)nullptr,
                                   (int64_t *)nullptr, (int64_t *)nullptr, cub::Sum(), n);
    return 2 * al(size_t(n) * 8) + al(std::max(t_sort, t_red)) + 1024;
}

int skm_coo_merge(const uint64_t *d_keys_in, const int64_t *d_vals_in, int64_t n, uint64_t *d_keys_out,
                  int64_t *d_vals_out, int64_t *d_n_out, void *workspace, size_t workspace_bytes,
                  skm_stream_t stream) {
    using namespace skm;
    if (n < 0 || !d_n_out) { set_error("skm_coo_merge: bad arguments"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_n_out, 0, 8, st));
    if (n == 0) return SKM_OK;
    if (!d_keys_in || !d_vals_in || !d_keys_out || !d_vals_out) { set_error("skm_coo_merge: NULL argument"); return SKM_ERR_INVALID; }
    const size_t need = skm_coo_merge_workspace(n);
    if (!workspace || workspace_bytes < need) { set_error("skm_coo_merge: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    const size_t seg = al(size_t(n) * 8);
    uint64_t *sk = (uint64_t *)p;
    int64_t *sv = (int64_t *)(p + seg);
    void *temp = p + 2 * seg;
    size_t temp_bytes = workspace_bytes - size_t((char *)temp - (char *)workspace);
    SKM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, d_keys_in, sk, d_vals_in, sv, n, 0, 64, st));
    temp_bytes = workspace_bytes - size_t((char *)temp - (char *)workspace);
    SKM_CUDA_TRY(cub::DeviceReduce::ReduceByKey(temp, temp_bytes, sk, d_keys_out, sv, d_vals_out, d_n_out, cub::Sum(), n, st));
    return SKM_OK;
}

size_t skm_csc_build_workspace(int64_t nnz, int64_t n_ann) {
    using namespace skm;
    if (nnz <= 0) return 256 + al(size_t(std::max<int64_t>(n_ann, 1)) * 8);
    size_t t_sort = 0, t_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int64_t *)nullptr,
                                    (int64_t *)nullptr, nnz, 0, 64);
    cub::DeviceScan::InclusiveSum(nullptr, t_scan, (const int64_t *)nullptr, (int64_t *)nullptr, (int64_t)1 << 27);
    return 4 * al(size_t(nnz) * 8) + al(size_t(n_ann) * 8) + al(std::max(t_sort, t_scan)) + 1024;
}

int skm_csc_build(const uint64_t *d_keys, const int64_t *d_vals, int64_t nnz, int64_t S, int64_t n_ann,
                  int64_t *d_colptr, int32_t *d_rows, float *d_w, double *d_mnorm2, void *workspace,
                  size_t workspace_bytes, skm_stream_t stream) {
    using namespace skm;
    if (nnz < 0 || S <= 0 || S > SKM_DENSE_MAX_SPACE || n_ann < 0 || !d_colptr) { set_error("skm_csc_build: bad arguments"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_colptr, 0, size_t(S + 1) * 8, st));
    if (d_mnorm2 && n_ann > 0) SKM_CUDA_TRY(cudaMemsetAsync(d_mnorm2, 0, size_t(n_ann) * 8, st));
    if (nnz == 0) return SKM_OK;
    if (!d_keys || !d_vals || !d_rows || !d_w) { set_error("skm_csc_build: NULL argument"); return SKM_ERR_INVALID; }
    const size_t need = skm_csc_build_workspace(nnz, n_ann);
    if (!workspace || workspace_bytes < need) { set_error("skm_csc_build: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    const size_t seg = al(size_t(nnz) * 8);
    uint64_t *k_in = (uint64_t *)p, *k_out = (uint64_t *)(p + seg);
    int64_t *perm_in = (int64_t *)(p + 2 * seg), *perm_out = (int64_t *)(p + 3 * seg);
    unsigned long long *norm2 = (unsigned long long *)(p + 4 * seg);
    void *temp = p + 4 * seg + al(size_t(n_ann) * 8);
    size_t temp_bytes = workspace_bytes - size_t((char *)temp - (char *)workspace);
    const int grid = (int)std::min<int64_t>((nnz + 255) / 256, int64_t(sm_count()) * 16);
    SKM_CUDA_TRY(cudaMemsetAsync(norm2, 0, size_t(n_ann) * 8, st));
    coo_row_norm2_kernel<<<grid, 256, 0, st>>>(d_keys, d_vals, nnz, (uint64_t)S, norm2);
    SKM_LAUNCH_CHECK("coo_row_norm2_kernel");
    csc_keys_kernel<<<grid, 256, 0, st>>>(d_keys, nnz, (uint64_t)S, (uint64_t)n_ann, k_in, perm_in);
    SKM_LAUNCH_CHECK("csc_keys_kernel");
    SKM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k_in, k_out, perm_in, perm_out, nnz, 0, 64, st));
    csc_emit_kernel<<<grid, 256, 0, st>>>(k_out, perm_out, d_vals, norm2, nnz, (uint64_t)n_ann, d_rows, d_w,
                                          reinterpret_cast<unsigned long long *>(d_colptr));
    SKM_LAUNCH_CHECK("csc_emit_kernel");
    temp_bytes = workspace_bytes - size_t((char *)temp - (char *)workspace);
    SKM_CUDA_TRY(cub::DeviceScan::InclusiveSum(temp, temp_bytes, d_colptr, d_colptr, S + 1, st));
    if (d_mnorm2 && n_ann > 0) {
        norm2_to_double_kernel<<<(int)std::min<int64_t>((n_ann + 255) / 256, 1024), 256, 0, st>>>(norm2, n_ann, d_mnorm2);
        SKM_LAUNCH_CHECK("norm2_to_double_kernel");
    }
    return SKM_OK;
}

int skm_apply_sparse(const int64_t *d_rowptr, const uint32_t *d_cols, const int32_t *d_vals, int64_t nq,
                     const int64_t *d_colptr, const int32_t *d_rows, const float *d_w, int64_t n_ann,
                     int32_t *d_top1, int32_t *d_top2, double *d_score1, double *d_score2, double *d_qnorm2,
                     skm_stream_t stream) {
    using namespace skm;
    if (nq < 0 || n_ann < 0 || n_ann > 50 * 1024) { set_error("skm_apply_sparse: n_ann=%lld outside [0, 51200] (shard the annotations)", (long long)n_ann); return n_ann > 50 * 1024 ? SKM_ERR_UNSUPPORTED : SKM_ERR_INVALID; }
    if (nq == 0) return SKM_OK;
    if (!d_rowptr || !d_colptr || !d_top1 || !d_top2 || !d_score1 || !d_score2) { set_error("skm_apply_sparse: NULL argument"); return SKM_ERR_INVALID; }
    // warps per CTA: as many score vectors as fit ~200 KB, at most 8
    const size_t per_warp = size_t(std::max<int64_t>(n_ann, 1)) * 4;
    int nw = int((200 * 1024) / per_warp);
    if (nw > 8) nw = 8;
    if (nw < 1) nw = 1;
    const size_t smem = per_warp * nw;
    int per_sm = int((227 * 1024) / (smem + 1024));
    if (per_sm > 2048 / (32 * nw)) per_sm = 2048 / (32 * nw);
    if (per_sm < 1) per_sm = 1;
    const int grid = (int)std::min<int64_t>((nq + nw - 1) / nw, int64_t(sm_count()) * per_sm);
    SKM_CUDA_TRY(cudaFuncSetAttribute(apply_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    apply_sparse_kernel<<<grid, 32 * nw, smem, (cudaStream_t)stream>>>(d_rowptr, d_cols, d_vals, nq, d_colptr, d_rows, d_w, (int)n_ann, d_top1,
                                                                       d_top2, d_score1, d_score2, d_qnorm2);
    SKM_LAUNCH_CHECK("apply_sparse_kernel");
    return SKM_OK;
}

}  // extern "C"

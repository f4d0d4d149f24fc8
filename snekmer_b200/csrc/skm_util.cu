// skm_util.cu — small data-movement kernels around the hot path.
//
// skm_gather_columns: KmerBasis.transform / KmerVec.harmonize (vectorize.py:54-119,
//   330-345) — out[r, j] = in[r, idx[j]], or 0 where idx[j] == n (k-mer absent from
//   the source basis).  Also used to expand counts over the distinct codes of a
//   supplied k-mer list to one column per list entry.
// skm_scatter_add_i64: Merge.merge_dataframes (learn.smk:467-494) — outer join of
//   count matrices on k-mer columns + row-wise sum, as dst[row_map[r], col_map[c]] +=
//   src[r, c] (integer adds: order-independent, bit-reproducible).
#include "skm_common.cuh"

namespace skm {

template <typename T>
__global__ void __launch_bounds__(256) gather_columns_kernel(const T *__restrict__ in, int64_t rows, int64_t n,
                                                             const int64_t *__restrict__ idx, int64_t p,
                                                             T *__restrict__ out) {
    const int64_t total = rows * p;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / p, j = i - r * p;
        const int64_t c = __ldg(idx + j);
        out[i] = (c >= 0 && c < n) ? in[r * n + c] : T(0);
    }
}

__global__ void __launch_bounds__(256) scatter_add_i64_kernel(const int64_t *__restrict__ src, int64_t rows, int64_t cols,
                                                              const int64_t *__restrict__ row_map,
                                                              const int64_t *__restrict__ col_map,
                                                              unsigned long long *__restrict__ dst, int64_t dst_cols) {
    const int64_t total = rows * cols;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t v = src[i];
        if (v == 0) continue;
        const int64_t r = i / cols, c = i - r * cols;
        const int64_t dr = __ldg(row_map + r), dc = __ldg(col_map + c);
        if (dr >= 0 && dc >= 0) atomicAdd(dst + dr * dst_cols + dc, (unsigned long long)v);
    }
}

// one warp per selected sequence: out_res[out_off[i] ...] = res[off[sel[i]] ...]
__global__ void __launch_bounds__(256) gather_sequences_kernel(const uint8_t *__restrict__ res, const int64_t *__restrict__ off,
                                                               const int64_t *__restrict__ sel, int64_t n_sel,
                                                               uint8_t *__restrict__ out_res, const int64_t *__restrict__ out_off) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t i = warp; i < n_sel; i += nwarps) {
        const int64_t s = __ldg(sel + i);
        const int64_t b = __ldg(off + s), len = __ldg(off + s + 1) - b, d = __ldg(out_off + i);
        // head bytes up to the first 4-byte boundary of the destination, then one word per lane and step
        const int64_t head = min(len, (4 - (d & 3)) & 3);
        if (lane < head) out_res[d + lane] = res[b + lane];
        const int64_t words = (len - head) >> 2;
        for (int64_t w = lane; w < words; w += 32) {
            const uint8_t *src = res + b + head + 4 * w;
            const uint32_t v = uint32_t(src[0]) | (uint32_t(src[1]) << 8) | (uint32_t(src[2]) << 16) | (uint32_t(src[3]) << 24);
            *reinterpret_cast<uint32_t *>(out_res + d + head + 4 * w) = v;
        }
        const int64_t done = head + 4 * words;
        if (lane < len - done) out_res[d + done + lane] = res[b + done + lane];
    }
}

}  // namespace skm

extern "C" {

int skm_gather_sequences(const uint8_t *d_residues, const int64_t *d_offsets, const int64_t *d_sel, int64_t n_sel,
                         uint8_t *d_out_residues, const int64_t *d_out_offsets, skm_stream_t stream) {
    using namespace skm;
    if (n_sel < 0) { set_error("skm_gather_sequences: negative size"); return SKM_ERR_INVALID; }
    if (n_sel == 0) return SKM_OK;
    if (!d_residues || !d_offsets || !d_sel || !d_out_residues || !d_out_offsets) { set_error("skm_gather_sequences: NULL argument"); return SKM_ERR_INVALID; }
    const int grid = (int)std::min<int64_t>((n_sel + 7) / 8, int64_t(sm_count()) * 8);
    gather_sequences_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_residues, d_offsets, d_sel, n_sel, d_out_residues, d_out_offsets);
    SKM_LAUNCH_CHECK("gather_sequences_kernel");
    return SKM_OK;
}

int skm_gather_columns(const void *d_in, int64_t rows, int64_t n, int elem_bytes, const int64_t *d_idx, int64_t p,
                       void *d_out, skm_stream_t stream) {
    using namespace skm;
    if (rows < 0 || n < 0 || p < 0 || (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8)) {
        set_error("skm_gather_columns: bad arguments (elem_bytes must be 1, 2, 4 or 8)");
        return SKM_ERR_INVALID;
    }
    if (rows == 0 || p == 0) return SKM_OK;
    if (!d_out || !d_idx || (n > 0 && !d_in)) { set_error("skm_gather_columns: NULL argument"); return SKM_ERR_INVALID; }
    const int grid = (int)std::min<int64_t>((rows * p + 255) / 256, int64_t(sm_count()) * 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: gather_columns_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)d_in, rows, n, d_idx, p, (uint8_t *)d_out); break;
        case 2: gather_columns_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)d_in, rows, n, d_idx, p, (uint16_t *)d_out); break;
        case 4: gather_columns_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)d_in, rows, n, d_idx, p, (uint32_t *)d_out); break;
        default: gather_columns_kernel<uint64_t><<<grid, 256, 0, st>>>((const uint64_t *)d_in, rows, n, d_idx, p, (uint64_t *)d_out); break;
    }
    SKM_LAUNCH_CHECK("gather_columns_kernel");
    return SKM_OK;
}

int skm_scatter_add_i64(const int64_t *d_src, int64_t rows, int64_t cols, const int64_t *d_row_map,
                        const int64_t *d_col_map, int64_t *d_dst, int64_t dst_rows, int64_t dst_cols,
                        skm_stream_t stream) {
    using namespace skm;
    if (rows < 0 || cols < 0 || dst_rows < 0 || dst_cols < 0) { set_error("skm_scatter_add_i64: negative size"); return SKM_ERR_INVALID; }
    if (rows == 0 || cols == 0) return SKM_OK;
    if (!d_src || !d_row_map || !d_col_map || !d_dst) { set_error("skm_scatter_add_i64: NULL argument"); return SKM_ERR_INVALID; }
    const int grid = (int)std::min<int64_t>((rows * cols + 255) / 256, int64_t(sm_count()) * 16);
    scatter_add_i64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, rows, cols, d_row_map, d_col_map,
                                                                  reinterpret_cast<unsigned long long *>(d_dst), dst_cols);
    SKM_LAUNCH_CHECK("scatter_add_i64_kernel");
    return SKM_OK;
}

}  // extern "C"

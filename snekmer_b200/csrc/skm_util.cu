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


// ---- compact transports of a dense count matrix (pipeline.vectorize_host) ------------------------------------
// uint8 counts saturated at 255 + an escape list (row, col, count) of the entries that do not fit: lossless, a
// quarter of the int32 bytes.  One thread per 4 consecutive elements of the flattened matrix (one output word).
template <typename T>
__global__ void __launch_bounds__(256) pack_u8_kernel(const T *__restrict__ in, int64_t total, int64_t cols,
                                                      uint8_t *__restrict__ out, int32_t *__restrict__ esc_row,
                                                      int32_t *__restrict__ esc_col, int32_t *__restrict__ esc_val,
                                                      int64_t cap, unsigned long long *__restrict__ n_esc) {
    const int64_t words = (total + 3) >> 2;
    for (int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; w < words; w += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i0 = w << 2;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = i0 + j;
            if (i >= total) break;
            const uint32_t v = uint32_t(in[i]);
            if (v > 254u) {                      // 255 marks an escaped entry (also the exact value 255)
                const unsigned long long slot = atomicAdd(n_esc, 1ull);
                if ((int64_t)slot < cap) {
                    esc_row[slot] = int32_t(i / cols);
                    esc_col[slot] = int32_t(i % cols);
                    esc_val[slot] = int32_t(v);
                }
                packed |= 255u << (8 * j);
            } else {
                packed |= v << (8 * j);
            }
        }
        if (i0 + 4 <= total) {
            reinterpret_cast<uint32_t *>(out)[w] = packed;
        } else {
            for (int j = 0; i0 + j < total; ++j) out[i0 + j] = uint8_t(packed >> (8 * j));
        }
    }
}

// presence bits (kmerize.smk:112-120: vecs = counts > 0): bit (c & 7) of byte c >> 3 of a row of ceil(cols / 8) bytes
// (numpy.unpackbits(..., bitorder="little")).  One warp per 32 columns of a row -> one ballot word.
template <typename T>
__global__ void __launch_bounds__(256) pack_bits_kernel(const T *__restrict__ in, int64_t rows, int64_t cols, int64_t row_bytes,
                                                        uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t wpr = (cols + 31) >> 5;                     // ballot words per row
    const int64_t nw = rows * wpr;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t w = warp; w < nw; w += nwarps) {
        const int64_t r = w / wpr, c = ((w - r * wpr) << 5) + lane;
        const bool on = (c < cols) && in[r * cols + c] != T(0);
        const unsigned m = __ballot_sync(FULL, on);
        const int64_t byte0 = r * row_bytes + ((w - r * wpr) << 2);
        if (lane < 4 && ((w - r * wpr) << 2) + lane < row_bytes) out[byte0 + lane] = uint8_t(m >> (8 * lane));
    }
}

// rows of an int32 matrix with an element outside [lo, hi] (apply_tc: query counts that do not fit 8 bits)
__global__ void __launch_bounds__(256) rows_out_of_range_kernel(const int32_t *__restrict__ X, int64_t rows, int64_t cols,
                                                                int32_t lo, int32_t hi, int32_t *__restrict__ out,
                                                                int64_t cap, unsigned long long *__restrict__ n_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        bool bad = false;
        for (int64_t c = lane; c < cols; c += 32) {
            const int32_t v = __ldg(X + r * cols + c);
            bad |= (v < lo || v > hi);
        }
        if (__any_sync(FULL, bad) && lane == 0) {
            const unsigned long long slot = atomicAdd(n_out, 1ull);
            if ((int64_t)slot < cap) out[slot] = int32_t(r);
        }
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

int skm_pack_counts_u8(const void *d_counts, int64_t rows, int64_t cols, int in_bits, uint8_t *d_out, int32_t *d_esc_row,
                       int32_t *d_esc_col, int32_t *d_esc_val, int64_t esc_capacity, int64_t *d_n_esc, skm_stream_t stream) {
    using namespace skm;
    if (rows < 0 || cols < 0 || (in_bits != 32 && in_bits != 16) || esc_capacity < 0 || rows >= (1ll << 31) || cols >= (1ll << 31)) {
        set_error("skm_pack_counts_u8: bad arguments (in_bits 32 or 16, rows / cols < 2^31)");
        return SKM_ERR_INVALID;
    }
    if (!d_n_esc) { set_error("skm_pack_counts_u8: NULL d_n_esc"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_n_esc, 0, 8, st));
    const int64_t total = rows * cols;
    if (total == 0) return SKM_OK;
    if (!d_counts || !d_out || (esc_capacity > 0 && (!d_esc_row || !d_esc_col || !d_esc_val))) { set_error("skm_pack_counts_u8: NULL argument"); return SKM_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(d_out) & 3u) != 0) { set_error("skm_pack_counts_u8: d_out must be 4-byte aligned"); return SKM_ERR_INVALID; }
    const int grid = (int)std::min<int64_t>(((total + 3) / 4 + 255) / 256, int64_t(sm_count()) * 16);
    auto *n = reinterpret_cast<unsigned long long *>(d_n_esc);
    if (in_bits == 32) pack_u8_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)d_counts, total, cols, d_out, d_esc_row, d_esc_col, d_esc_val, esc_capacity, n);
    else pack_u8_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)d_counts, total, cols, d_out, d_esc_row, d_esc_col, d_esc_val, esc_capacity, n);
    SKM_LAUNCH_CHECK("pack_u8_kernel");
    return SKM_OK;
}

int skm_pack_presence_bits(const void *d_counts, int64_t rows, int64_t cols, int in_bits, uint8_t *d_out, skm_stream_t stream) {
    using namespace skm;
    if (rows < 0 || cols < 0 || (in_bits != 32 && in_bits != 16)) { set_error("skm_pack_presence_bits: bad arguments (in_bits 32 or 16)"); return SKM_ERR_INVALID; }
    if (rows == 0 || cols == 0) return SKM_OK;
    if (!d_counts || !d_out) { set_error("skm_pack_presence_bits: NULL argument"); return SKM_ERR_INVALID; }
    const int64_t row_bytes = (cols + 7) >> 3;
    const int64_t nw = rows * ((cols + 31) >> 5);
    const int grid = (int)std::min<int64_t>((nw + 7) / 8, int64_t(sm_count()) * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (in_bits == 32) pack_bits_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)d_counts, rows, cols, row_bytes, d_out);
    else pack_bits_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)d_counts, rows, cols, row_bytes, d_out);
    SKM_LAUNCH_CHECK("pack_bits_kernel");
    return SKM_OK;
}

int skm_rows_out_of_range_i32(const int32_t *d_X, int64_t rows, int64_t cols, int32_t lo, int32_t hi, int32_t *d_rows_out,
                              int64_t capacity, int64_t *d_n_out, skm_stream_t stream) {
    using namespace skm;
    if (rows < 0 || cols < 0 || capacity < 0 || rows >= (1ll << 31) || !d_n_out) { set_error("skm_rows_out_of_range_i32: bad arguments"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_n_out, 0, 8, st));
    if (rows == 0 || cols == 0) return SKM_OK;
    if (!d_X || (capacity > 0 && !d_rows_out)) { set_error("skm_rows_out_of_range_i32: NULL argument"); return SKM_ERR_INVALID; }
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, int64_t(sm_count()) * 16);
    rows_out_of_range_kernel<<<grid, 256, 0, st>>>(d_X, rows, cols, lo, hi, d_rows_out, capacity, reinterpret_cast<unsigned long long *>(d_n_out));
    SKM_LAUNCH_CHECK("rows_out_of_range_kernel");
    return SKM_OK;
}

}  // extern "C"

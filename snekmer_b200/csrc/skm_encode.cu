// skm_encode.cu — kernel (a): packed-residue reader + LUT recode + rolling base-|A| codes.
//
// skm_encode_windows materialises the code of the window starting at every
// residue position (all-ones where invalid).  It is the stand-alone form of the
// enumerator that the fused kernels (basis / count / learn) inline; it exists
// for parity tests of reduce_vectorize (vectorize.py:292-328) and as the input
// stage of the sort-based sparse paths.
// Algorithmic bytes per residue: 1 read + sizeof(code) written (+8 per sequence).
#include "skm_common.cuh"

namespace skm {

template <typename CodeT, int NW>
__global__ void __launch_bounds__(256) encode_windows_kernel(const uint8_t *__restrict__ res, int64_t nres,
                                                             const int64_t *__restrict__ off, int64_t nseq,
                                                             const uint8_t *__restrict__ lut, int nsym, int k,
                                                             CodeT *__restrict__ out) {
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t s = warp; s < nseq; s += nwarps) {
        const int64_t b = __ldg(off + s), e = __ldg(off + s + 1);
        // positions whose window would run past the end (and whole short sequences)
        const int64_t tail0 = (e - b >= k) ? e - k + 1 : b;
        for (int64_t p = tail0 + lane; p < e; p += 32) out[p] = code_traits<CodeT>::none;
        warp_scan_sequence<CodeT, NW>(res, nres, b, e, s_lut, nsym, k,
                                      [&](int64_t g, CodeT code, bool ok) {
                                          if (g >= b && g <= e - k) out[g] = ok ? code : code_traits<CodeT>::none;
                                      });
    }
}

template <typename CodeT>
__global__ void fill_kernel(CodeT *out, int64_t n, CodeT v) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) out[i] = v;
}

__global__ void __launch_bounds__(256) reduce_bytes_kernel(const uint8_t *__restrict__ res, int64_t nres,
                                                           const uint8_t *__restrict__ cmap, uint8_t *__restrict__ out) {
    __shared__ uint8_t s_map[256];
    s_map[threadIdx.x] = cmap[threadIdx.x];
    __syncthreads();
    const int64_t nvec = nres >> 4;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(res) + i);
        uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t x = w[q], y = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) y |= uint32_t(s_map[(x >> (8 * j)) & 0xFF]) << (8 * j);
            w[q] = y;
        }
        reinterpret_cast<uint4 *>(out)[i] = v;
    }
    for (int64_t i = (nvec << 4) + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nres; i += stride)
        out[i] = s_map[res[i]];
}

}  // namespace skm

extern "C" {

int skm_encode_windows(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq,
                       const uint8_t *d_lut, int nsym, int k, int code_bits, void *d_codes_out,
                       skm_stream_t stream) {
    using namespace skm;
    int rc = check_common(d_residues, nres, d_offsets, nseq, d_lut, nsym, k);
    if (rc) return rc;
    if (code_bits != 32 && code_bits != 64) { set_error("code_bits must be 32 or 64"); return SKM_ERR_INVALID; }
    unsigned __int128 S;
    code_space(nsym, k, &S);
    if (code_bits == 32 && S >= ((unsigned __int128)1 << 32)) { set_error("nsym^k does not fit 32-bit codes"); return SKM_ERR_INVALID; }
    if (code_bits == 64 && S > (((unsigned __int128)1 << 64) - 1)) { set_error("nsym^k = 2^64 collides with the invalid sentinel"); return SKM_ERR_UNSUPPORTED; }
    if (nres == 0) return SKM_OK;
    if (!d_codes_out) { set_error("d_codes_out is NULL"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = sm_count() * 8;
    // positions that belong to no sequence (gaps before offsets[0] / after offsets[nseq]) read as invalid
    if (code_bits == 32) fill_kernel<uint32_t><<<grid, 256, 0, st>>>((uint32_t *)d_codes_out, nres, 0xFFFFFFFFu);
    else fill_kernel<uint64_t><<<grid, 256, 0, st>>>((uint64_t *)d_codes_out, nres, ~0ull);
    SKM_LAUNCH_CHECK("fill_kernel");
    if (nseq == 0) return SKM_OK;
    const int nw = neighbour_words(k);
    if (code_bits == 32) {
        SKM_DISPATCH_NW(nw, (encode_windows_kernel<uint32_t, NW><<<grid, 256, 0, st>>>(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, (uint32_t *)d_codes_out)));
    } else {
        SKM_DISPATCH_NW(nw, (encode_windows_kernel<uint64_t, NW><<<grid, 256, 0, st>>>(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, (uint64_t *)d_codes_out)));
    }
    SKM_LAUNCH_CHECK("encode_windows_kernel");
    return SKM_OK;
}

int skm_reduce_bytes(const uint8_t *d_residues, int64_t nres, const uint8_t *d_charmap, uint8_t *d_out,
                     skm_stream_t stream) {
    using namespace skm;
    if (nres < 0 || !d_charmap || (nres > 0 && (!d_residues || !d_out))) { set_error("skm_reduce_bytes: bad arguments"); return SKM_ERR_INVALID; }
    if (((reinterpret_cast<uintptr_t>(d_residues) | reinterpret_cast<uintptr_t>(d_out)) & 15u) != 0) { set_error("skm_reduce_bytes: buffers must be 16-byte aligned"); return SKM_ERR_INVALID; }
    if (nres == 0) return SKM_OK;
    reduce_bytes_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(d_residues, nres, d_charmap, d_out);
    SKM_LAUNCH_CHECK("reduce_bytes_kernel");
    return SKM_OK;
}

}  // extern "C"

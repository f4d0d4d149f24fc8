// skm_count.cu — kernel (b): per-sequence k-mer counts over the basis.
//
// Dense path (skm_count_dense) — replaces the per-sequence loops of
// kmerize.smk:117-120 (np.isin presence), learn.smk:359-383 and
// apply.smk:195-206 (dict counts gathered in basis order):
//   a CTA owns a tile of T consecutive sequences and keeps all T count rows in
//   shared memory as packed 16-bit (or 32-bit) counters; its warps pull
//   sequences off a shared-memory ticket, walk them with warp_scan_sequence and
//   bump row[col_of_code[code]] with shared-memory atomics; the tile is then
//   streamed to HBM as ONE contiguous range (rows are adjacent in the [N,K]
//   output) with 16-byte stores while the counters are re-zeroed in place.
//   HBM traffic is the algorithmic minimum: every residue read once, every
//   output element written once.  Algorithmic bytes per sequence:
//   L + 8 + K*sizeof(out).
//
// Sparse path (skm_count_csr): window codes -> columns, segmented radix sort of
// each sequence's keys, run-length encode into CSR.
#include <cub/cub.cuh>

#include "skm_common.cuh"

namespace skm {

// ---------------------------------------------------------------------------
// dense
// ---------------------------------------------------------------------------
template <typename CntT> struct cnt_ops;
template <> struct cnt_ops<uint16_t> {
    // two 16-bit counters per 32-bit word; no carry while a count stays < 65536
    static __device__ __forceinline__ void add(uint32_t *words, uint32_t idx) {
        atomicAdd(words + (idx >> 1), 1u << ((idx & 1u) * 16u));
    }
};
template <> struct cnt_ops<uint32_t> {
    static __device__ __forceinline__ void add(uint32_t *words, uint32_t idx) { atomicAdd(words + idx, 1u); }
};

template <typename CntT, typename OutT>
__device__ __forceinline__ void flush_tile(uint32_t *s_words, OutT *__restrict__ out, int64_t n_elems) {
    // s_words holds n_elems counters of CntT, out is the contiguous destination (element 0 of the tile).
    CntT *s_cnt = reinterpret_cast<CntT *>(s_words);
    constexpr int VEC = 16 / sizeof(OutT);          // output elements per 16-byte store
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (aligned) {
        const int64_t nvec = n_elems / VEC;
        for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
            OutT tmp[VEC];
            if constexpr (sizeof(CntT) == 2 && VEC == 4) {
                const uint2 c = *reinterpret_cast<const uint2 *>(s_cnt + v * 4);
                *reinterpret_cast<uint2 *>(s_cnt + v * 4) = make_uint2(0u, 0u);
                tmp[0] = OutT(c.x & 0xFFFFu); tmp[1] = OutT(c.x >> 16);
                tmp[2] = OutT(c.y & 0xFFFFu); tmp[3] = OutT(c.y >> 16);
            } else if constexpr (sizeof(CntT) == 2 && VEC == 8) {
                const uint4 c = *reinterpret_cast<const uint4 *>(s_cnt + v * 8);
                *reinterpret_cast<uint4 *>(s_cnt + v * 8) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4 *>(tmp) = c;
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) { tmp[j] = OutT(s_cnt[v * VEC + j]); s_cnt[v * VEC + j] = 0; }
            }
            __stcs(reinterpret_cast<uint4 *>(out) + v, *reinterpret_cast<const uint4 *>(tmp));
        }
        for (int64_t i = nvec * VEC + threadIdx.x; i < n_elems; i += blockDim.x) { out[i] = OutT(s_cnt[i]); s_cnt[i] = 0; }
    } else {
        for (int64_t i = threadIdx.x; i < n_elems; i += blockDim.x) { out[i] = OutT(s_cnt[i]); s_cnt[i] = 0; }
    }
}

template <typename CodeT, int NW, typename CntT, typename OutT>
__global__ void __launch_bounds__(256) count_dense_kernel(const uint8_t *__restrict__ res, int64_t nres,
                                                          const int64_t *__restrict__ off, int64_t nseq,
                                                          const uint8_t *__restrict__ lut, int nsym, int k,
                                                          const int32_t *__restrict__ col_of_code, int K, int T,
                                                          OutT *__restrict__ out) {
    extern __shared__ __align__(16) uint32_t s_words[];   // T*K counters of CntT (rounded up to 16 B)
    __shared__ uint8_t s_lut[256];
    __shared__ unsigned int s_next;
    s_lut[threadIdx.x] = lut[threadIdx.x];
    const int64_t tile_elems = int64_t(T) * K;
    const int n_words = int((tile_elems * sizeof(CntT) + 3) / 4);
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) s_words[i] = 0;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t ntiles = (nseq + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t s0 = tile * T;
        const int rows = (nseq - s0 < T) ? int(nseq - s0) : T;
        for (;;) {
            unsigned int r = 0;
            if (lane == 0) r = atomicAdd(&s_next, 1u);
            r = __shfl_sync(FULL, r, 0);
            if (r >= (unsigned)rows) break;
            const int64_t b = __ldg(off + s0 + r), e = __ldg(off + s0 + r + 1);
            const uint32_t row_base = r * uint32_t(K);
            warp_scan_sequence<CodeT, NW>(res, nres, b, e, s_lut, nsym, k, [&](int64_t, CodeT code, bool ok) {
                if (ok) {
                    const int32_t col = col_of_code ? __ldg(col_of_code + code) : int32_t(code);
                    if (col >= 0) cnt_ops<CntT>::add(s_words, row_base + uint32_t(col));
                }
            });
        }
        __syncthreads();
        if (threadIdx.x == 0) s_next = 0;
        flush_tile<CntT, OutT>(s_words, out + s0 * K, int64_t(rows) * K);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// sparse (CSR)
// ---------------------------------------------------------------------------
// keys[p] = column (or code) of the window starting at p, all-ones if invalid / not in the basis
template <int NW>
__global__ void __launch_bounds__(256) csr_keys_kernel(const uint8_t *__restrict__ res, int64_t nres,
                                                       const int64_t *__restrict__ off, int64_t nseq,
                                                       const uint8_t *__restrict__ lut, int nsym, int k,
                                                       const int32_t *__restrict__ col_of_code, int64_t res0,
                                                       uint32_t *__restrict__ keys) {
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t s = warp; s < nseq; s += nwarps) {
        const int64_t b = __ldg(off + s), e = __ldg(off + s + 1);
        const int64_t tail0 = (e - b >= k) ? e - k + 1 : b;
        for (int64_t p = tail0 + lane; p < e; p += 32) keys[p - res0] = 0xFFFFFFFFu;
        warp_scan_sequence<uint32_t, NW>(res, nres, b, e, s_lut, nsym, k, [&](int64_t g, uint32_t code, bool ok) {
            if (g >= b && g <= e - k) {
                uint32_t key = 0xFFFFFFFFu;
                if (ok) {
                    if (col_of_code) { const int32_t c = __ldg(col_of_code + code); if (c >= 0) key = uint32_t(c); }
                    else key = code;
                }
                keys[g - res0] = key;
            }
        });
    }
}

// one warp per sequence over its sorted keys: count (and optionally write) the runs
__global__ void __launch_bounds__(256) csr_runs_kernel(const uint32_t *__restrict__ keys, const int64_t *__restrict__ off,
                                                       int64_t seq0, int64_t nseq, int64_t res0,
                                                       int64_t *__restrict__ rowptr, uint32_t *__restrict__ cols,
                                                       int32_t *__restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t s = seq0 + warp; s < seq0 + nseq; s += nwarps) {
        const int64_t b = __ldg(off + s) - res0, e = __ldg(off + s + 1) - res0;
        int64_t nrun = 0;
        const int64_t wbase = cols ? rowptr[s] : 0;
        for (int64_t p0 = b; p0 < e; p0 += 32) {
            const int64_t p = p0 + lane;
            const uint32_t key = (p < e) ? keys[p] : 0xFFFFFFFFu;
            const uint32_t prevk = (p > b && p < e) ? keys[p - 1] : 0xFFFFFFFFu;
            const bool head = (p < e) && key != 0xFFFFFFFFu && (p == b || key != prevk);
            const unsigned m = __ballot_sync(FULL, head);
            if (cols && head) {
                // run length: scan forward (runs are short: a k-mer repeated within one protein)
                int64_t q = p + 1;
                while (q < e && keys[q] == key) ++q;
                const int64_t slot = wbase + nrun + __popc(m & ((1u << lane) - 1u));
                cols[slot] = key;
                vals[slot] = int32_t(q - p);
            }
            nrun += __popc(m);
        }
        if (!cols && lane == 0) rowptr[s + 1] = nrun;
    }
}

}  // namespace skm

extern "C" {

int skm_count_dense(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq,
                    const uint8_t *d_lut, int nsym, int k, const int32_t *d_col_of_code, int64_t S, int64_t K,
                    int out_bits, void *d_counts, int64_t max_len, skm_stream_t stream) {
    using namespace skm;
    int rc = check_common(d_residues, nres, d_offsets, nseq, d_lut, nsym, k);
    if (rc) return rc;
    unsigned __int128 S128;
    code_space(nsym, k, &S128);
    if (S128 > (unsigned __int128)SKM_DENSE_MAX_SPACE || (int64_t)S128 != S) {
        set_error("skm_count_dense: S=%lld must equal nsym^k and be <= 2^27", (long long)S);
        return S128 > (unsigned __int128)SKM_DENSE_MAX_SPACE ? SKM_ERR_UNSUPPORTED : SKM_ERR_INVALID;
    }
    if (out_bits != 32 && out_bits != 16) { set_error("skm_count_dense: out_bits must be 32 or 16"); return SKM_ERR_INVALID; }
    if (!d_col_of_code && K != S) { set_error("skm_count_dense: identity basis needs K == S"); return SKM_ERR_INVALID; }
    if (K < 0 || K > S) { set_error("skm_count_dense: K=%lld out of range", (long long)K); return SKM_ERR_INVALID; }
    if (out_bits == 16 && (max_len <= 0 || max_len > 65535)) { set_error("skm_count_dense: uint16 output needs 0 < max_len <= 65535"); return SKM_ERR_INVALID; }
    if (nseq == 0 || K == 0) return SKM_OK;
    if (!d_counts) { set_error("skm_count_dense: d_counts is NULL"); return SKM_ERR_INVALID; }
    const bool cnt16 = (max_len > 0 && max_len <= 65535);
    const size_t cnt_bytes = cnt16 ? 2 : 4;
    const size_t smem_cap = 200 * 1024;
    if (size_t(K) * cnt_bytes > smem_cap) {
        set_error("skm_count_dense: K=%lld rows do not fit shared memory; use skm_count_csr", (long long)K);
        return SKM_ERR_UNSUPPORTED;
    }
    // tile: as many rows as fit ~64 KB (3 CTAs/SM), multiple of 8 so every tile start is 16-byte aligned
    int64_t T = int64_t(64 * 1024 / (size_t(K) * cnt_bytes));
    if (T >= 8) T &= ~int64_t(7);
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    const size_t smem = ((size_t(T) * K * cnt_bytes + 15) / 16) * 16;
    const int64_t ntiles = (nseq + T - 1) / T;
    const int per_sm = smem <= 72 * 1024 ? 3 : (smem <= 110 * 1024 ? 2 : 1);
    const int grid = (int)std::min<int64_t>(ntiles, int64_t(sm_count()) * per_sm);
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = neighbour_words(k);
#define SKM_LAUNCH_DENSE(CNT, OUT)                                                                                   \
    SKM_DISPATCH_NW(nw, {                                                                                            \
        auto kern = count_dense_kernel<uint32_t, NW, CNT, OUT>;                                                      \
        SKM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        kern<<<grid, 256, smem, st>>>(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, d_col_of_code, (int)K,      \
                                      (int)T, (OUT *)d_counts);                                                      \
    })
    if (cnt16 && out_bits == 32) { SKM_LAUNCH_DENSE(uint16_t, int32_t); }
    else if (cnt16 && out_bits == 16) { SKM_LAUNCH_DENSE(uint16_t, uint16_t); }
    else { SKM_LAUNCH_DENSE(uint32_t, int32_t); }
#undef SKM_LAUNCH_DENSE
    SKM_LAUNCH_CHECK("count_dense_kernel");
    return SKM_OK;
}

// cub::DeviceSegmentedSort takes 32-bit item counts: one call handles < 2^30 residues (callers shard above that).
static const int64_t CSR_MAX_RES = (1ll << 30);

static size_t csr_align(size_t x) { return (x + 255) & ~size_t(255); }

size_t skm_count_csr_workspace(int64_t nres, int64_t nseq) {
    if (nres <= 0 || nseq <= 0) return 256;
    const int64_t items = std::min<int64_t>(nres, CSR_MAX_RES);
    size_t t_sort = 0, t_scan = 0;
    cub::DeviceSegmentedSort::SortKeys(nullptr, t_sort, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)items,
                                       (int)std::min<int64_t>(nseq, (1ll << 31) - 1), (const int64_t *)nullptr,
                                       (const int64_t *)nullptr);
    cub::DeviceScan::InclusiveSum(nullptr, t_scan, (const int64_t *)nullptr, (int64_t *)nullptr, nseq);
    return 2 * csr_align((size_t)nres * 4) + csr_align(std::max(t_sort, t_scan)) + 1024;
}

int skm_count_csr(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq,
                  const uint8_t *d_lut, int nsym, int k, const int32_t *d_col_of_code, int64_t S,
                  int64_t *d_rowptr, uint32_t *d_cols, int32_t *d_vals, void *workspace, size_t workspace_bytes,
                  skm_stream_t stream) {
    using namespace skm;
    int rc = check_common(d_residues, nres, d_offsets, nseq, d_lut, nsym, k);
    if (rc) return rc;
    unsigned __int128 S128;
    code_space(nsym, k, &S128);
    if (S128 >= ((unsigned __int128)1 << 32)) { set_error("skm_count_csr: code space must be < 2^32 (32-bit keys)"); return SKM_ERR_UNSUPPORTED; }
    if (d_col_of_code && (int64_t)S128 != S) { set_error("skm_count_csr: S must equal nsym^k"); return SKM_ERR_INVALID; }
    if (!d_rowptr) { set_error("skm_count_csr: d_rowptr is NULL"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_rowptr, 0, 8 * (size_t)(nseq + 1), st));
    if (nseq == 0 || nres == 0) return SKM_OK;
    if (!d_cols || !d_vals) { set_error("skm_count_csr: d_cols / d_vals are NULL"); return SKM_ERR_INVALID; }
    if (nres >= CSR_MAX_RES || nseq >= (1ll << 31)) { set_error("skm_count_csr: more than 2^30 residues per call; split the shard"); return SKM_ERR_UNSUPPORTED; }
    const size_t need = skm_count_csr_workspace(nres, nseq);
    if (!workspace || workspace_bytes < need) { set_error("skm_count_csr: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    const size_t seg = csr_align((size_t)nres * 4);
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    uint32_t *keys_a = (uint32_t *)p, *keys_b = (uint32_t *)(p + seg);
    void *temp = p + 2 * seg;
    size_t temp_bytes = workspace_bytes - (size_t)((char *)temp - (char *)workspace);
    const int grid = sm_count() * 8;
    const int nw = neighbour_words(k);
    // keys are indexed by absolute position in d_residues (positions outside every sequence are never read)
    SKM_DISPATCH_NW(nw, (csr_keys_kernel<NW><<<grid, 256, 0, st>>>(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, d_col_of_code, 0, keys_a)));
    SKM_LAUNCH_CHECK("csr_keys_kernel");
    SKM_CUDA_TRY(cub::DeviceSegmentedSort::SortKeys(temp, temp_bytes, keys_a, keys_b, (int)nres, (int)nseq, d_offsets,
                                                    d_offsets + 1, st));
    csr_runs_kernel<<<grid, 256, 0, st>>>(keys_b, d_offsets, 0, nseq, 0, d_rowptr, nullptr, nullptr);
    SKM_LAUNCH_CHECK("csr_runs_kernel(count)");
    temp_bytes = workspace_bytes - (size_t)((char *)temp - (char *)workspace);
    SKM_CUDA_TRY(cub::DeviceScan::InclusiveSum(temp, temp_bytes, d_rowptr + 1, d_rowptr + 1, nseq, st));
    csr_runs_kernel<<<grid, 256, 0, st>>>(keys_b, d_offsets, 0, nseq, 0, d_rowptr, d_cols, d_vals);
    SKM_LAUNCH_CHECK("csr_runs_kernel(fill)");
    return SKM_OK;
}

}  // extern "C"

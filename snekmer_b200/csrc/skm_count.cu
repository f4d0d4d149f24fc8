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
#include <cstdlib>
#include <cstring>

#include <cub/cub.cuh>

#include "skm_common.cuh"
#include "skm_tile.cuh"

namespace skm {

// ---------------------------------------------------------------------------
// dense
// ---------------------------------------------------------------------------
constexpr int CD_MAX_ROWS = 32;   // rows (sequences) per tile

// counters live in shared memory in the OUTPUT layout (int32, or uint16 packed two per word), so a tile
// is flushed with one bulk shared->global copy
template <typename OutT> struct cnt_ops;
template <> struct cnt_ops<int32_t> {
    static __device__ __forceinline__ void add(uint32_t base, uint32_t idx) { reds_add_u32(base + (idx << 2), 1u); }
};
template <> struct cnt_ops<uint16_t> {
    // no carry into the upper half while a count stays < 65536 (host: max_len <= 65535)
    static __device__ __forceinline__ void add(uint32_t base, uint32_t idx) {
        reds_add_u32(base + ((idx >> 1) << 2), 1u << ((idx & 1u) << 4));
    }
};

// smem: [T*K counters of OutT | symbol buffer for seg_cap residues]
template <typename OutT, bool HAS_MAP>
__global__ void __launch_bounds__(TS_THREADS) count_dense_kernel(const uint8_t *__restrict__ res, int64_t nres,
                                                                 const int64_t *__restrict__ off, int64_t nseq,
                                                                 const uint8_t *__restrict__ lut, uint32_t nsym, int k,
                                                                 uint32_t pow_k1, const int32_t *__restrict__ col_of_code,
                                                                 int K, int T, int seg_cap, OutT *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ uint8_t s_lut[256];
    __shared__ int32_t s_off[CD_MAX_ROWS + 1];        // row starts relative to the tile
    const int64_t tile_elems = int64_t(T) * K;
    const uint32_t cnt_bytes = uint32_t((size_t(tile_elems) * sizeof(OutT) + 15) & ~size_t(15));
    uint4 *s_cnt4 = reinterpret_cast<uint4 *>(s_raw);
    uint8_t *s_sym = s_raw + cnt_bytes;
    uint32_t cnt_addr = smem_addr(s_raw), sym_addr = smem_addr(s_sym);
    asm volatile("" : "+r"(cnt_addr), "+r"(sym_addr));      // keep the window addresses in registers
    ts_lut_init(s_lut, lut);
    for (int i = threadIdx.x; i < int(cnt_bytes >> 4); i += blockDim.x) s_cnt4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const int tid = threadIdx.x;
    const int64_t ntiles = (nseq + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t s0 = tile * T;
        const int rows = (nseq - s0 < T) ? int(nseq - s0) : T;
        const int64_t r0 = __ldg(off + s0), r1 = __ldg(off + s0 + rows);
        for (int r = tid; r <= rows; r += blockDim.x) s_off[r] = int32_t(__ldg(off + s0 + r) - r0);
        bool first = true;
        uint32_t tail = 0;
        for (int64_t a = r0; a < r1;) {
            const int64_t b = min(r1, (a + seg_cap) & ~int64_t(15));
            if (!first) ts_tail_write(s_sym, tail, k);
            const TsSeg g = ts_stage(res, nres, a, b, s_lut, s_sym);
            __syncthreads();
            if (first) ts_invalidate_front(s_sym, g);
            const int a_rel = int(a - r0), b_rel = int(b - r0);
            for (int r = tid; r < rows; r += blockDim.x) {
                const int o = s_off[r];
                if (o >= a_rel && o < b_rel) s_sym[g.lo + (o - a_rel)] |= uint8_t(SYM_FLAG);
            }
            __syncthreads();
            const int C = ts_chunk(g.hi - g.lo);
            const int i0 = g.lo + tid * C, i1 = min(i0 + C, g.hi);
            if (i0 < i1) {
                const int shift = a_rel - g.lo;                 // tile-relative position = symbol index + shift
                const int rel0 = i0 + shift;
                int lo = 0, hi = rows;                          // rows that start strictly before rel0
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_off[mid] < rel0) lo = mid + 1; else hi = mid;
                }
                int row = lo - 1;                               // -1 only when i0 is the (flagged) tile start
                uint32_t row_base = uint32_t(row) * uint32_t(K);
                const int shift_addr = shift - int(sym_addr);   // tile-relative position = shared address + shift_addr
                ts_scan_chunk<uint32_t>(
                    sym_addr, i0, i1, k, nsym, pow_k1,
                    [&](uint32_t, uint32_t code, bool ok) {
                        int32_t col = -1;
                        if (ok) col = HAS_MAP ? __ldg(col_of_code + code) : int32_t(code);
                        if (col >= 0) cnt_ops<OutT>::add(cnt_addr, row_base + uint32_t(col));
                    },
                    [&](uint32_t p) {
                        const int rel = int(p) + shift_addr;
                        do { ++row; } while (s_off[row + 1] <= rel);   // skips empty sequences; s_off[rows] > rel
                        row_base = uint32_t(row) * uint32_t(K);
                    });
            }
            tail = ts_tail_read(s_sym, g, k);
            first = false;
            a = b;
            __syncthreads();
        }
        // ---- flush: the tile is one contiguous range of the [N, K] output ----
        ts_bulk_fence();
        __syncthreads();
        OutT *dst = out + s0 * K;
        const int64_t n_elems = int64_t(rows) * K;
        const OutT *s_cnt = reinterpret_cast<const OutT *>(s_raw);
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            const uint32_t bytes = uint32_t(n_elems * sizeof(OutT)) & ~15u;
            if (tid == 0 && bytes) ts_bulk_store(dst, cnt_addr, bytes);
            for (int64_t i = bytes / sizeof(OutT) + tid; i < n_elems; i += blockDim.x) dst[i] = s_cnt[i];
            if (tid == 0) ts_bulk_wait_read();
        } else {
            for (int64_t i = tid; i < n_elems; i += blockDim.x) dst[i] = s_cnt[i];
        }
        __syncthreads();
        for (int i = tid; i < int(cnt_bytes >> 4); i += blockDim.x) s_cnt4[i] = make_uint4(0u, 0u, 0u, 0u);
        // the next tile touches the counters only after two more barriers
    }
}

// ---------------------------------------------------------------------------
// dense, one warp per sequence (rows up to ~9 KB)
// ---------------------------------------------------------------------------
// The tile kernel above synchronises the whole CTA four times per tile (stage / flag / scan / flush): at K = 1000
// it spends 28 % of its time in barriers and reaches 61 % of HBM.  Here every WARP is its own pipeline: it owns a
// contiguous, residue-balanced range of sequences, a private counter row (the output layout) and a private symbol
// buffer in shared memory, and only ever executes __syncwarp.  A sequence is walked in segments of 384 residues
// (12 per lane; the mean protein fits one segment); the finished row leaves with one bulk shared->global copy.
// ~48 independent warps per SM hide each other's load, bulk-store and shared-atomic latencies.
constexpr int CW_WARPS = 8;
constexpr int CW_C = 12;                          // residues per lane and segment (4 * odd: conflict-free byte reads)
constexpr int CW_SEG = 32 * CW_C;
constexpr int CW_SYM = ts_sym_bytes(CW_SEG);      // 480

// MAP: 0 = identity basis (column = code), 1 = col_of_code staged in shared memory as uint16 columns (S <= 16384;
// filtered codes point at a dummy counter behind the row, so the scan needs no validity test), 2 = col_of_code read
// from global memory (large code spaces).
// The scan is written out here rather than through ts_scan_chunk: one sequence per warp needs no boundary handling,
// symbols are used raw (an invalid symbol's garbage digit is added and later subtracted with the same value, and no
// window containing it is ever emitted), and the column comes from shared memory: 13 SASS instructions per residue
// against 24 for the generic scanner.
template <typename OutT, int MAP>
__device__ __forceinline__ void cw_emit(uint32_t code, uint32_t cnt_addr, uint32_t col_addr, const int32_t *__restrict__ col_of_code, int K) {
    uint32_t col;
    if (MAP == 0) col = code;
    else if (MAP == 1) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(col) : "r"(col_addr + 2u * code));
    else { const int32_t c = __ldg(col_of_code + code); col = c >= 0 ? uint32_t(c) : uint32_t(K); }
    cnt_ops<OutT>::add(cnt_addr, col);
}

// Branch-free emission: an invalid window is counted into the dummy counter behind the row (MAP 0: column K; MAP 1:
// the extra map entry S points at it), so the compiler can issue LDS.U16 + ATOMS without a branch around them.
template <typename OutT, int MAP>
__device__ __forceinline__ void cw_emit_sel(uint32_t code, bool ok, uint32_t cnt_addr, uint32_t col_addr,
                                            const int32_t *__restrict__ col_of_code, int S, int K) {
    if (MAP == 2) {
        if (ok) cw_emit<OutT, 2>(code, cnt_addr, col_addr, col_of_code, K);
        return;
    }
    uint32_t col;
    if (MAP == 0) col = ok ? code : uint32_t(K);
    else asm volatile("ld.shared.u16 %0, [%1];" : "=r"(col) : "r"(col_addr + 2u * (ok ? code : uint32_t(S))));
    cnt_ops<OutT>::add(cnt_addr, col);
}

// First version of the warp kernel (round 1): translated symbols staged as bytes, two LDS.U8 per residue, a run
// counter.  Still the path for k - 1 > 4 (and the A/B reference: SKM_CDW_BYTES=1).
template <typename OutT, int MAP>
__global__ void __launch_bounds__(32 * CW_WARPS, 6)
count_dense_warp_bytes_kernel(const uint8_t *__restrict__ res, int64_t nres, const int64_t *__restrict__ off, int64_t nseq,
                              const uint8_t *__restrict__ lut, uint32_t nsym, int k, uint32_t pow_k1,
                              const int32_t *__restrict__ col_of_code, int S, int K, uint32_t row_bytes, uint32_t map_bytes, int bulk_ok,
                              OutT *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ uint8_t s_lut[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t *s_col = reinterpret_cast<uint16_t *>(s_raw);                       // [S] when MAP == 1
    uint8_t *mine = s_raw + map_bytes + size_t(warp) * (row_bytes + CW_SYM);
    uint4 *s_cnt4 = reinterpret_cast<uint4 *>(mine);
    uint8_t *s_sym = mine + row_bytes;
    uint32_t cnt_addr = smem_addr(mine), sym_addr = smem_addr(s_sym), col_addr = smem_addr(s_col);
    asm volatile("" : "+r"(cnt_addr), "+r"(sym_addr), "+r"(col_addr));      // keep the window addresses in registers
    ts_lut_init(s_lut, lut);
    if (MAP == 1)
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            const int32_t c = __ldg(col_of_code + i);
            s_col[i] = uint16_t(c >= 0 ? c : K);                                // K = the dummy counter
        }
    for (int i = lane; i < int(row_bytes >> 4); i += 32) s_cnt4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                                        // LUT and column map; the only CTA-wide barrier
    // this warp's sequences: those that start in the w-th 1/W of the residue buffer
    int64_t lo = 0, hi = 0;
    if (lane == 0) {
        const int64_t W = int64_t(gridDim.x) * CW_WARPS, w = int64_t(blockIdx.x) * CW_WARPS + warp;
        const int64_t r0 = __ldg(off), span = __ldg(off + nseq) - r0;
        lo = (w == 0) ? 0 : lower_bound_off(off, nseq, r0 + (int64_t)(((__int128)span * w) / W));
        hi = (w + 1 == W) ? nseq : lower_bound_off(off, nseq, r0 + (int64_t)(((__int128)span * (w + 1)) / W));
    }
    lo = __shfl_sync(FULL, lo, 0);
    hi = __shfl_sync(FULL, hi, 0);
    const uint32_t uk = uint32_t(k);
    int64_t e = (lo < hi) ? __ldg(off + lo) : 0;
    for (int64_t s = lo; s < hi; ++s) {
        const int64_t b = e;
        e = __ldg(off + s + 1);
        bool first = true;
        uint32_t tail0 = 0, tail1 = 0;                      // the last k-1 symbols of the previous segment (k-1 <= 63)
        for (int64_t a = b; a < e;) {
            const int64_t a2 = min(e, (a + CW_SEG) & ~int64_t(15));
            if (!first) {
                if (lane < k - 1) s_sym[TS_PAD - (k - 1) + lane] = uint8_t(tail0);
                if (lane + 32 < k - 1) s_sym[TS_PAD - (k - 1) + lane + 32] = uint8_t(tail1);
            }
            // stage: [a & ~15, a2) translated into s_sym[TS_PAD ...), one 16-byte vector per lane (at most 25)
            const int64_t base = a & ~int64_t(15);
            const int lo_i = TS_PAD + int(a - base), hi_i = lo_i + int(a2 - a);
            const int nvec = int((a2 - base + 15) >> 4);
            if (lane < nvec) {
                const int64_t p = base + 16 * int64_t(lane);
                uint4 x;
                if (p + 16 <= nres) {
                    x = __ldg(reinterpret_cast<const uint4 *>(res + p));
                } else {
                    uint32_t w4[4] = {0, 0, 0, 0};
                    for (int j = 0; j < 16; ++j)
                        if (p + j < nres) w4[j >> 2] |= uint32_t(res[p + j]) << (8 * (j & 3));
                    x = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
                x.x = ts_translate4(x.x, s_lut);
                x.y = ts_translate4(x.y, s_lut);
                x.z = ts_translate4(x.z, s_lut);
                x.w = ts_translate4(x.w, s_lut);
                reinterpret_cast<uint4 *>(s_sym + TS_PAD)[lane] = x;
            }
            __syncwarp();
            if (first) {                                    // nothing in front of the sequence start is a residue of it
                for (int i = lane; i < lo_i; i += 32) s_sym[i] = uint8_t(SYM_BAD);
                __syncwarp();
            }
            // scan: lane owns symbols [i0, i1), warms up on the k-1 symbols in front of them
            const int n = hi_i - lo_i;
            const int C = ((((n + 31) >> 5) + 3) >> 2 | 1) << 2;     // smallest 4 * odd >= ceil(n / 32)
            const int i0 = lo_i + lane * C, i1 = min(i0 + C, hi_i);
            if (i0 < i1) {
                uint32_t run = 0, code = 0;
                uint32_t p = sym_addr + uint32_t(i0) - (uk - 1u);
                for (uint32_t j = 1; j < uk; ++j, ++p) {
                    const uint32_t sy = lds_u8<0>(p);
                    run = (sy >= SYM_BAD) ? 0u : run + 1u;
                    code = code * nsym + sy;
                }
                const uint32_t pend = sym_addr + uint32_t(i1);
                const uint32_t back = uk - 1u;
#pragma unroll 4
                for (; p < pend; ++p) {
                    const uint32_t sy = lds_u8<0>(p);
                    run = (sy >= SYM_BAD) ? 0u : run + 1u;
                    code = code * nsym + sy;
                    if (run >= uk) cw_emit<OutT, MAP>(code, cnt_addr, col_addr, col_of_code, K);
                    code -= lds_u8<0>(p - back) * pow_k1;
                }
            }
            __syncwarp();
            if (lane < k - 1) tail0 = s_sym[hi_i - (k - 1) + lane];
            if (lane + 32 < k - 1) tail1 = s_sym[hi_i - (k - 1) + lane + 32];
            __syncwarp();
            first = false;
            a = a2;
        }
        // ---- flush: the row is one contiguous K * sizeof(OutT) range of the output ----
        OutT *dst = out + s * K;
        if (bulk_ok) {
            ts_bulk_fence();
            __syncwarp();
            if (lane == 0) { ts_bulk_store(dst, cnt_addr, uint32_t(K) * uint32_t(sizeof(OutT))); ts_bulk_wait_read(); }
            __syncwarp();
        } else {
            __syncwarp();
            const OutT *s_cnt = reinterpret_cast<const OutT *>(mine);
            for (int i = lane; i < K; i += 32) dst[i] = s_cnt[i];
            __syncwarp();
        }
        for (int i = lane; i < int(row_bytes >> 4); i += 32) s_cnt4[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
    }
}

// Round-2 warp kernel (k - 1 <= 4, i.e. every dense basis of practical size).  What changed against the first version
// (profiles/r1m_* -> R2b_* -> R2c_*): 700 warp instructions and 221 shared-memory wavefronts per sequence, 59 % of the
// instructions OUTSIDE the scan loop (staging the translated symbols: 96; 64-bit position arithmetic per sequence and
// segment: ~170).
//   * the RAW residue bytes are staged (one LDG.128 + STS.128 per lane) and translated inside the scan, one LDS.U8 from
//     the LUT per residue, straight into the register that feeds the rolling code: no translate-and-repack pass;
//   * a lane owns a 4-ALIGNED chunk and reads it as 32-bit words; the translated previous word stays in a register and
//     one funnel shift of (previous : current) gives the four symbols that leave the window: no second read per residue;
//   * no run counter and no branch: the 8 validity bits of (previous : current) are tested against a k-bit window mask
//     and the emission is predicated (a branch on "any invalid byte" diverges in 54 % of the warp iterations — first
//     word, last word, an X somewhere in 192 bytes — and made the first word-based version execute MORE instructions);
//   * positions are 32-bit offsets from the warp's own (16-byte aligned) base, and the offsets of 31 sequences are
//     fetched with one coalesced load and handed out by shuffles;
// Measured and dropped (profiles/R2e_*): re-zeroing the flushed row with the copy engine instead of 8 STS.128 per lane
// (zeroing is 32 of the 189 LSU wavefronts per sequence and the LSU wavefront pipe is what bounds the kernel: 85 % busy,
// the SRAM banks 14 %).  cp.async.bulk shared -> shared from a zero row of the CTA on a per-warp mbarrier needs a
// cluster launch (illegal instruction otherwise) and ran at 1.21 ms against 0.85 ms; global (L2) -> shared 0.92 ms.
template <typename OutT, int MAP>
__global__ void __launch_bounds__(32 * CW_WARPS, 6)
count_dense_warp_kernel(const uint8_t *__restrict__ res, int64_t nres, const int64_t *__restrict__ off, int64_t nseq,
                        const uint8_t *__restrict__ lut, uint32_t nsym, int k, uint32_t pow_k1,
                        const int32_t *__restrict__ col_of_code, int S, int K, uint32_t row_bytes, uint32_t map_bytes, int bulk_ok,
                        OutT *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(16) uint8_t s_lut[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t *s_col = reinterpret_cast<uint16_t *>(s_raw);                       // [S] when MAP == 1
    uint8_t *mine = s_raw + map_bytes + size_t(warp) * (row_bytes + CW_SYM);
    uint4 *s_cnt4 = reinterpret_cast<uint4 *>(mine);
    uint8_t *s_res = mine + row_bytes;                                           // raw residue bytes of the segment
    uint32_t cnt_addr = smem_addr(mine), raw_addr = smem_addr(s_res), col_addr = smem_addr(s_col), lut_addr = smem_addr(s_lut);
    asm volatile("" : "+r"(cnt_addr), "+r"(raw_addr), "+r"(col_addr), "+r"(lut_addr));
    ts_lut_init(s_lut, lut);                                                     // invalid bytes -> SYM_BAD (0x40)
    if (threadIdx.x == 0) s_lut[0] = uint8_t(SYM_BAD);                           // byte 0 pads the staged segment: never a symbol
    if (MAP == 1)
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            const int32_t c = __ldg(col_of_code + i);
            s_col[i] = uint16_t(c >= 0 ? c : K);                                // K = the dummy counter
        }
    if (MAP == 1 && threadIdx.x == 0) s_col[S] = uint16_t(K);                   // entry S: where invalid windows go
    for (int i = lane; i < int(row_bytes >> 4); i += 32) s_cnt4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                                        // LUT and column map; the only CTA-wide barrier
    // this warp's sequences: those that start in the w-th 1/W of the residue buffer
    int64_t lo = 0, hi = 0;
    if (lane == 0) {
        const int64_t W = int64_t(gridDim.x) * CW_WARPS, w = int64_t(blockIdx.x) * CW_WARPS + warp;
        const int64_t r0 = __ldg(off), span = __ldg(off + nseq) - r0;
        lo = (w == 0) ? 0 : lower_bound_off(off, nseq, r0 + (int64_t)(((__int128)span * w) / W));
        hi = (w + 1 == W) ? nseq : lower_bound_off(off, nseq, r0 + (int64_t)(((__int128)span * (w + 1)) / W));
    }
    lo = __shfl_sync(FULL, lo, 0);
    hi = __shfl_sync(FULL, hi, 0);
    if (lo >= hi) return;
    // 32-bit positions relative to the 16-byte aligned base of this warp's range (alignment of a position is preserved)
    const int64_t r_base = __ldg(off + lo) & ~int64_t(15);
    const uint8_t *__restrict__ resw = res + r_base;
    const int64_t lim64 = nres - r_base;                    // bytes that may be read from resw
    const uint32_t lim = lim64 > int64_t(0xFFFFFFF0u) ? 0xFFFFFFF0u : uint32_t(lim64);
    const uint32_t uk = uint32_t(k), km1 = uk - 1u;
    const uint32_t out_shift = 32u - 8u * km1;              // (previous : current) >> out_shift = the four outgoing symbols
    const uint32_t win0 = ((1u << uk) - 1u) << (4u - km1);  // validity bits a window ending at byte 0 of `current` covers
    const uint32_t npow = 0u - pow_k1;
    uint32_t tailw = 0;
    for (int64_t s0 = lo; s0 < hi; s0 += 31) {
        // one coalesced load of 32 offsets: sequences s0 .. s0+30 (b = rel[i], e = rel[i+1])
        const int64_t si = min(s0 + lane, hi);
        const uint32_t rel = uint32_t(__ldg(off + si) - r_base);
        const int ns = (hi - s0 < 31) ? int(hi - s0) : 31;
        for (int i = 0; i < ns; ++i) {
            const uint32_t b = __shfl_sync(FULL, rel, i), e = __shfl_sync(FULL, rel, i + 1);
            bool first = true;
            for (uint32_t a = b; a < e;) {
                const uint32_t a2 = min(e, (a + uint32_t(CW_SEG)) & ~15u);
                const uint32_t base = a & ~15u;
                const int lo_i = TS_PAD + int(a - base), hi_i = lo_i + int(a2 - a);
                const int nvec = int((a2 - base + 15u) >> 4);
                // stage the raw bytes [base, a2) at s_res[TS_PAD ...): one 16-byte vector per lane (at most 25)
                if (lane < nvec) {
                    const uint32_t p = base + 16u * uint32_t(lane);
                    uint4 x;
                    if (p + 16u <= lim) {
                        x = __ldg(reinterpret_cast<const uint4 *>(resw + p));
                    } else {
                        uint32_t w4[4] = {0, 0, 0, 0};
                        for (int j = 0; j < 16; ++j)
                            if (p + uint32_t(j) < lim) w4[j >> 2] |= uint32_t(resw[p + j]) << (8 * (j & 3));
                        x = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    }
                    reinterpret_cast<uint4 *>(s_res + TS_PAD)[lane] = x;
                }
                if (!first && lane == 0) reinterpret_cast<uint32_t *>(s_res)[TS_PAD / 4 - 1] = tailw;   // the 4 bytes in front of this segment
                __syncwarp();
                // byte 0 translates to SYM_BAD: everything in front of the sequence and behind the segment's last word is invalid
                if (first) { const int f0 = (lo_i & ~3) - 4; if (f0 + lane < lo_i) s_res[f0 + lane] = 0; }   // previous word + front of the first word
                if (lane < ((4 - (hi_i & 3)) & 3)) s_res[hi_i + lane] = 0;
                __syncwarp();
                // lane owns cw words (cw odd -> conflict-free LDS.32) from word index w0
                const int A = lo_i & ~3;
                const int n = hi_i - A;
                const int cw = (((n + 31) >> 5) + 3) >> 2 | 1;
                const int w0 = (A >> 2) + lane * cw, w1 = min(w0 + cw, (hi_i + 3) >> 2);
                if (w0 < w1) {
                    uint32_t pa = raw_addr + 4u * uint32_t(w0);
                    uint32_t raw;
                    asm volatile("ld.shared.u32 %0, [%1+-4];" : "=r"(raw) : "r"(pa));
                    uint32_t t0, t1, t2, t3;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4440u)));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4441u)));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t2) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4442u)));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t3) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4443u)));
                    uint32_t prevT = t0 | (t1 << 8) | (t2 << 16) | (t3 << 24);
                    uint32_t pinv = (((prevT >> 6) & 0x01010101u) * 0x01020408u) >> 24;      // bit i: byte i of prevT is no symbol
                    uint32_t code = 0;
                    for (uint32_t t = 0; t < km1; ++t) code = code * nsym + ((prevT >> (out_shift + 8u * t)) & 0xFFu);
                    for (int w = w0; w < w1; ++w, pa += 4u) {
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw) : "r"(pa));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4440u)));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4441u)));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t2) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4442u)));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t3) : "r"(lut_addr + __byte_perm(raw, 0u, 0x4443u)));
                        const uint32_t curT = t0 | (t1 << 8) | (t2 << 16) | (t3 << 24);
                        const uint32_t outw = __funnelshift_rc(prevT, curT, out_shift);
                        const uint32_t cinv = (((curT >> 6) & 0x01010101u) * 0x01020408u) >> 24;
                        const uint32_t inv8 = pinv | (cinv << 4);
                        code = code * nsym + t0;
                        cw_emit_sel<OutT, MAP>(code, (inv8 & win0) == 0u, cnt_addr, col_addr, col_of_code, S, K);
                        code = __byte_perm(outw, 0u, 0x4440u) * npow + code;
                        code = code * nsym + t1;
                        cw_emit_sel<OutT, MAP>(code, (inv8 & (win0 << 1)) == 0u, cnt_addr, col_addr, col_of_code, S, K);
                        code = __byte_perm(outw, 0u, 0x4441u) * npow + code;
                        code = code * nsym + t2;
                        cw_emit_sel<OutT, MAP>(code, (inv8 & (win0 << 2)) == 0u, cnt_addr, col_addr, col_of_code, S, K);
                        code = __byte_perm(outw, 0u, 0x4442u) * npow + code;
                        code = code * nsym + t3;
                        cw_emit_sel<OutT, MAP>(code, (inv8 & (win0 << 3)) == 0u, cnt_addr, col_addr, col_of_code, S, K);
                        code = __byte_perm(outw, 0u, 0x4443u) * npow + code;
                        prevT = curT;
                        pinv = cinv;
                    }
                }
                __syncwarp();
                if (a2 < e && lane == 0) tailw = reinterpret_cast<const uint32_t *>(s_res)[(hi_i >> 2) - 1];   // hi_i is 16-aligned here
                __syncwarp();
                first = false;
                a = a2;
            }
            // ---- flush: the row is one contiguous K * sizeof(OutT) range of the output ----
            OutT *dst = out + (s0 + i) * K;
            if (bulk_ok) {
                ts_bulk_fence();
                __syncwarp();
                if (lane == 0) { ts_bulk_store(dst, cnt_addr, uint32_t(K) * uint32_t(sizeof(OutT))); ts_bulk_wait_read(); }
                __syncwarp();
            } else {
                __syncwarp();
                const OutT *s_cnt = reinterpret_cast<const OutT *>(mine);
                for (int j = lane; j < K; j += 32) dst[j] = s_cnt[j];
                __syncwarp();
            }
            for (int j = lane; j < int(row_bytes >> 4); j += 32) s_cnt4[j] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// sparse (CSR)
// ---------------------------------------------------------------------------
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

extern "C" int skm_window_keys_u32(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq,
                                   const uint8_t *d_lut, int nsym, int k, const int32_t *d_col_of_code, uint32_t *d_keys,
                                   cudaStream_t st);   // skm_sparse.cu

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
    if (!ts_supported(nsym, k)) { set_error("skm_count_dense: nsym=%d k=%d outside the kernel envelope (nsym <= %d, k <= %d)", nsym, k, TS_MAX_NSYM, TS_MAX_K); return SKM_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(d_counts) & 15u) != 0) { set_error("skm_count_dense: d_counts must be 16-byte aligned"); return SKM_ERR_INVALID; }
    const size_t out_bytes = out_bits / 8;
    const size_t smem_cap = 200 * 1024;
    if (size_t(K) * out_bytes > smem_cap) {
        set_error("skm_count_dense: K=%lld rows do not fit shared memory; use skm_count_csr", (long long)K);
        return SKM_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t pow_k1 = 1;
    for (int i = 0; i + 1 < k; ++i) pow_k1 *= (uint32_t)nsym;
    // Rows up to ~9 KB: one warp per sequence with a private counter row (>= 24 warps per SM).  One extra (dummy)
    // counter behind the row absorbs the windows whose code is filtered out of the basis.
    const uint32_t row_bytes = uint32_t((size_t(K + 1) * out_bytes + 15) & ~size_t(15));
    const char *force_tile = getenv("SKM_CD_TILE");
    if (row_bytes + CW_SYM <= 9 * 1024 + 256 && K < 65535 && !(force_tile && atoi(force_tile))) {
        const int map_mode = !d_col_of_code ? 0 : (S <= 16384 ? 1 : 2);
        const uint32_t map_bytes = map_mode == 1 ? uint32_t((size_t(S + 1) * 2 + 127) & ~size_t(127)) : 0u;     // + the entry for invalid windows
        const char *force_bytes = getenv("SKM_CDW_BYTES");                      // A/B switch: the byte-wise scan of round 1
        const bool words = k - 1 <= 4 && !(force_bytes && atoi(force_bytes));
        const size_t smem_w = map_bytes + size_t(CW_WARPS) * (row_bytes + CW_SYM);
        int per_sm_w = int((227 * 1024) / (smem_w + 1024 + 256));
        if (per_sm_w > 6) per_sm_w = 6;
        if (per_sm_w < 1) per_sm_w = 1;
        const int grid_w = (int)std::min<int64_t>((nseq + CW_WARPS - 1) / CW_WARPS, int64_t(sm_count()) * per_sm_w);
        const int bulk_ok = ((size_t(K) * out_bytes) % 16 == 0) ? 1 : 0;
#define SKM_LAUNCH_DENSE_W(OUT, MAP)                                                                                 \
    {                                                                                                                \
        if (!words) {                                                                                                \
            auto kern = count_dense_warp_bytes_kernel<OUT, MAP>;                                                     \
            SKM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));      \
            kern<<<grid_w, 32 * CW_WARPS, smem_w, st>>>(d_residues, nres, d_offsets, nseq, d_lut, (uint32_t)nsym, k, pow_k1, \
                                                        d_col_of_code, (int)S, (int)K, row_bytes, map_bytes, bulk_ok, (OUT *)d_counts); \
        } else {                                                                                                     \
            auto kern = count_dense_warp_kernel<OUT, MAP>;                                                           \
            SKM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));      \
            kern<<<grid_w, 32 * CW_WARPS, smem_w, st>>>(d_residues, nres, d_offsets, nseq, d_lut, (uint32_t)nsym, k, pow_k1, \
                                                        d_col_of_code, (int)S, (int)K, row_bytes, map_bytes, bulk_ok, (OUT *)d_counts); \
        }                                                                                                            \
    }
#define SKM_LAUNCH_DENSE_WM(OUT)                                                                                     \
    { if (map_mode == 0) SKM_LAUNCH_DENSE_W(OUT, 0) else if (map_mode == 1) SKM_LAUNCH_DENSE_W(OUT, 1) else SKM_LAUNCH_DENSE_W(OUT, 2) }
        if (out_bits == 32) SKM_LAUNCH_DENSE_WM(int32_t) else SKM_LAUNCH_DENSE_WM(uint16_t)
#undef SKM_LAUNCH_DENSE_WM
#undef SKM_LAUNCH_DENSE_W
        SKM_LAUNCH_CHECK("count_dense_warp_kernel");
        return SKM_OK;
    }
    // Tile shape.  Many small CTAs per SM (each with its own tile in flight) hide the staging-load, barrier
    // and bulk-store latencies of one another: 128 threads, ~16 KB of counters, rows a multiple of
    // 4 (int32) / 8 (uint16) so that every tile start is 16-byte aligned and leaves as one bulk copy.
    int64_t align_rows = 16 / (int64_t)out_bytes;           // rows per 16 bytes of output in the worst case ...
    while (align_rows > 1 && (size_t(align_rows / 2) * K * out_bytes) % 16 == 0) align_rows /= 2;   // ... fewer when K allows
    int threads = 128;
    int64_t T = int64_t((16 * 1024) / (size_t(K) * out_bytes));
    if (const char *e = getenv("SKM_CD_THREADS")) { const int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0) threads = v; }
    if (const char *e = getenv("SKM_CD_ROWS")) { const int v = atoi(e); if (v > 0) T = v; }
    if (T >= align_rows) T -= T % align_rows;
    if (T < align_rows) T = (size_t(align_rows) * K * out_bytes <= 64 * 1024) ? align_rows : 1;
    if (T > CD_MAX_ROWS) T = CD_MAX_ROWS;
    const int seg_cap = ts_seg_cap(28, threads);
    const size_t sym_bytes = (size_t)ts_sym_bytes(seg_cap);
    const size_t sm_bytes = 227 * 1024, per_cta_reserved = 1024 + 512;
    const size_t smem = ((size_t(T) * K * out_bytes + 15) & ~size_t(15)) + sym_bytes;
    const int64_t ntiles = (nseq + T - 1) / T;
    int per_sm = int(sm_bytes / (smem + per_cta_reserved));
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm > 32) per_sm = 32;
    if (per_sm < 1) per_sm = 1;
    const int grid = (int)std::min<int64_t>(ntiles, int64_t(sm_count()) * per_sm);
#define SKM_LAUNCH_DENSE(OUT, MAP)                                                                                   \
    {                                                                                                                \
        auto kern = count_dense_kernel<OUT, MAP>;                                                                    \
        SKM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        kern<<<grid, threads, smem, st>>>(d_residues, nres, d_offsets, nseq, d_lut, (uint32_t)nsym, k, pow_k1,       \
                                             d_col_of_code, (int)K, (int)T, seg_cap, (OUT *)d_counts);               \
    }
    if (out_bits == 32) { if (d_col_of_code) SKM_LAUNCH_DENSE(int32_t, true) else SKM_LAUNCH_DENSE(int32_t, false) }
    else { if (d_col_of_code) SKM_LAUNCH_DENSE(uint16_t, true) else SKM_LAUNCH_DENSE(uint16_t, false) }
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
    if (!ts_supported(nsym, k)) { set_error("skm_count_csr: nsym=%d k=%d outside the kernel envelope", nsym, k); return SKM_ERR_UNSUPPORTED; }
    const int grid = sm_count() * 8;
    // keys are indexed by the absolute position of the window's last residue (positions outside every sequence
    // are never read by the segmented sort)
    rc = skm_window_keys_u32(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, d_col_of_code, keys_a, st);
    if (rc) return rc;
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

// skm_tile.cuh — CTA-level segment scanner: the inner loop of kernels (a), (b), (b').
//
// A CTA walks a CONTIGUOUS range of the packed residue buffer in segments of at
// most `seg_cap` residues.  Per segment:
//   1. stage   residues are read once from HBM with aligned 16-byte loads,
//              translated through the 256-byte alphabet LUT and written to
//              shared memory as one symbol byte per residue
//                 bits 0-5  digit (index of the reduced symbol, 0 if invalid)
//                 bit  6    SYM_BAD   residue has no symbol (X, B, Z, *, lower case ...)
//                 bit  7    SYM_FLAG  first residue of a sequence
//   2. flag    the sequence starts that fall in the segment are marked;
//   3. scan    every thread owns a contiguous chunk of C positions (C = 4*odd,
//              so the 32 lanes of a warp hit 32 different banks) and rolls the
//              base-|A| code of the window ENDING at each position:
//                 code = code*|A| + digit[i]            -> emit if run >= k
//                 code -= digit[i-k+1] * |A|^(k-1)
//              `run` counts the valid residues since the last invalid residue
//              or sequence start, so windows never cross a sequence boundary
//              or an unmapped residue (vectorize.py:239-249).  A chunk warms up
//              on the k-1 symbols in front of it; the k-1 symbols in front of a
//              segment are carried over from the previous one in shared memory.
// The rare path (invalid residue / sequence start, ~0.4 % of positions) is a
// divergent branch; the common path is a dozen instructions per residue.
#pragma once

#include "skm_common.cuh"

namespace skm {

constexpr int TS_THREADS = 256;                    // upper bound of the CTA size of scanner kernels
constexpr int TS_PAD = 64;                         // >= k-1 halo symbols in front of the segment
constexpr int TS_MAX_K = TS_PAD;                   // k-1 <= 63
constexpr int TS_MAX_NSYM = 64;                    // digits are 6 bits
constexpr uint32_t SYM_BAD = 0x40u, SYM_FLAG = 0x80u, SYM_DIGIT = 0x3Fu;

// shared bytes of a symbol buffer that stages up to seg_cap residues (seg_cap % 16 == 0)
__host__ __device__ constexpr int ts_sym_bytes(int seg_cap) { return TS_PAD + seg_cap + 32; }
// segment capacity when every one of `threads` threads owns at most c residues (c = 4 * odd)
__host__ __device__ constexpr int ts_seg_cap(int c, int threads = TS_THREADS) { return threads * c; }

// chunk length for n positions over the CTA's threads: smallest 4*odd >= ceil(n / threads)
__device__ __forceinline__ int ts_chunk(int n) {
    int c = (n + int(blockDim.x) - 1) / int(blockDim.x);
    c = (c + 3) >> 2;            // words
    c |= 1;                      // odd
    return c << 2;
}

// ---- shared-memory accessors on 32-bit shared-window addresses --------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
__device__ __forceinline__ void reds_add_u32(uint32_t a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void reds_min_u32(uint32_t a, uint32_t v) {
    asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// LUT with the in-kernel symbol encoding (digit | SYM_BAD)
__device__ __forceinline__ void ts_lut_init(uint8_t *s_lut, const uint8_t *__restrict__ lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint32_t v = lut[i];
        s_lut[i] = (v == SYM_INVALID) ? uint8_t(SYM_BAD) : uint8_t(v);
    }
}

__device__ __forceinline__ uint32_t ts_translate4(uint32_t w, const uint8_t *s_lut) {
    return uint32_t(s_lut[w & 0xFFu]) | (uint32_t(s_lut[(w >> 8) & 0xFFu]) << 8) |
           (uint32_t(s_lut[(w >> 16) & 0xFFu]) << 16) | (uint32_t(s_lut[w >> 24]) << 24);
}

// Geometry of one staged segment [a, b) (absolute positions in `res`).
struct TsSeg {
    int64_t base;   // absolute position of symbol index TS_PAD (16-byte aligned)
    int lo, hi;     // symbol indices of a and b
};

// Step 1 of a segment: translate [a & ~15, b) into s_sym[TS_PAD ...).  For every segment
// but the first of a range, `a` is 16-byte aligned and the caller has already moved the
// previous tail into the pad (ts_tail_read / ts_tail_write).
__device__ __forceinline__ TsSeg ts_stage(const uint8_t *__restrict__ res, int64_t nres, int64_t a, int64_t b,
                                          const uint8_t *s_lut, uint8_t *s_sym) {
    TsSeg g;
    g.base = a & ~int64_t(15);
    g.lo = TS_PAD + int(a - g.base);
    g.hi = g.lo + int(b - a);
    const int nvec = int((b - g.base + 15) >> 4);
    uint4 *dst = reinterpret_cast<uint4 *>(s_sym + TS_PAD);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const int64_t p = g.base + 16 * int64_t(v);
        uint4 x;
        if (p + 16 <= nres) {
            x = __ldg(reinterpret_cast<const uint4 *>(res + p));
        } else {
            uint32_t w[4] = {0, 0, 0, 0};
            for (int j = 0; j < 16; ++j)
                if (p + j < nres) w[j >> 2] |= uint32_t(res[p + j]) << (8 * (j & 3));
            x = make_uint4(w[0], w[1], w[2], w[3]);
        }
        x.x = ts_translate4(x.x, s_lut);
        x.y = ts_translate4(x.y, s_lut);
        x.z = ts_translate4(x.z, s_lut);
        x.w = ts_translate4(x.w, s_lut);
        dst[v] = x;
    }
    return g;
}

// After the staging stores are visible (one __syncthreads): for the first segment of a
// range, everything in front of `a` reads as invalid.
__device__ __forceinline__ void ts_invalidate_front(uint8_t *s_sym, const TsSeg &g) {
    for (int i = threadIdx.x; i < g.lo; i += blockDim.x) s_sym[i] = uint8_t(SYM_BAD);
}

// Between two segments: read the last k-1 symbols of the finished segment ...
__device__ __forceinline__ uint32_t ts_tail_read(const uint8_t *s_sym, const TsSeg &g, int k) {
    const int t = threadIdx.x;
    return (t < k - 1) ? s_sym[g.hi - (k - 1) + t] : 0u;
}
// ... and, after a __syncthreads, put them in front of the next one.
__device__ __forceinline__ void ts_tail_write(uint8_t *s_sym, uint32_t v, int k) {
    const int t = threadIdx.x;
    if (t < k - 1) s_sym[TS_PAD - (k - 1) + t] = uint8_t(v);
}

// One position of the scan.  p = shared address of symbol i, q = p - (k-1).
//   emit(p, code, ok)  window ENDING at p; ok = it is valid (starts at p-(k-1));
//                      called for every position so that it can stay branch-free
//   boundary(p)        symbol at p is the first residue of a sequence (before emit)
template <int OFF, typename CodeT, typename Emit, typename Boundary>
__device__ __forceinline__ void ts_step(uint32_t p, uint32_t q, uint32_t k, CodeT nsym, CodeT neg_pow_k1, uint32_t &run,
                                        CodeT &code, Emit &emit, Boundary &boundary) {
    uint32_t s = lds_u8<OFF>(p);
    if (s >= SYM_BAD) {
        if (s & SYM_FLAG) { run = 0u; boundary(p + OFF); }
        if (s & SYM_BAD) run = 0xFFFFFFFFu;
        s &= SYM_DIGIT;
    }
    run += 1;
    code = code * nsym + CodeT(s);
    emit(p + OFF, code, run >= k);
    code += CodeT(lds_u8<OFF>(q) & SYM_DIGIT) * neg_pow_k1;
}

// Step 3 for one thread: symbol indices [i0, i1) of the staged segment at shared address `sym`.
template <typename CodeT, typename Emit, typename Boundary>
__device__ __forceinline__ void ts_scan_chunk(uint32_t sym, int i0, int i1, int k, CodeT nsym, CodeT pow_k1, Emit &&emit,
                                              Boundary &&boundary) {
    if (i0 >= i1) return;
    uint32_t run = 0;
    CodeT code = 0;
    uint32_t p = sym + uint32_t(i0 - (k - 1));
    for (int j = 0; j < k - 1; ++j, ++p) {
        uint32_t s = lds_u8<0>(p);
        if (s >= SYM_BAD) {
            run = (s & SYM_BAD) ? 0xFFFFFFFFu : 0u;
            s &= SYM_DIGIT;
        }
        run += 1;
        code = code * nsym + CodeT(s);
    }
    const CodeT npow = CodeT(0) - pow_k1;
    const uint32_t uk = uint32_t(k);
    uint32_t q = p - (uk - 1u);
    const uint32_t pend = sym + uint32_t(i1);
    for (; p + 4 <= pend; p += 4, q += 4) {
        ts_step<0>(p, q, uk, nsym, npow, run, code, emit, boundary);
        ts_step<1>(p, q, uk, nsym, npow, run, code, emit, boundary);
        ts_step<2>(p, q, uk, nsym, npow, run, code, emit, boundary);
        ts_step<3>(p, q, uk, nsym, npow, run, code, emit, boundary);
    }
    for (; p < pend; ++p, ++q) ts_step<0>(p, q, uk, nsym, npow, run, code, emit, boundary);
}

inline bool ts_supported(int nsym, int k) { return nsym <= TS_MAX_NSYM && k <= TS_MAX_K; }

// ---- bulk (TMA engine) shared -> global store -------------------------------------
// All threads that wrote the source call ts_bulk_fence() and then __syncthreads(); ONE
// thread calls ts_bulk_store() (dst, src 16-byte aligned, bytes % 16 == 0) and
// ts_bulk_wait_read() before the source is overwritten.
__device__ __forceinline__ void ts_bulk_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ts_bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void ts_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

}  // namespace skm

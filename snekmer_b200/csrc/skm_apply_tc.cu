// skm_apply_tc.cu — kernel (d), dense basis: cosine scoring on the 5th-gen tensor cores.
//
// Replaces sklearn.metrics.pairwise.cosine_similarity(M, Q).T + argsort top-2
// (apply.smk:278-335, learn.smk:811-849) when the k-mer basis is dense (K <= 32768).
//
// Counts are integers, so the GEMM is done EXACTLY on the integer tensor-core path
// (tcgen05.mma kind::i8, int32 accumulators in TMEM) instead of TF32:
//   Q (query counts, 0..255)  -> one uint8 plane            A operand, K-major
//   M (annotation sums)       -> n_planes base-256 digits   B operands, K-major
//   dot[q, a] = sum_j 256^j * (Q . M_j^T)[q, a]             exact while K * 255^2 < 2^31
// and only the final scaling by 1 / (||q|| ||m||) is floating point (float64, like the
// reference).  One CTA owns 128 queries and walks all annotation tiles of 128:
//   warp 0      TMA producer: 128 x 128-byte boxes of Q and of every M plane (SWIZZLE_128B)
//               into a multi-stage shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer: one elected lane issues 4 x n_planes tcgen05.mma (M=128, N=128,
//               K=32) per stage, tcgen05.commit releases the stage / publishes the tile
//   warps 2-5   epilogue: tcgen05.ld the int32 accumulators (lane = query row); float32 screening
//               against the running runner-up, exact int64 / float64 path for the survivors,
//               running top-2 per query in registers
// TMEM: 4 accumulator slots of 128 columns; a tile takes one slot per digit plane (ring).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <vector>

#include "skm_common.cuh"

namespace skm {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 128, UK = 32;
constexpr int MAX_PLANES = 4;
constexpr int ACC_SLOTS = 4;                  // TMEM accumulator slots of BN columns
constexpr int TILE_BYTES = BM * BK;            // 16 KB: one 128 x 128-byte operand tile
constexpr int THREADS = 192;
constexpr int64_t MAX_K = 32768;               // K * 255 * 255 < 2^31

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Bounded wait: a pipeline bug must surface as a launch failure (trap -> cudaErrorLaunchFailure reported by the
// next SKM_CUDA_TRY), never as a hung GPU.  ~4 s at 2 GHz is far beyond any legitimate wait in this kernel.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x10000;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 8000000000ll) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#define SKM_R4(a, o) "=r"(a[o]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3])
// tcgen05.ld 32x32b: lane l of warp w reads TMEM lane 32*(w%4)+l, N consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : SKM_R4(r, 0) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : SKM_R4(r, 0), SKM_R4(r, 4), SKM_R4(r, 8), SKM_R4(r, 12) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld(uint32_t addr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : SKM_R4(r, 0), SKM_R4(r, 4), SKM_R4(r, 8), SKM_R4(r, 12), SKM_R4(r, 16), SKM_R4(r, 20), SKM_R4(r, 24), SKM_R4(r, 28)
                 : "r"(addr));
}
#undef SKM_R4
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return uint64_t((addr >> 4) & 0x3FFFu) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}
// kind::i8: D = S32, A = B = unsigned 8-bit, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (2u << 4) | (0u << 7) | (0u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);

struct Top2d {
    double s1, s2;
    int i1, i2;
};
// higher score first; equal scores -> lower ORIGINAL annotation index (np.argsort(-S) on exact ties)
__device__ __forceinline__ void top2d_push(Top2d &t, double s, int i) {
    if (t.i1 < 0 || s > t.s1 || (s == t.s1 && i < t.i1)) { t.s2 = t.s1; t.i2 = t.i1; t.s1 = s; t.i1 = i; }
    else if (t.i2 < 0 || s > t.s2 || (s == t.s2 && i < t.i2)) { t.s2 = s; t.i2 = i; }
}

// per-tile epilogue operands staged in shared memory by the epilogue warps (double-buffered by tile parity)
struct EpiTile {
    double inv_mn[BN];      // 1 / ||m_a||  (0 for zero rows and padding)
    float inv_m32[BN];      // the same in float: the screening pass
    int32_t orig[BN];       // original annotation index, -1 = padding row
};

// exact dot of one (query, annotation) pair from its digit-plane accumulators
template <int NP, int W>
__device__ __forceinline__ int64_t plane_dot(const uint32_t (&r)[NP][W], int i) {
    int64_t dot = 0;
#pragma unroll
    for (int j = NP - 1; j >= 0; --j) dot = (dot << 8) + int64_t(int32_t(r[j][i]));
    return dot;
}
template <int NP, int W>
__device__ __forceinline__ float plane_dot_f32(const uint32_t (&r)[NP][W], int i) {
    float v = __int2float_rn(int32_t(r[NP - 1][i]));
#pragma unroll
    for (int j = NP - 2; j >= 0; --j) v = fmaf(v, 256.0f, __int2float_rn(int32_t(r[j][i])));
    return v;
}

// Epilogue of one annotation tile WITHOUT the full score matrix.  The scores only feed a top-2, so every
// accumulator is first screened in float32 — s32 = float(dot) / ||m|| against a threshold a little below the
// running runner-up — and only survivors (~2 ln A per query) take the exact path: int64 dot, float64 scaling,
// tie -> lowest index.  The screening keeps the float64 / 64-bit conversion (XU pipe, 16 lanes/clk) out of the
// common path: the first version converted every accumulator and was bound by that pipe at 7 % tensor activity.
// Zero dots never pass (thr > 0); the caller fills a missing runner-up with the lowest-index zero-score row.
template <int NP>
__device__ __forceinline__ void epi_tile_screen(uint32_t lane_addr, uint32_t cursor, const EpiTile &et, double inv_qn, double qn,
                                                Top2d &best, float &thr) {
    constexpr int CW = (NP == 1) ? 32 : 16;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CW) {
        uint32_t r[NP][CW];
#pragma unroll
        for (int j = 0; j < NP; ++j) tmem_ld(lane_addr + ((cursor + j) & 3u) * BN + c0, r[j]);
        tmem_ld_wait();
        uint32_t gm = 0;                                   // bit g: one of columns c0 + 4g .. 4g+3 may enter this query's top-2
#pragma unroll
        for (int i = 0; i < CW; i += 4) {
            const float4 im = *reinterpret_cast<const float4 *>(&et.inv_m32[c0 + i]);
            const float f0 = plane_dot_f32<NP, CW>(r, i) * im.x, f1 = plane_dot_f32<NP, CW>(r, i + 1) * im.y;
            const float f2 = plane_dot_f32<NP, CW>(r, i + 2) * im.z, f3 = plane_dot_f32<NP, CW>(r, i + 3) * im.w;
            if (fmaxf(fmaxf(f0, f1), fmaxf(f2, f3)) >= thr) gm |= 1u << (i >> 2);
        }
        uint32_t wm = __reduce_or_sync(FULL, gm);          // warp-uniform: groups some lane wants to look at
        while (wm) {
            const int g = __ffs(wm) - 1;
            wm &= wm - 1;
            uint32_t v[NP][4];
#pragma unroll
            for (int j = 0; j < NP; ++j) tmem_ld(lane_addr + ((cursor + j) & 3u) * BN + c0 + 4 * g, v[j]);   // warp-collective re-read
            tmem_ld_wait();
            if ((gm >> g) & 1u) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = c0 + 4 * g + e;
                    const int orig = et.orig[col];
                    if (orig >= 0 && plane_dot_f32<NP, 4>(v, e) * et.inv_m32[col] >= thr)
                        top2d_push(best, double(plane_dot<NP, 4>(v, e)) * (inv_qn * et.inv_mn[col]), orig);
                }
                // below the runner-up by more than the float32 error of the screening product (3 * 2^-24)
                if (best.i2 >= 0) thr = fmaxf(thr, __double2float_rd(best.s2 * qn) * 0.99999f);
            }
        }
    }
}

// Epilogue of one annotation tile WITH the full score matrix (save_apply_associations): every score is needed in
// float64, so everything takes the exact path.
template <int NP>
__device__ __forceinline__ void epi_tile_full(uint32_t lane_addr, uint32_t cursor, const EpiTile &et, double inv_qn, Top2d &best,
                                              double *__restrict__ full_row, bool q_ok) {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[NP][16];
#pragma unroll
        for (int j = 0; j < NP; ++j) tmem_ld(lane_addr + ((cursor + j) & 3u) * BN + c0, r[j]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int orig = et.orig[c0 + i];
            const double s = double(plane_dot<NP, 16>(r, i)) * (inv_qn * et.inv_mn[c0 + i]);
            if (orig >= 0) {
                if (q_ok) full_row[orig] = s;
                top2d_push(best, s, orig);
            }
        }
    }
}

// Annotation rows are stored sorted by magnitude class (perm[sorted] = original index), so that a tile of 128
// rows only carries the digit planes its largest entry needs (tile_planes[t] <= 4): small annotations — the
// bulk of a Zipf-like family size distribution — cost one int8 GEMM pass instead of several.
// TMEM: 512 columns = 4 accumulator slots of 128 columns; a tile with np planes takes np consecutive slots
// (mod 4), so 1-plane tiles are 4 deep in flight and the epilogue of tile t overlaps the MMAs of t+1 .. t+3.
// RESIDENT: the whole 128-query operand (k_chunks x 16 KB) stays in shared memory for the CTA's lifetime and
// only annotation tiles stream through the ring (K <= 1024); otherwise query chunks travel through the ring too.
template <bool RESIDENT, bool FULLOUT>
__global__ void __launch_bounds__(THREADS, 1)
apply_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_m, int ann_pad,
                int64_t nq, int n_ann, int k_chunks, int slots, const int32_t *__restrict__ perm,
                const int32_t *__restrict__ tile_planes, const double *__restrict__ qnorm2,
                const double *__restrict__ mnorm2, int32_t *__restrict__ top1, int32_t *__restrict__ top2,
                double *__restrict__ sc1, double *__restrict__ sc2, double *__restrict__ full) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * 8 + 2 * ACC_SLOTS + 1];   // full[8], empty[8], tmem_full[4], tmem_empty[4], q_full
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) EpiTile s_epi[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (s2u(smem) + 1023u) & ~1023u;
    const uint32_t q_bytes = RESIDENT ? uint32_t(k_chunks) * TILE_BYTES : 0u;
    const uint32_t ring_base = smem_base + q_bytes;
    const uint32_t full0 = s2u(&bars[0]), empty0 = s2u(&bars[8]), tfull0 = s2u(&bars[16]), tempty0 = s2u(&bars[16 + ACC_SLOTS]),
                   qfull = s2u(&bars[16 + 2 * ACC_SLOTS]);
    constexpr uint32_t tmem_cols = ACC_SLOTS * BN;          // 512: the whole TMEM of the SM (one CTA per SM)
    const int n_tiles = (n_ann + BN - 1) / BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < ACC_SLOTS; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 4); }
        mbar_init(qfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {        // TMEM allocation (whole warp), same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int64_t q0 = int64_t(blockIdx.x) * BM;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (RESIDENT) {
                mbar_arrive_expect_tx(qfull, q_bytes);
                for (int kc = 0; kc < k_chunks; ++kc) tma_load_2d(smem_base + kc * TILE_BYTES, &map_q, qfull, kc * BK, int(q0));
            }
            int slot = 0;
            uint32_t phase = 0;
            auto push = [&](const CUtensorMap *map, int c0, int c1) {
                mbar_wait(empty0 + 8 * slot, phase ^ 1);
                mbar_arrive_expect_tx(full0 + 8 * slot, TILE_BYTES);
                tma_load_2d(ring_base + slot * TILE_BYTES, map, full0 + 8 * slot, c0, c1);
                if (++slot == slots) { slot = 0; phase ^= 1; }
            };
            for (int t = 0; t < n_tiles; ++t) {
                const int np = __ldg(tile_planes + t);
                for (int kc = 0; kc < k_chunks; ++kc) {
                    if (!RESIDENT) push(&map_q, kc * BK, int(q0));
                    for (int j = 0; j < np; ++j) push(&map_m, kc * BK, j * ann_pad + t * BN);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            uint32_t cursor = 0, par_empty = 0;                 // bit s of par_empty: parity of the uses of accumulator slot s so far
            if (RESIDENT) { mbar_wait(qfull, 0); tc_fence_after(); }
            for (int t = 0; t < n_tiles; ++t) {
                const int np = __ldg(tile_planes + t);
                for (int j = 0; j < np; ++j) {                  // the epilogue has drained the slots this tile takes
                    const uint32_t s = (cursor + j) & 3u;
                    mbar_wait(tempty0 + 8 * s, ((par_empty >> s) & 1u) ^ 1u);
                    par_empty ^= 1u << s;
                }
                tc_fence_after();
                for (int kc = 0; kc < k_chunks; ++kc) {
                    uint32_t a_addr = smem_base + kc * TILE_BYTES;
                    int a_slot = -1;
                    if (!RESIDENT) {
                        mbar_wait(full0 + 8 * slot, phase);
                        a_addr = ring_base + slot * TILE_BYTES;
                        a_slot = slot;
                        if (++slot == slots) { slot = 0; phase ^= 1; }
                    }
                    for (int j = 0; j < np; ++j) {
                        mbar_wait(full0 + 8 * slot, phase);
                        tc_fence_after();
                        const uint32_t b_addr = ring_base + slot * TILE_BYTES;
                        const uint32_t d = tmem_base + ((cursor + j) & 3u) * BN;
#pragma unroll
                        for (int kk = 0; kk < BK / UK; ++kk)
                            umma_i8(d, smem_desc(a_addr + kk * UK), smem_desc(b_addr + kk * UK), IDESC, (kc | kk) ? 1u : 0u);
                        umma_commit(empty0 + 8 * slot);         // slot free once these MMAs have read it
                        if (++slot == slots) { slot = 0; phase ^= 1; }
                    }
                    if (a_slot >= 0) umma_commit(empty0 + 8 * a_slot);
                }
                umma_commit(tfull0 + 8 * cursor);               // accumulators of tile t complete (keyed by its first slot)
                cursor = (cursor + np) & 3u;
            }
        }
    } else {
        // ===== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int64_t q = q0 + row;
        const int et = threadIdx.x - 64;                          // 0..127 among the epilogue threads
        const double qn2 = (q < nq) ? qnorm2[q] : 0.0;
        const double qn = sqrt(qn2);
        const double inv_qn = qn2 > 0.0 ? 1.0 / qn : 0.0;
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        Top2d best{0.0, 0.0, -1, -1};
        float thr = 1e-30f;                                       // > 0: zero dots are never candidates (see the final fill)
        uint32_t cursor = 0, par_full = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int np = __ldg(tile_planes + t);
            EpiTile &tile = s_epi[t & 1];
            {   // original index and 1 / ||m_a|| of this tile's rows
                const int a = t * BN + et;
                const int orig = (a < n_ann) ? __ldg(perm + a) : -1;
                const double m2 = (orig >= 0) ? mnorm2[orig] : 0.0;
                const double inv = m2 > 0.0 ? 1.0 / sqrt(m2) : 0.0;
                tile.orig[et] = orig;
                tile.inv_mn[et] = inv;
                tile.inv_m32[et] = float(inv);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(tfull0 + 8 * cursor, (par_full >> cursor) & 1u);
            par_full ^= 1u << cursor;
            tc_fence_after();
            if (FULLOUT) {
                double *full_row = full + (q < nq ? q : 0) * int64_t(n_ann);
                switch (np) {
                    case 1: epi_tile_full<1>(lane_addr, cursor, tile, inv_qn, best, full_row, q < nq); break;
                    case 2: epi_tile_full<2>(lane_addr, cursor, tile, inv_qn, best, full_row, q < nq); break;
                    case 3: epi_tile_full<3>(lane_addr, cursor, tile, inv_qn, best, full_row, q < nq); break;
                    default: epi_tile_full<4>(lane_addr, cursor, tile, inv_qn, best, full_row, q < nq); break;
                }
            } else {
                switch (np) {
                    case 1: epi_tile_screen<1>(lane_addr, cursor, tile, inv_qn, qn, best, thr); break;
                    case 2: epi_tile_screen<2>(lane_addr, cursor, tile, inv_qn, qn, best, thr); break;
                    case 3: epi_tile_screen<3>(lane_addr, cursor, tile, inv_qn, qn, best, thr); break;
                    default: epi_tile_screen<4>(lane_addr, cursor, tile, inv_qn, qn, best, thr); break;
                }
            }
            tc_fence_before();
            if (lane == 0)
                for (int j = 0; j < np; ++j) mbar_arrive(tempty0 + 8 * ((cursor + j) & 3u));   // 4 arrivals (one per warp) free a slot
            cursor = (cursor + np) & 3u;
        }
        if (q < nq) {
            // rows that were never candidates score exactly 0: a missing winner / runner-up is the lowest-index one of them
            if (best.i1 < 0) { best.i1 = 0; best.s1 = 0.0; }
            if (best.i2 < 0 && n_ann > 1) { best.i2 = (best.i1 == 0) ? 1 : 0; best.s2 = 0.0; }
            top1[q] = best.i1; sc1[q] = best.s1;
            top2[q] = best.i2; sc2[q] = best.i2 >= 0 ? best.s2 : nan("");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---- operand preparation ------------------------------------------------------------------------
// Q int32 [nq, K] -> uint8 [nq, Kp] (zero padded); *flag |= 1 if a count does not fit 8 bits
__global__ void __launch_bounds__(256) split_q_kernel(const int32_t *__restrict__ Q, int64_t nq, int64_t K, int64_t Kp,
                                                      uint8_t *__restrict__ out, int *__restrict__ flag) {
    const int64_t total = nq * (Kp >> 2);
    bool bad = false;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / (Kp >> 2), c = (i - r * (Kp >> 2)) << 2;
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int32_t v = (c + j < K) ? __ldg(Q + r * K + c + j) : 0;
            bad |= (v < 0 || v > 255);
            w |= uint32_t(v & 0xFF) << (8 * j);
        }
        reinterpret_cast<uint32_t *>(out + r * Kp)[c >> 2] = w;
    }
    if (bad) atomicOr(flag, 1);
}
// M int64 [A, K] -> planes uint8 [n_planes, Apad, Kp]: digit j (base 256) of row perm[a]; rows / columns beyond
// A, K are zero
__global__ void __launch_bounds__(256) split_m_kernel(const int64_t *__restrict__ M, int64_t A, int64_t K, int64_t Apad,
                                                      int64_t Kp, int n_planes, const int32_t *__restrict__ perm,
                                                      uint8_t *__restrict__ out) {
    const int64_t total = Apad * Kp;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t a = i / Kp, c = i - a * Kp;
        const uint64_t v = (a < A && c < K) ? uint64_t(__ldg(M + int64_t(__ldg(perm + a)) * K + c)) : 0ull;
        for (int j = 0; j < n_planes; ++j) out[(int64_t(j) * Apad + a) * Kp + c] = uint8_t(v >> (8 * j));
    }
}
// largest entry of every row (one warp per row); negative entries are not representable -> all-ones
__global__ void __launch_bounds__(256) row_max_kernel(const int64_t *__restrict__ X, int64_t rows, int64_t cols,
                                                      unsigned long long *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        unsigned long long m = 0;
        for (int64_t c = lane; c < cols; c += 32) {
            const int64_t v = X[r * cols + c];
            const unsigned long long u = v < 0 ? ~0ull : (unsigned long long)v;
            m = u > m ? u : m;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(FULL, m, o); m = x > m ? x : m; }
        if (lane == 0) out[r] = m;
    }
}

static PFN_cuTensorMapEncodeTiled encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}
// uint8 matrix [rows, Kp] row-major, boxes of 128 rows x 128 bytes, 128-byte swizzle, zero fill outside
static int make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t Kp) {
    PFN_cuTensorMapEncodeTiled fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return SKM_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp};
    cuuint32_t box[2] = {BK, BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return SKM_ERR_CUDA; }
    return SKM_OK;
}

static int64_t pad128(int64_t x) { return (x + 127) & ~int64_t(127); }
// prepared-matrix blob: [perm int32 Apad | tile_planes int32 Apad/128 | pad to 1024] [planes MAXP x Apad x Kp]
static size_t meta_bytes(int64_t Apad) { return (size_t(Apad) * 4 + size_t(Apad / 128) * 4 + 1023) & ~size_t(1023); }

}  // namespace tc
}  // namespace skm

extern "C" {

size_t skm_apply_tc_planes_bytes(int64_t n_ann, int64_t K) {
    using namespace skm::tc;
    if (n_ann <= 0 || K <= 0) return 1024;
    const int64_t Apad = pad128(n_ann);
    return meta_bytes(Apad) + size_t(MAX_PLANES) * size_t(Apad) * size_t(pad128(K)) + 256;
}

int skm_apply_tc_prepare(const int64_t *d_M, int64_t n_ann, int64_t K, uint8_t *d_planes, size_t planes_bytes,
                         int *n_planes_out, skm_stream_t stream) {
    using namespace skm;
    using namespace skm::tc;
    if (!n_planes_out) { set_error("skm_apply_tc_prepare: n_planes_out is NULL"); return SKM_ERR_INVALID; }
    *n_planes_out = 0;
    if (n_ann <= 0 || K <= 0) { set_error("skm_apply_tc_prepare: empty matrix"); return SKM_ERR_INVALID; }
    if (K > MAX_K) { set_error("skm_apply_tc_prepare: K=%lld > %lld (int32 accumulators would overflow); use skm_apply_dense", (long long)K, (long long)MAX_K); return SKM_ERR_UNSUPPORTED; }
    if (!d_M || !d_planes || planes_bytes < skm_apply_tc_planes_bytes(n_ann, K)) { set_error("skm_apply_tc_prepare: bad buffers"); return SKM_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(d_planes) & 127u) != 0) { set_error("skm_apply_tc_prepare: d_planes must be 128-byte aligned"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Apad = pad128(n_ann), Kp = pad128(K), n_tiles = Apad / 128;
    // per-row maxima decide the digit planes a row needs (the plane area of the blob is scratch until the split)
    uint8_t *planes = d_planes + meta_bytes(Apad);
    unsigned long long *d_rowmax = reinterpret_cast<unsigned long long *>(planes);
    row_max_kernel<<<(int)std::min<int64_t>((n_ann + 7) / 8, int64_t(sm_count()) * 8), 256, 0, st>>>(d_M, n_ann, K, d_rowmax);
    SKM_LAUNCH_CHECK("row_max_kernel");
    std::vector<unsigned long long> rowmax((size_t)n_ann);
    SKM_CUDA_TRY(cudaMemcpyAsync(rowmax.data(), d_rowmax, size_t(n_ann) * 8, cudaMemcpyDeviceToHost, st));
    SKM_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int> cls((size_t)n_ann);
    int max_planes = 1;
    for (int64_t a = 0; a < n_ann; ++a) {
        int p = 1;
        while (p < 8 && (rowmax[a] >> (8 * p)) != 0) ++p;
        cls[a] = p;
        max_planes = std::max(max_planes, p);
    }
    if (max_planes > MAX_PLANES) { set_error("skm_apply_tc_prepare: entries need %d base-256 digit planes (> %d); use skm_apply_dense", max_planes, MAX_PLANES); return SKM_ERR_UNSUPPORTED; }
    // rows sorted by class, largest first (stable: original order inside a class)
    std::vector<int32_t> meta((size_t)Apad + (size_t)n_tiles, 0);
    {
        std::vector<int32_t> order((size_t)n_ann);
        for (int64_t a = 0; a < n_ann; ++a) order[a] = (int32_t)a;
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cls[x] > cls[y]; });
        for (int64_t a = 0; a < n_ann; ++a) meta[a] = order[a];
        for (int64_t t = 0; t < n_tiles; ++t) meta[Apad + t] = (t * 128 < n_ann) ? cls[order[t * 128]] : 1;
    }
    SKM_CUDA_TRY(cudaMemcpyAsync(d_planes, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, st));
    SKM_CUDA_TRY(cudaStreamSynchronize(st));          // `meta` is a local
    split_m_kernel<<<(int)std::min<int64_t>((Apad * Kp + 255) / 256, int64_t(sm_count()) * 16), 256, 0, st>>>(
        d_M, n_ann, K, Apad, Kp, max_planes, reinterpret_cast<const int32_t *>(d_planes), planes);
    SKM_LAUNCH_CHECK("split_m_kernel");
    *n_planes_out = max_planes;
    return SKM_OK;
}

size_t skm_apply_tc_workspace(int64_t nq, int64_t K) {
    using namespace skm::tc;
    if (nq <= 0 || K <= 0) return 512;
    return size_t(nq) * size_t(pad128(K)) + 512;
}

int skm_apply_tc(const int32_t *d_Q, int64_t nq, int64_t K, const uint8_t *d_planes, int n_planes, int64_t n_ann,
                 const double *d_qnorm2, const double *d_mnorm2, int32_t *d_top1, int32_t *d_top2, double *d_score1,
                 double *d_score2, double *d_scores_full, int *d_status, void *workspace, size_t workspace_bytes,
                 skm_stream_t stream) {
    using namespace skm;
    using namespace skm::tc;
    if (nq < 0 || K <= 0 || K > MAX_K || n_ann <= 0 || n_ann > 0x7FFFFF00ll || n_planes < 1 || n_planes > MAX_PLANES) { set_error("skm_apply_tc: bad sizes (K=%lld n_ann=%lld planes=%d)", (long long)K, (long long)n_ann, n_planes); return SKM_ERR_INVALID; }
    if (nq == 0) return SKM_OK;
    if (nq > 0x7FFFFF00ll) { set_error("skm_apply_tc: more than 2^31 queries per call"); return SKM_ERR_UNSUPPORTED; }
    if (!d_Q || !d_planes || !d_qnorm2 || !d_mnorm2 || !d_top1 || !d_top2 || !d_score1 || !d_score2 || !d_status) { set_error("skm_apply_tc: NULL argument"); return SKM_ERR_INVALID; }
    const size_t need = skm_apply_tc_workspace(nq, K);
    if (!workspace || workspace_bytes < need) { set_error("skm_apply_tc: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Kp = pad128(K), Apad = pad128(n_ann);
    const int32_t *perm = reinterpret_cast<const int32_t *>(d_planes);
    const int32_t *tile_planes = perm + Apad;
    const uint8_t *planes = d_planes + meta_bytes(Apad);
    uint8_t *q8 = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    SKM_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    split_q_kernel<<<(int)std::min<int64_t>((nq * (Kp >> 2) + 255) / 256, int64_t(sm_count()) * 16), 256, 0, st>>>(d_Q, nq, K, Kp, q8, d_status);
    SKM_LAUNCH_CHECK("split_q_kernel");
    alignas(64) CUtensorMap map_q, map_m;
    int rc = make_map(&map_q, q8, nq, Kp);
    if (rc) return rc;
    rc = make_map(&map_m, planes, int64_t(n_planes) * Apad, Kp);
    if (rc) return rc;
    const int k_chunks = int(Kp / BK);
    const bool resident = k_chunks <= 8;                       // 128 queries x 1024 bytes = 128 KB of the 227 KB
    const size_t budget = 227 * 1024 - 6144 - 1024;           // static shared (barriers + 2 epilogue tiles) + alignment slack
    const size_t q_bytes = resident ? size_t(k_chunks) * TILE_BYTES : 0;
    int slots = int((budget - q_bytes) / TILE_BYTES);
    if (slots > 8) slots = 8;
    if (slots < 2) { set_error("skm_apply_tc: pipeline does not fit shared memory"); return SKM_ERR_UNSUPPORTED; }
    const size_t smem = q_bytes + size_t(slots) * TILE_BYTES + 1024;
    const unsigned grid = (unsigned)((nq + BM - 1) / BM);
#define SKM_LAUNCH_TC(RES, FULLOUT)                                                                                          \
    {                                                                                                                        \
        auto kern = apply_tc_kernel<RES, FULLOUT>;                                                                           \
        SKM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        kern<<<grid, THREADS, smem, st>>>(map_q, map_m, (int)Apad, nq, (int)n_ann, k_chunks, slots, perm, tile_planes,       \
                                          d_qnorm2, d_mnorm2, d_top1, d_top2, d_score1, d_score2, d_scores_full);            \
    }
    if (resident) { if (d_scores_full) SKM_LAUNCH_TC(true, true) else SKM_LAUNCH_TC(true, false) }
    else { if (d_scores_full) SKM_LAUNCH_TC(false, true) else SKM_LAUNCH_TC(false, false) }
#undef SKM_LAUNCH_TC
    SKM_LAUNCH_CHECK("apply_tc_kernel");
    return SKM_OK;
}

}  // extern "C"

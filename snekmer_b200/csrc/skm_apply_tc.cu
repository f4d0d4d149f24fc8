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
// reference).  A CTA PAIR (cluster of 2, cta_group::2) owns 256 queries and walks all annotation
// tiles; per CTA:
//   warp 0      TMA producer: this CTA's query rows, its HALF of every annotation tile
//               (SWIZZLE_128B boxes) and the tile's epilogue operands into shared-memory rings,
//               completion on mbarriers (the operand bytes of both CTAs land on the leader's)
//   warp 1      MMA issuer (leader CTA): one elected lane issues 4 tcgen05.mma (M=256, N=256 or 128,
//               K=32) per ring slot; tcgen05.commit (multicast) releases the slot / publishes the tile
//   warps 2-9   epilogue: tcgen05.ld the int32 accumulators (lane = query row; two warps per TMEM
//               lane quarter, each on half of the tile's columns); float32 screening against the
//               running runner-up, exact int64 / float64 path for the survivors, running top-2 per
//               query in registers, merged at the end
// TMEM: two halves of 256 columns; a 1-plane tile of 256 annotations takes one half, so its epilogue
// overlaps the MMAs of the next tile.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <vector>

#include "skm_common.cuh"

namespace skm {
namespace tc {

constexpr int BM = 128, BK = 128, UK = 32;     // query rows per CTA, bytes of K per ring slot, bytes of K per MMA
constexpr int WIDE = 256, NARROW = 128;        // annotation rows per tile
constexpr int MAX_PLANES = 4;
constexpr int SLOT_BYTES = BM * BK;            // 16 KB: 128 rows x 128 bytes (query chunk, or a CTA's half of a wide tile)
constexpr int MAX_SLOTS = 16;                  // ring slots (barrier arrays)
constexpr int META_RING = 3;                   // tile-meta buffers in shared memory
constexpr int EPI_WARPS = 8;                   // two per TMEM lane quarter: each takes half of a tile's columns
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int64_t MAX_K = 32768;               // K * 255 * 255 < 2^31

// one annotation tile of the prepared matrix (rows in sorted order)
struct TileDesc {
    int32_t row0;       // first sorted row
    int32_t width;      // WIDE or NARROW
    int32_t np;         // digit planes this tile carries (1..4; WIDE tiles: 1..2)
    int32_t pad;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of a converged warp; the compiler then knows the guarded code runs on a single lane and issues the
// uniform-datapath instructions (UTMALDG, UTCIMMA, UTCBAR) directly instead of a per-lane serialisation loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xFFFFFFFF;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
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
// 1-D bulk copy global -> this CTA's shared memory (16-byte aligned, size % 16 == 0), bytes complete on a local mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// ---- CTA pair (cluster of 2, cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope): the accumulator reads it orders are TMEM loads fenced by
    // tcgen05.fence::before_thread_sync; a cluster-scope release costs a MEMBAR.GPU per arrival (12 % of the
    // epilogue's time in the first pair version)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose bytes complete on an mbarrier of either CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
// one MMA over both SMs of the pair: M = 256 (128 query rows from each CTA's shared memory), N annotation rows (N / 2
// from each), D = 128 lanes x N columns in each CTA's TMEM.  Issued by the leader CTA (rank 0) only.
template <bool ACCUMULATE>
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACCUMULATE ? 1 : 0) : "memory");
}
// arrives on the mbarrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#define SKM_R4(a, o) "=r"(a[o]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3])
// tcgen05.ld 32x32b: lane l of warp w reads TMEM lane 32*(w%4)+l, N consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : SKM_R4(r, 0) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld(uint32_t addr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : SKM_R4(r, 0), SKM_R4(r, 4) : "r"(addr));
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
// kind::i8: D = S32, A = B = unsigned 8-bit, both K-major, M = 256 (the CTA pair), N = tile width
constexpr uint32_t IDESC_WIDE = (2u << 4) | (0u << 7) | (0u << 10) | (uint32_t(WIDE >> 3) << 17) | (uint32_t((2 * BM) >> 4) << 24);
constexpr uint32_t IDESC_NARROW = (2u << 4) | (0u << 7) | (0u << 10) | (uint32_t(NARROW >> 3) << 17) | (uint32_t((2 * BM) >> 4) << 24);

#ifdef SKM_TC_PROF
__device__ long long g_tl[6][512];     // debug timeline of pair 0: see skm_debug_tc_timeline
#define TL(k, t) do { if (blockIdx.x == 0 && (t) < 512) g_tl[k][t] = clock64(); } while (0)
#else
#define TL(k, t) do { } while (0)
#endif

struct Top2d {
    double s1, s2;
    int i1, i2;
};
// higher score first; equal scores -> lower ORIGINAL annotation index (np.argsort(-S) on exact ties)
__device__ __forceinline__ void top2d_push(Top2d &t, double s, int i) {
    if (t.i1 < 0 || s > t.s1 || (s == t.s1 && i < t.i1)) { t.s2 = t.s1; t.i2 = t.i1; t.s1 = s; t.i1 = i; }
    else if (t.i2 < 0 || s > t.s2 || (s == t.s2 && i < t.i2)) { t.s2 = s; t.i2 = i; }
}

// Running top-2 of the screening epilogue.  Candidates are ordered by their float32 screening score whenever the two
// scores differ by more than 1e-5 relative (the float32 error is < 4e-7, so the exact order is the same); only
// near-ties are settled by the exact float64 scores, tie -> lower original index.  float64 arithmetic issued while the
// tensor pipe is saturated costs thousands of cycles per candidate (measured: one candidate = 3900 cycles at K = 1024
// against 600 with the tensor pipe idle), so the exact scores are otherwise formed once, after the last tile.
struct Top2x {
    float f1, f2;           // screening scores dot / ||m|| (float32)
    int64_t d1, d2;         // exact dots
    double im1, im2;        // 1 / ||m|| (float64) of the two rows
    int i1, i2;             // original annotation indices, -1 = empty
};
__device__ __forceinline__ double exact_score(int64_t dot, double inv_qn, double inv_mn) { return double(dot) * (inv_qn * inv_mn); }
// is the candidate (f, dot, im, i) ranked above the entry (fk, dk, imk, ik)?
__device__ __forceinline__ bool ranks_above(float f, int64_t dot, double im, int i, float fk, int64_t dk, double imk, int ik, double inv_qn) {
    if (f > fk * 1.00001f) return true;
    if (f < fk * 0.99999f) return false;
    const double s = exact_score(dot, inv_qn, im), sk = exact_score(dk, inv_qn, imk);
    return s > sk || (s == sk && i < ik);
}
__device__ __forceinline__ void top2x_push(Top2x &t, float f, int64_t dot, double im, int i, double inv_qn) {
    if (t.i1 < 0 || ranks_above(f, dot, im, i, t.f1, t.d1, t.im1, t.i1, inv_qn)) {
        t.f2 = t.f1; t.d2 = t.d1; t.im2 = t.im1; t.i2 = t.i1;
        t.f1 = f; t.d1 = dot; t.im1 = im; t.i1 = i;
    } else if (t.i2 < 0 || ranks_above(f, dot, im, i, t.f2, t.d2, t.im2, t.i2, inv_qn)) {
        t.f2 = f; t.d2 = dot; t.im2 = im; t.i2 = i;
    }
}

// Per-tile epilogue operands ("tile meta"), precomputed per call by tile_meta_kernel in the layout the epilogue reads
// them in, one contiguous block per tile so that the producer brings it in with ONE bulk copy:
//   [inv_mn float64 x width | inv_m32 float x width | orig int32 x width | inv_g8 float x width/8]
// = 16.5 bytes per row; the block of the tile starting at sorted row row0 sits at byte row0 * 16.5.
__host__ __device__ constexpr size_t meta_block_bytes(int width) { return size_t(width) * 16 + size_t(width) / 2; }
__host__ __device__ constexpr size_t meta_block_offset(int64_t row0) { return size_t(row0) * 16 + size_t(row0) / 2; }
struct EpiTile {
    const double *inv_mn;      // 1 / ||m_a||, 0 for zero rows and padding
    const float *inv_m32;      // the same in float: the screening pass
    const int32_t *orig;       // original annotation index, -1 = padding row
    const float *inv_g8;       // largest inv_m32 of every group of 8 rows (rows of a class are sorted by norm: nearly equal)
    __device__ __forceinline__ EpiTile(const uint8_t *base, int width)
        : inv_mn(reinterpret_cast<const double *>(base)), inv_m32(reinterpret_cast<const float *>(base + 8 * width)),
          orig(reinterpret_cast<const int32_t *>(base + 12 * width)), inv_g8(reinterpret_cast<const float *>(base + 16 * width)) {}
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

// Epilogue of one annotation tile WITHOUT the full score matrix.  The scores only feed a top-2, so the accumulators
// are first screened in float32 against a threshold a little below the running runner-up, and only survivors
// (~2 ln A per query) take the exact path: int64 dot, float64 scaling, tie -> lowest index.  1-plane tiles screen
// 8 accumulators with ONE conversion: max(dot) * max(1/||m||) bounds every score of the group from above
// (int -> float conversions run on the 16-lane XU pipe; one per accumulator made that pipe the limiter).
// Zero dots never pass (thr > 0); the caller fills a missing runner-up with the lowest-index zero-score row.
// tmem = this thread's TMEM lane address + first column of the tile; plane j starts `pstride` columns further.
template <int NP>
__device__ __forceinline__ void epi_tile_screen(uint32_t tmem, uint32_t pstride, int cbeg, int cend, const EpiTile &et, double inv_qn,
                                                Top2x &best, float &thr) {
    constexpr int CW = (NP == 1) ? 32 : 16;       // columns per TMEM load
    constexpr int GW = (NP == 1) ? 8 : 4;         // columns per screening group
#pragma unroll 1
    for (int c0 = cbeg; c0 < cend; c0 += CW) {
        uint32_t r[NP][CW];
#pragma unroll
        for (int j = 0; j < NP; ++j) tmem_ld(tmem + j * pstride + c0, r[j]);
        tmem_ld_wait();
        uint32_t gm = 0;                                   // bit g: a column of group g may enter this query's top-2
        if constexpr (NP == 1) {
#pragma unroll
            for (int g = 0; g < CW / 8; ++g) {
                const int i = 8 * g;
                const int m = max(max(max(int(r[0][i]), int(r[0][i + 1])), max(int(r[0][i + 2]), int(r[0][i + 3]))),
                                  max(max(int(r[0][i + 4]), int(r[0][i + 5])), max(int(r[0][i + 6]), int(r[0][i + 7]))));
                if (__int2float_ru(m) * et.inv_g8[(c0 >> 3) + g] >= thr) gm |= 1u << g;
            }
        } else {
#pragma unroll
            for (int i = 0; i < CW; i += 4) {
                const float4 im = *reinterpret_cast<const float4 *>(&et.inv_m32[c0 + i]);
                const float f0 = plane_dot_f32<NP, CW>(r, i) * im.x, f1 = plane_dot_f32<NP, CW>(r, i + 1) * im.y;
                const float f2 = plane_dot_f32<NP, CW>(r, i + 2) * im.z, f3 = plane_dot_f32<NP, CW>(r, i + 3) * im.w;
                if (fmaxf(fmaxf(f0, f1), fmaxf(f2, f3)) >= thr) gm |= 1u << (i >> 2);
            }
        }
        uint32_t wm = __reduce_or_sync(FULL, gm);          // warp-uniform: groups some lane wants to look at
        while (wm) {
            const int g = __ffs(wm) - 1;
            wm &= wm - 1;
            uint32_t v[NP][GW];
#pragma unroll
            for (int j = 0; j < NP; ++j) tmem_ld(tmem + j * pstride + c0 + GW * g, v[j]);   // warp-collective re-read
            tmem_ld_wait();
            if ((gm >> g) & 1u) {
                uint32_t pm = 0;                           // columns of the group that pass the float32 screen
                float fe[GW];
#pragma unroll
                for (int e = 0; e < GW; ++e) {
                    fe[e] = plane_dot_f32<NP, GW>(v, e) * et.inv_m32[c0 + GW * g + e];
                    if (fe[e] >= thr) pm |= 1u << e;
                }
                while (pm) {                               // usually one: ONE copy of the insertion, operands picked by select chains
                    const int e = __ffs(pm) - 1;
                    pm &= pm - 1;
                    int64_t dot = 0;
#pragma unroll
                    for (int j = NP - 1; j >= 0; --j) {
                        uint32_t x = v[j][0];
#pragma unroll
                        for (int i = 1; i < GW; ++i) x = (e == i) ? v[j][i] : x;
                        dot = (dot << 8) + int64_t(int32_t(x));
                    }
                    float f = fe[0];
#pragma unroll
                    for (int i = 1; i < GW; ++i) f = (e == i) ? fe[i] : f;
                    const int col = c0 + GW * g + e;
                    const int orig = et.orig[col];
                    if (orig >= 0) top2x_push(best, f, dot, et.inv_mn[col], orig, inv_qn);
                }
                // below the runner-up by more than the float32 error of the screening products
                if (best.i2 >= 0) thr = fmaxf(thr, best.f2 * 0.99999f);
            }
        }
    }
}

// Epilogue of one annotation tile WITH the full score matrix (save_apply_associations): every score is needed in
// float64, so everything takes the exact path.
template <int NP>
__device__ __forceinline__ void epi_tile_full(uint32_t tmem, uint32_t pstride, int cbeg, int cend, const EpiTile &et, double inv_qn,
                                              Top2d &best, double *__restrict__ full_row, bool q_ok) {
#pragma unroll 1
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
        uint32_t r[NP][16];
#pragma unroll
        for (int j = 0; j < NP; ++j) tmem_ld(tmem + j * pstride + c0, r[j]);
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

// Annotation rows are stored sorted by magnitude class, so that a tile only carries the digit planes its largest entry
// needs: the few rows with entries >= 2^16 form NARROW tiles (128 rows, 3-4 planes), everything else WIDE tiles
// (256 rows, 1-2 planes) — small annotations, the bulk of a Zipf-like family size distribution, cost one int8 GEMM pass.
//
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns 256 queries: each CTA keeps its 128 query rows and the
// accumulators of those rows, and loads only HALF of every annotation tile; the MMA (M = 256, N = tile width),
// issued by the leader CTA, reads both halves from both SMs.  That is 32 bytes of operand per SM and tensor-core
// cycle, and 128 tensor cycles per instruction (N = 256), which one issuing lane can sustain; the first versions
// (one CTA per tile, N = 128) sat at 39 % tensor activity, bound first by the L2 -> SM latency x bandwidth the ring
// could cover and then by the issue rate of the MMA lane.
// TMEM: 512 columns = two halves of 256; a 1-plane WIDE tile takes one half (its epilogue overlaps the next tile's
// MMAs), every other tile takes both (plane j at column j * width).
// RESIDENT: the CTA's whole 128-query operand (k_chunks x 16 KB) stays in shared memory for the CTA's lifetime and
// only annotation half-tiles stream through the ring (K <= 1024); otherwise query chunks travel through the ring too.
// Barriers: full[s] (leader only; bytes of both CTAs' loads), empty[s] (each CTA; multicast commit), tfull[h]
// (each CTA; multicast commit), tempty[h] (leader only; 8 epilogue warps x 2 CTAs), mfull / mempty[b] (each CTA;
// tile-meta ring between the producer and the 8 epilogue warps, which therefore never wait for one another).
template <bool RESIDENT, bool FULLOUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
apply_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_wide,
                const __grid_constant__ CUtensorMap map_narrow, int plane_rows, int64_t nq, int n_ann, int k_chunks, int slots,
                const uint8_t *__restrict__ tile_meta, const int32_t *__restrict__ header, const TileDesc *__restrict__ tiles,
                const double *__restrict__ qnorm2, int32_t *__restrict__ top1, int32_t *__restrict__ top2,
                double *__restrict__ sc1, double *__restrict__ sc2, double *__restrict__ full) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // full, empty, tmem_full[2], tmem_empty[2], q_full, meta_full, meta_empty
    __shared__ __align__(8) uint64_t bars[2 * MAX_SLOTS + 2 * 2 + 1 + 2 * META_RING];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(128) uint8_t s_meta[META_RING][meta_block_bytes(WIDE)];
    __shared__ float s_thr[2][BM];                           // screening thresholds, exchanged between the two warps of a query row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                  // 0 = leader
    const uint32_t smem_base = (s2u(smem) + 1023u) & ~1023u;   // same offset in both CTAs (same kernel, same layout)
    const uint32_t q_bytes = RESIDENT ? uint32_t(k_chunks) * SLOT_BYTES : 0u;
    const uint32_t ring_base = smem_base + q_bytes;
    const uint32_t full0 = s2u(&bars[0]), empty0 = s2u(&bars[MAX_SLOTS]), tfull0 = s2u(&bars[2 * MAX_SLOTS]),
                   tempty0 = s2u(&bars[2 * MAX_SLOTS + 2]), qfull = s2u(&bars[2 * MAX_SLOTS + 4]),
                   mfull0 = s2u(&bars[2 * MAX_SLOTS + 5]), mempty0 = s2u(&bars[2 * MAX_SLOTS + 5 + META_RING]);
    constexpr uint32_t tmem_cols = 512;                     // the whole TMEM of the SM (one CTA per SM)
    const int n_tiles = __ldg(header);

    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, 2 * EPI_WARPS); }
        mbar_init(qfull, 1);
        for (int i = 0; i < META_RING; ++i) { mbar_init(mfull0 + 8 * i, 1); mbar_init(mempty0 + 8 * i, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {        // TMEM allocation for the pair: one warp of each CTA, the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // the peer's barriers are initialised before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int64_t q0 = (int64_t(blockIdx.x >> 1) * 2 + rank) * BM;       // this CTA's 128 queries

    // TMEM halves a tile takes: [half, half + nh); a 2-half tile starts at half 0 (a free half 1 is skipped)
    auto tile_halves = [](const TileDesc &td, uint32_t &half) -> uint32_t {
        const uint32_t nh = (td.width == WIDE && td.np == 1) ? 1u : 2u;
        if (nh == 2u) half = 0u;
        return nh;
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own query rows, own half of every annotation tile; bytes land on the LEADER's barriers.
        // The whole warp walks the loop (waits are warp-uniform); one elected lane issues. =====
        const uint32_t qfull_leader = mapa_rank(qfull, 0);
        const uint32_t full0_leader = mapa_rank(full0, 0);
        if (RESIDENT) {
            if (elect_one()) {
                if (rank == 0) mbar_arrive_expect_tx(qfull, 2 * q_bytes);
                for (int kc = 0; kc < k_chunks; ++kc) tma_load_2d_pair(smem_base + kc * SLOT_BYTES, &map_q, qfull_leader, kc * BK, int(q0));
            }
            __syncwarp();
        }
        int slot = 0;
        uint32_t phase = 0;
        auto push = [&](const CUtensorMap *map, int c0, int c1, uint32_t bytes) {
            mbar_wait(empty0 + 8 * slot, phase ^ 1);
            if (elect_one()) {
                if (rank == 0) mbar_arrive_expect_tx(full0 + 8 * slot, 2 * bytes);
                tma_load_2d_pair(ring_base + slot * SLOT_BYTES, map, full0_leader + 8 * slot, c0, c1);
            }
            __syncwarp();
            if (++slot == slots) { slot = 0; phase ^= 1; }
        };
        int mbuf = 0;
        uint32_t mphase = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const TileDesc td = tiles[t];
            const CUtensorMap *map = td.width == WIDE ? &map_wide : &map_narrow;
            const int half_rows = td.width / 2;
            // this tile's epilogue operands: one bulk copy into the next tile-meta buffer (each CTA keeps its own copy)
            mbar_wait(mempty0 + 8 * mbuf, mphase ^ 1);
            if (elect_one()) {
                const uint32_t bytes = uint32_t(meta_block_bytes(td.width));
                mbar_arrive_expect_tx(mfull0 + 8 * mbuf, bytes);
                bulk_load(s2u(&s_meta[mbuf][0]), tile_meta + meta_block_offset(td.row0), bytes, mfull0 + 8 * mbuf);
            }
            __syncwarp();
            if (++mbuf == META_RING) { mbuf = 0; mphase ^= 1; }
            for (int kc = 0; kc < k_chunks; ++kc) {
                if (!RESIDENT) push(&map_q, kc * BK, int(q0), SLOT_BYTES);
                for (int j = 0; j < td.np; ++j) push(map, kc * BK, j * plane_rows + td.row0 + int(rank) * half_rows, uint32_t(half_rows) * BK);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only): the whole warp walks the loop, one elected lane issues.  Descriptors differ
        // only in their 14-bit start-address field, so an operand K-step is one add on a precomputed descriptor. =====
        if (rank == 0) {
            int slot = 0;
            uint32_t phase = 0;
            uint32_t half = 0, par_empty = 0;                   // bit h of par_empty: parity of the uses of TMEM half h so far
            const uint64_t desc_q0 = smem_desc(smem_base), desc_ring0 = smem_desc(ring_base);
            constexpr uint64_t slot_step = SLOT_BYTES >> 4, k_step = UK >> 4;
            if (RESIDENT) { mbar_wait(qfull, 0); tc_fence_after(); }
            for (int t = 0; t < n_tiles; ++t) {
                const TileDesc td = tiles[t];
                const uint32_t nh = tile_halves(td, half);
                for (uint32_t h = half; h < half + nh; ++h) {   // both CTAs' epilogues have drained the halves this tile takes
                    mbar_wait(tempty0 + 8 * h, ((par_empty >> h) & 1u) ^ 1u);
                    par_empty ^= 1u << h;
                }
                tc_fence_after();
                if (lane == 0) TL(0, t);
                const uint32_t idesc = td.width == WIDE ? IDESC_WIDE : IDESC_NARROW;
                const uint32_t d0 = tmem_base + half * WIDE;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    uint64_t a_desc = desc_q0 + uint64_t(kc) * slot_step;
                    int a_slot = -1;
                    if (!RESIDENT) {
                        mbar_wait(full0 + 8 * slot, phase);
                        a_desc = desc_ring0 + uint64_t(slot) * slot_step;
                        a_slot = slot;
                        if (++slot == slots) { slot = 0; phase ^= 1; }
                    }
                    for (int j = 0; j < td.np; ++j) {
                        mbar_wait(full0 + 8 * slot, phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t b_desc = desc_ring0 + uint64_t(slot) * slot_step;
                            const uint32_t d = d0 + uint32_t(j * td.width);
                            if (kc == 0) {
                                umma_i8_pair<false>(d, a_desc, b_desc, idesc);
                            } else {
                                umma_i8_pair<true>(d, a_desc, b_desc, idesc);
                            }
#pragma unroll
                            for (int kk = 1; kk < BK / UK; ++kk) umma_i8_pair<true>(d, a_desc + kk * k_step, b_desc + kk * k_step, idesc);
                            umma_commit_pair(empty0 + 8 * slot);    // slot free in both CTAs once these MMAs have read it
                        }
                        __syncwarp();
                        if (++slot == slots) { slot = 0; phase ^= 1; }
                    }
                    if (a_slot >= 0) {
                        if (elect_one()) umma_commit_pair(empty0 + 8 * a_slot);
                        __syncwarp();
                    }
                }
                if (elect_one()) umma_commit_pair(tfull0 + 8 * half);       // accumulators of tile t complete (keyed by its first half)
                __syncwarp();
                if (lane == 0) TL(1, t);
                half = (half + nh) & 1u;
            }
        }
    } else {
        // ===== epilogue (both CTAs): 8 warps, TMEM lane quarter = warp % 4, column half = (warp - 2) / 4; the warps run
        // independently.  The two warps of a row only share their screening thresholds (a bound from either half holds). =====
        const int quarter = warp & 3;
        const int ch = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int64_t q = q0 + row;
        const double qn2 = (q < nq) ? qnorm2[q] : 0.0;
        const double inv_qn = qn2 > 0.0 ? 1.0 / sqrt(qn2) : 0.0;
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        const uint32_t tempty_leader = mapa_rank(tempty0, 0);
        Top2d best{0.0, 0.0, -1, -1};                             // FULLOUT: every score is formed in float64 anyway
        Top2x bx{0.f, 0.f, 0, 0, 0.0, 0.0, -1, -1};               // screening mode
        float thr = 1e-30f;                                       // > 0: zero dots are never candidates (see the final fill)
        volatile float *my_thr = &s_thr[ch][row], *other_thr = &s_thr[ch ^ 1][row];
        *my_thr = thr;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t half = 0, par_full = 0;
        int mbuf = 0;
        uint32_t mphase = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const TileDesc td = tiles[t];
            mbar_wait(mfull0 + 8 * mbuf, mphase);                  // the producer's bulk copy of this tile's meta has landed
            const EpiTile tile(&s_meta[mbuf][0], td.width);
            const uint32_t nh = tile_halves(td, half);
            mbar_wait(tfull0 + 8 * half, (par_full >> half) & 1u);
            par_full ^= 1u << half;
            tc_fence_after();
            if (lane == 0 && warp == 2) TL(2, t);
            if (lane == 0 && warp == 6) TL(4, t);
            const uint32_t tmem = lane_addr + half * WIDE;
            const uint32_t pstride = uint32_t(td.width);
            const int cbeg = ch * (td.width >> 1), cend = cbeg + (td.width >> 1);
            *my_thr = thr;
            thr = fmaxf(thr, *other_thr);
            if (FULLOUT) {
                double *full_row = full + (q < nq ? q : 0) * int64_t(n_ann);
                switch (td.np) {
                    case 1: epi_tile_full<1>(tmem, pstride, cbeg, cend, tile, inv_qn, best, full_row, q < nq); break;
                    case 2: epi_tile_full<2>(tmem, pstride, cbeg, cend, tile, inv_qn, best, full_row, q < nq); break;
                    case 3: epi_tile_full<3>(tmem, pstride, cbeg, cend, tile, inv_qn, best, full_row, q < nq); break;
                    default: epi_tile_full<4>(tmem, pstride, cbeg, cend, tile, inv_qn, best, full_row, q < nq); break;
                }
            } else {
                switch (td.np) {
                    case 1: epi_tile_screen<1>(tmem, pstride, cbeg, cend, tile, inv_qn, bx, thr); break;
                    case 2: epi_tile_screen<2>(tmem, pstride, cbeg, cend, tile, inv_qn, bx, thr); break;
                    case 3: epi_tile_screen<3>(tmem, pstride, cbeg, cend, tile, inv_qn, bx, thr); break;
                    default: epi_tile_screen<4>(tmem, pstride, cbeg, cend, tile, inv_qn, bx, thr); break;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && warp == 2) TL(3, t);
            if (lane == 0 && warp == 6) TL(5, t);
            if (lane == 0) {
                for (uint32_t h = half; h < half + nh; ++h) mbar_arrive_cluster(tempty_leader + 8 * h);   // 16 arrivals free a half
                mbar_arrive(mempty0 + 8 * mbuf);                                                          // 8 arrivals free the meta buffer
            }
            if (++mbuf == META_RING) { mbuf = 0; mphase ^= 1; }
            half = (half + nh) & 1u;
        }
        // the warps of the second column half hand their top-2 over (the tile-meta buffers are free by now)
        if (!FULLOUT) {   // the exact scores of the two survivors, with the tensor pipe idle
            best.i1 = bx.i1; best.s1 = bx.i1 >= 0 ? exact_score(bx.d1, inv_qn, bx.im1) : 0.0;
            best.i2 = bx.i2; best.s2 = bx.i2 >= 0 ? exact_score(bx.d2, inv_qn, bx.im2) : 0.0;
        }
        Top2d *xch = reinterpret_cast<Top2d *>(&s_meta[0][0]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ch == 1) xch[row] = best;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ch == 0) {
            const Top2d o = xch[row];
            if (o.i1 >= 0) top2d_push(best, o.s1, o.i1);
            if (o.i2 >= 0) top2d_push(best, o.s2, o.i2);
        }
        if (ch == 0 && q < nq) {
            // rows that were never candidates score exactly 0: a missing winner / runner-up is the lowest-index one of them
            if (best.i1 < 0) { best.i1 = 0; best.s1 = 0.0; }
            if (best.i2 < 0 && n_ann > 1) { best.i2 = (best.i1 == 0) ? 1 : 0; best.s2 = 0.0; }
            top1[q] = best.i1; sc1[q] = best.s1;
            top2[q] = best.i2; sc2[q] = best.i2 >= 0 ? best.s2 : nan("");
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // neither CTA leaves (or frees TMEM) while the pair's MMAs / remote arrivals may still touch it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// tile meta (see EpiTile): one block of 256 threads per tile
__global__ void __launch_bounds__(WIDE) tile_meta_kernel(const int32_t *__restrict__ header, const TileDesc *__restrict__ tiles,
                                                         const int32_t *__restrict__ perm, const double *__restrict__ mnorm2,
                                                         uint8_t *__restrict__ out) {
    if (int(blockIdx.x) >= __ldg(header)) return;
    const TileDesc td = tiles[blockIdx.x];
    const int i = threadIdx.x;
    if (i >= td.width) return;                                   // whole warps (width is 128 or 256)
    uint8_t *base = out + meta_block_offset(td.row0);
    const int32_t orig = perm[td.row0 + i];
    const double m2 = (orig >= 0) ? mnorm2[orig] : 0.0;
    const double inv = m2 > 0.0 ? 1.0 / sqrt(m2) : 0.0;
    const float inv32 = float(inv);
    reinterpret_cast<double *>(base)[i] = inv;
    reinterpret_cast<float *>(base + 8 * td.width)[i] = inv32;
    reinterpret_cast<int32_t *>(base + 12 * td.width)[i] = orig;
    float g8 = inv32;
    g8 = fmaxf(g8, __shfl_xor_sync(FULL, g8, 1));
    g8 = fmaxf(g8, __shfl_xor_sync(FULL, g8, 2));
    g8 = fmaxf(g8, __shfl_xor_sync(FULL, g8, 4));
    if ((i & 7) == 0) reinterpret_cast<float *>(base + 16 * td.width)[i >> 3] = g8;
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
// M int64 [A, K] -> planes uint8 [n_planes, rows, Kp]: digit j (base 256) of row perm[a]; padding rows (perm < 0) and
// columns beyond K are zero
__global__ void __launch_bounds__(256) split_m_kernel(const int64_t *__restrict__ M, int64_t K, int64_t rows, int64_t Kp,
                                                      int n_planes, const int32_t *__restrict__ perm, uint8_t *__restrict__ out) {
    const int64_t total = rows * Kp;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t a = i / Kp, c = i - a * Kp;
        const int32_t o = __ldg(perm + a);
        const uint64_t v = (o >= 0 && c < K) ? uint64_t(__ldg(M + int64_t(o) * K + c)) : 0ull;
        for (int j = 0; j < n_planes; ++j) out[(int64_t(j) * rows + a) * Kp + c] = uint8_t(v >> (8 * j));
    }
}
// largest entry (negative entries are not representable -> all-ones) and squared norm of every row, one warp per row
__global__ void __launch_bounds__(256) row_max_kernel(const int64_t *__restrict__ X, int64_t rows, int64_t cols,
                                                      unsigned long long *__restrict__ out, double *__restrict__ norm2) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        unsigned long long m = 0;
        double n2 = 0.0;
        for (int64_t c = lane; c < cols; c += 32) {
            const int64_t v = X[r * cols + c];
            const unsigned long long u = v < 0 ? ~0ull : (unsigned long long)v;
            m = u > m ? u : m;
            n2 += double(v) * double(v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long x = __shfl_xor_sync(FULL, m, o);
            m = x > m ? x : m;
            n2 += __shfl_xor_sync(FULL, n2, o);
        }
        if (lane == 0) { out[r] = m; norm2[r] = n2; }
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
// uint8 matrix [rows, Kp] row-major, boxes of box_rows rows x 128 bytes, 128-byte swizzle, zero fill outside
static int make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t Kp, int box_rows) {
    PFN_cuTensorMapEncodeTiled fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return SKM_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp};
    cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return SKM_ERR_CUDA; }
    return SKM_OK;
}

static int64_t pad128(int64_t x) { return (x + 127) & ~int64_t(127); }
// rows of the prepared matrix: the narrow region is padded to 128, the wide region to 256
static int64_t plane_rows_for(int64_t n_ann) { return pad128(n_ann + 127 + 255); }
// prepared-matrix blob: [header 1024 B: n_tiles, plane_rows | perm int32 rows | TileDesc rows/128 | pad to 1024]
//                       [planes MAX_PLANES x rows x Kp]
static size_t perm_offset() { return 1024; }
static size_t tiles_offset(int64_t rows) { return perm_offset() + size_t(rows) * 4; }
static size_t meta_bytes(int64_t rows) { return (tiles_offset(rows) + size_t(rows / 128) * sizeof(TileDesc) + 1023) & ~size_t(1023); }

}  // namespace tc
}  // namespace skm

extern "C" {

#ifdef SKM_TC_PROF
// debug builds only: timeline of CTA pair 0 (clock64 per tile): MMA tile start / issued, epilogue warp 2 and 6 start / end
SKM_API int skm_debug_tc_timeline(long long *host_out) {
    return cudaMemcpyFromSymbol(host_out, skm::tc::g_tl, sizeof(skm::tc::g_tl)) == cudaSuccess ? 0 : -1;
}
#endif

size_t skm_apply_tc_planes_bytes(int64_t n_ann, int64_t K) {
    using namespace skm::tc;
    if (n_ann <= 0 || K <= 0) return 2048;
    const int64_t rows = plane_rows_for(n_ann);
    return meta_bytes(rows) + size_t(MAX_PLANES) * size_t(rows) * size_t(pad128(K)) + 256;
}

int skm_apply_tc_prepare(const int64_t *d_M, int64_t n_ann, int64_t K, uint8_t *d_planes, size_t planes_bytes,
                         int *n_planes_out, skm_stream_t stream) {
    using namespace skm;
    using namespace skm::tc;
    if (!n_planes_out) { set_error("skm_apply_tc_prepare: n_planes_out is NULL"); return SKM_ERR_INVALID; }
    *n_planes_out = 0;
    if (n_ann <= 0 || K <= 0) { set_error("skm_apply_tc_prepare: empty matrix"); return SKM_ERR_INVALID; }
    if (K > MAX_K) { set_error("skm_apply_tc_prepare: K=%lld > %lld (int32 accumulators would overflow); use skm_apply_dense", (long long)K, (long long)MAX_K); return SKM_ERR_UNSUPPORTED; }
    if (n_ann > 0x7FFFF000ll) { set_error("skm_apply_tc_prepare: too many annotations"); return SKM_ERR_UNSUPPORTED; }
    if (!d_M || !d_planes || planes_bytes < skm_apply_tc_planes_bytes(n_ann, K)) { set_error("skm_apply_tc_prepare: bad buffers"); return SKM_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(d_planes) & 127u) != 0) { set_error("skm_apply_tc_prepare: d_planes must be 128-byte aligned"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = plane_rows_for(n_ann), Kp = pad128(K);
    // per-row maxima decide the digit planes a row needs (the plane area of the blob is scratch until the split)
    uint8_t *planes = d_planes + meta_bytes(rows);
    unsigned long long *d_rowmax = reinterpret_cast<unsigned long long *>(planes);
    double *d_rownorm2 = reinterpret_cast<double *>(d_rowmax + n_ann);
    row_max_kernel<<<(int)std::min<int64_t>((n_ann + 7) / 8, int64_t(sm_count()) * 8), 256, 0, st>>>(d_M, n_ann, K, d_rowmax, d_rownorm2);
    SKM_LAUNCH_CHECK("row_max_kernel");
    std::vector<unsigned long long> rowmax((size_t)n_ann);
    std::vector<double> rownorm2((size_t)n_ann);
    SKM_CUDA_TRY(cudaMemcpyAsync(rowmax.data(), d_rowmax, size_t(n_ann) * 8, cudaMemcpyDeviceToHost, st));
    SKM_CUDA_TRY(cudaMemcpyAsync(rownorm2.data(), d_rownorm2, size_t(n_ann) * 8, cudaMemcpyDeviceToHost, st));
    SKM_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int> cls((size_t)n_ann);
    int max_planes = 1;
    int64_t n_narrow = 0;
    for (int64_t a = 0; a < n_ann; ++a) {
        int p = 1;
        while (p < 8 && (rowmax[a] >> (8 * p)) != 0) ++p;
        cls[a] = p;
        max_planes = std::max(max_planes, p);
        n_narrow += p > 2;
    }
    if (max_planes > MAX_PLANES) { set_error("skm_apply_tc_prepare: entries need %d base-256 digit planes (> %d); use skm_apply_dense", max_planes, MAX_PLANES); return SKM_ERR_UNSUPPORTED; }
    // rows sorted by class, largest first, and by norm inside a class: neighbouring rows then have nearly equal norms,
    // which is what lets the epilogue screen 8 accumulators with one bound (epi_tile_screen).  Classes 3-4 (entries
    // >= 2^16) form the narrow region, classes 1-2 the wide region.
    std::vector<int32_t> order((size_t)n_ann);
    for (int64_t a = 0; a < n_ann; ++a) order[a] = (int32_t)a;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return cls[x] != cls[y] ? cls[x] > cls[y] : rownorm2[x] > rownorm2[y];
    });
    const int64_t narrow_rows = pad128(n_narrow), n_wide = n_ann - n_narrow;
    std::vector<int32_t> meta(meta_bytes(rows) / 4, 0);
    int32_t *perm = meta.data() + perm_offset() / 4;
    TileDesc *tiles = reinterpret_cast<TileDesc *>(meta.data() + tiles_offset(rows) / 4);
    for (int64_t r = 0; r < rows; ++r) perm[r] = -1;
    for (int64_t i = 0; i < n_narrow; ++i) perm[i] = order[i];
    for (int64_t i = 0; i < n_wide; ++i) perm[narrow_rows + i] = order[n_narrow + i];
    int32_t n_tiles = 0;
    for (int64_t r0 = 0; r0 < n_narrow; r0 += NARROW) tiles[n_tiles++] = TileDesc{(int32_t)r0, NARROW, cls[perm[r0]], 0};
    for (int64_t r0 = 0; r0 < n_wide; r0 += WIDE) tiles[n_tiles++] = TileDesc{(int32_t)(narrow_rows + r0), WIDE, cls[perm[narrow_rows + r0]], 0};
    meta[0] = n_tiles;
    meta[1] = (int32_t)rows;
    SKM_CUDA_TRY(cudaMemcpyAsync(d_planes, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, st));
    SKM_CUDA_TRY(cudaStreamSynchronize(st));          // `meta` is a local
    split_m_kernel<<<(int)std::min<int64_t>((rows * Kp + 255) / 256, int64_t(sm_count()) * 16), 256, 0, st>>>(
        d_M, K, rows, Kp, max_planes, reinterpret_cast<const int32_t *>(d_planes + perm_offset()), planes);
    SKM_LAUNCH_CHECK("split_m_kernel");
    *n_planes_out = max_planes;
    return SKM_OK;
}

size_t skm_apply_tc_workspace(int64_t nq, int64_t K, int64_t n_ann) {
    using namespace skm::tc;
    if (nq <= 0 || K <= 0 || n_ann <= 0) return 512;
    return size_t(nq) * size_t(pad128(K)) + meta_block_offset(plane_rows_for(n_ann)) + 1024;
}

int skm_apply_tc(const int32_t *d_Q, int64_t nq, int64_t K, const uint8_t *d_planes, int n_planes, int64_t n_ann,
                 const double *d_qnorm2, const double *d_mnorm2, int32_t *d_top1, int32_t *d_top2, double *d_score1,
                 double *d_score2, double *d_scores_full, int *d_status, void *workspace, size_t workspace_bytes,
                 skm_stream_t stream) {
    using namespace skm;
    using namespace skm::tc;
    if (nq < 0 || K <= 0 || K > MAX_K || n_ann <= 0 || n_ann > 0x7FFFF000ll || n_planes < 1 || n_planes > MAX_PLANES) { set_error("skm_apply_tc: bad sizes (K=%lld n_ann=%lld planes=%d)", (long long)K, (long long)n_ann, n_planes); return SKM_ERR_INVALID; }
    if (nq == 0) return SKM_OK;
    if (nq > 0x7FFFFF00ll) { set_error("skm_apply_tc: more than 2^31 queries per call"); return SKM_ERR_UNSUPPORTED; }
    if (!d_Q || !d_planes || !d_qnorm2 || !d_mnorm2 || !d_top1 || !d_top2 || !d_score1 || !d_score2 || !d_status) { set_error("skm_apply_tc: NULL argument"); return SKM_ERR_INVALID; }
    const size_t need = skm_apply_tc_workspace(nq, K, n_ann);
    if (!workspace || workspace_bytes < need) { set_error("skm_apply_tc: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Kp = pad128(K), rows = plane_rows_for(n_ann);
    const int32_t *header = reinterpret_cast<const int32_t *>(d_planes);
    const int32_t *perm = reinterpret_cast<const int32_t *>(d_planes + perm_offset());
    const TileDesc *tiles = reinterpret_cast<const TileDesc *>(d_planes + tiles_offset(rows));
    const uint8_t *planes = d_planes + meta_bytes(rows);
    uint8_t *tile_meta = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    uint8_t *q8 = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tile_meta + meta_block_offset(rows)) + 255) & ~uintptr_t(255));
    SKM_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    tile_meta_kernel<<<(unsigned)(rows / NARROW), WIDE, 0, st>>>(header, tiles, perm, d_mnorm2, tile_meta);
    SKM_LAUNCH_CHECK("tile_meta_kernel");
    split_q_kernel<<<(int)std::min<int64_t>((nq * (Kp >> 2) + 255) / 256, int64_t(sm_count()) * 16), 256, 0, st>>>(d_Q, nq, K, Kp, q8, d_status);
    SKM_LAUNCH_CHECK("split_q_kernel");
    alignas(64) CUtensorMap map_q, map_wide, map_narrow;
    int rc = make_map(&map_q, q8, nq, Kp, BM);
    if (rc) return rc;
    rc = make_map(&map_wide, planes, int64_t(n_planes) * rows, Kp, WIDE / 2);
    if (rc) return rc;
    rc = make_map(&map_narrow, planes, int64_t(n_planes) * rows, Kp, NARROW / 2);
    if (rc) return rc;
    const int k_chunks = int(Kp / BK);
    const bool resident = k_chunks <= 8;                       // 128 queries x 1024 bytes = 128 KB of the 227 KB
    const size_t budget = 227 * 1024 - 14336 - 1024;          // static shared (barriers + tile-meta ring) + alignment slack
    const size_t q_bytes = resident ? size_t(k_chunks) * SLOT_BYTES : 0;
    int slots = int((budget - q_bytes) / SLOT_BYTES);
    if (slots > MAX_SLOTS) slots = MAX_SLOTS;
    if (slots < 2) { set_error("skm_apply_tc: pipeline does not fit shared memory"); return SKM_ERR_UNSUPPORTED; }
    // > half of the SM's shared memory: one CTA per SM, which the 512-column TMEM allocation relies on
    const size_t smem = std::max<size_t>(q_bytes + size_t(slots) * SLOT_BYTES + 1024, 120 * 1024);
    const unsigned grid = 2u * (unsigned)((nq + 2 * BM - 1) / (2 * BM));          // CTA pairs of 256 queries
#define SKM_LAUNCH_TC(RES, FULLOUT)                                                                                          \
    {                                                                                                                        \
        auto kern = apply_tc_kernel<RES, FULLOUT>;                                                                           \
        SKM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        kern<<<grid, THREADS, smem, st>>>(map_q, map_wide, map_narrow, (int)rows, nq, (int)n_ann, k_chunks, slots, tile_meta, \
                                          header, tiles, d_qnorm2, d_top1, d_top2, d_score1, d_score2, d_scores_full);       \
    }
    if (resident) { if (d_scores_full) SKM_LAUNCH_TC(true, true) else SKM_LAUNCH_TC(true, false) }
    else { if (d_scores_full) SKM_LAUNCH_TC(false, true) else SKM_LAUNCH_TC(false, false) }
#undef SKM_LAUNCH_TC
    SKM_LAUNCH_CHECK("apply_tc_kernel");
    return SKM_OK;
}

}  // extern "C"

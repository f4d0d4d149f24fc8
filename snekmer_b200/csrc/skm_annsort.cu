// skm_annsort.cu — learn (c) for the light annotations: one CTA sorts one task in shared memory.
//
// filter_and_construct / the annotation sums (learn.smk:306-326, 385-408) for a COO matrix: rank r's sequences are
// gathered by annotation, so the windows of an annotation are one contiguous range of residues.  Instead of writing a
// key per window to HBM and radix-sorting all of them device-wide (skm_learn_sparse_group: ~40 bytes per window), a TASK
// = (annotation, code range [code_lo, code_hi)) with at most AS_CAP windows is handled by ONE CTA entirely on chip:
//   1. scan     the annotation's residues are staged and scanned (skm_tile.cuh); windows whose code lies in the task's
//               range are appended to a shared-memory key buffer (warp-aggregated cursor);
//   2. sort     LSD radix sort of the (code - code_lo) keys in shared memory, 8 bits per pass, only the bits the range
//               needs (21 bits = 3 passes for the whole 6-letter k = 8 space); ranks come from warp match_any + per-warp
//               digit histograms, so there are no shared-memory atomics;
//   3. encode   run-length encode; the task's place in the output is a decoupled look-back over the per-task entry counts
//               (tasks take tickets in (annotation, code_lo) order, so the output is one globally sorted COO list and is
//               written exactly once); the Totals row (learn.smk:380) gets one integer atomic per ENTRY.
// An annotation with no more than AS_CAP residues is one task; larger ("medium") ones are cut into code ranges from an
// exact per-annotation histogram of the leading code digits (ann_hist_kernel), so a task never overflows; each of its
// tasks re-scans the annotation's residues (L2-resident).  Heavy annotations stay on the dense-row path (skm_rows.cu).
// Integer counting only: bit-identical to the sorted path.
#include <algorithm>

#include "skm_common.cuh"
#include "skm_tile.cuh"

namespace skm {

constexpr int AS_THREADS = 1024;
constexpr int AS_WARPS = AS_THREADS / 32;
constexpr int AS_CAP = 26 * 1024;                                   // windows per task
constexpr int AS_SEG = ts_seg_cap(12, AS_THREADS);                  // 12,288 positions per staged segment
constexpr int AS_SYM_BYTES = (ts_sym_bytes(AS_SEG) + 15) & ~15;
constexpr int AS_BINS = 256;                                        // 8-bit digits
constexpr int AS_HIST_BYTES = AS_WARPS * AS_BINS * 2;               // uint16 per (warp, digit)
constexpr int AS_SMEM = 2 * 4 * AS_CAP + AS_HIST_BYTES;             // two key buffers + histograms (symbols alias buffer B)
static_assert(AS_SYM_BYTES <= 4 * AS_CAP, "the staged symbols alias key buffer B");
static_assert(AS_CAP < 65536, "histogram offsets are 16-bit");
constexpr unsigned long long AS_VAL = (1ull << 62) - 1;

struct AsTasks {
    const int32_t *seq_lo, *seq_hi;       // sequences [seq_lo, seq_hi) of the gathered batch = the task's annotation
    const uint32_t *code_lo, *code_hi;    // code range of the task
    const int64_t *ann;                   // annotation id (key = ann * S + code)
    const int64_t *out_base;              // entries that other paths place in front of this annotation
};

__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ void st_vol(unsigned long long *p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long *>(p) = v; }

// One CTA scans the sequences [s_lo, s_hi) (contiguous residues) segment by segment; emit(code, ok) for every position.
template <typename Emit>
__device__ __forceinline__ void as_scan_seqs(const uint8_t *__restrict__ res, int64_t nres, const int64_t *__restrict__ off, int64_t s_lo,
                                             int64_t s_hi, const uint8_t *s_lut, uint8_t *s_sym, int seg_cap, int k, uint32_t nsym,
                                             uint32_t pow_k1, unsigned *s_nstart, Emit &&emit) {
    const int tid = threadIdx.x;
    uint32_t sym_addr = smem_addr(s_sym);
    const int64_t r_lo = __ldg(off + s_lo), r_hi = __ldg(off + s_hi);
    int64_t cur = s_lo;
    bool first = true;
    uint32_t tail = 0;
    for (int64_t a = r_lo; a < r_hi;) {
        const int64_t b = min(r_hi, (a + seg_cap) & ~int64_t(15));
        if (!first) ts_tail_write(s_sym, tail, k);
        const TsSeg g = ts_stage(res, nres, a, b, s_lut, s_sym);
        __syncthreads();
        if (first) ts_invalidate_front(s_sym, g);
        unsigned mine = 0;
        for (int64_t s = cur + tid; s < s_hi; s += blockDim.x) {
            const int64_t o = __ldg(off + s);
            if (o >= b) break;
            s_sym[g.lo + int(o - a)] |= uint8_t(SYM_FLAG);
            ++mine;
        }
        if (mine) atomicAdd(s_nstart, mine);
        __syncthreads();
        const int64_t nstart = *s_nstart;
        const int C = ts_chunk(g.hi - g.lo);
        const int i0 = g.lo + tid * C, i1 = min(i0 + C, g.hi);
        if (i0 < i1)
            ts_scan_chunk<uint32_t>(sym_addr, i0, i1, k, nsym, pow_k1, [&](uint32_t, uint32_t code, bool ok) { emit(code, ok); }, [](uint32_t) {});
        tail = ts_tail_read(s_sym, g, k);
        __syncthreads();
        first = false;
        cur += nstart;
        a = b;
        if (tid == 0) *s_nstart = 0;          // ordered before the next atomicAdd by the barrier after staging
    }
}

// exclusive scan of one value per thread over the CTA (1024 threads); *total gets the sum.  s_w: 33 words of scratch.
__device__ __forceinline__ uint32_t as_block_scan(uint32_t v, uint32_t *s_w, uint32_t *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += x; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t t = s_w[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, ti, o); if (lane >= o) ti += x; }
        s_w[lane] = ti - t;
        if (lane == 31) s_w[32] = ti;
    }
    __syncthreads();
    const uint32_t r = s_w[w] + incl - v;
    *total = s_w[32];
    __syncthreads();                          // s_w is reused by the next scan
    return r;
}

__global__ void __launch_bounds__(AS_THREADS, 1) ann_sort_kernel(const uint8_t *__restrict__ res, int64_t nres, const int64_t *__restrict__ off,
                                                                 const uint8_t *__restrict__ lut, uint32_t nsym, int k, uint32_t pow_k1,
                                                                 const AsTasks t, int n_tasks, uint64_t S, unsigned *__restrict__ ticket,
                                                                 unsigned long long *__restrict__ chain, uint64_t *__restrict__ keys_out,
                                                                 int64_t *__restrict__ vals_out, int64_t capacity, int64_t *__restrict__ task_prefix,
                                                                 unsigned long long *__restrict__ totals, int64_t *__restrict__ total_out,
                                                                 int *__restrict__ overflow) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *buf_a = reinterpret_cast<uint32_t *>(smem);
    uint32_t *buf_b = buf_a + AS_CAP;
    uint16_t *hist = reinterpret_cast<uint16_t *>(buf_b + AS_CAP);
    uint8_t *s_sym = reinterpret_cast<uint8_t *>(buf_b);
    __shared__ uint8_t s_lut[256];
    __shared__ unsigned s_nstart, s_cnt;
    __shared__ int s_task;
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_w[33];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    ts_lut_init(s_lut, lut);
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_task = (int)atomicAdd(ticket, 1u); s_cnt = 0; s_nstart = 0; }
        __syncthreads();
        const int task = s_task;
        if (task >= n_tasks) break;
        const uint32_t c_lo = __ldg(t.code_lo + task), range = __ldg(t.code_hi + task) - c_lo;
        // ---- 1. scan: keys of the task's windows -> buf_a ---------------------------------------------------------
        as_scan_seqs(res, nres, off, __ldg(t.seq_lo + task), __ldg(t.seq_hi + task), s_lut, s_sym, AS_SEG, k, nsym, pow_k1, &s_nstart,
                     [&](uint32_t code, bool ok) {
                         const uint32_t rel = code - c_lo;
                         const bool pred = ok && rel < range;
                         const unsigned am = __activemask();
                         const unsigned b = __ballot_sync(am, pred);
                         if (b) {
                             const int leader = __ffs(b) - 1;
                             unsigned base = 0;
                             if (lane == leader) base = atomicAdd(&s_cnt, (unsigned)__popc(b));
                             base = __shfl_sync(am, base, leader);
                             if (pred) {
                                 const unsigned pos = base + __popc(b & lt);
                                 if (pos < (unsigned)AS_CAP) buf_a[pos] = rel;
                             }
                         }
                     });
        __syncthreads();
        int n = (int)s_cnt;
        if (n > AS_CAP) {                     // cannot happen with tasks cut from the exact histogram; the host then redoes the step
            if (tid == 0) atomicOr(overflow, 1);
            n = 0;
        }
        // ---- 2. sort: LSD radix, 8 bits per pass, over the bits of range - 1 -------------------------------------------
        const int bits = range > 1u ? 32 - __clz(range - 1u) : 0;
        uint32_t *src = buf_a, *dst = buf_b;
        const int m = (((n + AS_WARPS - 1) / AS_WARPS) + 31) & ~31;      // keys per warp, whole rows of 32
        const int beg = w * m, end = min(beg + m, n);
        for (int sh = 0; sh < bits && n > 1; sh += 8) {
            for (int i = tid; i < AS_HIST_BYTES / 4; i += AS_THREADS) reinterpret_cast<uint32_t *>(hist)[i] = 0u;
            __syncthreads();
            uint16_t *my = hist + w * AS_BINS;
            for (int i0 = beg; i0 < end; i0 += 32) {                     // per-warp digit counts
                const int i = i0 + lane;
                const bool v = i < end;
                const unsigned vm = __ballot_sync(FULL, v);
                if (v) {
                    const uint32_t d = (src[i] >> sh) & 255u;
                    const unsigned peers = __match_any_sync(vm, d);
                    if (lane == __ffs(peers) - 1) my[d] = uint16_t(my[d] + __popc(peers));
                }
                __syncwarp();
            }
            __syncthreads();
            {   // exclusive scan in (digit, warp) order: thread = (digit, group of 8 warps)
                const int d = tid >> 2, w0 = (tid & 3) * 8;
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = hist[(w0 + j) * AS_BINS + d]; sum += v[j]; }
                uint32_t total;
                uint32_t run = as_block_scan(sum, s_w, &total);
#pragma unroll
                for (int j = 0; j < 8; ++j) { hist[(w0 + j) * AS_BINS + d] = uint16_t(run); run += v[j]; }
            }
            __syncthreads();
            for (int i0 = beg; i0 < end; i0 += 32) {                     // stable scatter
                const int i = i0 + lane;
                const bool v = i < end;
                const unsigned vm = __ballot_sync(FULL, v);
                uint32_t d = 0, base = 0;
                unsigned peers = 0;
                if (v) {
                    const uint32_t key = src[i];
                    d = (key >> sh) & 255u;
                    peers = __match_any_sync(vm, d);
                    base = my[d];
                    dst[base + __popc(peers & lt)] = key;
                }
                __syncwarp();
                if (v && lane == __ffs(peers) - 1) my[d] = uint16_t(base + __popc(peers));
                __syncwarp();
            }
            __syncthreads();
            uint32_t *x = src; src = dst; dst = x;
        }
        // ---- 3. run-length encode: positions of the run heads -> dst, then the entries go out ----------------------------
        const int c = (n + AS_THREADS - 1) / AS_THREADS;
        const int j0 = min(tid * c, n), j1 = min(j0 + c, n);
        uint32_t heads = 0;
        for (int i = j0; i < j1; ++i) heads += (i == 0 || src[i] != src[i - 1]) ? 1u : 0u;
        uint32_t nnz;
        uint32_t at = as_block_scan(heads, s_w, &nnz);
        for (int i = j0; i < j1; ++i)
            if (i == 0 || src[i] != src[i - 1]) dst[at++] = (uint32_t)i;
        if (w == 0) {                         // decoupled look-back over the tasks in front (chain word: flag << 62 | entries)
            unsigned long long excl = 0;
            if (task > 0) {
                if (lane == 0) st_vol(chain + task, (1ull << 62) | nnz);
                for (int j = task - 1;; j -= 32) {
                    const int idx = j - lane;
                    unsigned long long s = 2ull << 62;                   // in front of task 0: a full prefix of 0
                    if (idx >= 0) { do { s = ld_vol(chain + idx); } while ((s >> 62) == 0ull); }
                    __syncwarp();
                    const unsigned full = __ballot_sync(FULL, (s >> 62) == 2ull);
                    const int stop = full ? __ffs(full) - 1 : 31;        // nearest task that already knows its prefix
                    unsigned long long v = lane <= stop ? (s & AS_VAL) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                    excl += v;
                    if (full) break;
                }
            }
            if (lane == 0) {
                st_vol(chain + task, (2ull << 62) | (excl + nnz));
                s_base = excl;
                task_prefix[task] = (int64_t)excl;
                if (task == n_tasks - 1) *total_out = (int64_t)(excl + nnz);
            }
        }
        __syncthreads();
        const int64_t o0 = __ldg(t.out_base + task) + (int64_t)s_base;
        const uint64_t key0 = uint64_t(__ldg(t.ann + task)) * S + c_lo;
        for (uint32_t e = tid; e < nnz; e += AS_THREADS) {
            const uint32_t pos = dst[e], nxt = e + 1 < nnz ? dst[e + 1] : (uint32_t)n;
            const uint32_t rel = src[pos];
            const int64_t o = o0 + e;
            if (o < capacity) { keys_out[o] = key0 + rel; vals_out[o] = int64_t(nxt - pos); }
            if (totals) atomicAdd(totals + c_lo + rel, (unsigned long long)(nxt - pos));
        }
    }
}

// hist[r][code / bin_width] += 1 for every valid window of the sequences [seq_lo[r], seq_hi[r]): one CTA per row
__global__ void __launch_bounds__(AS_THREADS) ann_hist_kernel(const uint8_t *__restrict__ res, int64_t nres, const int64_t *__restrict__ off,
                                                              const uint8_t *__restrict__ lut, uint32_t nsym, int k, uint32_t pow_k1,
                                                              const int32_t *__restrict__ seq_lo, const int32_t *__restrict__ seq_hi,
                                                              uint32_t bin_width, int n_bins, uint32_t *__restrict__ hist) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_sym = smem;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + AS_SYM_BYTES);
    __shared__ uint8_t s_lut[256];
    __shared__ unsigned s_nstart;
    const int r = blockIdx.x;
    ts_lut_init(s_lut, lut);
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_hist[i] = 0u;
    if (threadIdx.x == 0) s_nstart = 0;
    __syncthreads();
    as_scan_seqs(res, nres, off, __ldg(seq_lo + r), __ldg(seq_hi + r), s_lut, s_sym, AS_SEG, k, nsym, pow_k1, &s_nstart,
                 [&](uint32_t code, bool ok) { if (ok) atomicAdd(&s_hist[code / bin_width], 1u); });
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) hist[int64_t(r) * n_bins + i] = s_hist[i];
}

static int as_check(const void *d_residues, int64_t nres, const void *d_offsets, int64_t nseq, const void *d_lut, int nsym, int k,
                    const char *who, uint64_t *S_out, uint32_t *pow_k1) {
    int rc = check_common(d_residues, nres, d_offsets, nseq, d_lut, nsym, k);
    if (rc) return rc;
    if (!ts_supported(nsym, k)) { set_error("%s: nsym=%d k=%d outside the kernel envelope", who, nsym, k); return SKM_ERR_UNSUPPORTED; }
    unsigned __int128 S128;
    code_space(nsym, k, &S128);
    if (S128 > (unsigned __int128)SKM_DENSE_MAX_SPACE) { set_error("%s: code space must be <= 2^27", who); return SKM_ERR_UNSUPPORTED; }
    *S_out = (uint64_t)S128;
    uint32_t p = 1;
    for (int i = 0; i + 1 < k; ++i) p *= (uint32_t)nsym;
    *pow_k1 = p;
    return SKM_OK;
}

}  // namespace skm

extern "C" {

int skm_ann_sort_cap(void) { return skm::AS_CAP; }

int skm_ann_hist(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq, const uint8_t *d_lut, int nsym, int k,
                 const int32_t *d_seq_lo, const int32_t *d_seq_hi, int64_t n_rows, uint32_t bin_width, int n_bins, uint32_t *d_hist,
                 skm_stream_t stream) {
    using namespace skm;
    uint64_t S;
    uint32_t pow_k1;
    int rc = as_check(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, "skm_ann_hist", &S, &pow_k1);
    if (rc) return rc;
    if (n_rows < 0 || n_rows > (1ll << 30) || n_bins < 1 || n_bins > 8192 || bin_width < 1 || uint64_t(bin_width) * uint64_t(n_bins) < S) {
        set_error("skm_ann_hist: need 1 <= n_bins <= 8192 and bin_width * n_bins >= nsym^k");
        return SKM_ERR_INVALID;
    }
    if (n_rows == 0) return SKM_OK;
    if (!d_seq_lo || !d_seq_hi || !d_hist) { set_error("skm_ann_hist: NULL argument"); return SKM_ERR_INVALID; }
    const int smem = AS_SYM_BYTES + 4 * n_bins;
    SKM_CUDA_TRY(cudaFuncSetAttribute(ann_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ann_hist_kernel<<<(unsigned)n_rows, AS_THREADS, smem, (cudaStream_t)stream>>>(d_residues, nres, d_offsets, d_lut, (uint32_t)nsym, k, pow_k1, d_seq_lo,
                                                                                 d_seq_hi, bin_width, n_bins, d_hist);
    SKM_LAUNCH_CHECK("ann_hist_kernel");
    return SKM_OK;
}

size_t skm_ann_sort_workspace(int64_t n_tasks) { return size_t(std::max<int64_t>(n_tasks, 0)) * 8 + 512; }

int skm_ann_sort(const uint8_t *d_residues, int64_t nres, const int64_t *d_offsets, int64_t nseq, const uint8_t *d_lut, int nsym, int k,
                 const int32_t *d_seq_lo, const int32_t *d_seq_hi, const uint32_t *d_code_lo, const uint32_t *d_code_hi, const int64_t *d_ann,
                 const int64_t *d_out_base, int64_t n_tasks, uint64_t *d_keys_out, int64_t *d_vals_out, int64_t capacity,
                 int64_t *d_task_prefix, int64_t *d_totals, int64_t *d_total_out, int *d_overflow, void *workspace, size_t workspace_bytes,
                 skm_stream_t stream) {
    using namespace skm;
    uint64_t S;
    uint32_t pow_k1;
    int rc = as_check(d_residues, nres, d_offsets, nseq, d_lut, nsym, k, "skm_ann_sort", &S, &pow_k1);
    if (rc) return rc;
    if (n_tasks < 0 || n_tasks >= (1ll << 31) || capacity < 0 || !d_total_out || !d_overflow) { set_error("skm_ann_sort: bad arguments"); return SKM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_total_out, 0, 8, st));
    SKM_CUDA_TRY(cudaMemsetAsync(d_overflow, 0, sizeof(int), st));
    if (n_tasks == 0) return SKM_OK;
    if (!d_seq_lo || !d_seq_hi || !d_code_lo || !d_code_hi || !d_ann || !d_out_base || !d_keys_out || !d_vals_out || !d_task_prefix) {
        set_error("skm_ann_sort: NULL argument");
        return SKM_ERR_INVALID;
    }
    const size_t need = skm_ann_sort_workspace(n_tasks);
    if (!workspace || workspace_bytes < need) { set_error("skm_ann_sort: workspace %zu < %zu", workspace_bytes, need); return SKM_ERR_WORKSPACE; }
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    unsigned *ticket = reinterpret_cast<unsigned *>(p);
    unsigned long long *chain = reinterpret_cast<unsigned long long *>(p + 256);
    SKM_CUDA_TRY(cudaMemsetAsync(p, 0, 256 + size_t(n_tasks) * 8, st));
    SKM_CUDA_TRY(cudaFuncSetAttribute(ann_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AS_SMEM));
    AsTasks t{d_seq_lo, d_seq_hi, d_code_lo, d_code_hi, d_ann, d_out_base};
    const int grid = (int)std::min<int64_t>(n_tasks, sm_count());           // persistent: one CTA per SM, tasks by ticket
    ann_sort_kernel<<<grid, AS_THREADS, AS_SMEM, st>>>(d_residues, nres, d_offsets, d_lut, (uint32_t)nsym, k, pow_k1, t, (int)n_tasks, S, ticket, chain,
                                                       d_keys_out, d_vals_out, capacity, d_task_prefix,
                                                       reinterpret_cast<unsigned long long *>(d_totals), d_total_out, d_overflow);
    SKM_LAUNCH_CHECK("ann_sort_kernel");
    return SKM_OK;
}

}  // extern "C"

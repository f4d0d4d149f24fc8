// skm_peer.cu — the multi-GPU fan-in of sparse learn over NVLink peer memory (one process per GPU).
//
// Replaces Merge.merge_dataframes across jobs (learn.smk:467-494) for COO count matrices sharded by sequence: rank r must
// end up with the entries of ITS annotation range from every rank.  Instead of packing, handing the list to NCCL's
// send/recv and unpacking, ONE kernel packs every entry (key << count_bits | count, 8 bytes) and stores it straight into
// the receive buffer of the rank that owns its annotation, through the NVSwitch (plain coalesced 8-byte stores to a
// peer-mapped pointer).  The receive buffers are cudaMalloc'ed by this library (CUDA IPC needs whole allocations) and
// opened by the peers through cudaIpcOpenMemHandle; the handles travel through the host-side process group.
#include <algorithm>
#include <cstring>

#include "skm_common.cuh"

namespace skm {

constexpr int PEER_MAX_WORLD = 16;

struct PushArgs {
    uint64_t *dst[PEER_MAX_WORLD];      // receive buffer of rank r (peer-mapped; own buffer for r == rank)
    int64_t dst_off[PEER_MAX_WORLD];    // where this rank's run starts in rank r's buffer (entries)
    int64_t cut[PEER_MAX_WORLD + 1];    // entries [cut[r], cut[r+1]) of the local sorted list belong to rank r
};

// blockIdx.y = destination rank, so a CTA streams one contiguous run to one peer.  The stores are 16 bytes per thread (two
// entries; a warp's store instruction is 512 contiguous bytes on the NVLink) once the destination is 16-byte aligned;
// the loads are local.
__global__ void __launch_bounds__(256) coo_pack_push_kernel(const uint64_t *__restrict__ keys, const int64_t *__restrict__ vals,
                                                            const PushArgs a, int world, int count_bits, int *__restrict__ overflow) {
    const uint64_t vmax = 1ull << count_bits, kmax = 1ull << (64 - count_bits);
    bool bad = false;
    const int r = blockIdx.y;
    const int64_t lo = a.cut[r], n = a.cut[r + 1] - lo;
    uint64_t *__restrict__ out = a.dst[r] + a.dst_off[r];
    const uint64_t *__restrict__ k = keys + lo;
    const int64_t *__restrict__ v = vals + lo;
    auto pack = [&](int64_t i) {
        const uint64_t kk = k[i], vv = uint64_t(v[i]);
        bad |= (vv >= vmax) | (kk >= kmax);
        return (kk << count_bits) | (vv & (vmax - 1));
    };
    const int64_t head = (n > 0 && (reinterpret_cast<uintptr_t>(out) & 8)) ? 1 : 0;      // entries in front of the first aligned pair
    const int64_t pairs = (n - head) >> 1;
    const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x, stride = int64_t(gridDim.x) * blockDim.x;
    if (tid == 0 && head) out[0] = pack(0);
    if (tid == 1 && ((n - head) & 1)) out[n - 1] = pack(n - 1);
    ulonglong2 *__restrict__ out2 = reinterpret_cast<ulonglong2 *>(out + head);
    for (int64_t p = tid; p < pairs; p += stride) {
        const int64_t i = head + 2 * p;
        ulonglong2 w;
        w.x = pack(i);
        w.y = pack(i + 1);
        out2[p] = w;
    }
    if (bad) atomicOr(overflow, 1);
}

}  // namespace skm

int skm_peer_alloc(size_t bytes, void **d_ptr, void *handle_out) {
    using namespace skm;
    if (!d_ptr || !handle_out || bytes == 0) { set_error("skm_peer_alloc: bad arguments"); return SKM_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == SKM_PEER_HANDLE_BYTES, "handle size");
    void *p = nullptr;
    SKM_CUDA_TRY(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(handle_out, &h, sizeof(h));
    *d_ptr = p;
    return SKM_OK;
}

int skm_peer_open(const void *handle, void **d_ptr) {
    using namespace skm;
    if (!handle || !d_ptr) { set_error("skm_peer_open: bad arguments"); return SKM_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    SKM_CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SKM_OK;
}

int skm_peer_close(void *d_ptr) {
    using namespace skm;
    if (!d_ptr) return SKM_OK;
    SKM_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return SKM_OK;
}

int skm_peer_free(void *d_ptr) {
    using namespace skm;
    if (!d_ptr) return SKM_OK;
    SKM_CUDA_TRY(cudaFree(d_ptr));
    return SKM_OK;
}

int skm_coo_pack_push(const uint64_t *d_keys, const int64_t *d_vals, const int64_t *cut_host, int world, int count_bits,
                      void *const *peer_bufs_host, const int64_t *dst_off_host, int *d_overflow, skm_stream_t stream) {
    using namespace skm;
    if (world < 1 || world > PEER_MAX_WORLD || !cut_host || !peer_bufs_host || !dst_off_host || !d_overflow || count_bits < 1 ||
        count_bits > 62) {
        set_error("skm_coo_pack_push: bad arguments (world 1..%d)", PEER_MAX_WORLD);
        return SKM_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SKM_CUDA_TRY(cudaMemsetAsync(d_overflow, 0, sizeof(int), st));
    PushArgs a{};
    int64_t longest = 0;
    for (int r = 0; r < world; ++r) {
        if (cut_host[r + 1] < cut_host[r] || dst_off_host[r] < 0) { set_error("skm_coo_pack_push: cuts must not decrease, offsets must be >= 0"); return SKM_ERR_INVALID; }
        if (cut_host[r + 1] > cut_host[r] && !peer_bufs_host[r]) { set_error("skm_coo_pack_push: NULL buffer for rank %d", r); return SKM_ERR_INVALID; }
        a.dst[r] = (uint64_t *)peer_bufs_host[r];
        a.dst_off[r] = dst_off_host[r];
        a.cut[r] = cut_host[r];
        longest = std::max(longest, cut_host[r + 1] - cut_host[r]);
    }
    a.cut[world] = cut_host[world];
    if (longest == 0) return SKM_OK;
    if (!d_keys || !d_vals) { set_error("skm_coo_pack_push: NULL argument"); return SKM_ERR_INVALID; }
    // enough CTAs in flight per destination to cover the NVLink round trip; the grid's y dimension is the destination
    const int per_dst = (int)std::min<int64_t>((longest / 2 + 255) / 256 + 1, std::max(1, sm_count() * 8 / world));
    coo_pack_push_kernel<<<dim3(per_dst, world), 256, 0, st>>>(d_keys, d_vals, a, world, count_bits, d_overflow);
    SKM_LAUNCH_CHECK("coo_pack_push_kernel");
    return SKM_OK;
}

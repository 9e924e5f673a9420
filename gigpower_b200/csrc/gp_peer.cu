// Peer memory for the path's one exchange: the all-gather of per-scenario
// voltage magnitudes and losses across the GPUs of one box, one process per GPU.
//
// A rank allocates its gather buffer here, sends the CUDA IPC handle to its peers (over whatever
// control plane the host uses - torch.distributed in the Python host) and maps theirs.  The solve
// kernels then store a scenario's |V| rows and loss straight into every rank's buffer (NVLink /
// NVSwitch posted writes issued from the kernel epilogue), or gp_peer_push moves finished rows with
// the copy engines.  Either way there is no collective kernel competing with the solve for SMs.
#include "gp_common.cuh"

using namespace gp;

static_assert(sizeof(cudaIpcMemHandle_t) == GP_PEER_HANDLE_BYTES, "GP_PEER_HANDLE_BYTES");

extern "C" int gp_peer_alloc(int device, int64_t bytes, void** ptr_out, void* handle_out) {
    if (bytes <= 0 || !ptr_out || !handle_out) return fail(GP_EINVAL, "gp_peer_alloc: bad argument");
    GP_CUDA(cudaSetDevice(device));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);      // plain cudaMalloc: pool / VMM memory has no legacy IPC handle
    if (e != cudaSuccess) return fail(GP_ENOMEM, "gp_peer_alloc: cudaMalloc(%lld) -> %s", (long long)bytes, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(GP_ECUDA, "gp_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return GP_OK;
}

extern "C" int gp_peer_free(int device, void* ptr) {
    if (!ptr) return GP_OK;
    GP_CUDA(cudaSetDevice(device));
    GP_CUDA(cudaFree(ptr));
    return GP_OK;
}

extern "C" int gp_peer_open(int device, const void* handle, void** ptr_out) {
    if (!handle || !ptr_out) return fail(GP_EINVAL, "gp_peer_open: bad argument");
    GP_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    GP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return GP_OK;
}

extern "C" int gp_peer_close(int device, void* ptr) {
    if (!ptr) return GP_OK;
    GP_CUDA(cudaSetDevice(device));
    GP_CUDA(cudaIpcCloseMemHandle(ptr));
    return GP_OK;
}

// gp_set_option("peer_push_streams", n): copy streams gp_peer_push spreads its destinations over (1..8, default 1).
// Measured on 8 x B200 (round 2, 123-bus FBS, 2.1 GB into every GPU per step; profiles/r02_scaling.jsonl): seven
// peer copies in a row on ONE stream deliver 373 GB/s per GPU and the step takes 5.62 ms against a 4.1 ms solve;
// dealt round-robin over 4 / 7 streams (several copy engines at once) the same bytes take 5.91 / 6.04 ms.  The
// all-to-all pattern, not the engine count, bounds the exchange on this box (NCCL and stores from the solve
// kernel's epilogue reach the same rate), so one stream stays the default.
int g_peer_push_streams = 1;

namespace {
struct PushLanes {
    cudaStream_t lane[8] = {};
    cudaEvent_t fork = nullptr, join[8] = {};
    bool ready = false;
};
PushLanes g_lanes[64];      // per device
}  // namespace

extern "C" int gp_peer_push(int device, int32_t n_dst, void* const* dst, const void* src_dev, int64_t bytes,
                            void* stream) {
    if (n_dst < 0 || (n_dst > 0 && !dst) || bytes < 0 || (bytes > 0 && !src_dev))
        return fail(GP_EINVAL, "gp_peer_push: bad argument");
    if (bytes == 0) return GP_OK;
    GP_CUDA(cudaSetDevice(device));
    for (int q = 0; q < n_dst; ++q)
        if (!dst[q]) return fail(GP_EINVAL, "gp_peer_push: null destination %d", q);
    cudaStream_t st = (cudaStream_t)stream;
    int lanes = g_peer_push_streams;
    if (lanes > n_dst) lanes = n_dst;
    if (lanes <= 1 || device < 0 || device >= 64) {
        for (int q = 0; q < n_dst; ++q) {
            if (dst[q] == src_dev) continue;
            GP_CUDA(cudaMemcpyAsync(dst[q], src_dev, (size_t)bytes, cudaMemcpyDefault, st));
        }
        return GP_OK;
    }
    PushLanes& L = g_lanes[device];
    if (!L.ready) {
        GP_CUDA(cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming));
        for (int i = 0; i < 8; ++i) {
            GP_CUDA(cudaStreamCreateWithFlags(&L.lane[i], cudaStreamNonBlocking));
            GP_CUDA(cudaEventCreateWithFlags(&L.join[i], cudaEventDisableTiming));
        }
        L.ready = true;
    }
    GP_CUDA(cudaEventRecord(L.fork, st));
    for (int i = 0; i < lanes; ++i) GP_CUDA(cudaStreamWaitEvent(L.lane[i], L.fork, 0));
    for (int q = 0; q < n_dst; ++q) {
        if (dst[q] == src_dev) continue;
        GP_CUDA(cudaMemcpyAsync(dst[q], src_dev, (size_t)bytes, cudaMemcpyDefault, L.lane[q % lanes]));
    }
    for (int i = 0; i < lanes; ++i) {
        GP_CUDA(cudaEventRecord(L.join[i], L.lane[i]));
        GP_CUDA(cudaStreamWaitEvent(st, L.join[i], 0));
    }
    return GP_OK;
}

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

extern "C" int gp_peer_push(int device, int32_t n_dst, void* const* dst, const void* src_dev, int64_t bytes,
                            void* stream) {
    if (n_dst < 0 || (n_dst > 0 && !dst) || bytes < 0 || (bytes > 0 && !src_dev))
        return fail(GP_EINVAL, "gp_peer_push: bad argument");
    if (bytes == 0) return GP_OK;
    GP_CUDA(cudaSetDevice(device));
    for (int q = 0; q < n_dst; ++q) {
        if (!dst[q]) return fail(GP_EINVAL, "gp_peer_push: null destination %d", q);
        if (dst[q] == src_dev) continue;
        GP_CUDA(cudaMemcpyAsync(dst[q], src_dev, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    }
    return GP_OK;
}

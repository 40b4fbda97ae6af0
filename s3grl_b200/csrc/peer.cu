// Peer memory for the fused gather + all-gather (s3_gather_peers): one cudaMalloc'ed buffer per
// GPU, shared with the other ranks of the node through CUDA runtime IPC handles. Opening a handle
// maps the peer's buffer into this process and enables peer access lazily, so kernel 3 can store
// its output rows straight into every GPU's operator matrices over NVLink (the
// one exchange step of the path; the reference has no distributed code).
#include "common.cuh"

namespace s3 {

cudaError_t peer_alloc(int64_t bytes, void** ptr) { return cudaMalloc(ptr, (size_t)bytes); }
cudaError_t peer_free(void* ptr) { return cudaFree(ptr); }

cudaError_t peer_export(void* ptr, unsigned char* handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == S3_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) return e;
    memcpy(handle, &h, sizeof(h));
    return cudaSuccess;
}

cudaError_t peer_open(const unsigned char* handle, void** ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

cudaError_t peer_close(void* ptr) { return cudaIpcCloseMemHandle(ptr); }

}  // namespace s3

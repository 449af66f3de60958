// Peer-memory plumbing for the multi-GPU detection gather (the reference itself never shards
// evaluation, scripts/val.py:170-172).  One process per GPU; the gather root allocates the gathered detection buffers
// with dsn_peer_alloc, exports them as CUDA IPC handles and every other rank maps them (dsn_peer_open).  The NMS kernel
// of a rank is then simply given `det` / `det_count` pointers INTO the root's buffer: its result rows travel over
// NVLink / NVSwitch as the kernel produces them (24 B per detection + 4 B per image instead of a fixed-capacity
// all-gather), and a one-thread signal kernel publishes "step s of rank r has landed" with a system-scope release
// store that the root's wait kernel acquires.  No NCCL call on the per-batch path.
#include <cstring>

#include "common.cuh"

namespace dsn {

// *local_step += 1, then publish the new value to the (remote) flag.  Everything earlier kernels of this stream wrote
// to peer memory happens-before this kernel (stream order); the system-scope fence makes those writes visible to the
// flag's observer before the flag itself (fence cumulativity).
__global__ void peer_signal_kernel(unsigned int *local_step, unsigned int *remote_flag) {
    const unsigned int v = *local_step + 1u;
    *local_step = v;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote_flag), "r"(v) : "memory");
}

// Spin until flags[i] - targets[i] >= 0 for every i < n (acquire loads at system scope), at most timeout_ns; on a
// timeout *status is set to 1 + the index of a flag that never arrived (the caller raises), otherwise left untouched.
__global__ void peer_wait_kernel(const unsigned int *flags, const unsigned int *targets, int n, unsigned long long timeout_ns, int *status) {
    const int i = threadIdx.x;
    if (i >= n) return;
    const unsigned int target = targets[i];
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        if ((int)(v - target) >= 0) return;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            atomicMax(status, i + 1);
            return;
        }
        __nanosleep(200);
    }
}

}  // namespace dsn

using namespace dsn;

extern "C" __attribute__((visibility("default"))) int dsn_peer_alloc(int64_t bytes, void **ptr) {
    if (!ptr || bytes <= 0) return DSN_E_INVALID;
    DSN_CUDA(cudaMalloc(ptr, (size_t)bytes));
    DSN_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_peer_free(void *ptr) {
    if (!ptr) return DSN_OK;
    DSN_CUDA(cudaFree(ptr));
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_peer_export(const void *ptr, unsigned char *handle64) {
    if (!ptr || !handle64) return DSN_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    DSN_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)));
    memcpy(handle64, &h, 64);
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_peer_open(const unsigned char *handle64, void **ptr) {
    if (!ptr || !handle64) return DSN_E_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    DSN_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_peer_close(void *ptr) {
    if (!ptr) return DSN_OK;
    DSN_CUDA(cudaIpcCloseMemHandle(ptr));
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_peer_signal(unsigned int *local_step, unsigned int *remote_flag, void *stream) {
    if (!local_step || !remote_flag) return DSN_E_INVALID;
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(local_step, remote_flag);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

extern "C" __attribute__((visibility("default"))) int dsn_peer_wait(const unsigned int *flags, const unsigned int *targets, int n, int64_t timeout_ms,
                                                                    int *status, void *stream) {
    if (!flags || !targets || !status || n < 1 || n > 1024 || timeout_ms < 0) return DSN_E_INVALID;
    peer_wait_kernel<<<1, ((n + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(flags, targets, n, (unsigned long long)timeout_ms * 1000000ull, status);
    DSN_CHECK_LAUNCH();
    return DSN_OK;
}

// Sum of a few dozen doubles across the GPUs of one box, every step, written for LATENCY: the (C, 18) accumulators that
// the global sites' reverse mode needs are 144 bytes, and an NCCL all-reduce of them inside the captured
// step cost 120-160 us at 8 GPUs (measured: 0.188 ms/step without it, 0.31-0.35 ms with it).  Here every rank PUSHES its
// values straight into a slot of every peer's buffer over NVLink (peer pointers from CUDA IPC), then a flag; each rank
// then only polls its OWN memory and adds the slots in rank order -- the same order everywhere, so all ranks get
// bit-identical sums.  Two small kernels in the step's graph, no host involvement, no NCCL on the data path.
//
// Buffer of one rank (tq_p2p_bytes()): sequence number, flags[2][kMaxRanks], slots[2][kMaxRanks][kMaxValues]; double
// buffered by call parity (a rank can be at most one call ahead of a peer: it cannot finish call s+1 without that peer's
// push of s+1, which the peer issues only after consuming call s).
#include <stddef.h>
#include <string.h>

#include "common.cuh"

namespace tq {

constexpr int kMaxRanks = 16;
constexpr int kMaxValues = 128;

struct P2PBuffer {
    unsigned long long seq;                       // calls completed by THIS rank's push kernel
    unsigned long long timeout;                   // set when a wait gave up (a peer died): results are garbage, no hang
    unsigned long long flags[2][kMaxRanks];       // flags[par][q] = sequence number of rank q's latest push into par
    double slots[2][kMaxRanks][kMaxValues];
};

__global__ void p2p_push_kernel(const double* __restrict__ values, int n, int rank, int world, P2PBuffer* const* __restrict__ peers) {
    __shared__ unsigned long long s_seq;
    P2PBuffer* own = peers[rank];
    if (threadIdx.x == 0) s_seq = own->seq + 1ull;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ull);
    for (int t = threadIdx.x; t < n * world; t += blockDim.x) {
        const int p = t / n, i = t - p * n;
        peers[p]->slots[par][rank][i] = values[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        volatile unsigned long long* flag = &peers[threadIdx.x]->flags[par][rank];
        *flag = seq;
    }
    if (threadIdx.x == 0) own->seq = seq;
}

__global__ void p2p_wait_sum_kernel(P2PBuffer* __restrict__ own, int n, int world, double* __restrict__ out) {
    const unsigned long long seq = own->seq;      // written by this rank's push kernel, earlier on the same stream
    const int par = (int)(seq & 1ull);
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < world) {
        volatile unsigned long long* flag = &own->flags[par][threadIdx.x];
        long long spins = 0;
        while (*flag < seq) {
            if (++spins > (1ll << 27)) { ok = 0; break; }    // ~ seconds: a peer is gone; do not hang the GPU
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (!ok && threadIdx.x == 0) own->timeout = seq;
    __threadfence_system();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < world; ++q) s += ((volatile double*)own->slots[par][q])[i];
        out[i] = s;
    }
}

}  // namespace tq

using namespace tq;

extern "C" int64_t tq_p2p_bytes(void) { return (int64_t)sizeof(P2PBuffer); }
extern "C" int tq_p2p_max_values(void) { return kMaxValues; }
extern "C" int tq_p2p_max_ranks(void) { return kMaxRanks; }

// cudaMalloc'ed (not from a caching allocator: IPC handles name whole allocations), zeroed buffer + its IPC handle (64 bytes)
extern "C" int tq_p2p_alloc(void** buffer, unsigned char* handle64) {
    TQ_CHECK_ARG(buffer && handle64, "NULL pointer");
    int st = cuda_status(cudaMalloc(buffer, sizeof(P2PBuffer)), "cudaMalloc(P2PBuffer)");
    if (st != TQ_OK) return st;
    st = cuda_status(cudaMemset(*buffer, 0, sizeof(P2PBuffer)), "cudaMemset(P2PBuffer)");
    if (st != TQ_OK) return st;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    st = cuda_status(cudaIpcGetMemHandle(&h, *buffer), "cudaIpcGetMemHandle");
    if (st != TQ_OK) return st;
    memcpy(handle64, &h, 64);
    return cuda_status(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
}

extern "C" int tq_p2p_open(const unsigned char* handle64, void** buffer) {
    TQ_CHECK_ARG(buffer && handle64, "NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return cuda_status(cudaIpcOpenMemHandle(buffer, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

extern "C" int tq_p2p_close(void* buffer) { return cuda_status(cudaIpcCloseMemHandle(buffer), "cudaIpcCloseMemHandle"); }
extern "C" int tq_p2p_free(void* buffer) { return cuda_status(cudaFree(buffer), "cudaFree"); }

// peers: DEVICE array of `world` buffer pointers (own buffer at index `rank`)
extern "C" int tq_p2p_push(const double* values, int n, int rank, int world, const void* peers, void* stream) {
    TQ_CHECK_ARG(values && peers, "NULL pointer");
    TQ_CHECK_ARG(n >= 1 && n <= kMaxValues && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad sizes");
    p2p_push_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(values, n, rank, world, (P2PBuffer* const*)peers);
    TQ_LAUNCH_CHECK("p2p_push_kernel launch");
    return TQ_OK;
}

extern "C" int tq_p2p_wait_sum(void* own_buffer, int n, int world, double* out, void* stream) {
    TQ_CHECK_ARG(own_buffer && out, "NULL pointer");
    TQ_CHECK_ARG(n >= 1 && n <= kMaxValues && world >= 1 && world <= kMaxRanks, "bad sizes");
    p2p_wait_sum_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((P2PBuffer*)own_buffer, n, world, out);
    TQ_LAUNCH_CHECK("p2p_wait_sum_kernel launch");
    return TQ_OK;
}

// host read of the timeout marker (0 = never timed out)
extern "C" int tq_p2p_timed_out(const void* own_buffer, unsigned long long* seq) {
    TQ_CHECK_ARG(own_buffer && seq, "NULL pointer");
    return cuda_status(cudaMemcpy(seq, (const char*)own_buffer + offsetof(P2PBuffer, timeout), sizeof(unsigned long long), cudaMemcpyDeviceToHost),
                       "cudaMemcpy(timeout)");
}

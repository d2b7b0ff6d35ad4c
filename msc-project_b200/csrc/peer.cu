// peer.cu — the exchange steps of the multi-GPU engine, over NVLink/NVSwitch peer memory.
//
// Replaces the per-step messages of the reference's master/worker protocol
// (/root/reference/src/master.cpp:268-292, src/worker.cpp:109-142): MPI_Bcast(MPI_Crds) (master.cpp:279)
// becomes a pull of the few non-owned tetrad blocks a rank actually reads, straight out of the owners'
// memory; the dense MPI_Reduce of N x 758 doubles (master.cpp:288) became one 64-bit block mask per
// flagged pair stored into every rank's buffer by the distance kernel itself (nb.cu: publish_hit); the
// TAG_FORCE "go" messages (master.cpp:276-278) become a flag barrier.  Every GPU maps the others' buffers
// (CUDA IPC between processes, plain peer access inside one process), so the kernels below are ordinary
// loads and stores; they are stream-ordered with the compute kernels and captured in the same CUDA graph.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// All-to-all flag barrier: thread r tells rank r "I have arrived at barrier number e" by a release store
// into slot [my rank] of rank r's flag array, then waits until slot [r] of its own array shows that rank
// r has arrived too.  The barrier number lives in device memory and is advanced by the kernel itself, so
// a captured graph replays correctly; every rank runs the same sequence of barriers.  What was written
// before the barrier (by earlier kernels of the stream, into local or peer memory) is visible to the
// kernels the other ranks launch after it.  A rank that never arrives (a crashed peer) would hang the
// GPU: the wait gives up after `timeout_ns` and raises the error flag the host checks at its next sync.
struct BarrierArgs {
    unsigned long long* const* flags;   // [world] flags[r] -> rank r's array of `world` arrival counters
    unsigned long long* epoch;          // this rank's barrier counter
    int* err;
    int world, rank;
    unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(32) peer_barrier_kernel(BarrierArgs A) {
    const int r = threadIdx.x;
    const unsigned long long e = *A.epoch + 1;
    __syncwarp();
    if (r < A.world) {
        __threadfence_system();
        st_release_sys(A.flags[r] + A.rank, e);
        const unsigned long long* mine = A.flags[A.rank] + r;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(mine) < e) {
            if (global_ns() - t0 > A.timeout_ns) { atomicExch(A.err, 1); break; }
        }
    }
    __syncwarp();
    if (r == 0) *A.epoch = e;
}

cudaError_t launch_peer_barrier(unsigned long long* const* flags, unsigned long long* epoch, int* err, int world,
                                int rank, unsigned long long timeout_ns, cudaStream_t st) {
    BarrierArgs A{flags, epoch, err, world, rank, timeout_ns};
    peer_barrier_kernel<<<1, 32, 0, st>>>(A);
    return cudaGetLastError();
}

// Halo pull: copy the listed tetrad blocks (3*S doubles each) from their owners' coordinate buffers into
// this rank's buffer at the same index.  Owners are contiguous blocks of `chunk` tetrads.  The loads bypass
// the caches of this GPU (ld.cv): the lines belong to another GPU's memory and change every step.
struct PullArgs {
    double* local;
    double* const* peer;   // [world] base pointers of the same buffer on every rank
    const int* list;
    int count, chunk, words;   // words = 3*S
};

__global__ void __launch_bounds__(128) halo_pull_kernel(PullArgs A) {
    for (int i = blockIdx.x; i < A.count; i += gridDim.x) {
        const int t = A.list[i];
        const double2* src = reinterpret_cast<const double2*>(A.peer[t / A.chunk] + (size_t)t * A.words);
        double2* dst = reinterpret_cast<double2*>(A.local + (size_t)t * A.words);
        for (int k = threadIdx.x; k < A.words / 2; k += blockDim.x) dst[k] = __ldcv(src + k);
    }
}

cudaError_t launch_halo_pull(double* local, double* const* peer, const int* list, int count, int chunk, int S, int sms,
                             cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    PullArgs A{local, peer, list, count, chunk, 3 * S};
    const int grid = count < 8 * sms ? count : 8 * sms;
    halo_pull_kernel<<<grid, 128, 0, st>>>(A);
    return cudaGetLastError();
}

// Which non-owned tetrads this rank reads during a step: both ends of the pairs in its slice of the
// distance pass, and the partners of its owned tetrads in the force pass.  flag[t] = 1, then a compaction.
__global__ void halo_flag_kernel(const int2* pairs, long long np, long long p0, long long p1, int t0, int t1,
                                 unsigned char* flag) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        const int2 pr = pairs[p];
        const bool mine = p >= p0 && p < p1;
        const bool ox = pr.x >= t0 && pr.x < t1, oy = pr.y >= t0 && pr.y < t1;
        if (mine || oy) { if (!ox) flag[pr.x] = 1; }
        if (mine || ox) { if (!oy) flag[pr.y] = 1; }
    }
}
__global__ void halo_compact_kernel(const unsigned char* flag, int n, int* list, int* count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && flag[t]) list[atomicAdd(count, 1)] = t;
}

cudaError_t build_halo_list(const int2* pairs, long long np, long long p0, long long p1, int t0, int t1, int n,
                            unsigned char* flag, int* list, int* count, cudaStream_t st) {
    cudaError_t err = cudaMemsetAsync(flag, 0, (size_t)n, st);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(count, 0, sizeof(int), st);
    if (err != cudaSuccess) return err;
    if (np > 0) {
        long long blocks = (np + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        halo_flag_kernel<<<(int)blocks, 256, 0, st>>>(pairs, np, p0, p1, t0, t1, flag);
    }
    halo_compact_kernel<<<(n + 255) / 256, 256, 0, st>>>(flag, n, list, count);
    return cudaGetLastError();
}

}  // namespace edmd

// peer.cu -- peer-GPU memory over NVLink for the two exchange steps of the path: the sharded
// similarity matrix (every rank's tensor-core kernel streams the other ranks' video-embedding operand images straight out
// of their HBM, cmhar_similarity_img) and the scalar loss reduction.  One process per GPU: a buffer allocated here is
// exported as a CUDA IPC handle, the other ranks map it (cmhar_peer_open) and hand the mapped pointers to the kernels.
//
// cmhar_peer_barrier is the only cross-GPU synchronisation: rank r release-stores a monotonically increasing epoch into slot r
// of every rank's flag block and acquire-spins on its own block until all slots carry that epoch.  The epoch counter lives in
// device memory, so the launch can be captured in a CUDA graph and replayed.  The kernel waits on kernels of OTHER GPUs
// only (one rank per GPU, like NCCL); a rank that never arrives trips the watchdog
// (%globaltimer, 20 s by default, CMHAR_PEER_TIMEOUT_MS) and traps instead of hanging the box.
#include "common.cuh"

namespace cmhar {
namespace peer {

struct BarrierArgs {
    unsigned long long* flags[CMHAR_MAX_PEERS];      // flags[p] = rank p's flag block (CMHAR_MAX_PEERS words), peer-mapped
    int rank, world;
    unsigned long long* epoch;                       // local counter
    const double* slots;                             // optional: after the barrier, sum_out = scale * sum(slots[0..n_slots))
    int n_slots;
    double scale;
    double* sum_out;
    unsigned long long timeout_ns;                   // watchdog: a rank that never arrives traps this kernel instead of hanging the box
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void peer_barrier_kernel(const BarrierArgs a) {
    const int t = threadIdx.x;
    const unsigned long long e = *a.epoch + 1ull;
    __syncwarp();
    if (t < a.world) {
        st_release_sys(a.flags[t] + a.rank, e);                     // tell rank t that this rank has arrived
        const unsigned long long t0 = global_timer_ns();
        unsigned int spins = 0;
        while (ld_acquire_sys(a.flags[a.rank] + t) < e) {           // wait for rank t's arrival in the local block
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > a.timeout_ns) {
                printf("cmhar peer barrier: rank %d timed out after %llu ms waiting for rank %d (epoch %llu, saw %llu)\n", a.rank,
                       a.timeout_ns / 1000000ull, t, e, ld_acquire_sys(a.flags[a.rank] + t));
                __trap();
            }
        }
    }
    __syncwarp();
    if (t == 0) {
        *a.epoch = e;
        if (a.sum_out) {
            double s = 0.0;
            for (int i = 0; i < a.n_slots; ++i) s += *reinterpret_cast<const volatile double*>(a.slots + i);    // fixed order: identical on every rank
            *a.sum_out = s * a.scale;
        }
    }
}

}  // namespace peer
}  // namespace cmhar

using namespace cmhar;

extern "C" {

int cmhar_peer_alloc(size_t bytes, void** ptr_out) {
    CMHAR_REQUIRE(ptr_out && bytes > 0, "cmhar_peer_alloc: bad argument");
    // Operand images inside the buffer must start on a 1 KiB boundary on EVERY rank that maps it.  cudaMalloc promises 256 bytes only
    // (a 1 MiB + 1 KiB request came back 512-byte aligned once other allocations had moved the heap: the 2-GPU bench failed in
    // cmhar_mlp2_forward_img); requests of whole 2 MiB pages are page aligned, locally and through cudaIpcOpenMemHandle.  Checked, not assumed.
    const size_t page = (size_t)2 << 20;
    const size_t rounded = (bytes + page - 1) / page * page;
    void* p = nullptr;
    CMHAR_CHECK_CUDA(cudaMalloc(&p, rounded));
    if (((uintptr_t)p & 1023) != 0) {
        cudaFree(p);
        CMHAR_REQUIRE(false, "cmhar_peer_alloc: cudaMalloc(%zu) returned a pointer that is not 1 KiB aligned", rounded);
    }
    CMHAR_CHECK_CUDA(cudaMemset(p, 0, rounded));
    *ptr_out = p;
    return CMHAR_OK;
}

int cmhar_peer_free(void* ptr) {
    if (ptr) CMHAR_CHECK_CUDA(cudaFree(ptr));
    return CMHAR_OK;
}

int cmhar_peer_export(const void* ptr, void* handle_out) {
    CMHAR_REQUIRE(ptr && handle_out, "cmhar_peer_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == CMHAR_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CMHAR_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle_out, &h, sizeof(h));
    return CMHAR_OK;
}

int cmhar_peer_open(const void* handle, void** ptr_out) {
    CMHAR_REQUIRE(handle && ptr_out, "cmhar_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    CMHAR_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    if (((uintptr_t)p & 1023) != 0) {
        cudaIpcCloseMemHandle(p);
        CMHAR_REQUIRE(false, "cmhar_peer_open: the mapping of the peer buffer is not 1 KiB aligned");
    }
    *ptr_out = p;
    return CMHAR_OK;
}

int cmhar_peer_close(void* ptr) {
    if (ptr) CMHAR_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return CMHAR_OK;
}

int cmhar_peer_barrier(void* const* flag_blocks, int32_t rank, int32_t world, void* local_epoch, const double* slots,
                       int32_t n_slots, double scale, double* sum_out, cmhar_stream_t s) {
    CMHAR_REQUIRE(flag_blocks && local_epoch && world >= 1 && world <= CMHAR_MAX_PEERS && rank >= 0 && rank < world,
                  "cmhar_peer_barrier: bad argument (world %d, rank %d)", world, rank);
    CMHAR_REQUIRE(!sum_out || (slots && n_slots > 0), "cmhar_peer_barrier: sum_out needs slots");
    peer::BarrierArgs a{};
    for (int i = 0; i < world; ++i) {
        CMHAR_REQUIRE(flag_blocks[i], "cmhar_peer_barrier: null flag block %d", i);
        a.flags[i] = reinterpret_cast<unsigned long long*>(flag_blocks[i]);
    }
    a.rank = rank; a.world = world; a.epoch = reinterpret_cast<unsigned long long*>(local_epoch);
    a.slots = slots; a.n_slots = n_slots; a.scale = scale; a.sum_out = sum_out;
    static unsigned long long timeout_ms = 0;
    if (timeout_ms == 0) {
        const char* e = getenv("CMHAR_PEER_TIMEOUT_MS");        // default 20 s: ranks may be seconds apart (first-use compilation, host jitter)
        timeout_ms = (e && atoll(e) > 0) ? (unsigned long long)atoll(e) : 20000ull;
    }
    a.timeout_ns = timeout_ms * 1000000ull;
    peer::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)s>>>(a);
    CMHAR_LAUNCH_CHECK();
    return CMHAR_OK;
}

}  // extern "C"

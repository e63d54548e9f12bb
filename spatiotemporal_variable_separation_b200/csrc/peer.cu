// Gradient all-reduce over NVLink / NVSwitch peer memory (one process per GPU, one node).
//
// The gradient arena of every rank is mapped into every other rank's address space (CUDA IPC; parallel.py), so the exchange
// step of data-parallel training is ONE kernel per rank instead of a library collective:
// rank r owns the r-th slice of the arena, reads that slice from all `world` arenas (world-1 of them over NVLink),
// adds them in a FIXED rank order (every rank obtains bit-identical sums: replicas never drift), and writes the sum back
// into all `world` arenas — a reduce-scatter and an all-gather fused into one pass, 2 * (world-1)/world * bytes per
// direction and GPU, no staging buffers, no SM left idle waiting for a ring neighbour.  Two tiny barrier kernels (flags in
// peer memory, monotonically increasing epoch kept on the device so that a captured CUDA graph replays correctly)
// bracket it: before = every rank has finished its backward pass, after = every rank's slice has landed everywhere.
#include "common.cuh"

namespace vs {

constexpr int PEER_MAX_WORLD = 16;
struct PeerPtrs { float* p[PEER_MAX_WORLD]; };
struct PeerFlags { unsigned int* f[PEER_MAX_WORLD]; };

__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;      // volatile: never served from a stale L1 line of a previous step
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_peer(float* p, const float4& v) {
    asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flags[r] of rank q = the last epoch rank r has announced to rank q.  One block; thread r talks to rank r.
__global__ void peer_barrier_kernel(PeerFlags flags, int rank, int world, unsigned int* __restrict__ epoch_dev) {
    __shared__ unsigned int epoch;
    if (threadIdx.x == 0) epoch = *epoch_dev + 1;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int r = threadIdx.x;
        __threadfence_system();                                    // everything this GPU wrote before is visible system-wide
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.f[r] + rank), "r"(epoch) : "memory");
        unsigned int seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags.f[rank] + r) : "memory");
        } while ((int)(seen - epoch) < 0);
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch_dev = epoch;
}

template <int WORLD>
__global__ void __launch_bounds__(512) peer_allreduce_kernel(PeerPtrs ptrs, int rank, long long n4_per_rank, long long n4_total) {
    const long long begin = (long long)rank * n4_per_rank;
    long long end = begin + n4_per_rank;
    if (end > n4_total) end = n4_total;
    for (long long i = begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.x * blockDim.x) {
        float4 v[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; ++r) v[r] = ld_peer(ptrs.p[r] + 4 * i);          // all loads in flight before the first add
        float4 s = v[0];
#pragma unroll
        for (int r = 1; r < WORLD; ++r) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
#pragma unroll
        for (int r = 0; r < WORLD; ++r) st_peer(ptrs.p[r] + 4 * i, s);
    }
}

}  // namespace vs

using namespace vs;

extern "C" int vs_peer_barrier(void* const* flag_ptrs_host, int32_t rank, int32_t world, uint32_t* epoch_dev, void* stream) {
    VS_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "peer_barrier: bad rank %d / world %d", rank, world);
    PeerFlags f;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) f.f[r] = r < world ? reinterpret_cast<unsigned int*>(flag_ptrs_host[r]) : nullptr;
    peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(f, rank, world, epoch_dev);
    return launched("peer_barrier_kernel");
}

extern "C" int vs_peer_allreduce(void* const* arena_ptrs_host, int32_t rank, int32_t world, int64_t n, int32_t max_blocks,
                                 void* stream) {
    VS_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "peer_allreduce: bad rank %d / world %d", rank, world);
    VS_REQUIRE(n % 4 == 0, "peer_allreduce: element count must be a multiple of 4 (got %lld)", (long long)n);
    PeerPtrs p;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        p.p[r] = r < world ? reinterpret_cast<float*>(arena_ptrs_host[r]) : nullptr;
        VS_REQUIRE(r >= world || (reinterpret_cast<uintptr_t>(p.p[r]) & 15) == 0, "peer_allreduce: arenas must be 16-byte aligned");
    }
    if (n == 0 || world == 1) return 0;
    const long long n4 = n / 4, per = cdiv(n4, world);
    long long blocks = cdiv(per, 512);
    const long long cap = max_blocks > 0 ? max_blocks : 2LL * num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
#define VS_PEER_CASE(W) case W: peer_allreduce_kernel<W><<<(unsigned)blocks, 512, 0, as_stream(stream)>>>(p, rank, per, n4); break;
    switch (world) {
        VS_PEER_CASE(2) VS_PEER_CASE(3) VS_PEER_CASE(4) VS_PEER_CASE(5) VS_PEER_CASE(6) VS_PEER_CASE(7) VS_PEER_CASE(8)
        default: return fail("peer_allreduce: world sizes 2..8 are instantiated (got %d)", world);
    }
#undef VS_PEER_CASE
    return launched("peer_allreduce_kernel");
}

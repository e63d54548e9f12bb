// Gradient all-reduce over NVLink / NVSwitch peer memory (one process per GPU, one node).
//
// The gradient arena of every rank is mapped into every other rank's address space (CUDA IPC; parallel.py), so the exchange
// step of data-parallel training is ONE kernel per rank instead of a library collective:
// rank r owns the r-th slice of the arena, reads that slice from all `world` arenas (world-1 of them over NVLink),
// adds them (every element is summed by one rank only, so all ranks obtain bit-identical sums: replicas never drift), and writes the sum back
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
    float4 v;      // .cg: not through the (incoherent) local L1, which could serve a line of the previous step
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_peer(float* p, const float4& v) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
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
    // the peers are visited starting from this rank's right neighbour, so that at any moment the `world` ranks pull from
    // (and push to) `world` different GPUs instead of all hammering rank 0 first.  The sum is taken in visiting order:
    // every element is summed by exactly one rank (its slice's owner) and broadcast, so all ranks still hold identical
    // bits, and the order is a fixed function of the owner (run-to-run deterministic)
    constexpr int U = WORLD <= 4 ? 2 : 1;                   // float4s per thread and iteration
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < end; i0 += U * stride) {
        float4 v[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < WORLD; ++k) {
                const int r = (rank + 1 + k) % WORLD;
                const long long i = i0 + u * stride;
                v[u][k] = i < end ? ld_peer(ptrs.p[r] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);   // all loads in flight
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i >= end) break;
            float4 s = v[u][0];
#pragma unroll
            for (int r = 1; r < WORLD; ++r) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
#pragma unroll
            for (int k = 0; k < WORLD; ++k) st_peer(ptrs.p[(rank + 1 + k) % WORLD] + 4 * i, s);
        }
    }
}

// bf16 transport: every rank first rounds its own gradients to bf16 (vs_grad_compress), the owner of a slice sums the
// `world` bf16 copies in fp32 and writes the bf16-rounded sum to everybody, and every rank widens its buffer back into the
// fp32 arena (vs_grad_expand).  Half the NVLink bytes of the fp32 exchange; all ranks still hold identical bits.
struct PeerPtrs16 { __nv_bfloat16* p[PEER_MAX_WORLD]; };

__device__ __forceinline__ uint4 ld_peer16(const __nv_bfloat16* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_peer16(__nv_bfloat16* p, const uint4& v) {
    asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int WORLD>
__global__ void __launch_bounds__(512) peer_allreduce_bf16_kernel(PeerPtrs16 ptrs, int rank, long long n8_per_rank, long long n8_total) {
    const long long begin = (long long)rank * n8_per_rank;
    long long end = begin + n8_per_rank;
    if (end > n8_total) end = n8_total;
    for (long long i = begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.x * blockDim.x) {
        uint4 v[WORLD];
#pragma unroll
        for (int k = 0; k < WORLD; ++k) v[k] = ld_peer16(ptrs.p[(rank + 1 + k) % WORLD] + 8 * i);
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < WORLD; ++k) {
            const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { s[2 * e] += __uint_as_float(w[e] << 16); s[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u); }
        }
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(s[2 * e], s[2 * e + 1]);
            ow[e] = *reinterpret_cast<uint32_t*>(&b2);
        }
#pragma unroll
        for (int k = 0; k < WORLD; ++k) st_peer16(ptrs.p[(rank + 1 + k) % WORLD] + 8 * i, o);
    }
}

__global__ void grad_compress_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ out, long long n8) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(g)[2 * i], b = reinterpret_cast<const float4*>(g)[2 * i + 1];
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        reinterpret_cast<uint4*>(out)[i] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                       *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
    }
}
__global__ void grad_expand_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ g, long long n8) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4*>(in)[i];
        reinterpret_cast<float4*>(g)[2 * i] = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                                                          __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
        reinterpret_cast<float4*>(g)[2 * i + 1] = make_float4(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u),
                                                              __uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u));
    }
}

}  // namespace vs

using namespace vs;

extern "C" int vs_grad_compress(const float* grad, void* out_bf16, int64_t n, void* stream) {
    VS_REQUIRE(n % 8 == 0 && ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0,
               "grad_compress: n %% 8 == 0 and 16-byte aligned buffers");
    if (n == 0) return 0;
    long long blocks = cdiv(n / 8, 256);
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    grad_compress_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(grad, reinterpret_cast<__nv_bfloat16*>(out_bf16), n / 8);
    return launched("grad_compress_kernel");
}

extern "C" int vs_grad_expand(const void* in_bf16, float* grad, int64_t n, void* stream) {
    VS_REQUIRE(n % 8 == 0 && ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(in_bf16)) & 15) == 0,
               "grad_expand: n %% 8 == 0 and 16-byte aligned buffers");
    if (n == 0) return 0;
    long long blocks = cdiv(n / 8, 256);
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    grad_expand_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in_bf16), grad, n / 8);
    return launched("grad_expand_kernel");
}

extern "C" int vs_peer_allreduce_bf16(void* const* buf_ptrs_host, int32_t rank, int32_t world, int64_t n, int32_t max_blocks,
                                      void* stream) {
    VS_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "peer_allreduce_bf16: bad rank %d / world %d", rank, world);
    VS_REQUIRE(n % 8 == 0, "peer_allreduce_bf16: element count must be a multiple of 8 (got %lld)", (long long)n);
    PeerPtrs16 p;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        p.p[r] = r < world ? reinterpret_cast<__nv_bfloat16*>(buf_ptrs_host[r]) : nullptr;
        VS_REQUIRE(r >= world || (reinterpret_cast<uintptr_t>(p.p[r]) & 15) == 0, "peer_allreduce_bf16: buffers must be 16-byte aligned");
    }
    if (n == 0 || world == 1) return 0;
    const long long n8 = n / 8, per = cdiv(n8, world);
    long long blocks = cdiv(per, 512);
    const long long cap = max_blocks > 0 ? max_blocks : 2LL * num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
#define VS_PEER16_CASE(W) case W: peer_allreduce_bf16_kernel<W><<<(unsigned)blocks, 512, 0, as_stream(stream)>>>(p, rank, per, n8); break;
    switch (world) {
        VS_PEER16_CASE(2) VS_PEER16_CASE(3) VS_PEER16_CASE(4) VS_PEER16_CASE(5) VS_PEER16_CASE(6) VS_PEER16_CASE(7) VS_PEER16_CASE(8)
        default: return fail("peer_allreduce_bf16: world sizes 2..8 are instantiated (got %d)", world);
    }
#undef VS_PEER16_CASE
    return launched("peer_allreduce_bf16_kernel");
}

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

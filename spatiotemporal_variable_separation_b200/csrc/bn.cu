// Grouped training-mode BatchNorm + activation, NHWC.  HBM-bound streaming kernels.
// Replaces aten::native_batch_norm(_backward) + the in-place activation of
// /root/reference/var_sep/networks/conv.py:56-59 (nn.BatchNorm2d: eps 1e-5, momentum 0.1, affine).
//
// "Grouped": rows [g*rpg, (g+1)*rpg) belong to reference call g; statistics are per (g, channel),
// so one launch over G batched decoder time steps equals G sequential reference calls.
#include <type_traits>

#include "common.cuh"

namespace vs {

// block = 32 channels x up to 32 groups: thread (c, g) turns the fp64 sums of its (group, channel) into mean / invstd
// (the fp64 divisions and the square root run in parallel over the groups); the running-statistics EMA is then applied
// group by group, in the order of the reference's calls, by the g == 0 thread of each channel
constexpr int FIN_C = 32, FIN_G = 32;
__global__ void __launch_bounds__(FIN_C * FIN_G) bn_finalize_kernel(const double* __restrict__ stats, int G, int C, double count, float eps,
                                                                    float momentum, float* __restrict__ mean, float* __restrict__ invstd,
                                                                    float* __restrict__ rmean, float* __restrict__ rvar, long long* nbt) {
    __shared__ float s_mu[FIN_G][FIN_C], s_ub[FIN_G][FIN_C];
    const int cl = threadIdx.x, gl = threadIdx.y;
    const int c = blockIdx.x * FIN_C + cl;
    if (blockIdx.x == 0 && cl == 0 && gl == 0 && nbt != nullptr) *nbt += G;
    float rm = 0.f, rv = 0.f;
    if (gl == 0 && c < C) { rm = rmean ? rmean[c] : 0.f; rv = rvar ? rvar[c] : 0.f; }
    for (int g0 = 0; g0 < G; g0 += blockDim.y) {
        const int g = g0 + gl;
        if (g < G && c < C) {
            const double s1 = stats[((long long)g * C + c) * 2], s2 = stats[((long long)g * C + c) * 2 + 1];
            const double mu = s1 / count;
            double var = s2 / count - mu * mu;          // biased variance, fp64: no cancellation issue at these sizes
            if (var < 0.0) var = 0.0;
            mean[g * C + c] = (float)mu;
            invstd[g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
            s_mu[gl][cl] = (float)mu;
            s_ub[gl][cl] = (float)(count > 1.0 ? var * count / (count - 1.0) : var);
        }
        __syncthreads();
        if (gl == 0 && c < C) {
            for (int k = 0; k < (int)blockDim.y && g0 + k < G; ++k) {
                rm = (1.f - momentum) * rm + momentum * s_mu[k][cl];
                rv = (1.f - momentum) * rv + momentum * s_ub[k][cl];
            }
        }
        __syncthreads();
    }
    if (gl == 0 && c < C) {
        if (rmean) rmean[c] = rm;
        if (rvar) rvar[c] = rv;
    }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = rmean[c];
    invstd[c] = 1.f / sqrtf(rvar[c] + eps);
}

// out = act(gamma*(y-mean)*invstd + beta); 4 channels per thread when C % 4 == 0
template <typename T, bool VEC>
__global__ void bn_act_fwd_kernel(const T* __restrict__ y, T* __restrict__ out, long long rows, int C, long long rpg,
                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int act) {
    const int W = VEC ? 4 : 1;
    const long long total = rows * C / W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * W;
        const long long row = e / C;
        const int c = (int)(e - row * C);
        const int g = (int)(row / rpg);
        if (VEC) {
            const float4 v = ld4<T>(y + e);
            const float4 mu = *reinterpret_cast<const float4*>(mean + g * C + c);
            const float4 is = *reinterpret_cast<const float4*>(invstd + g * C + c);
            const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
            const float4 be = *reinterpret_cast<const float4*>(beta + c);
            float4 o;
            o.x = act_fwd(ga.x * ((v.x - mu.x) * is.x) + be.x, act);
            o.y = act_fwd(ga.y * ((v.y - mu.y) * is.y) + be.y, act);
            o.z = act_fwd(ga.z * ((v.z - mu.z) * is.z) + be.z, act);
            o.w = act_fwd(ga.w * ((v.w - mu.w) * is.w) + be.w, act);
            st4<T>(out + e, o);
        } else {
            const float xh = (ld<T>(y + e) - mean[g * C + c]) * invstd[g * C + c];
            st<T>(out + e, act_fwd(gamma[c] * xh + beta[c], act));
        }
    }
}

// sums[g][c] += { sum dz, sum dz*xhat };  block = 32 channels x 8 row lanes, grid.y = G * chunks
template <typename T>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ y, int C, long long rpg,
                                     int chunks, const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                     double* __restrict__ sums) {
    __shared__ double r1[8][33], r2[8][33];
    constexpr bool EXACT = sizeof(T) == 4;           // parity mode: fp64 element arithmetic and sums
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int g = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
    const long long per = (rpg + chunks - 1) / chunks;
    const long long r0 = (long long)g * rpg + (long long)chunk * per;
    long long r_end = r0 + per;
    if (r_end > (long long)(g + 1) * rpg) r_end = (long long)(g + 1) * rpg;
    float s1 = 0.f, s2 = 0.f;
    double d1 = 0.0, d2 = 0.0;
    if (c < C) {
        const float mu = mean[g * C + c], is = invstd[g * C + c], ga = gamma[c], be = beta[c];
        for (long long r = r0 + threadIdx.y; r < r_end; r += 8) {
            if (EXACT) {
                const double xh = ((double)ld<T>(y + r * C + c) - (double)mu) * (double)is;
                const double dz = (double)ld<T>(dout + r * C + c) * (double)act_grad_from_in((float)((double)ga * xh + (double)be), act);
                d1 += dz;
                d2 += dz * xh;
            } else {
                const float xh = (ld<T>(y + r * C + c) - mu) * is;
                const float dz = ld<T>(dout + r * C + c) * act_grad_from_in(ga * xh + be, act);
                s1 += dz;
                s2 += dz * xh;
            }
        }
    }
    r1[threadIdx.y][threadIdx.x] = EXACT ? d1 : (double)s1;
    r2[threadIdx.y][threadIdx.x] = EXACT ? d2 : (double)s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t1 += r1[k][threadIdx.x]; t2 += r2[k][threadIdx.x]; }
        atomicAdd(&sums[((long long)g * C + c) * 2], t1);
        atomicAdd(&sums[((long long)g * C + c) * 2 + 1], t2);
    }
}

template <typename T, bool VEC>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ y, T* __restrict__ dy,
                                    long long rows, int C, long long rpg, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, int act, const double* __restrict__ sums,
                                    int train) {
    const int W = VEC ? 4 : 1;
    const long long total = rows * C / W;
    const float inv_count = 1.f / (float)rpg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * W;
        const long long row = e / C;
        const int c0 = (int)(e - row * C);
        const int g = (int)(row / rpg);
        float yv[4], dv[4], o[4];
        if (VEC) {
            const float4 a = ld4<T>(y + e), b = ld4<T>(dout + e);
            yv[0] = a.x; yv[1] = a.y; yv[2] = a.z; yv[3] = a.w;
            dv[0] = b.x; dv[1] = b.y; dv[2] = b.z; dv[3] = b.w;
        } else {
            yv[0] = ld<T>(y + e); dv[0] = ld<T>(dout + e);
        }
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const int c = c0 + k;
            const float is = invstd[g * C + c], ga = gamma[c];
            if (sizeof(T) == 4) {          // parity mode: fp64 element arithmetic (see bn_bwd_apply_col_kernel)
                const double xh = ((double)yv[k] - (double)mean[g * C + c]) * (double)is;
                const double dz = (double)dv[k] * (double)act_grad_from_in((float)((double)ga * xh + (double)beta[c]), act);
                const double m1 = train ? sums[((long long)g * C + c) * 2] / (double)rpg : 0.0;
                const double m2 = train ? sums[((long long)g * C + c) * 2 + 1] / (double)rpg : 0.0;
                o[k] = (float)((double)ga * (double)is * (dz - m1 - xh * m2));
                continue;
            }
            const float xh = (yv[k] - mean[g * C + c]) * is;
            const float dz = dv[k] * act_grad_from_in(ga * xh + beta[c], act);
            if (train) {
                const float m1 = (float)sums[((long long)g * C + c) * 2] * inv_count;
                const float m2 = (float)sums[((long long)g * C + c) * 2 + 1] * inv_count;
                o[k] = ga * is * (dz - m1 - xh * m2);
            } else {
                o[k] = ga * is * dz;
            }
        }
        if (VEC) st4<T>(dy + e, make_float4(o[0], o[1], o[2], o[3]));
        else st<T>(dy + e, o[0]);
    }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, int G, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
    for (int g = 0; g < G; ++g) { s1 += sums[((long long)g * C + c) * 2]; s2 += sums[((long long)g * C + c) * 2 + 1]; }
    if (dbeta) dbeta[c] += (float)s1;
    if (dgamma) dgamma[c] += (float)s2;
}

void bn_param_grad(const double* sums, int G, int C, float* dgamma, float* dbeta, cudaStream_t stream) {
    bn_param_grad_kernel<<<(int)cdiv(C, 128), 128, 0, stream>>>(sums, G, C, dgamma, dbeta);
}

// ------------------------------------------------------------------------------------------------
// Fast path ("column-stationary"): every thread owns 16 bytes of channels (8 bf16 / 4 fp32) for the whole
// kernel, keeps that channel slice's statistics and affine parameters in registers and walks down the
// rows of one (group, row-chunk): no integer division and no parameter reload in the streaming loop,
// 16-byte coalesced accesses.  Eligible when C*sizeof(T)/16 divides 256.
// ------------------------------------------------------------------------------------------------
template <typename T> struct Vec;
constexpr int COL_ROWS_IN_FLIGHT = 4;     // 16-byte loads per operand a thread issues before it consumes any of them
template <> struct Vec<float> {
    static constexpr int W = 4;
    static __device__ __forceinline__ void unpack(const uint4& u, float* v) {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ void load(const float* p, float* v) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    static __device__ __forceinline__ void store(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int W = 8;
    static __device__ __forceinline__ void unpack(const uint4& u, float* v) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&b);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// compile-time activation for the common cases (the identity, ReLU, LeakyReLU(0.2)), runtime switch otherwise
#define VS_DISPATCH_ACT(act, A, ...)                                              \
    switch (act) {                                                                \
        case VS_ACT_NONE: { constexpr int A = VS_ACT_NONE; __VA_ARGS__ } break;   \
        case VS_ACT_RELU: { constexpr int A = VS_ACT_RELU; __VA_ARGS__ } break;   \
        case VS_ACT_LEAKY: { constexpr int A = VS_ACT_LEAKY; __VA_ARGS__ } break; \
        default: { constexpr int A = -1; __VA_ARGS__ } break;                     \
    }

struct ColPlan {
    int tpr;             // threads per row = C / W
    int rows_per_iter;   // 256 / tpr
    int chunks;          // row chunks per group
    long long rpg, per;  // rows per group, rows per chunk
};

template <typename T>
__device__ __forceinline__ uint4 raw16(const T* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// cp.async (LDGSTS) staging for the streaming loops: every thread copies its own 16-byte pieces of the next rows into
// its own shared-memory slots and reads them back after cp.async.wait_group -- no registers are held by loads in flight
// and the compiler cannot shrink the number of outstanding requests (it sinks plain loads next to their uses).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// staging (2 buffers x 2 tensors x U rows x 256 threads x 16 B) is also large enough for the reduction scratch
constexpr int REDUCE_SMEM = 2 * 2 * COL_ROWS_IN_FLIGHT * 256 * 16;
static_assert(REDUCE_SMEM >= 2 * 256 * 8 * 4, "reduction scratch");
template <typename K>
static int reduce_smem_attr(K kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, REDUCE_SMEM);
    if (e != cudaSuccess) return fail("bn_reduce_col_kernel smem attribute: %s", cudaGetErrorString(e));
    return 0;
}

// Empty asm that "uses" a loaded vector: placed after the load loop of a group it keeps every load of the group above
// this point (the compiler otherwise sinks loads next to their uses and leaves two 16-byte requests in flight per
// thread instead of eight).
__device__ __forceinline__ void keep_loaded(uint4& v) { asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)); }

template <typename T>
static bool col_plan(long long rows, int C, int G, ColPlan& pl, int blocks_per_sm = 8) {
    constexpr int W = Vec<T>::W;
    if (C % W != 0) return false;
    pl.tpr = C / W;
    if (pl.tpr > 256 || 256 % pl.tpr != 0) return false;
    pl.rows_per_iter = 256 / pl.tpr;
    pl.rpg = rows / G;
    long long chunks = cdiv((long long)blocks_per_sm * num_sms(), G);
    const long long max_chunks = cdiv(pl.rpg, 2LL * COL_ROWS_IN_FLIGHT * pl.rows_per_iter);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    pl.per = cdiv(cdiv(pl.rpg, chunks), pl.rows_per_iter) * pl.rows_per_iter;
    pl.chunks = (int)cdiv(pl.rpg, pl.per);
    return (long long)G * pl.chunks <= 2147483647LL;
}

#define VS_COL_SETUP                                                                      \
    constexpr int W = Vec<T>::W;                                                          \
    const int c = (threadIdx.x % pl.tpr) * W, rl = threadIdx.x / pl.tpr;                  \
    const int g = blockIdx.x / pl.chunks, chunk = blockIdx.x % pl.chunks;                 \
    const long long r0 = (long long)g * pl.rpg + (long long)chunk * pl.per;               \
    long long r1 = r0 + pl.per;                                                           \
    if (r1 > (long long)(g + 1) * pl.rpg) r1 = (long long)(g + 1) * pl.rpg;

// Optional fused vs_bn_finalize: with fin.stats != nullptr every thread derives mean / invstd of its channels from the fp64
// sums itself (the arithmetic of bn_finalize_kernel), the chunk-0 block of every group publishes them for the backward
// pass, and the (group 0, chunk 0) block applies the running-statistics EMA group by group: one launch less per layer.
constexpr int BN_FIN_MAX_C = 1024;
struct BnFinArgs {
    const double* stats; double count; float eps, momentum;
    float* mean_out; float* invstd_out; float* rmean; float* rvar; long long* nbt; int G;
};

// ACT >= 0: activation known at compile time (no per-element switch); ACT < 0: use the runtime argument
template <typename T, int ACT>
__global__ void __launch_bounds__(256) bn_act_fwd_col_kernel(const T* __restrict__ y, T* __restrict__ out, int C, ColPlan pl,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta, int act_rt,
                                                             const BnFinArgs fin) {
    const int act = ACT >= 0 ? ACT : act_rt;
    VS_COL_SETUP
    float mu[W], is[W], ga[W], be[W];
    if (fin.stats != nullptr) {
        // one thread per CHANNEL does the fp64 division / square root (not one per channel slice and row lane: the fp64
        // sequences are a few hundred cycles each and would dominate the short launches), the block shares the result
        __shared__ float s_mu[BN_FIN_MAX_C], s_is[BN_FIN_MAX_C];
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            const double s1 = fin.stats[((long long)g * C + ch) * 2], s2 = fin.stats[((long long)g * C + ch) * 2 + 1];
            const double m = s1 / fin.count;
            double var = s2 / fin.count - m * m;
            if (var < 0.0) var = 0.0;
            const float mf = (float)m, isf = (float)(1.0 / sqrt(var + (double)fin.eps));
            s_mu[ch] = mf; s_is[ch] = isf;
            if (chunk == 0) {
                fin.mean_out[g * C + ch] = mf; fin.invstd_out[g * C + ch] = isf;
                if (g == 0) {
                    if (ch == 0 && fin.nbt != nullptr) *fin.nbt += fin.G;
                    float rm = fin.rmean ? fin.rmean[ch] : 0.f, rv = fin.rvar ? fin.rvar[ch] : 0.f;
                    for (int gg = 0; gg < fin.G; ++gg) {           // in the order of the reference's calls
                        const double t1 = fin.stats[((long long)gg * C + ch) * 2], t2 = fin.stats[((long long)gg * C + ch) * 2 + 1];
                        const double mg = t1 / fin.count;
                        double vg = t2 / fin.count - mg * mg;
                        if (vg < 0.0) vg = 0.0;
                        rm = (1.f - fin.momentum) * rm + fin.momentum * (float)mg;
                        rv = (1.f - fin.momentum) * rv + fin.momentum * (float)(fin.count > 1.0 ? vg * fin.count / (fin.count - 1.0) : vg);
                    }
                    if (fin.rmean) fin.rmean[ch] = rm;
                    if (fin.rvar) fin.rvar[ch] = rv;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < W; ++k) { mu[k] = s_mu[c + k]; is[k] = s_is[c + k]; ga[k] = gamma[c + k]; be[k] = beta[c + k]; }
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) { mu[k] = mean[g * C + c + k]; is[k] = invstd[g * C + c + k]; ga[k] = gamma[c + k]; be[k] = beta[c + k]; }
    }
    // several independent rows in flight per thread (all raw 16-byte loads of a group are issued before any is
    // consumed; the group loop is branch-free so that the compiler keeps them back to back): the kernel is bound by
    // the bytes in flight per SM, not by arithmetic
    constexpr int U = COL_ROWS_IN_FLIGHT;
    const long long step = pl.rows_per_iter;
    long long r = r0 + rl;
    for (; r + (U - 1) * step < r1; r += U * step) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = raw16(y + (r + u * step) * C + c);
#pragma unroll
        for (int u = 0; u < U; ++u) keep_loaded(raw[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v[W];
            Vec<T>::unpack(raw[u], v);
#pragma unroll
            for (int k = 0; k < W; ++k) v[k] = act_fwd(ga[k] * ((v[k] - mu[k]) * is[k]) + be[k], act);
            Vec<T>::store(out + (r + u * step) * C + c, v);
        }
    }
    for (; r < r1; r += step) {
        float v[W];
        Vec<T>::unpack(raw16(y + r * C + c), v);
#pragma unroll
        for (int k = 0; k < W; ++k) v[k] = act_fwd(ga[k] * ((v[k] - mu[k]) * is[k]) + be[k], act);
        Vec<T>::store(out + r * C + c, v);
    }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256) bn_bwd_apply_col_kernel(const T* __restrict__ dout, const T* __restrict__ y, T* __restrict__ dy,
                                                               int C, ColPlan pl, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int act_rt,
                                                               const double* __restrict__ sums, int train,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta, int G) {
    const int act = ACT >= 0 ? ACT : act_rt;
    VS_COL_SETUP
    // fp32 storage = parity mode: the element arithmetic runs in fp64, as ATen's CPU batch-norm backward does (accumulate
    // type double).  dz - mean(dz) - xhat * mean(dz * xhat) cancels to a small fraction of dz in the layers next to a
    // nearly-linear output (measured: 1e5-fold in the last BatchNorm layer of the mnist-small-mul case, where fp32
    // element arithmetic left the input gradient 1.3e-2 off while every kernel matched its specification on random data).
    constexpr bool EXACT = sizeof(T) == 4;
    float mu[W], is[W], ga[W], be[W], m1[W], m2[W];
    double m1d[EXACT ? W : 1], m2d[EXACT ? W : 1];
    const float inv_count = 1.f / (float)pl.rpg;
    if ((dgamma != nullptr || dbeta != nullptr) && g == 0 && chunk == 0 && rl == 0) {
        // the affine gradients (bn_param_grad_kernel folded in): sum of the per-group sums, one thread per channel slice
#pragma unroll
        for (int k = 0; k < W; ++k) {
            double s1 = 0.0, s2 = 0.0;
            for (int gg = 0; gg < G; ++gg) { s1 += sums[((long long)gg * C + c + k) * 2]; s2 += sums[((long long)gg * C + c + k) * 2 + 1]; }
            if (dbeta) dbeta[c + k] += (float)s1;
            if (dgamma) dgamma[c + k] += (float)s2;
        }
    }
#pragma unroll
    for (int k = 0; k < W; ++k) {
        mu[k] = mean[g * C + c + k]; is[k] = invstd[g * C + c + k]; ga[k] = gamma[c + k]; be[k] = beta[c + k];
        m1[k] = train ? (float)sums[((long long)g * C + c + k) * 2] * inv_count : 0.f;
        m2[k] = train ? (float)sums[((long long)g * C + c + k) * 2 + 1] * inv_count : 0.f;
        if (EXACT) {
            m1d[k] = train ? sums[((long long)g * C + c + k) * 2] / (double)pl.rpg : 0.0;
            m2d[k] = train ? sums[((long long)g * C + c + k) * 2 + 1] / (double)pl.rpg : 0.0;
        }
    }
    constexpr int U = COL_ROWS_IN_FLIGHT;
    const long long step = pl.rows_per_iter;
    auto apply = [&](const uint4& qy, const uint4& qd, long long rr) {
        float v[W], d[W];
        Vec<T>::unpack(qy, v);
        Vec<T>::unpack(qd, d);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            if (EXACT) {
                const double xh = ((double)v[k] - (double)mu[k]) * (double)is[k];
                const double dz = (double)d[k] * (double)act_grad_from_in((float)((double)ga[k] * xh + (double)be[k]), act);
                d[k] = (float)((double)ga[k] * (double)is[k] * (dz - m1d[k] - xh * m2d[k]));
            } else {
                const float xh = (v[k] - mu[k]) * is[k];
                const float dz = d[k] * act_grad_from_in(ga[k] * xh + be[k], act);
                d[k] = ga[k] * is[k] * (dz - m1[k] - xh * m2[k]);
            }
        }
        Vec<T>::store(dy + rr * C + c, d);
    };
    // cp.async staging as in bn_reduce_col_kernel: [2 buffers][2 tensors][U rows][256 threads] x 16 bytes
    extern __shared__ uint4 stage[];
    const long long nrows = r1 > r0 + rl ? (r1 - r0 - rl + step - 1) / step : 0;
    const int ngroups = (int)((nrows + U - 1) / U);
    auto slot = [&](int buf, int tensor, int u) -> uint4* { return stage + ((buf * 2 + tensor) * U + u) * 256 + threadIdx.x; };
    auto issue = [&](int grp) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = (long long)grp * U + u;
            if (i < nrows) {
                const long long rr = r0 + rl + i * step;
                cp_async16(slot(grp & 1, 0, u), y + rr * C + c);
                cp_async16(slot(grp & 1, 1, u), dout + rr * C + c);
            }
        }
        cp_async_commit();
    };
    if (ngroups > 0) issue(0);
    for (int grp = 0; grp < ngroups; ++grp) {
        if (grp + 1 < ngroups) { issue(grp + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = (long long)grp * U + u;
            if (i < nrows) apply(*slot(grp & 1, 0, u), *slot(grp & 1, 1, u), r0 + rl + i * step);
        }
    }
}

// MODE 0: BatchNorm backward sums {sum dz, sum dz*xhat};  MODE 1: forward statistics {sum y, sum y^2}
template <typename T, int MODE, int ACT>
__global__ void __launch_bounds__(256) bn_reduce_col_kernel(const T* __restrict__ dout, const T* __restrict__ y, int C, ColPlan pl,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, int act_rt,
                                                            double* __restrict__ sums) {
    const int act = ACT >= 0 ? ACT : act_rt;
    VS_COL_SETUP
    // dynamic shared memory: staging slots [2 buffers][NT tensors][U rows][256 threads] x 16 bytes, reused as the
    // block-reduction scratch [2][256 * W] floats at the end
    extern __shared__ uint4 stage[];
    constexpr int NT = MODE == 0 ? 2 : 1;
    constexpr bool EXACT = sizeof(T) == 4;
    using acc_t = typename std::conditional<EXACT, double, float>::type;
    float mu[W], is[W], ga[W], be[W];
    acc_t s1[W], s2[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
        s1[k] = 0; s2[k] = 0;
        if (MODE == 0) { mu[k] = mean[g * C + c + k]; is[k] = invstd[g * C + c + k]; ga[k] = gamma[c + k]; be[k] = beta[c + k]; }
    }
    constexpr int U = COL_ROWS_IN_FLIGHT;
    const long long step = pl.rows_per_iter;
    auto accumulate = [&](const uint4& qy, const uint4& qd) {
        float v[W];
        Vec<T>::unpack(qy, v);
        if (MODE == 0) {
            float d[W];
            Vec<T>::unpack(qd, d);
#pragma unroll
            for (int k = 0; k < W; ++k) {
                if (EXACT) {       // parity mode: fp64 element arithmetic and fp64 running sums (see bn_bwd_apply_col_kernel)
                    const double xh = ((double)v[k] - (double)mu[k]) * (double)is[k];
                    const double dz = (double)d[k] * (double)act_grad_from_in((float)((double)ga[k] * xh + (double)be[k]), act);
                    s1[k] += dz; s2[k] += dz * xh;
                } else {
                    const float xh = (v[k] - mu[k]) * is[k];
                    const float dz = d[k] * act_grad_from_in(ga[k] * xh + be[k], act);
                    s1[k] += dz; s2[k] += dz * xh;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < W; ++k) { s1[k] += v[k]; s2[k] += (acc_t)v[k] * (acc_t)v[k]; }
        }
    };
    // rows of this thread: r0 + rl + i * step; groups of U rows, group g+1 is in flight while group g is consumed
    const long long nrows = r1 > r0 + rl ? (r1 - r0 - rl + step - 1) / step : 0;
    const int ngroups = (int)((nrows + U - 1) / U);
    auto slot = [&](int buf, int tensor, int u) -> uint4* { return stage + ((buf * NT + tensor) * U + u) * 256 + threadIdx.x; };
    auto issue = [&](int grp) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = (long long)grp * U + u;
            if (i < nrows) {
                const long long rr = r0 + rl + i * step;
                cp_async16(slot(grp & 1, 0, u), y + rr * C + c);
                if (MODE == 0) cp_async16(slot(grp & 1, 1, u), dout + rr * C + c);
            }
        }
        cp_async_commit();
    };
    if (ngroups > 0) issue(0);
    for (int grp = 0; grp < ngroups; ++grp) {
        if (grp + 1 < ngroups) { issue(grp + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if ((long long)grp * U + u < nrows) accumulate(*slot(grp & 1, 0, u), MODE == 0 ? *slot(grp & 1, 1, u) : make_uint4(0, 0, 0, 0));
    }
    __syncthreads();        // staging memory becomes the reduction scratch
    acc_t (*red)[256 * W] = reinterpret_cast<acc_t (*)[256 * W]>(stage);
    static_assert(2 * 256 * W * sizeof(acc_t) <= REDUCE_SMEM, "reduction scratch");
    // reduce the row lanes of the block (threads with equal threadIdx.x % tpr), then fp64 atomics
#pragma unroll
    for (int k = 0; k < W; ++k) { red[0][threadIdx.x * W + k] = s1[k]; red[1][threadIdx.x * W + k] = s2[k]; }
    __syncthreads();
    if (rl == 0) {
#pragma unroll
        for (int k = 0; k < W; ++k) {
            double t1 = 0.0, t2 = 0.0;
            for (int j = 0; j < pl.rows_per_iter; ++j) {
                t1 += (double)red[0][(j * pl.tpr + threadIdx.x) * W + k];
                t2 += (double)red[1][(j * pl.tpr + threadIdx.x) * W + k];
            }
            atomicAdd(&sums[((long long)g * C + c + k) * 2], t1);
            atomicAdd(&sums[((long long)g * C + c + k) * 2 + 1], t2);
        }
    }
}

// per-(group, channel) sum / sum of squares of a stored [rows, C] tensor (second-pass BatchNorm statistics of
// the tensor-core and thin convolution paths)
int column_stats(const void* y, int dtype, long long rows, int C, int G, double* stats, cudaStream_t stream) {
    VS_DISPATCH_DTYPE(dtype, T, {
        ColPlan pl;
        if (!col_plan<T>(rows, C, G, pl, 8)) return -1;
        if (int rc = reduce_smem_attr(bn_reduce_col_kernel<T, 1, 0>)) return rc;
        bn_reduce_col_kernel<T, 1, 0><<<(unsigned)(G * pl.chunks), 256, REDUCE_SMEM, stream>>>(nullptr, (const T*)y, C, pl, nullptr,
                                                                                             nullptr, nullptr, nullptr, 0, stats);
    });
    return launched("bn_reduce_col_kernel");
}

static int ew_blocks(long long work) {
    long long b = cdiv(work, 256);
    const long long cap = 8LL * num_sms();
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace vs

using namespace vs;

extern "C" int vs_bn_finalize(const double* stats, int32_t G, int32_t C, int64_t count, float eps, float momentum,
                              float* mean, float* invstd, float* running_mean, float* running_var,
                              int64_t* num_batches_tracked, void* stream) {
    VS_REQUIRE(G >= 1 && C >= 1 && count >= 1, "bn_finalize: bad sizes");
    int gy = 1;
    while (gy < G && gy < FIN_G) gy <<= 1;
    bn_finalize_kernel<<<(int)cdiv(C, FIN_C), dim3(FIN_C, gy), 0, as_stream(stream)>>>(
        stats, G, C, (double)count, eps, momentum, mean, invstd, running_mean, running_var,
        reinterpret_cast<long long*>(num_batches_tracked));
    return launched("bn_finalize_kernel");
}

extern "C" int vs_bn_eval_stats(const float* running_mean, const float* running_var, int32_t C, float eps, float* mean,
                                float* invstd, void* stream) {
    bn_eval_stats_kernel<<<(int)cdiv(C, 128), 128, 0, as_stream(stream)>>>(running_mean, running_var, C, eps, mean, invstd);
    return launched("bn_eval_stats_kernel");
}

extern "C" int vs_bn_act_forward(const void* y, void* out, int32_t dtype, int64_t rows, int32_t C, int32_t G,
                                 const float* mean, const float* invstd, const float* gamma, const float* beta,
                                 int32_t act, void* stream) {
    VS_REQUIRE(G >= 1 && rows % G == 0, "bn_act_forward: rows %lld not divisible by groups %d", (long long)rows, G);
    if (rows == 0) return 0;
    const long long rpg = rows / G;
    VS_DISPATCH_DTYPE(dtype, T, {
        ColPlan pl;
        if (col_plan<T>(rows, C, G, pl)) {
            VS_DISPATCH_ACT(act, A, {
                BnFinArgs none;
                memset(&none, 0, sizeof(none));
                bn_act_fwd_col_kernel<T, A><<<(unsigned)(G * pl.chunks), 256, 0, as_stream(stream)>>>((const T*)y, (T*)out, C, pl, mean, invstd, gamma, beta, act, none);
            });
            return launched("bn_act_fwd_col_kernel");
        }
    });
    const bool vec = (C % 4) == 0;
    const int blocks = ew_blocks(rows * C / (vec ? 4 : 1));
    VS_DISPATCH_DTYPE(dtype, T, {
        if (vec) bn_act_fwd_kernel<T, true><<<blocks, 256, 0, as_stream(stream)>>>((const T*)y, (T*)out, rows, C, rpg, mean, invstd, gamma, beta, act);
        else bn_act_fwd_kernel<T, false><<<blocks, 256, 0, as_stream(stream)>>>((const T*)y, (T*)out, rows, C, rpg, mean, invstd, gamma, beta, act);
    });
    return launched("bn_act_fwd_kernel");
}

extern "C" int vs_bn_finalize_act_forward(const double* stats, int32_t G, int32_t C, int64_t count, float eps, float momentum,
                                          float* mean, float* invstd, float* running_mean, float* running_var,
                                          int64_t* num_batches_tracked, const void* y, void* out, int32_t dtype, int64_t rows,
                                          const float* gamma, const float* beta, int32_t act, void* stream) {
    VS_REQUIRE(G >= 1 && C >= 1 && count >= 1 && rows % G == 0, "bn_finalize_act_forward: bad sizes");
    if (rows > 0) {
        bool done = false;
        VS_DISPATCH_DTYPE(dtype, T, {
            ColPlan pl;
            if (C <= BN_FIN_MAX_C && col_plan<T>(rows, C, G, pl)) {
                BnFinArgs fin = {stats, (double)count, eps, momentum, mean, invstd, running_mean, running_var,
                                 reinterpret_cast<long long*>(num_batches_tracked), G};
                VS_DISPATCH_ACT(act, A, {
                    bn_act_fwd_col_kernel<T, A><<<(unsigned)(G * pl.chunks), 256, 0, as_stream(stream)>>>((const T*)y, (T*)out, C, pl, mean, invstd, gamma, beta, act, fin);
                });
                done = true;
            }
        });
        if (done) return launched("bn_act_fwd_col_kernel");
    }
    if (int rc = vs_bn_finalize(stats, G, C, count, eps, momentum, mean, invstd, running_mean, running_var, num_batches_tracked, stream)) return rc;
    return vs_bn_act_forward(y, out, dtype, rows, C, G, mean, invstd, gamma, beta, act, stream);
}

extern "C" int vs_bn_act_backward_reduce(const void* dout, const void* y, int32_t dtype, int64_t rows, int32_t C,
                                         int32_t G, const float* mean, const float* invstd, const float* gamma,
                                         const float* beta, int32_t act, double* sums, void* stream) {
    VS_REQUIRE(G >= 1 && rows % G == 0, "bn_act_backward_reduce: rows not divisible by groups");
    if (rows == 0) return 0;
    const long long rpg = rows / G;
    VS_DISPATCH_DTYPE(dtype, T, {
        ColPlan pl;
        if (col_plan<T>(rows, C, G, pl, 9)) {      // 3 resident blocks per SM (80 registers, 64 KB): 9 per SM = three full waves
            VS_DISPATCH_ACT(act, A, {
                if (int rc = reduce_smem_attr(bn_reduce_col_kernel<T, 0, A>)) return rc;
                bn_reduce_col_kernel<T, 0, A><<<(unsigned)(G * pl.chunks), 256, REDUCE_SMEM, as_stream(stream)>>>((const T*)dout, (const T*)y, C, pl, mean, invstd, gamma, beta, act, sums);
            });
            return launched("bn_reduce_col_kernel");
        }
    });
    const int cx = (int)cdiv(C, 32);
    long long chunks = cdiv(4LL * num_sms(), (long long)cx * G);
    if (chunks > cdiv(rpg, 64)) chunks = cdiv(rpg, 64);
    if (chunks < 1) chunks = 1;
    VS_REQUIRE((long long)G * chunks <= 65535, "bn reduce grid too large");
    dim3 grid(cx, (unsigned)(G * chunks)), block(32, 8);
    VS_DISPATCH_DTYPE(dtype, T, (bn_bwd_reduce_kernel<T><<<grid, block, 0, as_stream(stream)>>>(
        (const T*)dout, (const T*)y, C, rpg, (int)chunks, mean, invstd, gamma, beta, act, sums)));
    return launched("bn_bwd_reduce_kernel");
}

extern "C" int vs_bn_act_backward_apply(const void* dout, const void* y, void* dy, int32_t dtype, int64_t rows,
                                        int32_t C, int32_t G, const float* mean, const float* invstd,
                                        const float* gamma, const float* beta, int32_t act, const double* sums,
                                        int32_t train, float* dgamma, float* dbeta, void* stream) {
    VS_REQUIRE(G >= 1 && rows % G == 0, "bn_act_backward_apply: rows not divisible by groups");
    if (rows == 0) return 0;
    const long long rpg = rows / G;
    bool done = false;
    VS_DISPATCH_DTYPE(dtype, T, {
        ColPlan pl;
        if (col_plan<T>(rows, C, G, pl)) {
            VS_DISPATCH_ACT(act, A, {
                if (int rc = reduce_smem_attr(bn_bwd_apply_col_kernel<T, A>)) return rc;
                bn_bwd_apply_col_kernel<T, A><<<(unsigned)(G * pl.chunks), 256, REDUCE_SMEM, as_stream(stream)>>>((const T*)dout, (const T*)y, (T*)dy, C, pl, mean, invstd, gamma, beta, act, sums, train, dgamma, dbeta, G);
            });
            done = true;
        }
    });
    const bool vec = (C % 4) == 0;
    const int blocks = ew_blocks(rows * C / (vec ? 4 : 1));
    if (!done) VS_DISPATCH_DTYPE(dtype, T, {
        if (vec) bn_bwd_apply_kernel<T, true><<<blocks, 256, 0, as_stream(stream)>>>((const T*)dout, (const T*)y, (T*)dy, rows, C, rpg, mean, invstd, gamma, beta, act, sums, train);
        else bn_bwd_apply_kernel<T, false><<<blocks, 256, 0, as_stream(stream)>>>((const T*)dout, (const T*)y, (T*)dy, rows, C, rpg, mean, invstd, gamma, beta, act, sums, train);
    });
    int rc = launched("bn_bwd_apply_kernel");
    if (rc) return rc;
    if (!done && (dgamma != nullptr || dbeta != nullptr)) {          // (the column kernel adds them itself)
        bn_param_grad_kernel<<<(int)cdiv(C, 128), 128, 0, as_stream(stream)>>>(sums, G, C, dgamma, dbeta);
        rc = launched("bn_param_grad_kernel");
    }
    return rc;
}

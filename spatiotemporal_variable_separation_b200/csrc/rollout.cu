// Latent rollout of the MLP residual stepper as ONE launch per direction.
//
// Replaces the loop of /root/reference/var_sep/networks/model.py:78-83 over MLPResnet.forward
// (resnet.py:42-50; mlp.py:66-71):   x <- x + L3(relu(L2(relu(L1 x))))   n_blocks times per time step,
// for T-1 strictly sequential steps.  The recurrence couples time steps but not batch rows, so each CTA owns
// a few rows for the WHOLE rollout (no grid-wide synchronisation): it keeps its rows' state in shared
// memory, streams the (L2-resident, ~1 MB) weights with coalesced 16-byte loads once per step, and writes
// only what backprop-through-time needs (the two post-ReLU hiddens and the block inputs).
// The backward kernel runs the adjoint recurrence the same way and emits the per-step pre-activation
// gradients; the weight gradients are then three (T-1)*B-row GEMMs (one launch each, vs_conv_wgrad) instead of
// 3*(T-1) small ones.  All arithmetic is fp32 (latent dynamics are precision-sensitive and tiny: 0.2 % of FLOPs).
#include "common.cuh"

namespace vs {

constexpr int RO_MAX_BLOCKS = 8;
constexpr int RO_THREADS = 512;

struct RolloutWeights {
    const float* w1[RO_MAX_BLOCKS]; const float* b1[RO_MAX_BLOCKS];
    const float* w2[RO_MAX_BLOCKS]; const float* b2[RO_MAX_BLOCKS];
    const float* w3[RO_MAX_BLOCKS]; const float* b3[RO_MAX_BLOCKS];
};

// out[r][n] = act(bias[n] + sum_k W[n][k] * in[r][k])   for RB rows held in shared memory.
// One warp per group of NO outputs: lanes split k, the NO weight rows are fetched with back-to-back coalesced
// float4 loads (NO*nin/128 independent requests in flight per lane - the layer is L2-latency bound, not FLOP
// bound), then NO*RB shuffle reductions.
template <int RB, bool RELU, int NO>
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias, int nout, int nin,
                                           const float* __restrict__ in_s, int in_ld, float* __restrict__ out_s, int out_ld) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = RO_THREADS / 32;
    for (int n0 = warp * NO; n0 < nout; n0 += nwarps * NO) {
        float acc[NO][RB];
#pragma unroll
        for (int o = 0; o < NO; ++o)
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[o][r] = 0.f;
        if ((nin & 3) == 0) {
            for (int k = lane * 4; k < nin; k += 128) {
                float4 w[NO];
#pragma unroll
                for (int o = 0; o < NO; ++o)
                    w[o] = n0 + o < nout ? __ldg(reinterpret_cast<const float4*>(W + (long long)(n0 + o) * nin + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float4 x = *reinterpret_cast<const float4*>(in_s + r * in_ld + k);
#pragma unroll
                    for (int o = 0; o < NO; ++o) acc[o][r] += w[o].x * x.x + w[o].y * x.y + w[o].z * x.z + w[o].w * x.w;
                }
            }
        } else {
            for (int k = lane; k < nin; k += 32) {
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    const float w = n0 + o < nout ? __ldg(W + (long long)(n0 + o) * nin + k) : 0.f;
#pragma unroll
                    for (int r = 0; r < RB; ++r) acc[o][r] = fmaf(w, in_s[r * in_ld + k], acc[o][r]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NO; ++o)
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[o][r] = warp_sum(acc[o][r]);
        if (lane == 0) {
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                if (n0 + o >= nout) break;
                const float b = bias ? bias[n0 + o] : 0.f;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float v = acc[o][r] + b;
                    out_s[r * out_ld + n0 + o] = RELU ? fmaxf(v, 0.f) : v;
                }
            }
        }
    }
}

// shared memory: x[RB][d] | u[RB][h] | v[RB][h] | r[RB][d]
template <int RB>
__global__ void __launch_bounds__(RO_THREADS) rollout_fwd_kernel(float* __restrict__ codes, RolloutWeights w, int T, int B, int d, int h,
                                                                 int nb, float* __restrict__ hidden, float* __restrict__ xin,
                                                                 float* __restrict__ res) {
#define hslot(j, which, t) ((((long long)(j) * 2 + (which)) * (T - 1)) + ((t) - 1))
    extern __shared__ float sm[];
    float* x = sm; float* u = x + RB * d; float* v = u + RB * h; float* rr = v + RB * h;
    const int row0 = blockIdx.x * RB;
    const int nrow = min(RB, B - row0);
    for (int i = threadIdx.x; i < RB * d; i += RO_THREADS) {
        const int r = i / d, k = i % d;
        x[i] = r < nrow ? codes[(long long)(row0 + r) * d + k] : 0.f;
    }
    __syncthreads();
    for (int t = 1; t < T; ++t) {
        for (int j = 0; j < nb; ++j) {
            const long long slot = (long long)j * (T - 1) + (t - 1);
            if (xin) for (int i = threadIdx.x; i < nrow * d; i += RO_THREADS) xin[(slot * B + row0) * d + i] = x[i];
            dense_rows<RB, true, 4>(w.w1[j], w.b1[j], h, d, x, d, u, h);
            __syncthreads();
            dense_rows<RB, true, 8>(w.w2[j], w.b2[j], h, h, u, h, v, h);
            __syncthreads();
            dense_rows<RB, false, 2>(w.w3[j], w.b3[j], d, h, v, h, rr, d);
            __syncthreads();
            if (hidden)
                for (int i = threadIdx.x; i < nrow * h; i += RO_THREADS) {
                    hidden[(hslot(j, 0, t) * B + row0) * h + i] = u[i];
                    hidden[(hslot(j, 1, t) * B + row0) * h + i] = v[i];
                }
            for (int i = threadIdx.x; i < nrow * d; i += RO_THREADS) {
                if (res) res[(slot * B + row0) * d + i] = rr[i];
                x[i] += rr[i];
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < nrow * d; i += RO_THREADS) codes[((long long)t * B + row0) * d + i] = x[i];
    }
}

// adjoint recurrence.  g[RB][d] running gradient | a[RB][h] | b[RB][h] | tmp[RB][d]
template <int RB>
__global__ void __launch_bounds__(RO_THREADS) rollout_bwd_kernel(float* __restrict__ dcodes, RolloutWeights w, int T, int B, int d, int h,
                                                                 int nb, const float* __restrict__ hidden, float* __restrict__ dres,
                                                                 float* __restrict__ dhidden) {
    extern __shared__ float sm[];
    float* g = sm; float* a = g + RB * d; float* b = a + RB * h; float* tmp = b + RB * h;
    const int row0 = blockIdx.x * RB;
    const int nrow = min(RB, B - row0);
    for (int i = threadIdx.x; i < RB * d; i += RO_THREADS) {
        const int r = i / d, k = i % d;
        g[i] = r < nrow ? dcodes[((long long)(T - 1) * B + row0 + r) * d + k] : 0.f;
    }
    for (int i = threadIdx.x; i < RB * h; i += RO_THREADS) { a[i] = 0.f; b[i] = 0.f; }
    __syncthreads();
    for (int t = T - 1; t >= 1; --t) {
        for (int j = nb - 1; j >= 0; --j) {
            const long long slot = (long long)j * (T - 1) + (t - 1);
            // dr = g (gradient of the residual output);  d(a2) = (dr W3) * [h2 > 0]
            for (int i = threadIdx.x; i < nrow * d; i += RO_THREADS) dres[(slot * B + row0) * d + i] = g[i];
            dense_rows<RB, false, 4>(w.w3[j], nullptr, h, d, g, d, b, h);          // w3 = W3^T [h][d]
            __syncthreads();
            for (int i = threadIdx.x; i < nrow * h; i += RO_THREADS) {
                const float hv = hidden[(hslot(j, 1, t) * B + row0) * h + i];
                const float v = hv > 0.f ? b[i] : 0.f;
                b[i] = v;
                dhidden[(hslot(j, 1, t) * B + row0) * h + i] = v;
            }
            __syncthreads();
            // d(a1) = (d(a2) W2) * [h1 > 0]
            dense_rows<RB, false, 8>(w.w2[j], nullptr, h, h, b, h, a, h);          // w2 = W2^T [h][h]
            __syncthreads();
            for (int i = threadIdx.x; i < nrow * h; i += RO_THREADS) {
                const float hv = hidden[(hslot(j, 0, t) * B + row0) * h + i];
                const float v = hv > 0.f ? a[i] : 0.f;
                a[i] = v;
                dhidden[(hslot(j, 0, t) * B + row0) * h + i] = v;
            }
            __syncthreads();
            // dx_in = g + d(a1) W1
            dense_rows<RB, false, 2>(w.w1[j], nullptr, d, h, a, h, tmp, d);        // w1 = W1^T [d][h]
            __syncthreads();
            for (int i = threadIdx.x; i < RB * d; i += RO_THREADS) g[i] += tmp[i];
            __syncthreads();
        }
        // the decoder's gradient w.r.t. codes[t-1] joins the chain
        for (int i = threadIdx.x; i < nrow * d; i += RO_THREADS) {
            const long long o = ((long long)(t - 1) * B + row0) * d + i;
            g[i] += dcodes[o];
            if (t == 1) dcodes[o] = g[i];
        }
        __syncthreads();
    }
}

static int fill_weights(RolloutWeights& rw, const float* const* w_host, int nb) {
    VS_REQUIRE(nb >= 1 && nb <= RO_MAX_BLOCKS, "latent rollout: n_blocks must be in [1,%d]", RO_MAX_BLOCKS);
    for (int j = 0; j < nb; ++j) {
        rw.w1[j] = w_host[6 * j + 0]; rw.b1[j] = w_host[6 * j + 1];
        rw.w2[j] = w_host[6 * j + 2]; rw.b2[j] = w_host[6 * j + 3];
        rw.w3[j] = w_host[6 * j + 4]; rw.b3[j] = w_host[6 * j + 5];
    }
    return 0;
}

}  // namespace vs

using namespace vs;

constexpr int RB = 2;

extern "C" int vs_latent_rollout_forward(float* codes, const float* const* w_host, int32_t T, int32_t B, int32_t d, int32_t h,
                                         int32_t n_blocks, float* hidden, float* xin, float* res, void* stream) {
    VS_REQUIRE(T >= 1 && B >= 1 && d >= 1 && h >= 1, "latent rollout: bad sizes");
    if (T == 1) return 0;
    RolloutWeights rw;
    if (int rc = fill_weights(rw, w_host, n_blocks)) return rc;
    const int smem = (2 * RB * d + 2 * RB * h) * (int)sizeof(float);
    VS_REQUIRE(smem <= 200 * 1024 && (h % 4) == 0, "latent rollout: hidden size %d not supported", h);
    if (smem > 48 * 1024) cudaFuncSetAttribute(rollout_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rollout_fwd_kernel<RB><<<(unsigned)cdiv(B, RB), RO_THREADS, smem, as_stream(stream)>>>(codes, rw, T, B, d, h, n_blocks, hidden, xin, res);
    return launched("rollout_fwd_kernel");
}

extern "C" int vs_latent_rollout_backward(float* dcodes, const float* const* w_host, int32_t T, int32_t B, int32_t d, int32_t h,
                                          int32_t n_blocks, const float* hidden, float* dres, float* dhidden, void* stream) {
    VS_REQUIRE(T >= 1 && B >= 1 && d >= 1 && h >= 1, "latent rollout: bad sizes");
    VS_REQUIRE(hidden != nullptr && dres != nullptr && dhidden != nullptr, "latent rollout backward: null buffer");
    if (T == 1) return 0;
    RolloutWeights rw;
    if (int rc = fill_weights(rw, w_host, n_blocks)) return rc;
    const int smem = (2 * RB * d + 2 * RB * h) * (int)sizeof(float);
    VS_REQUIRE(smem <= 200 * 1024, "latent rollout: hidden size %d not supported", h);
    if (smem > 48 * 1024) cudaFuncSetAttribute(rollout_bwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rollout_bwd_kernel<RB><<<(unsigned)cdiv(B, RB), RO_THREADS, smem, as_stream(stream)>>>(dcodes, rw, T, B, d, h, n_blocks, hidden, dres, dhidden);
    return launched("rollout_bwd_kernel");
}

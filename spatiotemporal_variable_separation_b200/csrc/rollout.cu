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
#include "tc_ptx.cuh"      // warp_transpose_sum32

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

// =================================================================================================
// Weight-resident cluster variant (one block of 8 CTAs per 8 batch rows).
//
// The per-row kernels above stream every block's weights (h*h*4 bytes, ~1 MB) from L2 into each SM once per time step,
// which bounds a step at the per-SM L2 ingest rate.  Here the hidden units are split over the 8 CTAs of a thread-block
// cluster: CTA c keeps ITS rows of W1 and W2 and ITS columns of W3 in shared memory for the whole rollout (138 KB for
// h = 512, d = 20) and the cluster exchanges activations through distributed shared memory:
//   layer A (d -> h, slice of h per CTA)   -> all-gather of the slice into every CTA's copy of the h-vector
//   layer B (h -> h, slice of h per CTA)   -> stays local
//   layer C (h -> d, K-split over the slice) -> partial sums to every CTA, summed in rank order (deterministic)
// with two cluster barriers per block and step.  Forward and adjoint recurrences have the same shape (the backward
// takes the transposed weights), so one kernel template serves both.
// =================================================================================================
namespace vs {

constexpr int RC_THREADS = 512;
constexpr int RC_MAX_ITEMS = 2;   // (unit, row) items per thread in layer A, warp tasks per warp in layer B

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a location in this CTA's shared memory) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* p, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(rank));
    return out;
}
__device__ __forceinline__ void dsmem_store(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct RolloutClusterArgs {
    float* codes;            // fwd: codes [T][B][d];   bwd: dcodes
    const float* hidden;     // bwd: saved post-ReLU hiddens
    float* hidden_out;       // fwd: hidden;            bwd: dhidden
    float* xin;              // fwd only (may be null)
    float* res;              // fwd: res (may be null); bwd: dres
    int T, B, d, h, nb;
};

// CS = CTAs per cluster (8 portable, 16 where the device can co-schedule it), R = batch rows per cluster (8 or 16).
// The arithmetic of a row does not depend on R (rows never mix; every accumulator goes through the same lane
// butterfly), so R may follow the batch size; CS changes the order of the layer-C partial sums and is therefore a
// property of the device only (eval rollouts stay bit-identical across batch sizes).
// shared memory (floats): per block j: WA[hs][d+1] | WB[hs][h] | WC[d][hs+1] | bA[hs] | bB[hs] | bC[d]
//   (WA and WC rows padded by one float: conflict-free when lanes walk down a column)
//                         then x[R][d] | full[R][h] | outB[R][hs] | part[CS][R][d]
template <bool BWD, int CS, int R>
__global__ void __launch_bounds__(RC_THREADS, 1) rollout_cluster_kernel(const RolloutClusterArgs a, const RolloutWeights w) {
    constexpr int OUTS = 32 / R;      // hidden units per warp task in layer B (OUTS * R = 32 accumulators per lane)
    extern __shared__ float sm[];
    const int T = a.T, B = a.B, d = a.d, h = a.h, nb = a.nb;
    const int hs = h / CS;
    const int rank = (int)cluster_rank();
    const int cluster_id = blockIdx.x / CS;
    const int row0 = cluster_id * R;
    const int nrow = min(R, B - row0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wstride = (hs * (d + 1) + hs * h + d * (hs + 1) + 2 * hs + d + 3) & ~3;      // multiple of 4 floats (float4 reads)
    float* x = sm + (size_t)nb * wstride;
    float* full = x + R * d;
    float* outB = full + R * h;
    float* part = outB + R * hs;
#define hslot(j, which, t) ((((long long)(j) * 2 + (which)) * (T - 1)) + ((t) - 1))

    // ---- resident weights.  Layer A = first layer applied (fwd: W1 [h][d]; bwd: W3^T [h][d]), B = middle, C = last
    for (int j = 0; j < nb; ++j) {
        float* WA = sm + (size_t)j * wstride; float* WB = WA + hs * (d + 1); float* WC = WB + hs * h;
        float* bA = WC + d * (hs + 1); float* bB = bA + hs; float* bC = bB + hs;
        const float* gA = BWD ? w.w3[j] : w.w1[j];
        const float* gB = w.w2[j];
        const float* gC = BWD ? w.w1[j] : w.w3[j];
        for (int i = tid; i < hs * d; i += RC_THREADS) { const int o = i / d, k = i - o * d; WA[o * (d + 1) + k] = gA[(long long)(rank * hs + o) * d + k]; }
        for (int i = tid; i < hs * h; i += RC_THREADS) WB[i] = gB[(long long)rank * hs * h + i];
        for (int i = tid; i < d * hs; i += RC_THREADS) { const int n = i / hs, k = i - n * hs; WC[n * (hs + 1) + k] = gC[(long long)n * h + rank * hs + k]; }
        for (int i = tid; i < hs; i += RC_THREADS) { bA[i] = BWD ? 0.f : w.b1[j][rank * hs + i]; bB[i] = BWD ? 0.f : w.b2[j][rank * hs + i]; }
        for (int i = tid; i < d; i += RC_THREADS) bC[i] = BWD ? 0.f : w.b3[j][i];
    }
    // ---- initial state (every CTA of the cluster keeps a copy)
    for (int i = tid; i < R * d; i += RC_THREADS) {
        const int r = i / d, k = i - r * d;
        x[i] = r < nrow ? a.codes[((long long)(BWD ? T - 1 : 0) * B + row0 + r) * d + k] : 0.f;
    }
    cluster_sync_all();          // also: every CTA of the cluster has started (DSMEM is valid from here on)

    // work items (host checks hs*R <= 2*512, hs/OUTS <= 2*16 warps, d*R <= 512)
    const int nA = hs * R, nBt = hs / OUTS;
    const bool itemC = tid < d * R;
    const int nC = tid % d, rC = tid / d;
    for (int step = 1; step < T; ++step) {
        const int t = BWD ? T - step : step;
        for (int jj = 0; jj < nb; ++jj) {
            const int j = BWD ? nb - 1 - jj : jj;
            const long long slot = (long long)j * (T - 1) + (t - 1);
            float* WA = sm + (size_t)j * wstride; float* WB = WA + hs * (d + 1); float* WC = WB + hs * h;
            float* bA = WC + d * (hs + 1); float* bB = bA + hs; float* bC = bB + hs;
            const long long baseA = hslot(j, BWD ? 1 : 0, t) * B + row0, baseB = hslot(j, BWD ? 0 : 1, t) * B + row0;
            // backward: the saved hiddens that gate this step's gradients are fetched now and consumed after the GEMVs
            float gateA[RC_MAX_ITEMS], gateB[RC_MAX_ITEMS], vA[RC_MAX_ITEMS], vB[RC_MAX_ITEMS];
#pragma unroll
            for (int ii = 0; ii < RC_MAX_ITEMS; ++ii) {
                gateA[ii] = gateB[ii] = 1.f; vA[ii] = vB[ii] = 0.f;
                if (BWD) {
                    const int pA = tid + ii * RC_THREADS, oA = pA % hs, rA = pA / hs;
                    if (pA < nA && rA < nrow) gateA[ii] = a.hidden[(baseA + rA) * h + rank * hs + oA];
                    const int ot = warp + ii * (RC_THREADS / 32), oB = ot * OUTS + lane / R, rB = lane % R;
                    if (ot < nBt && rB < nrow) gateB[ii] = a.hidden[(baseB + rB) * h + rank * hs + oB];
                }
            }
            // ---- layer A: this CTA's hs hidden units for all rows, scattered into every CTA's `full`
#pragma unroll
            for (int ii = 0; ii < RC_MAX_ITEMS; ++ii) {
                const int pA = tid + ii * RC_THREADS, oA = pA % hs, rA = pA / hs;
                if (pA < nA) {
                    float acc = bA[oA];
                    const float* wr = WA + oA * (d + 1);
                    const float* xr = x + rA * d;
                    for (int k = 0; k < d; ++k) acc = fmaf(wr[k], xr[k], acc);
                    vA[ii] = BWD ? ((rA < nrow && gateA[ii] > 0.f) ? acc : 0.f) : fmaxf(acc, 0.f);
                    float* loc = full + rA * h + rank * hs + oA;
#pragma unroll
                    for (int c = 0; c < CS; ++c) dsmem_store(dsmem_addr(loc, (uint32_t)c), vA[ii]);
                }
            }
            cluster_sync_all();
            // global stores are issued after the barrier, next to the longest stretch of arithmetic
#pragma unroll
            for (int ii = 0; ii < RC_MAX_ITEMS; ++ii) {
                const int pA = tid + ii * RC_THREADS, oA = pA % hs, rA = pA / hs;
                if (pA < nA && rA < nrow && a.hidden_out) a.hidden_out[(baseA + rA) * h + rank * hs + oA] = vA[ii];
            }
            if (rank == 0) {
                // fwd: block input (for the W1 gradient);  bwd: gradient of the block's residual output
                float* dst = BWD ? a.res : a.xin;
                if (dst && tid < nrow * d) dst[(slot * B + row0) * d + tid] = x[tid];
            }
            // ---- layer B: warp task = OUTS hidden units x R rows, lanes split k, 32 accumulators per lane
#pragma unroll
            for (int ii = 0; ii < RC_MAX_ITEMS; ++ii) {
                const int ot = warp + ii * (RC_THREADS / 32);
                if (ot < nBt) {
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                    for (int k = lane * 4; k < h; k += 128) {
                        float4 wv[OUTS];
#pragma unroll
                        for (int o = 0; o < OUTS; ++o) wv[o] = *reinterpret_cast<const float4*>(WB + (ot * OUTS + o) * h + k);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float4 xv = *reinterpret_cast<const float4*>(full + r * h + k);
#pragma unroll
                            for (int o = 0; o < OUTS; ++o)
                                acc[o * R + r] += wv[o].x * xv.x + wv[o].y * xv.y + wv[o].z * xv.z + wv[o].w * xv.w;
                        }
                    }
                    const int oB = ot * OUTS + lane / R, rB = lane % R;
                    const float tot = warp_transpose_sum32(acc, lane) + bB[oB];       // lane i: unit ot*OUTS + i/R, row i%R
                    vB[ii] = BWD ? ((rB < nrow && gateB[ii] > 0.f) ? tot : 0.f) : fmaxf(tot, 0.f);
                    outB[rB * hs + oB] = vB[ii];
                }
            }
            __syncthreads();
            // ---- layer C: partial sums over this CTA's slice of the hidden units, sent to every CTA
            if (itemC) {
                float acc = 0.f;
                const float* wr = WC + nC * (hs + 1);
                const float* vr = outB + rC * hs;
                for (int k = 0; k < hs; ++k) acc = fmaf(wr[k], vr[k], acc);
                float* loc = part + (rank * R + rC) * d + nC;
#pragma unroll
                for (int c = 0; c < CS; ++c) dsmem_store(dsmem_addr(loc, (uint32_t)c), acc);
            }
            cluster_sync_all();
#pragma unroll
            for (int ii = 0; ii < RC_MAX_ITEMS; ++ii) {
                const int ot = warp + ii * (RC_THREADS / 32), oB = ot * OUTS + lane / R, rB = lane % R;
                if (ot < nBt && rB < nrow && a.hidden_out) a.hidden_out[(baseB + rB) * h + rank * hs + oB] = vB[ii];
            }
            // ---- residual update, identical in every CTA (fixed summation order over the ranks)
            if (itemC) {
                float rr = bC[nC];
#pragma unroll
                for (int c = 0; c < CS; ++c) rr += part[(c * R + rC) * d + nC];
                if (!BWD && rank == 0 && rC < nrow && a.res) a.res[(slot * B + row0 + rC) * d + nC] = rr;
                x[rC * d + nC] += rr;
            }
            __syncthreads();
        }
        if (BWD) {
            // the decoder's gradient w.r.t. codes[t-1] joins the chain
            if (tid < nrow * d) {
                const long long o = ((long long)(t - 1) * B + row0) * d + tid;
                x[tid] += a.codes[o];
                if (t == 1 && rank == 0) a.codes[o] = x[tid];
            }
            __syncthreads();
        } else if (rank == 0) {
            if (tid < nrow * d) a.codes[((long long)t * B + row0) * d + tid] = x[tid];
        }
    }
    cluster_sync_all();          // no CTA exits while a peer may still write into its shared memory
#undef hslot
}

static size_t rollout_cluster_smem(int d, int h, int nb, int CS, int R) {
    const int hs = h / CS;
    const size_t wstride = ((size_t)hs * (d + 1) + (size_t)hs * h + (size_t)d * (hs + 1) + 2 * hs + d + 3) & ~(size_t)3;
    return (nb * wstride + (size_t)R * d + (size_t)R * h + (size_t)R * hs + (size_t)CS * R * d) * sizeof(float);
}

static bool rollout_cluster_fits(int d, int h, int nb, int CS, int R) {
    const int hs = h / CS, outs = 32 / R;
    if (h % (CS * 4) != 0 || hs < 4 || hs % outs != 0) return false;
    if (hs * R > RC_MAX_ITEMS * RC_THREADS || hs / outs > RC_MAX_ITEMS * (RC_THREADS / 32) || d < 1 || d * R > RC_THREADS) return false;
    return rollout_cluster_smem(d, h, nb, CS, R) <= 220 * 1024;
}

template <bool BWD, int CS, int R>
static int rollout_cluster_max_active(int d, int h, int nb) {
    // clusters of this shape the device can keep resident at once (0: cannot launch)
    const size_t smem = rollout_cluster_smem(d, h, nb, CS, R);
    if (cudaFuncSetAttribute(rollout_cluster_kernel<BWD, CS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (CS > 8 && cudaFuncSetAttribute(rollout_cluster_kernel<BWD, CS, R>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(CS * 64));
    cfg.blockDim = dim3(RC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, rollout_cluster_kernel<BWD, CS, R>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <bool BWD, int CS, int R>
static int launch_rollout_cluster(const RolloutClusterArgs& a, const RolloutWeights& rw, cudaStream_t stream) {
    const size_t smem = rollout_cluster_smem(a.d, a.h, a.nb, CS, R);
    cudaError_t e = cudaFuncSetAttribute(rollout_cluster_kernel<BWD, CS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("rollout_cluster_kernel smem attribute: %s", cudaGetErrorString(e));
    if (CS > 8) {
        e = cudaFuncSetAttribute(rollout_cluster_kernel<BWD, CS, R>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return fail("rollout_cluster_kernel cluster attribute: %s", cudaGetErrorString(e));
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(cdiv(a.B, R) * CS));
    cfg.blockDim = dim3(RC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, rollout_cluster_kernel<BWD, CS, R>, a, rw);
    if (e != cudaSuccess) return fail("rollout_cluster_kernel launch: %s", cudaGetErrorString(e));
    return launched("rollout_cluster_kernel");
}

// returns 0 = launched, -1 = this (d, h, n_blocks) does not fit the cluster kernel, >0 = error.
// Cluster size: 16 if the device can keep at least 8 such clusters resident, else 8 (a device property, never the batch
// size); rows per cluster: 8, or 16 when that avoids a second wave of clusters.
template <bool BWD>
static int rollout_cluster(const RolloutClusterArgs& a, const RolloutWeights& rw, cudaStream_t stream) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("VARSEP_DISABLE_ROLLOUT_CLUSTER"); off = (e && e[0] == '1') ? 1 : 0; }
    if (off) return -1;
    struct Choice { int d, h, nb, cs, max8, max16, dev; };
    static Choice cache = {0, 0, 0, 0, 0, 0, -1};
    const int dev = current_device();          // occupancy is a property of the device the launch goes to
    if (cache.d != a.d || cache.h != a.h || cache.nb != a.nb || cache.cs == 0 || cache.dev != dev) {
        Choice c = {a.d, a.h, a.nb, -1, 0, 0, dev};
        if (rollout_cluster_fits(a.d, a.h, a.nb, 16, 8)) {
            c.max8 = rollout_cluster_max_active<BWD, 16, 8>(a.d, a.h, a.nb);
            if (c.max8 >= 8) {
                c.cs = 16;
                c.max16 = rollout_cluster_fits(a.d, a.h, a.nb, 16, 16) ? rollout_cluster_max_active<BWD, 16, 16>(a.d, a.h, a.nb) : 0;
            }
        }
        if (c.cs < 0 && rollout_cluster_fits(a.d, a.h, a.nb, 8, 8)) {
            c.max8 = rollout_cluster_max_active<BWD, 8, 8>(a.d, a.h, a.nb);
            if (c.max8 >= 1) {
                c.cs = 8;
                c.max16 = rollout_cluster_fits(a.d, a.h, a.nb, 8, 16) ? rollout_cluster_max_active<BWD, 8, 16>(a.d, a.h, a.nb) : 0;
            }
        }
        cache = c;
    }
    if (cache.cs < 0) return -1;
    const long long need8 = cdiv(a.B, 8), need16 = cdiv(a.B, 16);
    const bool rows16 = cache.max16 > 0 && cdiv(need8, cache.max8) > cdiv(need16, cache.max16);
    if (cache.cs == 16) return rows16 ? launch_rollout_cluster<BWD, 16, 16>(a, rw, stream) : launch_rollout_cluster<BWD, 16, 8>(a, rw, stream);
    return rows16 ? launch_rollout_cluster<BWD, 8, 16>(a, rw, stream) : launch_rollout_cluster<BWD, 8, 8>(a, rw, stream);
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
    {
        RolloutClusterArgs a = {codes, nullptr, hidden, xin, res, T, B, d, h, n_blocks};
        const int rc = rollout_cluster<false>(a, rw, as_stream(stream));
        if (rc >= 0) return rc;
    }
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
    {
        RolloutClusterArgs a = {dcodes, hidden, dhidden, nullptr, dres, T, B, d, h, n_blocks};
        const int rc = rollout_cluster<true>(a, rw, as_stream(stream));
        if (rc >= 0) return rc;
    }
    const int smem = (2 * RB * d + 2 * RB * h) * (int)sizeof(float);
    VS_REQUIRE(smem <= 200 * 1024, "latent rollout: hidden size %d not supported", h);
    if (smem > 48 * 1024) cudaFuncSetAttribute(rollout_bwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rollout_bwd_kernel<RB><<<(unsigned)cdiv(B, RB), RO_THREADS, smem, as_stream(stream)>>>(dcodes, rw, T, B, d, h, n_blocks, hidden, dres, dhidden);
    return launched("rollout_bwd_kernel");
}

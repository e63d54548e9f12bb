// HBM-bound "thin" convolutions: layers with a handful of channels on one side, where a 64x64 GEMM
// tile would be >90 % padding.  In the reference these are the first encoder layer
// (Conv 5->64, conv.py:119; AI 61 FLOP/B) and the last decoder layer (ConvTranspose 64->1..3,
// conv.py:263,318; AI 15 FLOP/B) with their gradients.  No tensor cores: the work is streaming the
// wide tensor once with coalesced 16-byte accesses; grids are sized from the pixel count.
#include "common.cuh"

namespace vs {

struct ThinArgs {
    int N, IH, IW, IC, OH, OW, OC, R, S, stride, pad, transposed, act;
};

__device__ __forceinline__ bool thin_src(const ThinArgs& a, int oh, int ow, int r, int s, int& ih, int& iw) {
    if (a.transposed) {
        const int th = oh + a.pad - r, tw = ow + a.pad - s;
        if (th < 0 || tw < 0 || th % a.stride || tw % a.stride) return false;
        ih = th / a.stride; iw = tw / a.stride;
    } else {
        ih = oh * a.stride - a.pad + r; iw = ow * a.stride - a.pad + s;
        if (ih < 0 || iw < 0) return false;
    }
    return ih < a.IH && iw < a.IW;
}

// ---- few OUTPUT channels (OC <= 4), IC % 64 == 0: 8 lanes per output pixel, 8 channels per lane and tap.
// wp: [OC][R*S][IC]
template <typename T, int OC>
__global__ void __launch_bounds__(256) thin_out_kernel(ThinArgs a, const T* __restrict__ in, const T* __restrict__ wp,
                                                       const float* __restrict__ bias, T* __restrict__ out) {
    const long long total = (long long)a.N * a.OH * a.OW;
    const int sub = threadIdx.x & 7;
    // 32 pixels per block iteration; the loop bound is block-uniform so the shuffles below are convergent
    for (long long base = (long long)blockIdx.x * 32; base < total; base += (long long)gridDim.x * 32) {
        const long long pix = base + (threadIdx.x >> 3);
        const bool live = pix < total;
        const int ow = (int)(pix % a.OW);
        const int oh = (int)((pix / a.OW) % a.OH);
        const int n = (int)(pix / ((long long)a.OW * a.OH));
        float acc[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) acc[o] = 0.f;
        // transposed: only the taps congruent to (o + pad) mod stride can contribute
        const int r0 = a.transposed ? (oh + a.pad) % a.stride : 0, s0 = a.transposed ? (ow + a.pad) % a.stride : 0;
        const int rs = a.transposed ? a.stride : 1;
        if (live)
            for (int r = r0; r < a.R; r += rs)
                for (int s = s0; s < a.S; s += rs) {
                    int ih, iw;
                    if (!thin_src(a, oh, ow, r, s, ih, iw)) continue;
                    const T* src = in + (((long long)n * a.IH + ih) * a.IW + iw) * a.IC;
                    const T* w = wp + (long long)(r * a.S + s) * a.IC;
                    for (int c = sub * 8; c < a.IC; c += 64) {
                        const float4 x0 = ld4<T>(src + c), x1 = ld4<T>(src + c + 4);
#pragma unroll
                        for (int o = 0; o < OC; ++o) {
                            const T* wo = w + (long long)o * a.R * a.S * a.IC + c;
                            const float4 w0 = ld4<T>(wo), w1 = ld4<T>(wo + 4);
                            acc[o] += x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w + x1.x * w1.x + x1.y * w1.y +
                                      x1.z * w1.z + x1.w * w1.w;
                        }
                    }
                }
#pragma unroll
        for (int o = 0; o < OC; ++o) {
            acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
            acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
            acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 4);
        }
        if (live && sub == 0) {
#pragma unroll
            for (int o = 0; o < OC; ++o) st<T>(out + pix * OC + o, act_fwd(acc[o] + (bias ? bias[o] : 0.f), a.act));
        }
    }
}

// ---- last DCGAN decoder layer: ConvTranspose k4 s2 p1, 64 -> 1 channel (conv.py:263).  HBM-bound (AI 15 FLOP/B):
// the kernel's job is to read every input pixel ONCE.  A block stages an 8x32 input tile + 1-pixel halo in shared
// memory; 8 lanes share one input position (i,j) - each owns 8 of the 64 channels and keeps its 16x8 filter
// taps in registers - and produce the 2x2 output pixels (2i+a, 2j+b) from the 3x3 neighbourhood (every
// neighbour feeds exactly the parity classes whose tap index r = a + 1 - 2*di is valid), then 3 shuffles.
template <typename T>
__global__ void __launch_bounds__(128) convT_k4s2_to1_kernel(const T* __restrict__ in, const T* __restrict__ wp, const float* __restrict__ bias,
                                                             T* __restrict__ out, int N, int H, int W, int act) {
    constexpr int TH = 8, TW = 32, PH = TH + 2, PW = TW + 2, IC = 64;
    extern __shared__ __align__(16) unsigned char patch_raw[];
    T* patch = reinterpret_cast<T*>(patch_raw);
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h;
    const int n = b / tiles_h;
    const int i0 = th * TH, j0 = tw * TW;
    // stage the patch (zero outside the image): 16-byte chunks, coalesced over channels then columns
    constexpr int CH16 = IC * (int)sizeof(T) / 16, EPC = 16 / (int)sizeof(T);
    for (int q = threadIdx.x; q < PH * PW * CH16; q += 128) {
        const int ch = q % CH16, px = q / CH16;
        const int pi = px / PW, pj = px % PW;
        const int ih = i0 - 1 + pi, iw = j0 - 1 + pj;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ih >= 0 && ih < H && iw >= 0 && iw < W)
            v = *reinterpret_cast<const uint4*>(in + (((long long)n * H + ih) * W + iw) * IC + ch * EPC);
        *reinterpret_cast<uint4*>(patch + (long long)px * IC + ch * EPC) = v;
    }
    // this lane's filter slice: w[tap][8 channels]
    const int sub = threadIdx.x & 7;
    float w[16][8];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const float4 a = ld4<T>(wp + t * IC + sub * 8), c = ld4<T>(wp + t * IC + sub * 8 + 4);
        w[t][0] = a.x; w[t][1] = a.y; w[t][2] = a.z; w[t][3] = a.w; w[t][4] = c.x; w[t][5] = c.y; w[t][6] = c.z; w[t][7] = c.w;
    }
    const float bv = bias ? bias[0] : 0.f;
    __syncthreads();
    const int OW = 2 * W;
    for (int pos = threadIdx.x >> 3; pos < TH * TW; pos += 16) {
        const int li = pos / TW, lj = pos % TW;
        float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                const T* px = patch + ((long long)(li + 1 + di) * PW + (lj + 1 + dj)) * IC + sub * 8;
                const float4 x0 = ld4<T>(px), x1 = ld4<T>(px + 4);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int r = a + 1 - 2 * di;
                    if (r < 0 || r > 3) continue;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int s = c + 1 - 2 * dj;
                        if (s < 0 || s > 3) continue;
                        const float* ww = w[r * 4 + s];
                        acc[a][c] += x0.x * ww[0] + x0.y * ww[1] + x0.z * ww[2] + x0.w * ww[3] + x1.x * ww[4] + x1.y * ww[5] +
                                     x1.z * ww[6] + x1.w * ww[7];
                    }
                }
            }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                acc[a][c] += __shfl_xor_sync(0xffffffffu, acc[a][c], 1);
                acc[a][c] += __shfl_xor_sync(0xffffffffu, acc[a][c], 2);
                acc[a][c] += __shfl_xor_sync(0xffffffffu, acc[a][c], 4);
            }
        const int i = i0 + li, j = j0 + lj;
        if (sub < 2 && i < H && j < W) {        // lane 0 writes output row 2i, lane 1 row 2i+1 (two adjacent pixels each)
            T* dst = out + ((long long)n * 2 * H + 2 * i + sub) * OW + 2 * j;
            st<T>(dst, act_fwd(acc[sub][0] + bv, act));
            st<T>(dst + 1, act_fwd(acc[sub][1] + bv, act));
        }
    }
}

// ---- first DCGAN encoder layer: Conv k4 s2 p1, IC <= 8 input channels -> OC (multiple of 8, <= 128) (conv.py:119).
// A block stages the (2*8+2) x (2*32+2) input patch of an 8x32 output tile and the whole filter in shared
// memory (fp32); each thread owns one output pixel, holds its 16*IC window values in registers and walks over
// the output channels 8 at a time with broadcast LDS.128 weight reads: no global-memory latency in the FMA loop.
template <typename T, int IC>
__global__ void __launch_bounds__(256) conv_k4s2_fewin_kernel(const T* __restrict__ in, const T* __restrict__ wp, const float* __restrict__ bias,
                                                              T* __restrict__ out, int N, int H, int W, int OC, int act) {
    constexpr int TH = 8, TW = 32, PH = 2 * TH + 2, PW = 2 * TW + 2;
    extern __shared__ __align__(16) float fsm[];
    float* wsm = fsm;                              // [16][IC][OC]
    float* patch = fsm + 16 * IC * OC;             // [PH][PW][IC]
    const int P = H / 2, Q = W / 2;
    const int tiles_w = (Q + TW - 1) / TW, tiles_h = (P + TH - 1) / TH;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h;
    const int n = b / tiles_h;
    const int p0 = th * TH, q0 = tw * TW;
    for (int i = threadIdx.x; i < 16 * IC * OC; i += 256) {
        const int oc = i % OC, c = (i / OC) % IC, tap = i / (OC * IC);
        wsm[i] = ld<T>(wp + ((long long)oc * 16 + tap) * IC + c);
    }
    for (int i = threadIdx.x; i < PH * PW * IC; i += 256) {
        const int c = i % IC, px = i / IC;
        const int ih = 2 * p0 - 1 + px / PW, iw = 2 * q0 - 1 + px % PW;
        patch[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? ld<T>(in + (((long long)n * H + ih) * W + iw) * IC + c) : 0.f;
    }
    __syncthreads();
    const int li = threadIdx.x / TW, lj = threadIdx.x % TW;
    const int p = p0 + li, q = q0 + lj;
    float x[16 * IC];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int c = 0; c < IC; ++c) x[(r * 4 + s) * IC + c] = patch[((2 * li + r) * PW + 2 * lj + s) * IC + c];
    if (p >= P || q >= Q) return;
    T* dst = out + (((long long)n * P + p) * Q + q) * OC;
    for (int o8 = 0; o8 < OC; o8 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[o8 + j] : 0.f;
#pragma unroll
        for (int kv = 0; kv < 16 * IC; ++kv) {
            const float4 w0 = *reinterpret_cast<const float4*>(wsm + kv * OC + o8);
            const float4 w1 = *reinterpret_cast<const float4*>(wsm + kv * OC + o8 + 4);
            acc[0] = fmaf(x[kv], w0.x, acc[0]); acc[1] = fmaf(x[kv], w0.y, acc[1]);
            acc[2] = fmaf(x[kv], w0.z, acc[2]); acc[3] = fmaf(x[kv], w0.w, acc[3]);
            acc[4] = fmaf(x[kv], w1.x, acc[4]); acc[5] = fmaf(x[kv], w1.y, acc[5]);
            acc[6] = fmaf(x[kv], w1.z, acc[6]); acc[7] = fmaf(x[kv], w1.w, acc[7]);
        }
        st4<T>(dst + o8, make_float4(act_fwd(acc[0], act), act_fwd(acc[1], act), act_fwd(acc[2], act), act_fwd(acc[3], act)));
        st4<T>(dst + o8 + 4, make_float4(act_fwd(acc[4], act), act_fwd(acc[5], act), act_fwd(acc[6], act), act_fwd(acc[7], act)));
    }
}

template <typename T, int IC>
static void launch_fewin(const ThinArgs& a, const T* in, const T* wp, const float* bias, T* out, cudaStream_t stream) {
    const int smem = (16 * IC * a.OC + 18 * 66 * IC) * (int)sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(conv_k4s2_fewin_kernel<T, IC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const long long blocks = (long long)a.N * cdiv(a.OH, 8) * cdiv(a.OW, 32);
    conv_k4s2_fewin_kernel<T, IC><<<(unsigned)blocks, 256, smem, stream>>>(in, wp, bias, out, a.N, a.IH, a.IW, a.OC, a.act);
}

// ---- few INPUT channels (IC <= 8), OC % 8 == 0: one thread = one output pixel x 8 output channels.
// weights staged in shared memory as fp32 [tap][c][oc]
template <typename T>
__global__ void __launch_bounds__(256) thin_in_kernel(ThinArgs a, const T* __restrict__ in, const T* __restrict__ wp,
                                                      const float* __restrict__ bias, T* __restrict__ out) {
    extern __shared__ float wsm[];
    const int RS = a.R * a.S;
    for (int i = threadIdx.x; i < RS * a.IC * a.OC; i += blockDim.x) {
        const int oc = i % a.OC, c = (i / a.OC) % a.IC, tap = i / (a.OC * a.IC);
        wsm[i] = ld<T>(wp + ((long long)oc * RS + tap) * a.IC + c);
    }
    __syncthreads();
    const unsigned groups = a.OC / 8;
    const unsigned total = (unsigned)a.N * a.OH * a.OW * groups;      // < 2^32, checked by the launcher
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int og = (int)(idx % groups);
        const unsigned pix = idx / groups;
        const int ow = (int)(pix % a.OW);
        const int oh = (int)((pix / a.OW) % a.OH);
        const int n = (int)(pix / ((unsigned)a.OW * a.OH));
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[og * 8 + j] : 0.f;
        for (int r = 0; r < a.R; ++r)
            for (int s = 0; s < a.S; ++s) {
                int ih, iw;
                if (!thin_src(a, oh, ow, r, s, ih, iw)) continue;
                const T* src = in + (((long long)n * a.IH + ih) * a.IW + iw) * a.IC;
                const float* w = wsm + (long long)(r * a.S + s) * a.IC * a.OC + og * 8;
                for (int c = 0; c < a.IC; ++c) {
                    const float x = ld<T>(src + c);
                    const float4 w0 = *reinterpret_cast<const float4*>(w + c * a.OC);
                    const float4 w1 = *reinterpret_cast<const float4*>(w + c * a.OC + 4);
                    acc[0] += x * w0.x; acc[1] += x * w0.y; acc[2] += x * w0.z; acc[3] += x * w0.w;
                    acc[4] += x * w1.x; acc[5] += x * w1.y; acc[6] += x * w1.z; acc[7] += x * w1.w;
                }
            }
        T* dst = out + (long long)pix * a.OC + og * 8;
        st4<T>(dst, make_float4(act_fwd(acc[0], a.act), act_fwd(acc[1], a.act), act_fwd(acc[2], a.act), act_fwd(acc[3], a.act)));
        st4<T>(dst + 4, make_float4(act_fwd(acc[4], a.act), act_fwd(acc[5], a.act), act_fwd(acc[6], a.act), act_fwd(acc[7], a.act)));
    }
}

// ---- few input values per pixel (R*S*IC <= 16, e.g. the dgrad of ConvTranspose 64->1): one thread computes ALL
// output channels of one pixel; the R*S*IC gathered inputs sit in registers, the weights are broadcast LDS.128s.
template <typename T, int KV, int OC>
__global__ void __launch_bounds__(128) thin_in_wide_kernel(ThinArgs a, const T* __restrict__ in, const T* __restrict__ wp,
                                                           const float* __restrict__ bias, T* __restrict__ out) {
    __shared__ __align__(16) float wsm[KV * OC];      // [kv = tap*IC + c][oc]
    const int RS = a.R * a.S;
    for (int i = threadIdx.x; i < KV * OC; i += blockDim.x) {
        const int oc = i % OC, kv = i / OC;
        wsm[i] = kv < RS * a.IC ? ld<T>(wp + (long long)oc * RS * a.IC + kv) : 0.f;
    }
    __syncthreads();
    const unsigned total = (unsigned)a.N * a.OH * a.OW;
    for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
        const int ow = (int)(pix % a.OW);
        const int oh = (int)((pix / a.OW) % a.OH);
        const int n = (int)(pix / ((unsigned)a.OW * a.OH));
        float x[KV];
#pragma unroll
        for (int kv = 0; kv < KV; ++kv) x[kv] = 0.f;
        int kv = 0;
        for (int r = 0; r < a.R; ++r)
            for (int s = 0; s < a.S; ++s) {
                int ih, iw;
                const bool ok = thin_src(a, oh, ow, r, s, ih, iw);
                const T* src = in + (((long long)n * a.IH + ih) * a.IW + iw) * a.IC;
                for (int c = 0; c < a.IC; ++c, ++kv) {
                    const float v = ok ? ld<T>(src + c) : 0.f;
#pragma unroll
                    for (int q = 0; q < KV; ++q) if (q == kv) x[q] = v;      // keeps x[] in registers
                }
            }
        T* dst = out + (long long)pix * OC;
#pragma unroll
        for (int o8 = 0; o8 < OC; o8 += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[o8 + j] : 0.f;
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const float4 w0 = *reinterpret_cast<const float4*>(&wsm[q * OC + o8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&wsm[q * OC + o8 + 4]);
                acc[0] = fmaf(x[q], w0.x, acc[0]); acc[1] = fmaf(x[q], w0.y, acc[1]);
                acc[2] = fmaf(x[q], w0.z, acc[2]); acc[3] = fmaf(x[q], w0.w, acc[3]);
                acc[4] = fmaf(x[q], w1.x, acc[4]); acc[5] = fmaf(x[q], w1.y, acc[5]);
                acc[6] = fmaf(x[q], w1.z, acc[6]); acc[7] = fmaf(x[q], w1.w, acc[7]);
            }
            st4<T>(dst + o8, make_float4(act_fwd(acc[0], a.act), act_fwd(acc[1], a.act), act_fwd(acc[2], a.act), act_fwd(acc[3], a.act)));
            st4<T>(dst + o8 + 4, make_float4(act_fwd(acc[4], a.act), act_fwd(acc[5], a.act), act_fwd(acc[6], a.act), act_fwd(acc[7], a.act)));
        }
    }
}

// ---- weight gradient with few channels on the gathered ("big") side: C <= 5, K % 64 == 0.
// dw[k][c][tap] += sum_pix small[pix][k] * big[src(pix,tap)][c].
// Per 64-pixel tile the R*R*C gathered values of every pixel are staged ONCE in shared memory (the address
// arithmetic and the scattered 2-byte loads are not repeated by the 64 channel lanes); then each thread
// (channel k, pixel lane) streams `small` coalesced and does R*R*C FMAs per pixel against broadcast LDS.128s.
template <typename T, int RR, int CC>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(vs_conv_geom g, const T* __restrict__ small_, const T* __restrict__ big,
                                                         float* __restrict__ dw, long long rows_per_block) {
    constexpr int NV = RR * RR * CC, NVP = (NV + 3) / 4 * 4, TP = 64;
    constexpr int KPT = NVP <= 32 ? 2 : 1;            // channels per thread (register budget: KPT * NVP accumulators)
    constexpr int KT = 64 / KPT, PL = 256 / KT;       // channel threads, pixel lanes
    __shared__ __align__(16) float gs[TP][NVP];
    __shared__ float red[PL][64];
    const int kl = (threadIdx.x % KT) * KPT, pl = threadIdx.x / KT;
    const int k = blockIdx.y * 64 + kl;
    const long long M = (long long)g.N * g.P * g.Q;
    const long long m0 = (long long)blockIdx.x * rows_per_block;
    const long long m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
    const unsigned PQ = (unsigned)g.P * g.Q;
    float acc[KPT][NVP];
#pragma unroll
    for (int j = 0; j < KPT; ++j)
#pragma unroll
        for (int i = 0; i < NVP; ++i) acc[j][i] = 0.f;
    for (long long mt = m0; mt < m1; mt += TP) {
        __syncthreads();
        // stage: TP pixels x R*R taps, one (pixel, tap) pair per thread iteration
#pragma unroll
        for (int i = threadIdx.x; i < TP * RR * RR; i += 256) {
            const int px = i / (RR * RR), tap = i - px * (RR * RR);
            const long long m = mt + px;
            float v[CC];
#pragma unroll
            for (int c = 0; c < CC; ++c) v[c] = 0.f;
            if (m < m1) {
                const unsigned mu = (unsigned)m;
                const int n = (int)(mu / PQ);
                const int rem = (int)(mu - (unsigned)n * PQ);
                const int ih = (rem / g.Q) * g.stride - g.pad + tap / RR, iw = (rem % g.Q) * g.stride - g.pad + tap % RR;
                if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
                    const T* src = big + (((long long)n * g.H + ih) * g.W + iw) * CC;
#pragma unroll
                    for (int c = 0; c < CC; ++c) v[c] = ld<T>(src + c);
                }
            }
#pragma unroll
            for (int c = 0; c < CC; ++c) gs[px][tap * CC + c] = v[c];
        }
        __syncthreads();
        // all loads of this thread's pixels are issued before the first FMA (the loop is latency-bound otherwise)
        float xv[TP / PL][KPT];
#pragma unroll
        for (int q = 0; q < TP / PL; ++q) {
            const long long m = mt + pl + PL * q;
#pragma unroll
            for (int j = 0; j < KPT; ++j) xv[q][j] = m < m1 ? ld<T>(small_ + m * g.K + k + j) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < TP / PL; ++q) {
            const int px = pl + PL * q;
#pragma unroll
            for (int i4 = 0; i4 < NVP / 4; ++i4) {
                const float4 b = *reinterpret_cast<const float4*>(&gs[px][i4 * 4]);
#pragma unroll
                for (int j = 0; j < KPT; ++j) {
                    acc[j][i4 * 4 + 0] = fmaf(xv[q][j], b.x, acc[j][i4 * 4 + 0]);
                    acc[j][i4 * 4 + 1] = fmaf(xv[q][j], b.y, acc[j][i4 * 4 + 1]);
                    acc[j][i4 * 4 + 2] = fmaf(xv[q][j], b.z, acc[j][i4 * 4 + 2]);
                    acc[j][i4 * 4 + 3] = fmaf(xv[q][j], b.w, acc[j][i4 * 4 + 3]);
                }
            }
        }
    }
    // reduce the pixel lanes through shared memory, then one atomic per (k, c, tap)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KPT; ++j) red[pl][kl + j] = acc[j][i];
        __syncthreads();
        if (threadIdx.x < 64) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < PL; ++l) t += red[l][threadIdx.x];
            atomicAdd(&dw[((long long)(blockIdx.y * 64 + threadIdx.x) * CC + (i % CC)) * (RR * RR) + i / CC], t);
        }
    }
}

template <typename T>
__global__ void colstats_any_kernel(const T* __restrict__ y, int C, long long rpg, int chunks, double* __restrict__ stats) {
    __shared__ double r1[8][33], r2[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int g = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
    const long long per = (rpg + chunks - 1) / chunks;
    const long long r0 = (long long)g * rpg + (long long)chunk * per;
    long long r_end = r0 + per;
    if (r_end > (long long)(g + 1) * rpg) r_end = (long long)(g + 1) * rpg;
    double s1 = 0.0, s2 = 0.0;
    if (c < C)
        for (long long r = r0 + threadIdx.y; r < r_end; r += 8) {
            const float v = ld<T>(y + r * C + c);
            s1 += v; s2 += (double)v * v;
        }
    r1[threadIdx.y][threadIdx.x] = s1; r2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t1 += r1[k][threadIdx.x]; t2 += r2[k][threadIdx.x]; }
        atomicAdd(&stats[((long long)g * C + c) * 2], t1);
        atomicAdd(&stats[((long long)g * C + c) * 2 + 1], t2);
    }
}

int column_stats(const void* y, int dtype, long long rows, int C, int G, double* stats, cudaStream_t stream);   // bn.cu

int stats_of_output(const vs_conv_geom* g, int dtype, const void* out, long long rows, int OC, double* stats,
                           cudaStream_t stream) {
    const int fast = column_stats(out, dtype, rows, OC, g->groups, stats, stream);
    if (fast >= 0) return fast;
    const long long rpg = rows / g->groups;
    const int cx = (int)cdiv(OC, 32);
    long long chunks = cdiv(4LL * num_sms(), (long long)cx * g->groups);
    if (chunks > cdiv(rpg, 64)) chunks = cdiv(rpg, 64);
    if (chunks < 1) chunks = 1;
    dim3 grid(cx, (unsigned)(g->groups * chunks)), block(32, 8);
    VS_DISPATCH_DTYPE(dtype, T, (colstats_any_kernel<T><<<grid, block, 0, stream>>>((const T*)out, OC, rpg, (int)chunks, stats)));
    return launched("colstats_any_kernel");
}

int conv_forward_thin_eligible(const vs_conv_geom* g, int mode) {
    const bool tr = mode == VS_CONV_TRANSPOSED;
    const int IC = tr ? g->K : g->C, OC = tr ? g->C : g->K, OH = tr ? g->H : g->P, OW = tr ? g->W : g->Q;
    if (OH * OW < 256) return 0;
    if (OC <= 4 && IC % 64 == 0) return 1;
    if (IC <= 8 && OC % 8 == 0) return 1;
    return 0;
}

// returns 0 = done, -1 = not a thin geometry, >0 error
int conv_forward_thin(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                      double* stats, cudaStream_t stream) {
    ThinArgs a;
    a.N = g->N; a.R = g->R; a.S = g->S; a.stride = g->stride; a.pad = g->pad; a.act = g->act;
    a.transposed = mode == VS_CONV_TRANSPOSED;
    if (!a.transposed) { a.IH = g->H; a.IW = g->W; a.IC = g->C; a.OH = g->P; a.OW = g->Q; a.OC = g->K; }
    else { a.IH = g->P; a.IW = g->Q; a.IC = g->K; a.OH = g->H; a.OW = g->W; a.OC = g->C; }
    const long long pixels = (long long)a.N * a.OH * a.OW;
    // kernel selection must not depend on the batch size: per-sample results have to be bit-identical
    // whatever else is in the batch, so the threshold is on the per-sample pixel count
    if (a.OH * a.OW < 256) return -1;
    int rc;
    if (a.transposed && a.OC == 1 && a.IC == 64 && a.R == 4 && a.S == 4 && a.stride == 2 && a.pad == 1 && a.OH == 2 * a.IH &&
        a.OW == 2 * a.IW) {
        const long long blocks = (long long)a.N * cdiv(a.IH, 8) * cdiv(a.IW, 32);
        VS_REQUIRE(blocks < 2147483647LL, "convT_k4s2_to1: grid too large");
        VS_DISPATCH_DTYPE(g->dtype, T, {
            const int smem = 10 * 34 * 64 * (int)sizeof(T);
            if (smem > 48 * 1024) cudaFuncSetAttribute(convT_k4s2_to1_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            convT_k4s2_to1_kernel<T><<<(unsigned)blocks, 128, smem, stream>>>((const T*)in, (const T*)wp, bias, (T*)out, a.N, a.IH, a.IW, a.act);
        });
        rc = launched("convT_k4s2_to1_kernel");
    } else if (a.OC <= 4 && a.IC % 64 == 0) {
        long long blocks = cdiv(pixels * 8, 256);
        if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
        VS_DISPATCH_DTYPE(g->dtype, T, {
            switch (a.OC) {
                case 1: thin_out_kernel<T, 1><<<(unsigned)blocks, 256, 0, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out); break;
                case 2: thin_out_kernel<T, 2><<<(unsigned)blocks, 256, 0, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out); break;
                case 3: thin_out_kernel<T, 3><<<(unsigned)blocks, 256, 0, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out); break;
                default: thin_out_kernel<T, 4><<<(unsigned)blocks, 256, 0, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out); break;
            }
        });
        rc = launched("thin_out_kernel");
    } else if (!a.transposed && a.IC <= 8 && a.OC % 8 == 0 && a.OC <= 128 && a.R == 4 && a.S == 4 && a.stride == 2 && a.pad == 1 &&
               a.IH == 2 * a.OH && a.IW == 2 * a.OW && (16 * a.IC * a.OC + 18 * 66 * a.IC) * 4 <= 160 * 1024) {
        VS_DISPATCH_DTYPE(g->dtype, T, {
            switch (a.IC) {
                case 1: launch_fewin<T, 1>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
                case 2: launch_fewin<T, 2>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
                case 3: launch_fewin<T, 3>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
                case 4: launch_fewin<T, 4>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
                case 5: launch_fewin<T, 5>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
                case 6: launch_fewin<T, 6>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
                case 7: launch_fewin<T, 7>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
                default: launch_fewin<T, 8>(a, (const T*)in, (const T*)wp, bias, (T*)out, stream); break;
            }
        });
        rc = launched("conv_k4s2_fewin_kernel");
    } else if (a.R * a.S * a.IC <= 16 && a.OC == 64 && pixels < 4000000000LL) {
        long long blocks = cdiv(pixels, 128);
        if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
        VS_DISPATCH_DTYPE(g->dtype, T, (thin_in_wide_kernel<T, 16, 64><<<(unsigned)blocks, 128, 0, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out)));
        rc = launched("thin_in_wide_kernel");
    } else if (a.IC <= 8 && a.OC % 8 == 0 && (long long)a.R * a.S * a.IC * a.OC * 4 <= 96 * 1024 &&
               pixels * (a.OC / 8) < 4000000000LL) {
        const int smem = a.R * a.S * a.IC * a.OC * 4;
        long long blocks = cdiv(pixels * (a.OC / 8), 256);
        if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
        VS_DISPATCH_DTYPE(g->dtype, T, {
            if (smem > 48 * 1024) cudaFuncSetAttribute(thin_in_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            thin_in_kernel<T><<<(unsigned)blocks, 256, smem, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out);
        });
        rc = launched("thin_in_kernel");
    } else {
        return -1;
    }
    if (rc) return rc;
    if (stats != nullptr) rc = stats_of_output(g, g->dtype, out, pixels, a.OC, stats, stream);
    return rc;
}

template <typename T, int RR>
static bool launch_thin_wgrad_c(const vs_conv_geom* g, dim3 grid, const T* small_, const T* big, float* dw, long long rpb,
                                cudaStream_t stream) {
    switch (g->C) {
        case 1: thin_wgrad_kernel<T, RR, 1><<<grid, 256, 0, stream>>>(*g, small_, big, dw, rpb); return true;
        case 2: thin_wgrad_kernel<T, RR, 2><<<grid, 256, 0, stream>>>(*g, small_, big, dw, rpb); return true;
        case 3: thin_wgrad_kernel<T, RR, 3><<<grid, 256, 0, stream>>>(*g, small_, big, dw, rpb); return true;
        case 4: thin_wgrad_kernel<T, RR, 4><<<grid, 256, 0, stream>>>(*g, small_, big, dw, rpb); return true;
        case 5: thin_wgrad_kernel<T, RR, 5><<<grid, 256, 0, stream>>>(*g, small_, big, dw, rpb); return true;
        default: return false;
    }
}

int conv_wgrad_thin(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream) {
    const long long M = (long long)g->N * g->P * g->Q;
    if (g->C > 5 || g->K % 64 != 0 || g->R != g->S || (g->R != 3 && g->R != 4) || g->P * g->Q < 256 || M >= 4000000000LL) return -1;
    long long blocks = 8LL * num_sms() / (g->K / 64);
    if (blocks > cdiv(M, 256)) blocks = cdiv(M, 256);
    if (blocks < 1) blocks = 1;
    const long long rpb = cdiv(cdiv(M, blocks), 64) * 64;
    dim3 grid((unsigned)cdiv(M, rpb), (unsigned)(g->K / 64));
    bool ok = false;
    VS_DISPATCH_DTYPE(g->dtype, T, {
        ok = g->R == 3 ? launch_thin_wgrad_c<T, 3>(g, grid, (const T*)small_, (const T*)big, dw, rpb, stream)
                       : launch_thin_wgrad_c<T, 4>(g, grid, (const T*)small_, (const T*)big, dw, rpb, stream);
    });
    if (!ok) return -1;
    return launched("thin_wgrad_kernel");
}

}  // namespace vs

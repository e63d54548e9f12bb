// tcgen05 convolutions for layers with a handful of channels on the image side (bf16 operands, fp32 accumulation).
//
// The first encoder convolution (nt_cond*nc = 5 input channels, /root/reference/var_sep/networks/conv.py:119-123) and
// the backward of the last decoder up-convolution (nc = 1 output channel, conv.py:258-263) have R*S*C <= 128: one
// 128-byte row of shared memory holds a pixel's whole receptive field.  A TMA box cannot express that gather
// (C*2 bytes is not a multiple of 16), so the CTA builds the im2col tile itself, directly in the 128-byte-swizzled
// layout tcgen05.mma reads, and the image-side tensor is streamed from HBM exactly once:
//   forward : out[pix][k]      = sum_kk xcol[pix][kk] * w[k][kk]          (K-major A built by 4 warps, B resident)
//   wgrad   : dw[k][c][tap]   += sum_pix xcol[pix][(tap,c)] * small[pix][k] (MN-major A built, B = TMA boxes of `small`)
// with kk = (r*S + s)*C + c and xcol[pix][kk] = big[n][p*stride-pad+r][q*stride-pad+s][c] (zero outside the image).
// The shared-memory image of a tile is the same in both cases: row = pixel, 128 bytes = 64 kk values, the 16-byte
// chunk index XORed with (row & 7).  The input patch a pixel box needs ((HT-1)*stride+R rows of ((WT-1)*stride+S)*C
// contiguous values) is itself a TMA box of the image viewed as [N][H][W*C] (zero-filled outside), prefetched several
// tiles ahead, so the builders never wait on global memory.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace vs {

constexpr int IM_MAX_KK = 128;

struct Im2colParams {
    int N, H, W, C, P, Q, K, R, S, stride, pad;
    int KK, nchunk16;            // R*S*C and the number of 16-byte chunks per pixel row that hold data
    int WT, HT, NT;              // pixel box over the small grid (powers of two)
    int wt_shift, ht_shift;
    int tiles_w, tiles_h, tiles_n, total_tiles;
    int ptiles_per_split;        // wgrad
    int PH, PWCp, patch_bytes;   // input patch of one pixel box: rows, row pitch in elements (multiple of 8), bytes
    int patch_stage;             // bytes between patch stages (patch_bytes rounded up to 128)
    int act, has_bias, partial, n_per_group;
};

constexpr int IM_PATCH_STAGE = 8192, IM_PATCH_STAGE_W = 4096, IM_PST = 3;   // bytes per patch stage (forward, wgrad), stages

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// one 16-byte chunk (8 consecutive kk) of the receptive field whose top-left element sits at pbuf[base]
__device__ __forceinline__ uint4 im2col_chunk(const Im2colParams& p, const unsigned short* pbuf, const uint32_t* tab, int base, int j) {
    unsigned short v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int kk = j * 8 + e;
        v[e] = kk < p.KK ? pbuf[base + (int)tab[kk]] : (unsigned short)0;
    }
    return make_uint4(v[0] | ((uint32_t)v[1] << 16), v[2] | ((uint32_t)v[3] << 16), v[4] | ((uint32_t)v[5] << 16),
                      v[6] | ((uint32_t)v[7] << 16));
}

// The patch offsets of a thread's chunks never change from tile to tile (same pixel of the box, same kk): threads that
// own at most two chunks keep them in registers, which cuts the builder from ~300 to ~40 instructions per tile.
struct ChunkOffsets { int off[16]; uint32_t valid; };
__device__ __forceinline__ void chunk_offsets(const Im2colParams& p, const uint32_t* tab, int base, int j0, int nch, ChunkOffsets& co) {
    co.valid = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int kk = (j0 + (i >> 3)) * 8 + (i & 7);
        const bool ok = (i >> 3) < nch && kk < p.KK;
        co.off[i] = ok ? base + (int)tab[kk] : 0;
        co.valid |= ok ? (1u << i) : 0u;
    }
}
// chunk c (0 or 1) of the thread's two cached chunks
__device__ __forceinline__ uint4 gather_cached(const unsigned short* pb, const ChunkOffsets& co, int c) {
    unsigned short v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const unsigned short x = pb[co.off[c * 8 + e]];
        v[e] = (co.valid >> (c * 8 + e)) & 1u ? x : (unsigned short)0;
    }
    return make_uint4(v[0] | ((uint32_t)v[1] << 16), v[2] | ((uint32_t)v[3] << 16), v[4] | ((uint32_t)v[5] << 16),
                      v[6] | ((uint32_t)v[7] << 16));
}

// tab[kk] = offset of (r, s, c) inside the patch relative to the receptive field's top-left element
__device__ __forceinline__ void im2col_table(const Im2colParams& p, uint32_t* tab) {
    for (int kk = threadIdx.x; kk < IM_MAX_KK; kk += blockDim.x) {
        uint32_t t = 0;
        if (kk < p.KK) {
            const int tap = kk / p.C, c = kk - tap * p.C;
            t = (uint32_t)((tap / p.S) * p.PWCp + (tap % p.S) * p.C + c);
        }
        tab[kk] = t;
    }
}
// =================================================================================================== forward
// warps: 0 = spare (weights are staged by everyone at start), 1 = TMEM allocator + MMA issuer, 2..5 = im2col builders,
// 6..9 = epilogue.  Persistent over 128-pixel tiles; KC = number of 64-wide kk chunks (1 or 2).
constexpr int IMF_THREADS = 320;

template <int BN, int KC, int STAGES, bool TS>
struct ImfSmem {
    static constexpr int A_BYTES = KC * TC_BM * 128, B_BYTES = KC * BN * 128;
    static constexpr int OUT_OFF = STAGES * A_BYTES + B_BYTES;
    static constexpr int OUT_BYTES = TS ? 2 * TC_BM * 128 : 0;       // two staged bf16 output tiles (BN = 64)
    static constexpr int BAR_OFF = OUT_OFF + OUT_BYTES;
    static constexpr int TAB_OFF = BAR_OFF + 256, STAT_OFF = TAB_OFF + IM_MAX_KK * 4;
    static constexpr int PRM_OFF = (STAT_OFF + BN * 2 * 4 + 15) / 16 * 16;             // [BN][8] floats (EPI != 0)
    static constexpr int PATCH_OFF = (PRM_OFF + BN * 8 * 4 + 127) / 128 * 128;
    static constexpr int total(int patch_stage) { return PATCH_OFF + IM_PST * patch_stage + 1024; }
    static_assert((2 * STAGES + 2 * IM_PST + 5) * 8 <= 256, "barrier area");
    static_assert(!TS || BN == 64, "the staged epilogue handles one 128-byte row per pixel");
};

// EPI selects what the epilogue does with the accumulator tile D[pixel][k]:
//   0  out = act(D + bias)  (+ BatchNorm statistics of D + bias)                                    -- a convolution
//   1  D is the gradient w.r.t. the OUTPUT of a BatchNorm + activation block whose pre-BatchNorm tensor is bb.y:
//      dz = D * act'(gamma * xhat + beta),  stats[g][k] += {sum dz, sum dz * xhat}; nothing is stored
//   2  same dz, then  out = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat))  (the BatchNorm backward)
// Modes 1 and 2 are the backward of the fused decoder tail: the gradient w.r.t. the normalised tensor (268 MB for the
// Moving-MNIST decoder) is recomputed from the 16x smaller frame gradient in both passes instead of being written
// once and read twice.
template <int BN, int KC, int STAGES, bool TS, int EPI>
__global__ void __launch_bounds__(IMF_THREADS, 2) im2col_fwd_kernel(const __grid_constant__ CUtensorMap map_big,
                                                                   const __grid_constant__ CUtensorMap map_out,
                                                                   const __grid_constant__ Im2colParams p,
                                                                   const __grid_constant__ BnBwdArgs bb,
                                                                   const unsigned short* __restrict__ wp,
                                                                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                                   double* __restrict__ stats) {
    static_assert(EPI == 0 || (BN == 64 && KC == 1), "the BatchNorm-backward epilogues handle one 64-channel tile");
    static_assert(EPI != 1 || !TS, "the reduction pass stores nothing");
    static_assert(EPI != 2 || TS, "the apply pass stores through the staged tile");
    using S = ImfSmem<BN, KC, STAGES, TS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* b_smem = smem + STAGES * S::A_BYTES;
    uint8_t* obuf = smem + S::OUT_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint64_t* pready = bars + 2 * STAGES + 4;        // [IM_PST] patch landed
    uint64_t* pempty = pready + IM_PST;              // [IM_PST] patch consumed by the four builder warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + IM_PST);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + S::TAB_OFF);
    float* sstat = reinterpret_cast<float*>(smem + S::STAT_OFF);
    float4* sprm = reinterpret_cast<float4*>(smem + S::PRM_OFF);      // EPI != 0: per channel {invstd, -mean*invstd, gamma, beta}, {gamma*invstd, mean dz, mean dz*xhat, 0}
    uint8_t* pbuf = smem + S::PATCH_OFF;             // [IM_PST][p.patch_stage]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // one contiguous range of tiles per CTA (few BatchNorm groups per CTA: statistics are flushed once per group)
    const int t_begin = (int)((long long)blockIdx.x * p.total_tiles / gridDim.x);
    const int t_end = (int)((long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x);

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_big);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        for (int s = 0; s < IM_PST; ++s) { mbar_init(&pready[s], 1); mbar_init(&pempty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
    im2col_table(p, tab);
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) sstat[i] = 0.f;
    // A stages: zero once (chunks past nchunk16 are never written again); B: the packed weights [K][KK], zero padded
    for (int i = threadIdx.x; i < STAGES * S::A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < BN * KC * 8; i += blockDim.x) {
        const int oc = i / (KC * 8), jj = i % (KC * 8);            // jj: 16-byte chunk along kk
        unsigned short v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kk = jj * 8 + e;
            v[e] = (oc < p.K && kk < p.KK) ? __ldg(wp + (long long)oc * p.KK + kk) : (unsigned short)0;
        }
        const int chunk = jj >> 3, j = jj & 7;
        *reinterpret_cast<uint4*>(b_smem + chunk * (BN * 128) + oc * 128 + ((j ^ (oc & 7)) << 4)) =
            make_uint4(v[0] | ((uint32_t)v[1] << 16), v[2] | ((uint32_t)v[3] << 16), v[4] | ((uint32_t)v[5] << 16),
                       v[6] | ((uint32_t)v[7] << 16));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

#define VS_IM_DECODE(idx)                                       \
    int t_ = (idx);                                             \
    const int tw = t_ % p.tiles_w; t_ /= p.tiles_w;             \
    const int th = t_ % p.tiles_h;                              \
    const int tn = t_ / p.tiles_h;                              \
    const int j0 = tw * p.WT, i0 = th * p.HT, b0 = tn * p.NT;

    if (warp == 0) {
        // ===== TMA producer: the input patch of every tile, IM_PST tiles ahead =====
        if (lane == 0) {
            int lt = 0;
            for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
                VS_IM_DECODE(idx)
                const int ps = lt % IM_PST;
                mbar_wait(&pempty[ps], ((lt / IM_PST) & 1) ^ 1);
                mbar_expect_tx(&pready[ps], (uint32_t)p.patch_bytes);
                tma_load_3d(pbuf + ps * p.patch_stage, &map_big, &pready[ps], ((j0 * p.stride - p.pad) * p.C) & ~7, i0 * p.stride - p.pad, b0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16_f32(BN);
            const int ksteps = (p.KK + 15) >> 4;
            int lt = 0;
            for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
                const int acc = lt & 1, s = lt % STAGES;
                mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
                mbar_wait(&full[s], (lt / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * S::A_BYTES), b_addr = smem_u32(b_smem);
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t ao = (uint32_t)((k >> 2) * (TC_BM * 128) + (k & 3) * 32);
                    const uint32_t bo = (uint32_t)((k >> 2) * (BN * 128) + (k & 3) * 32);
                    umma_bf16(tmem_d, kmajor_sw128_desc(a_addr + ao), kmajor_sw128_desc(b_addr + bo), idesc, k != 0 ? 1u : 0u);
                }
                umma_commit(&empty[s]);
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp >= 2 && warp < 6) {
        // ===== im2col builders: thread = pixel row of the tile =====
        const int m = threadIdx.x - 64;
        const int w = m & (p.WT - 1), h = (m >> p.wt_shift) & (p.HT - 1);
        const int base = h * p.stride * p.PWCp + w * p.stride * p.C;
        const bool cached = p.nchunk16 <= 2;
        ChunkOffsets co;
        chunk_offsets(p, tab, base, 0, p.nchunk16, co);
        int lt = 0;
        for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
            const int s = lt % STAGES, ps = lt % IM_PST;
            const int j0 = (idx % p.tiles_w) * p.WT;
            // the box starts at the 16-byte boundary below the patch's first element
            const unsigned short* pb = reinterpret_cast<const unsigned short*>(pbuf + ps * p.patch_stage) + (((j0 * p.stride - p.pad) * p.C) & 7);
            uint8_t* a_row = smem + s * S::A_BYTES + m * 128;
            mbar_wait(&pready[ps], (lt / IM_PST) & 1);
            mbar_wait(&empty[s], ((lt / STAGES) & 1) ^ 1);
            if (cached) {
                *reinterpret_cast<uint4*>(a_row + ((0 ^ (m & 7)) << 4)) = gather_cached(pb, co, 0);
                if (p.nchunk16 == 2) *reinterpret_cast<uint4*>(a_row + ((1 ^ (m & 7)) << 4)) = gather_cached(pb, co, 1);
            } else {
                for (int j = 0; j < p.nchunk16; ++j) {
                    const uint4 v = im2col_chunk(p, pb, tab, base, j);
                    *reinterpret_cast<uint4*>(a_row + (j >> 3) * (TC_BM * 128) + (((j & 7) ^ (m & 7)) << 4)) = v;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&full[s]); mbar_arrive(&pempty[ps]); }
        }
    } else if (warp >= 6) {
        // ===== epilogue: one warp per TMEM lane quarter =====
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int w = m & (p.WT - 1), h = (m >> p.wt_shift) & (p.HT - 1), n = m >> (p.wt_shift + p.ht_shift);
        int lt = 0;
        int stat_g = -1;
        int prm_g = -1;
        double stat_acc[2 * BN / 128] = {};
        for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
            VS_IM_DECODE(idx)
            const int acc = lt & 1;
            const int pp = i0 + h, qq = j0 + w, nn = b0 + n;
            const bool ok = pp < p.P && qq < p.Q && nn < p.N;
            __nv_bfloat16* dst = out + (((long long)nn * p.P + pp) * p.Q + qq) * p.K;
            uint4 yraw[EPI != 0 ? BN / 8 : 1];
            if (EPI != 0) {
                // this pixel's row of the pre-BatchNorm tensor (128 contiguous bytes), requested before the accumulator wait
                const uint4* yrow = reinterpret_cast<const uint4*>(bb.y + (((long long)nn * p.P + pp) * p.Q + qq) * p.K);
#pragma unroll
                for (int v = 0; v < BN / 8; ++v) yraw[v] = ok ? __ldg(yrow + v) : make_uint4(0, 0, 0, 0);
                // per-channel constants of this tile's BatchNorm group (a tile is one image: NT == 1)
                const int g = b0 / bb.n_per_group;
                if (g != prm_g) {                                       // uniform over the epilogue warps
                    asm volatile("bar.sync 1, 128;" ::: "memory");      // everybody is done with the previous table
                    const int c = threadIdx.x - 192;
                    if (c < BN && c < p.K) {
                        const float is = __ldg(bb.invstd + (long long)g * p.K + c), mu = __ldg(bb.mean + (long long)g * p.K + c);
                        const float ga = __ldg(bb.gamma + c), be = __ldg(bb.beta + c);
                        float m1 = 0.f, m2 = 0.f;
                        if (EPI == 2 && bb.train) {
                            m1 = (float)bb.sums[((long long)g * p.K + c) * 2] * bb.inv_count;
                            m2 = (float)bb.sums[((long long)g * p.K + c) * 2 + 1] * bb.inv_count;
                        }
                        sprm[2 * c] = make_float4(is, -mu * is, ga, be);
                        sprm[2 * c + 1] = make_float4(ga * is, m1, m2, 0.f);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    prm_g = g;
                }
            }
            mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
            tc_fence_after();
            // (fully unrolled for the BatchNorm-backward epilogues: the y row then stays in registers)
            constexpr int C0_UNROLL = EPI != 0 ? BN / 32 : 1;
#pragma unroll C0_UNROLL
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                float xs[32];
                float xh[EPI == 1 ? 32 : 1];          // EPI 1: dz * xhat
                if (EPI != 0) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const uint4 yq = yraw[(c0 + c) >> 3];
                        const uint32_t wv = ((c >> 1) & 3) == 0 ? yq.x : ((c >> 1) & 3) == 1 ? yq.y : ((c >> 1) & 3) == 2 ? yq.z : yq.w;
                        const float yv = (c & 1) ? __uint_as_float(wv & 0xffff0000u) : __uint_as_float(wv << 16);
                        const float4 pa = sprm[2 * (c0 + c)];
                        const float xhat = fmaf(yv, pa.x, pa.y);
                        const float dz = fmaf(pa.z, xhat, pa.w) > 0.f ? __uint_as_float(r[c]) : __uint_as_float(r[c]) * bb.neg_slope;
                        if (EPI == 2) {
                            const float4 pb = sprm[2 * (c0 + c) + 1];
                            xs[c] = pb.x * (dz - pb.y - xhat * pb.z);
                        } else {
                            xs[c] = ok ? dz : 0.f;
                            xh[c] = ok ? dz * xhat : 0.f;
                        }
                    }
                } else if (p.has_bias) {          // uniform: dgrad launches carry no bias and skip the 32 shuffles
                    const float bias_l = (p.has_bias && c0 + lane < p.K) ? __ldg(bias + c0 + lane) : 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]) + __shfl_sync(0xffffffffu, bias_l, c);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]);
                }
                if (EPI == 1) {
                    // reduction pass: nothing is stored
                } else if (TS) {
                    stage_row32(obuf + (lt & 1) * (TC_BM * 128), m, c0, xs, EPI == 2 ? (int)VS_ACT_NONE : p.act);
                } else if (p.partial || BN > p.K) {
                    if (ok) {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (c0 + c < p.K) dst[c0 + c] = __float2bfloat16_rn(act_fwd(xs[c], p.act));
                    }
                } else if (ok) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = v * 8 + e * 2;
                            __nv_bfloat162 b2 = p.act == VS_ACT_NONE ? __floats2bfloat162_rn(xs[c], xs[c + 1])
                                                                     : __floats2bfloat162_rn(act_fwd(xs[c], p.act), act_fwd(xs[c + 1], p.act));
                            pk[e] = *reinterpret_cast<uint32_t*>(&b2);
                        }
                        *reinterpret_cast<uint4*>(dst + c0 + v * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                if (EPI == 1) {
                    // {sum dz, sum dz * xhat} over the 32 rows of this warp (masked rows hold zeros), in place
                    const float s1 = warp_transpose_sum32(xs, lane);
                    const float s2 = warp_transpose_sum32(xh, lane);
                    atomicAdd(&sstat[(c0 + lane) * 2], s1);
                    atomicAdd(&sstat[(c0 + lane) * 2 + 1], s2);
                } else if (EPI == 0 && stats != nullptr) {
                    float wk[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) wk[c] = ok ? xs[c] : 0.f;
                    const float s1 = warp_transpose_sum32(wk, lane);
#pragma unroll
                    for (int c = 0; c < 32; ++c) wk[c] = ok ? xs[c] * xs[c] : 0.f;
                    const float s2 = warp_transpose_sum32(wk, lane);
                    atomicAdd(&sstat[(c0 + lane) * 2], s1);
                    atomicAdd(&sstat[(c0 + lane) * 2 + 1], s2);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (TS) {
                // staged tile complete -> one TMA store (clipped at the tensor edges); the previous tile's store is drained
                // first so that its buffer can be refilled by the next tile
                fence_proxy_async();
                if (threadIdx.x == 192) tma_store_wait_read();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 192) {
                    tma_store_4d(&map_out, obuf + (lt & 1) * (TC_BM * 128), 0, j0, i0, b0);
                    tma_store_commit();
                }
            }
            if (EPI != 2 && stats != nullptr) {
                // running fp64 sums per thread (entries t and t + 128), flushed when the BatchNorm group changes
                if (!TS) asm volatile("bar.sync 1, 128;" ::: "memory");
                const int g = b0 / p.n_per_group;
                const int t = threadIdx.x - 192;
                if (g != stat_g) {
                    if (stat_g >= 0) {
#pragma unroll
                        for (int e = 0; e < 2 * BN / 128; ++e) {
                            const int i = t + e * 128, col = i >> 1;
                            if (col < p.K) atomicAdd(&stats[((long long)stat_g * p.K + col) * 2 + (i & 1)], stat_acc[e]);
                            stat_acc[e] = 0.0;
                        }
                    }
                    stat_g = g;
                }
#pragma unroll
                for (int e = 0; e < 2 * BN / 128; ++e) { stat_acc[e] += (double)sstat[t + e * 128]; sstat[t + e * 128] = 0.f; }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        if (EPI != 2 && stats != nullptr && stat_g >= 0) {
            const int t = threadIdx.x - 192;
#pragma unroll
            for (int e = 0; e < 2 * BN / 128; ++e) {
                const int i = t + e * 128, col = i >> 1;
                if (col < p.K) atomicAdd(&stats[((long long)stat_g * p.K + col) * 2 + (i & 1)], stat_acc[e]);
            }
        }
        if (TS && threadIdx.x == 192) tma_store_wait_all();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

// geometry, pixel box and input patch of one tile; false if the layer does not fit this kernel family
static bool im2col_plan(Im2colParams& p, const vs_conv_geom* g, int pixels, int max_patch_bytes) {
    memset(&p, 0, sizeof(p));
    p.N = g->N; p.H = g->H; p.W = g->W; p.C = g->C; p.P = g->P; p.Q = g->Q; p.K = g->K; p.R = g->R; p.S = g->S;
    p.stride = g->stride; p.pad = g->pad;
    p.KK = g->R * g->S * g->C;
    p.nchunk16 = (p.KK + 7) / 8;
    if (g->C >= 32 || p.KK > IM_MAX_KK || g->R > 16 || g->S > 16) return false;
    if (g->P * g->Q < 256) return false;       // per-image size only: the choice of kernel must not depend on the batch
    if (((long long)g->W * g->C * 2) % 16 != 0) return false;      // TMA row pitch of the [N][H][W*C] view
    // widest pixel box whose patch rows still fit one TMA box (<= 256 elements)
    p.WT = pow2ceil(p.Q) < pixels ? pow2ceil(p.Q) : pixels;
    // (the box starts at a multiple of 8 elements -- TMA needs a 16-byte aligned innermost coordinate -- hence +7)
    while (p.WT > 1 && (((p.WT - 1) * p.stride + p.S) * p.C + 7 + 7) / 8 * 8 > 256) p.WT >>= 1;
    p.HT = pow2ceil(p.P) < pixels / p.WT ? pow2ceil(p.P) : pixels / p.WT;
    p.NT = pixels / (p.WT * p.HT);
    if (p.NT != 1) return false;
    p.wt_shift = 0; while ((1 << p.wt_shift) < p.WT) ++p.wt_shift;
    p.ht_shift = 0; while ((1 << p.ht_shift) < p.HT) ++p.ht_shift;
    p.tiles_w = (int)cdiv(p.Q, p.WT); p.tiles_h = (int)cdiv(p.P, p.HT); p.tiles_n = p.N;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.PH = (p.HT - 1) * p.stride + p.R;
    p.PWCp = (((p.WT - 1) * p.stride + p.S) * p.C + 7 + 7) / 8 * 8;
    p.patch_bytes = p.PH * p.PWCp * 2;
    p.patch_stage = (p.patch_bytes + 127) / 128 * 128;
    return p.PH <= 256 && p.PWCp <= 256 && p.patch_bytes <= max_patch_bytes;
}

static int im2col_big_map(CUtensorMap& mb, const Im2colParams& p, const void* big) {
    EncodeTiledFn enc = encode_fn();
    cuuint64_t dims[3] = {(cuuint64_t)p.W * p.C, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t strides[2] = {(cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.H * p.W * p.C * 2};
    cuuint32_t box[3] = {(cuuint32_t)p.PWCp, (cuuint32_t)p.PH, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(big), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(im2col patch) failed: %d", (int)rc);
    return 0;
}

static bool im2col_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_DISABLE_IM2COL"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1 || tc_disabled();
}

int conv_forward_im2col_eligible(const vs_conv_geom* g, int mode) {
    if (g->dtype != VS_BF16 || (g->flags & VS_FLAG_FORCE_SIMT) || im2col_disabled()) return 0;
    if (mode != VS_CONV_DIRECT || g->K > 128 || g->K < 16) return 0;
    Im2colParams p;
    return im2col_plan(p, g, 128, IM_PATCH_STAGE) ? 1 : 0;
}

template <int BN, int KC, int STAGES, bool TS, int EPI = 0>
static int launch_imf(const CUtensorMap& mb, const CUtensorMap& mo, const Im2colParams& p, const void* wp, const float* bias,
                      void* out, double* stats, cudaStream_t stream, const BnBwdArgs* bb = nullptr) {
    using S = ImfSmem<BN, KC, STAGES, TS>;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(im2col_fwd_kernel<BN, KC, STAGES, TS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             S::total(IM_PATCH_STAGE));
        if (e != cudaSuccess) return fail("im2col_fwd_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    BnBwdArgs none;
    memset(&none, 0, sizeof(none));
    const int resident = 2 * num_sms();
    const int grid = p.total_tiles < resident ? p.total_tiles : resident;
    im2col_fwd_kernel<BN, KC, STAGES, TS, EPI><<<grid, IMF_THREADS, S::total(p.patch_stage), stream>>>(
        mb, mo, p, bb ? *bb : none, (const unsigned short*)wp, bias, (__nv_bfloat16*)out, stats);
    return launched("im2col_fwd_kernel");
}

// returns 0 = done, -1 = geometry not eligible, >0 = error
int conv_forward_im2col(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                        double* stats, cudaStream_t stream) {
    if (!conv_forward_im2col_eligible(g, mode)) return -1;
    if (reinterpret_cast<uintptr_t>(in) & 15) return -1;
    if (!encode_fn()) return -1;
    Im2colParams p;
    if (!im2col_plan(p, g, 128, IM_PATCH_STAGE)) return -1;
    CUtensorMap mb;
    if (int rc = im2col_big_map(mb, p, in)) return rc;
    p.act = g->act; p.has_bias = bias != nullptr;
    p.partial = (g->K % 8 != 0) || (reinterpret_cast<uintptr_t>(out) & 15) ? 1 : 0;
    p.n_per_group = g->N / g->groups;
    const bool fuse_stats = stats != nullptr && (p.n_per_group % p.NT) == 0;
    double* st = fuse_stats ? stats : nullptr;
    const int KC = p.KK > 64 ? 2 : 1;
    // staged epilogue + TMA store for one 64-channel tile with 16-byte aligned rows and a single kk chunk
    CUtensorMap mo;
    memset(&mo, 0, sizeof(mo));
    const bool staged = g->K <= 64 && KC == 1 && !p.partial;
    if (staged) {
        cuuint64_t dims[4] = {(cuuint64_t)g->K, (cuuint64_t)g->Q, (cuuint64_t)g->P, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)g->K * 2, (cuuint64_t)g->K * g->Q * 2, (cuuint64_t)g->K * g->Q * g->P * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.WT, (cuuint32_t)p.HT, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode_fn()(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(im2col out) failed: %d", (int)r);
    }
    int rc;
    if (staged)          rc = launch_imf<64, 1, 3, true>(mb, mo, p, wp, bias, out, st, stream);
    else if (g->K <= 64) rc = KC == 1 ? launch_imf<64, 1, 4, false>(mb, mo, p, wp, bias, out, st, stream)
                                      : launch_imf<64, 2, 2, false>(mb, mo, p, wp, bias, out, st, stream);
    else                 rc = KC == 1 ? launch_imf<128, 1, 4, false>(mb, mo, p, wp, bias, out, st, stream)
                                      : launch_imf<128, 2, 2, false>(mb, mo, p, wp, bias, out, st, stream);
    if (rc) return rc;
    if (stats != nullptr && !fuse_stats) return stats_of_output(g, g->dtype, out, (long long)g->N * g->P * g->Q, g->K, stats, stream);
    return 0;
}

// =================================================================================================== weight gradient
// D[kk][k] += sum over 64-pixel blocks; warps: 0 = TMA producer (boxes of `small`), 1 = MMA issuer, 2..5 = im2col
// builders during the reduction, then the epilogue (fp32 reductions into the torch-layout gradient).
// XF: `small` is the PRE-BatchNorm tensor y of the layer below; the builder warps apply BatchNorm + activation to every
// landed box of it in shared memory (one fused multiply-add + max per element) before the MMA reads it, so the
// weight gradient of the fused decoder tail needs no materialised normalised tensor.
constexpr int IMW_THREADS = 192, IMW_XF_THREADS = 128;
template <int BN, int STAGES, bool XF>
__global__ void __launch_bounds__(IMW_THREADS + (XF ? IMW_XF_THREADS : 0), 2) im2col_wgrad_kernel(const __grid_constant__ CUtensorMap map_small,
                                                              const __grid_constant__ CUtensorMap map_big,
                                                              const __grid_constant__ Im2colParams p,
                                                              const __grid_constant__ BnApplyArgs bn, float* __restrict__ dw) {
    static_assert(!XF || BN == 64, "operand transform: one 64-channel box per stage");
    constexpr int PIX = 64, CHUNK = 64 * PIX * 2;
    constexpr int A_BYTES = 2 * CHUNK, B_BYTES = BN * PIX * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* pready = bars + 2 * STAGES + 1;
    uint64_t* pempty = pready + IM_PST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + IM_PST);
    uint64_t* bfull = pempty + IM_PST + 1;           // [STAGES] XF: the box of `small` has landed (transform may start)
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + 256);
    uint8_t* pbuf = smem + STAGES * STAGE_BYTES + 256 + IM_MAX_KK * 4;       // [IM_PST][IM_PATCH_STAGE_W], 128-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x;
    const int pt0 = blockIdx.z * p.ptiles_per_split;
    int pt1 = pt0 + p.ptiles_per_split;
    if (pt1 > p.total_tiles) pt1 = p.total_tiles;
    const int nkb = pt1 - pt0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_small);
        prefetch_tmap(&map_big);
        // full: TMA expect + 4 builder warps (XF: 4 builder warps + 4 transform warps, the latter after the box has landed)
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], XF ? 8 : 5); mbar_init(&empty[i], 1); if (XF) mbar_init(&bfull[i], 1); }
        for (int i = 0; i < IM_PST; ++i) { mbar_init(&pready[i], 1); mbar_init(&pempty[i], 4); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    im2col_table(p, tab);
    for (int s = 0; s < STAGES; ++s)
        for (int i = threadIdx.x; i < A_BYTES / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(smem + s * STAGE_BYTES)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (nkb <= 0) {
        __syncthreads();
        if (warp == 1) tmem_dealloc<BN>(tmem_base);
        return;
    }

#define VS_IMW_DECODE(t)                                        \
    int t_ = (t);                                               \
    const int tw = t_ % p.tiles_w; t_ /= p.tiles_w;             \
    const int th = t_ % p.tiles_h;                              \
    const int tn = t_ / p.tiles_h;                              \
    const int q0 = tw * p.WT, p0 = th * p.HT, b0 = tn * p.NT;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES, ps = kb % IM_PST;
                VS_IMW_DECODE(pt0 + kb)
                mbar_wait(&pempty[ps], ((kb / IM_PST) & 1) ^ 1);
                mbar_expect_tx(&pready[ps], (uint32_t)p.patch_bytes);
                tma_load_3d(pbuf + ps * IM_PATCH_STAGE_W, &map_big, &pready[ps], ((q0 * p.stride - p.pad) * p.C) & ~7, p0 * p.stride - p.pad, b0);
                mbar_wait(&empty[st], ((kb / STAGES) & 1) ^ 1);
                uint8_t* b_dst = smem + st * STAGE_BYTES + A_BYTES;
                uint64_t* bbar = XF ? &bfull[st] : &full[st];
                mbar_expect_tx(bbar, B_BYTES);
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) tma_load_4d(b_dst + j * CHUNK, &map_small, bbar, kt * BN + j * 64, q0, p0, b0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16_f32(BN) | (1u << 15) | (1u << 16);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES;
                mbar_wait(&full[st], (kb / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + st * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                for (int k = 0; k < PIX / 16; ++k)
                    umma_bf16(tmem_base, mnmajor_sw128_desc(a_addr + k * 2048, CHUNK), mnmajor_sw128_desc(b_addr + k * 2048, CHUNK),
                              idesc, (kb | k) != 0 ? 1u : 0u);
                umma_commit(&empty[st]);
            }
            umma_commit(tmem_full);
        }
    } else if (XF && warp >= IMW_THREADS / 32) {
        // ===== XF: BatchNorm + activation applied in place to every landed box of `small` =====
        // chunk id = t + 128 i of the [64 pixels][128 B] box: row t/8 + 16 i, physical chunk t%8 -> always the same eight
        // channels for this thread (the swizzle XORs the chunk index with row & 7 = (t/8) & 7)
        const int t = threadIdx.x - IMW_THREADS;
        const int xphys = t & 7, xrsub = t >> 3, xcg = xphys ^ (xrsub & 7);
        float xsc[8], xsh[8];          // z = y * sc + sh with sc = gamma * invstd, sh = beta - mean * sc
        int xg = -1;
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % STAGES;
            const int g = ((pt0 + kb) / (p.tiles_w * p.tiles_h)) * p.NT / bn.n_per_group;
            if (g != xg) {
                xg = g;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = kt * BN + xcg * 8 + e;
                    xsc[e] = __ldg(bn.gamma + c) * __ldg(bn.invstd + (long long)g * p.K + c);
                    xsh[e] = __ldg(bn.beta + c) - __ldg(bn.mean + (long long)g * p.K + c) * xsc[e];
                }
            }
            mbar_wait(&bfull[st], (kb / STAGES) & 1);
            uint8_t* base = smem + st * STAGE_BYTES + A_BYTES + xrsub * 128 + xphys * 16;
#pragma unroll
            for (int i = 0; i < PIX / 16; ++i) {
                uint4* qp = reinterpret_cast<uint4*>(base + i * 16 * 128);
                uint4 v = *qp;
                uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a0 = __uint_as_float(wv[e] << 16), a1 = __uint_as_float(wv[e] & 0xffff0000u);
                    const float z0 = fmaf(a0, xsc[2 * e], xsh[2 * e]), z1 = fmaf(a1, xsc[2 * e + 1], xsh[2 * e + 1]);
                    const float r0 = fmaxf(z0, bn.neg_slope * z0), r1 = fmaxf(z1, bn.neg_slope * z1);
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(r0, r1);
                    wv[e] = *reinterpret_cast<uint32_t*>(&b2);
                }
                *qp = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[st]);
        }
    } else {
        // ===== builders: items = (16-byte chunk j, pixel m), m fastest =====
        const int t = threadIdx.x - 64;                   // 0..127
        const int items = p.nchunk16 * PIX;
        // at most two items per thread (nchunk16 <= 4): item i of this thread = chunk (t >> 6) + 2 i of pixel t & 63
        const bool cached = p.nchunk16 <= 4;
        ChunkOffsets co;
        const int cm = t & (PIX - 1), cj = t >> 6;
        {
            const int w = cm & (p.WT - 1), h = (cm >> p.wt_shift) & (p.HT - 1);
            const int base = h * p.stride * p.PWCp + w * p.stride * p.C;
            ChunkOffsets c0, c1;
            chunk_offsets(p, tab, base, cj, cj < p.nchunk16 ? 1 : 0, c0);
            chunk_offsets(p, tab, base, cj + 2, cj + 2 < p.nchunk16 ? 1 : 0, c1);
#pragma unroll
            for (int i = 0; i < 8; ++i) { co.off[i] = c0.off[i]; co.off[8 + i] = c1.off[i]; }
            co.valid = (c0.valid & 0xffu) | ((c1.valid & 0xffu) << 8);
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % STAGES, ps = kb % IM_PST;
            const int q0 = ((pt0 + kb) % p.tiles_w) * p.WT;
            const unsigned short* pb = reinterpret_cast<const unsigned short*>(pbuf + ps * IM_PATCH_STAGE_W) + (((q0 * p.stride - p.pad) * p.C) & 7);
            uint8_t* a_dst = smem + st * STAGE_BYTES;
            mbar_wait(&pready[ps], (kb / IM_PST) & 1);
            mbar_wait(&empty[st], ((kb / STAGES) & 1) ^ 1);
            if (cached) {
                if (cj < p.nchunk16) *reinterpret_cast<uint4*>(a_dst + cm * 128 + ((cj ^ (cm & 7)) << 4)) = gather_cached(pb, co, 0);
                if (cj + 2 < p.nchunk16) *reinterpret_cast<uint4*>(a_dst + cm * 128 + (((cj + 2) ^ (cm & 7)) << 4)) = gather_cached(pb, co, 1);
            } else {
                for (int it = t; it < items; it += 128) {
                    const int m = it & (PIX - 1), j = it >> 6;
                    const int w = m & (p.WT - 1), h = (m >> p.wt_shift) & (p.HT - 1);
                    const int base = h * p.stride * p.PWCp + w * p.stride * p.C;
                    const uint4 v = im2col_chunk(p, pb, tab, base, j);
                    *reinterpret_cast<uint4*>(a_dst + (j >> 3) * CHUNK + m * 128 + (((j & 7) ^ (m & 7)) << 4)) = v;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&full[st]); mbar_arrive(&pempty[ps]); }
        }
        // ===== epilogue: TMEM lane = kk, column = channel of `small` =====
        const int q = warp & 3;
        const int kk = q * 32 + lane;
        const int RS = p.R * p.S;
        const int tap = kk / p.C, c = kk - tap * p.C;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (kk < p.KK) {
                float* dst = dw + ((long long)(kt * BN + c0) * p.C + c) * RS + tap;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (kt * BN + c0 + i < p.K) atomicAdd(dst + (long long)i * p.C * RS, __uint_as_float(v[i]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

template <int BN, int STAGES, bool XF>
static int launch_imw(const CUtensorMap& ms, const CUtensorMap& mb, const Im2colParams& p, const BnApplyArgs& bn, float* dw, int k_tiles,
                      int splits, cudaStream_t stream) {
    constexpr int SMEM = STAGES * (2 * 64 * 64 * 2 + BN * 64 * 2) + 1024 + 256 + IM_MAX_KK * 4 + IM_PST * IM_PATCH_STAGE_W;
    static_assert((3 * STAGES + 2 * IM_PST + 2) * 8 <= 256, "barrier area");
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(im2col_wgrad_kernel<BN, STAGES, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return fail("im2col_wgrad_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    dim3 grid((unsigned)k_tiles, 1, (unsigned)splits);
    im2col_wgrad_kernel<BN, STAGES, XF><<<grid, IMW_THREADS + (XF ? IMW_XF_THREADS : 0), SMEM, stream>>>(ms, mb, p, bn, dw);
    return launched("im2col_wgrad_kernel");
}

// bn != nullptr: `small_` is the pre-BatchNorm tensor y; BatchNorm + activation is applied on the operand path
int conv_wgrad_im2col(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream,
                      const BnApplyArgs* bn) {
    if (g->dtype != VS_BF16 || (g->flags & VS_FLAG_FORCE_SIMT) || im2col_disabled()) return -1;
    if (bn != nullptr && g->K != 64) return -1;
    if (g->K % 8 != 0 || g->K < 32) return -1;                         // TMA: 16-byte channel pitch of `small`
    if ((reinterpret_cast<uintptr_t>(small_) | reinterpret_cast<uintptr_t>(big)) & 15) return -1;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1;
    Im2colParams p;
    if (!im2col_plan(p, g, 64, IM_PATCH_STAGE_W)) return -1;
    // partial boxes in W/H would shift the pixel order inside the TMA box relative to the builder's decode
    if (g->Q % p.WT != 0 || g->P % p.HT != 0) return -1;
    CUtensorMap mb;
    if (int rc = im2col_big_map(mb, p, big)) return rc;
    const int BN = g->K > 64 ? 128 : 64;
    const int k_tiles = (int)cdiv(g->K, BN);
    long long splits = cdiv(2LL * num_sms(), k_tiles);       // two CTAs per SM; fewer CTAs = fewer colliding atomics
    if (splits > p.total_tiles) splits = p.total_tiles;
    if (splits > 65535) splits = 65535;
    p.ptiles_per_split = (int)cdiv(p.total_tiles, splits);
    splits = cdiv(p.total_tiles, p.ptiles_per_split);
    CUtensorMap ms;
    {
        cuuint64_t dims[4] = {(cuuint64_t)g->K, (cuuint64_t)g->Q, (cuuint64_t)g->P, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)g->K * 2, (cuuint64_t)g->K * g->Q * 2, (cuuint64_t)g->K * g->Q * g->P * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.WT, (cuuint32_t)p.HT, (cuuint32_t)p.NT};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult rc = enc(&ms, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(small_), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(small) failed: %d", (int)rc);
    }
    BnApplyArgs none;
    memset(&none, 0, sizeof(none));
    if (bn != nullptr) return launch_imw<64, 4, true>(ms, mb, p, *bn, dw, k_tiles, (int)splits, stream);
    return BN == 128 ? launch_imw<128, 3, false>(ms, mb, p, none, dw, k_tiles, (int)splits, stream)
                     : launch_imw<64, 4, false>(ms, mb, p, none, dw, k_tiles, (int)splits, stream);
}

// ------------------------------------------------------------------------------------------ fused decoder tail, backward
// The thin transposed convolution's input gradient (a direct convolution of the frame gradient `dout` [N,H,W,C] with the
// DIRECT-packed weights wp [K][R*S][C]) is produced tile by tile on the tensor cores and consumed in the epilogue:
// phase 0 accumulates the BatchNorm backward sums, phase 1 writes dy.  0 = done, -1 = not eligible, > 0 = error.
int tail_eligible(const vs_conv_geom* g) {
    if (g->dtype != VS_BF16 || g->K != 64 || im2col_disabled()) return 0;
    Im2colParams p;
    if (!im2col_plan(p, g, 128, IM_PATCH_STAGE) || p.KK > 64) return 0;
    if (g->P % p.HT != 0 || g->Q % p.WT != 0) return 0;            // whole tiles only: every epilogue row is a pixel
    Im2colParams pw;
    if (!im2col_plan(pw, g, 64, IM_PATCH_STAGE_W) || g->Q % pw.WT != 0 || g->P % pw.HT != 0) return 0;
    return 1;
}

int tail_bn_backward(const vs_conv_geom* g, const BnBwdArgs& bb, const void* dout, const void* wp, int phase, double* sums, void* dy,
                     cudaStream_t stream) {
    if (!tail_eligible(g) || !encode_fn()) return -1;
    if ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(bb.y) | reinterpret_cast<uintptr_t>(dy)) & 15) return -1;
    Im2colParams p;
    if (!im2col_plan(p, g, 128, IM_PATCH_STAGE)) return -1;
    CUtensorMap mb;
    if (int rc = im2col_big_map(mb, p, dout)) return rc;
    p.act = VS_ACT_NONE; p.has_bias = 0; p.partial = 0;
    p.n_per_group = bb.n_per_group;
    CUtensorMap mo;
    memset(&mo, 0, sizeof(mo));
    if (phase == 0) return launch_imf<64, 1, 4, false, 1>(mb, mo, p, wp, nullptr, nullptr, sums, stream, &bb);
    {
        cuuint64_t dims[4] = {(cuuint64_t)g->K, (cuuint64_t)g->Q, (cuuint64_t)g->P, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)g->K * 2, (cuuint64_t)g->K * g->Q * 2, (cuuint64_t)g->K * g->Q * g->P * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.WT, (cuuint32_t)p.HT, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode_fn()(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(tail dy) failed: %d", (int)r);
    }
    return launch_imf<64, 1, 3, true, 2>(mb, mo, p, wp, nullptr, dy, nullptr, stream, &bb);
}

}  // namespace vs

// Last decoder up-convolution (ConvTranspose2d(nf, nc, 4, 2, 1), nc <= 2; /root/reference/var_sep/networks/conv.py:258-263)
// as GEMM + col2im on the tensor cores.
//
// A tap GEMM would fetch every input pixel once per output-parity class and tap (16 times) to fill nc of the
// accumulator columns.  Here the input image is read ONCE: for every input pixel
//   G[pix][(c, r, s)] = sum_ic x[pix][ic] * w[ic][c][r][s]                 (tcgen05, M = 128 pixels, N = 16*nc)
// and an output pixel gathers the (at most) four G entries whose tap lands on it
//   out[2i-1+r][2j-1+s][c] += G[(i, j)][(c, r, s)]
// from a per-image fp32 copy of G in shared memory, followed by bias, activation and one coalesced bf16 store.
// One work item = one image (P*Q <= 1024 input pixels = up to 8 accumulator tiles in TMEM, double buffered so the
// next image's loads and MMAs overlap this image's col2im).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace vs {

struct Col2imParams {
    int N, P, Q, K, C;           // input [N,P,Q,K], output [N,2P,2Q,C]
    int kchunks, tiles;          // K/64, P*Q/128
    int HT;                      // input rows per 128-pixel tile (= 128/Q)
    int act, has_bias;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int C2I_THREADS = 320, C2I_XF_THREADS = 128, C2I_MAX_TILES = 8, C2I_MAX_KCHUNKS = 4;

template <int NB, int STAGES>
struct C2iSmem {
    static constexpr int A_BYTES = TC_BM * 128, B_BYTES = C2I_MAX_KCHUNKS * NB * 128;
    static constexpr int G_OFF = STAGES * A_BYTES + B_BYTES;
    static constexpr int G_BYTES = NB * C2I_MAX_TILES * 128 * 4;
    static constexpr int BAR_OFF = G_OFF + G_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static_assert((3 * STAGES + 6) * 8 <= 256, "barrier area");
};

// XF: the input tensor is the PRE-BatchNorm output y of the previous layer; four extra warps apply
//   act(gamma * (y - mean[g]) * invstd[g] + beta)   (as one fused multiply-add per element: bn_act_fwd_col_kernel's result
// up to the fp32 rounding before the bf16 rounding)
// to every A stage in shared memory between its TMA landing and the MMA, so the normalised tensor (268 MB for the
// Moving-MNIST decoder) is never written to or re-read from HBM.
template <int NB, int STAGES, bool XF>
__global__ void __launch_bounds__(C2I_THREADS + (XF ? C2I_XF_THREADS : 0), 1)
convT_col2im_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ Col2imParams p, const __grid_constant__ BnApplyArgs bn,
                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
    using S = C2iSmem<NB, STAGES>;
    constexpr int ACC_COLS = C2I_MAX_TILES * NB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* b_smem = smem + STAGES * S::A_BYTES;
    float* G = reinterpret_cast<float*>(smem + S::G_OFF);                 // [NB][P*Q]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint64_t* b_full = bars + 2 * STAGES + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);
    uint64_t* xready = bars + 2 * STAGES + 6;        // [STAGES] XF: stage transformed in place, ready for the MMA

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int PQ = p.P * p.Q;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
        mbar_init(b_full, 1);
        if (XF) for (int s = 0; s < STAGES; ++s) mbar_init(&xready[s], C2I_XF_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<2 * ACC_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: the weights once, then one box per (image, pixel tile, 64-channel chunk) =====
        if (lane == 0) {
            mbar_expect_tx(b_full, (uint32_t)(p.kchunks * NB * 128));
            for (int kc = 0; kc < p.kchunks; ++kc) tma_load_2d(b_smem + kc * NB * 128, &map_b, b_full, kc * 64, 0);
            int it = 0;
            for (int img = blockIdx.x; img < p.N; img += gridDim.x)
                for (int t = 0; t < p.tiles; ++t)
                    for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                        mbar_expect_tx(&full[s], S::A_BYTES);
                        tma_load_4d(smem + s * S::A_BYTES, &map_a, &full[s], kc * 64, 0, t * p.HT, img);
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16_f32(NB);
            mbar_wait(b_full, 0);
            int it = 0, li = 0;
            for (int img = blockIdx.x; img < p.N; img += gridDim.x, ++li) {
                const int acc = li & 1;
                mbar_wait(&tmem_empty[acc], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int t = 0; t < p.tiles; ++t) {
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS + t * NB);
                    for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(XF ? &xready[s] : &full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + s * S::A_BYTES), b_addr = smem_u32(b_smem + kc * NB * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_d, kmajor_sw128_desc(a_addr + k * 32), kmajor_sw128_desc(b_addr + k * 32), idesc,
                                      (kc | k) != 0 ? 1u : 0u);
                        umma_commit(&empty[s]);
                    }
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (XF && warp >= C2I_THREADS / 32) {
        // ===== BatchNorm + activation applied to the landed stage, in place =====
        // 16-byte chunk id = t + 128 i: pixel row t/8 + 16 i, physical chunk t%8; the 128-byte swizzle XORs the chunk
        // index with (row & 7) = ((t/8) & 7), so a thread always owns the SAME eight channels: parameters in registers
        const int t = threadIdx.x - C2I_THREADS;
        const int phys = t & 7, rsub = t >> 3;
        const int cg = phys ^ (rsub & 7);
        float sc[8], sh[8];                     // z = y * sc + sh with sc = gamma * invstd, sh = beta - mean * sc
        int cur_g = -1, cur_kc = -1, it = 0;
        for (int img = blockIdx.x; img < p.N; img += gridDim.x) {
            const int g = img / bn.n_per_group;
            for (int tl = 0; tl < p.tiles; ++tl)
                for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                    if (g != cur_g || kc != cur_kc) {
                        cur_g = g; cur_kc = kc;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = kc * 64 + cg * 8 + e;
                            sc[e] = __ldg(bn.gamma + c) * __ldg(bn.invstd + (long long)g * p.K + c);
                            sh[e] = __ldg(bn.beta + c) - __ldg(bn.mean + (long long)g * p.K + c) * sc[e];
                        }
                    }
                    const int s = it % STAGES;
                    mbar_wait(&full[s], (it / STAGES) & 1);
                    uint8_t* base = smem + s * S::A_BYTES + rsub * 128 + phys * 16;
#pragma unroll
                    for (int i = 0; i < TC_BM / 16; ++i) {
                        uint4* q = reinterpret_cast<uint4*>(base + i * 16 * 128);
                        uint4 v = *q;
                        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float a0 = __uint_as_float(w[e] << 16), a1 = __uint_as_float(w[e] & 0xffff0000u);
                            // act(z) = max(z, neg_slope * z) for neg_slope in [0, 1]: none / ReLU / LeakyReLU without a branch
                            const float z0 = fmaf(a0, sc[2 * e], sh[2 * e]), z1 = fmaf(a1, sc[2 * e + 1], sh[2 * e + 1]);
                            const float r0 = fmaxf(z0, bn.neg_slope * z0), r1 = fmaxf(z1, bn.neg_slope * z1);
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(r0, r1);
                            w[e] = *reinterpret_cast<uint32_t*>(&b2);
                        }
                        *q = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&xready[s]);
                }
        }
    } else {
        // ===== epilogue (8 warps): TMEM -> G in shared memory, then the output gather =====
        const int ew = warp - 2;                 // 0..7
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int half = ew >> 2;                // the two warps of a quarter alternate over the tiles
        const int et = threadIdx.x - 64;         // 0..255
        const int H = 2 * p.P, W = 2 * p.Q;
        int li = 0;
        for (int img = blockIdx.x; img < p.N; img += gridDim.x, ++li) {
            const int acc = li & 1;
            mbar_wait(&tmem_full[acc], (li >> 1) & 1);
            tc_fence_after();
            for (int t = half; t < p.tiles; t += 2) {
                float* g = G + t * 128 + q * 32 + lane;
#pragma unroll
                for (int c0 = 0; c0 < NB; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + t * NB + c0), r);
#pragma unroll
                    for (int c = 0; c < 16; ++c) g[(c0 + c) * PQ] = __uint_as_float(r[c]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // items: (channel c, output row h, group of 8 output columns)
            const int groups_w = W >> 3;
            const int items = p.C * H * groups_w;
            for (int item = et; item < items; item += 256) {
                const int gw = item % groups_w;
                const int h = (item / groups_w) % H;
                const int c = item / (groups_w * H);
                const float b = p.has_bias ? __ldg(bias + c) : 0.f;
                const int r0 = (h + 1) & 1;
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = b;
#pragma unroll
                for (int dr = 0; dr < 2; ++dr) {
                    const int r = r0 + 2 * dr;
                    const int i = (h + 1 - r) >> 1;              // h + 1 - r is even
                    if (h + 1 - r < 0 || i >= p.P) continue;
                    const float* grow = G + (c * 16 + r * 4) * PQ + i * p.Q;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int w = gw * 8 + e;
                        const int s0 = (w + 1) & 1;
#pragma unroll
                        for (int ds = 0; ds < 2; ++ds) {
                            const int s = s0 + 2 * ds;
                            const int j = (w + 1 - s) >> 1;
                            if (w + 1 - s >= 0 && j < p.Q) v[e] += grow[s * PQ + j];
                        }
                    }
                }
                __nv_bfloat16* dst = out + (((long long)img * H + h) * W + gw * 8) * p.C + c;
                if (p.C == 1) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(act_fwd(v[2 * e], p.act), act_fwd(v[2 * e + 1], p.act));
                        pk[e] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) dst[e * p.C] = __float2bfloat16_rn(act_fwd(v[e], p.act));
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // G is rewritten by the next image
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * ACC_COLS>(tmem_base);
    }
}

static bool col2im_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_DISABLE_COL2IM"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1 || tc_disabled();
}

int conv_forward_col2im_eligible(const vs_conv_geom* g, int mode) {
    if (g->dtype != VS_BF16 || (g->flags & VS_FLAG_FORCE_SIMT) || col2im_disabled()) return 0;
    if (mode != VS_CONV_TRANSPOSED || g->R != 4 || g->S != 4 || g->stride != 2 || g->pad != 1) return 0;
    if (g->C > 2 || g->K % 64 != 0 || g->K > 64 * C2I_MAX_KCHUNKS) return 0;
    if (g->Q > 128 || (g->Q & (g->Q - 1)) != 0 || g->Q < 4) return 0;
    const int HT = 128 / g->Q;
    if (g->P % HT != 0 || g->P * g->Q > 128 * C2I_MAX_TILES) return 0;
    if (g->H != 2 * g->P || g->W != 2 * g->Q) return 0;
    return 1;
}

template <int NB, int STAGES, bool XF>
static int launch_c2i(const CUtensorMap& ma, const CUtensorMap& mb, const Col2imParams& p, const BnApplyArgs& bn, const float* bias,
                      void* out, cudaStream_t stream) {
    using S = C2iSmem<NB, STAGES>;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(convT_col2im_kernel<NB, STAGES, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail("convT_col2im_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    const int grid = p.N < num_sms() ? p.N : num_sms();
    convT_col2im_kernel<NB, STAGES, XF><<<grid, C2I_THREADS + (XF ? C2I_XF_THREADS : 0), S::TOTAL, stream>>>(
        ma, mb, p, bn, bias, (__nv_bfloat16*)out);
    return launched("convT_col2im_kernel");
}

// returns 0 = done, -1 = geometry not eligible, >0 = error.  bn != nullptr: `in` is the pre-BatchNorm tensor y and the
// normalisation + activation described by *bn is applied on the operand path (the fused decoder tail)
int conv_forward_col2im(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                        double* stats, cudaStream_t stream, const BnApplyArgs* bn) {
    if (stats != nullptr || !conv_forward_col2im_eligible(g, mode)) return -1;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(wp) | reinterpret_cast<uintptr_t>(out)) & 15) return -1;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1;
    Col2imParams p;
    memset(&p, 0, sizeof(p));
    p.N = g->N; p.P = g->P; p.Q = g->Q; p.K = g->K; p.C = g->C;
    p.kchunks = g->K / 64; p.HT = 128 / g->Q; p.tiles = g->P * g->Q / 128;
    p.act = g->act; p.has_bias = bias != nullptr;
    const int NB = 16 * g->C;
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)g->K, (cuuint64_t)g->Q, (cuuint64_t)g->P, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)g->K * 2, (cuuint64_t)g->K * g->Q * 2, (cuuint64_t)g->K * g->Q * g->P * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)g->Q, (cuuint32_t)p.HT, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(col2im A) failed: %d", (int)r);
    }
    {
        // TRANSPOSED-operand pack [C][R*S][K] read as a [16*C][K] matrix: row = c*16 + r*4 + s
        cuuint64_t dims[2] = {(cuuint64_t)g->K, (cuuint64_t)NB};
        cuuint64_t strides[1] = {(cuuint64_t)g->K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)NB};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(col2im B) failed: %d", (int)r);
    }
    BnApplyArgs none;
    memset(&none, 0, sizeof(none));
    if (bn != nullptr)
        return NB == 16 ? launch_c2i<16, 6, true>(ma, mb, p, *bn, bias, out, stream) : launch_c2i<32, 4, true>(ma, mb, p, *bn, bias, out, stream);
    return NB == 16 ? launch_c2i<16, 6, false>(ma, mb, p, none, bias, out, stream) : launch_c2i<32, 4, false>(ma, mb, p, none, bias, out, stream);
}

}  // namespace vs

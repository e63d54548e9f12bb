// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Replaces the cuDNN kernels behind nn.Conv2d / nn.ConvTranspose2d of the DCGAN / VGG / SST stacks
// (/root/reference/var_sep/networks/conv.py:119-123,147-170,258-263,295-318) for the layers whose
// input channel count is a multiple of 8 (16-byte TMA pitch).  One kernel covers
//   - direct convolutions (any R x S, stride 1 or 2, zero padding),
//   - transposed convolutions, decomposed into stride*stride output-parity classes so that no
//     structurally-zero tap is ever multiplied (k4 s2 p1: four 2x2 stride-1 convolutions),
// as a "tap GEMM": for every output tile of 128 pixels and every (tap, 64-channel chunk) one TMA box
// of the NHWC input (shifted by the tap offset, zero-filled outside the image by the TMA unit, strided
// for stride-2 convolutions) is multiplied with a [BN x 64] slab of the packed weights.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 2..9 = epilogue (TMEM -> registers -> bias / activation / BatchNorm sums -> bf16 -> global, directly or
// through a staged shared-memory tile and a TMA store).  Persistent CTAs, two per SM, two TMEM accumulator stages.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace vs {

constexpr int TC_MAX_TAPS = 25, TC_MAX_CLASSES = 16, TC_TABLE = 32, TC_MAX_OUT_MAPS = 4;   // classes * taps <= TC_TABLE
constexpr int TC_EPI_WARPS = 8, TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // TMA warp + MMA warp + epilogue warps

struct TcParams {
    int N, OH, OW, OC;
    int OHc, OWc, ost;        // class grid (output pixels of one parity class) and output sub-sampling
    int WT, HT, NT;           // tile box in class-grid units, WT*HT*NT == 128
    int tiles_w, tiles_h, tiles_n, n_tiles, classes, total_tiles;
    int IC, kchunks, ntaps;
    int in_sh, in_sw;         // class-grid -> input coordinate multiplier
    int act, has_bias, partial, n_per_group;
    int res_taps;             // resident-weight pair kernel: filter taps R*S (slabs = res_taps * kchunks)
    signed char dh[TC_TABLE], dw[TC_TABLE];          // [class * ntaps + tap]: input offset of the tap on the class grid
    unsigned char wtap[TC_TABLE];                    // ... and its index r*S + s in the packed weights
    unsigned char ca[TC_MAX_CLASSES], cb[TC_MAX_CLASSES];
};

template <int BN, int STAGES, bool TS>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2, B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // two bf16 output tiles staged for TMA stores; a tile is BN/64 half-tiles of [128 pixels][128 bytes]
    static constexpr int TILE_BYTES = (BN / 64) * TC_BM * 128;
    static constexpr int OUT_BYTES = TS ? 2 * TILE_BYTES : 0;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES + OUT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 1024 /*align slack*/ + 256 /*barriers*/ + BN * 2 * 4 /*tile statistics*/;
    static_assert((2 * STAGES + 5) * 8 <= 256, "barrier area");
    static_assert(!TS || BN == 64 || BN == 128, "the staged epilogue handles 128-byte rows: 64-column half-tiles");
    static_assert(TOTAL <= 227 * 1024, "shared memory");
};

// output tensor maps of the staged epilogue, one per output-parity class: [OC, OW/ost, OH/ost, N] views of the NHWC output
struct TcOutMaps { CUtensorMap m[TC_MAX_OUT_MAPS]; };

// Persistent: every CTA walks a strided list of (pixel tile, output-channel tile, parity class) work items.
// The TMA producer and the MMA issuer run ahead across tile boundaries (one smem ring for the whole kernel);
// two TMEM accumulator stages let tile i+1's MMAs overlap tile i's epilogue.
// TS: the epilogue stages the bf16 tile in shared memory and writes it with one TMA store (coalesced, clipped at the
// tensor edges by the TMA unit) instead of 16-byte stores from every lane to 32 different lines.
// SS (with TS): the BatchNorm sums are taken from the staged tile -- thread (column pair, 16-row group) reads its 16
// bf16x2 words (one 128-byte row per warp instruction: conflict-free) and keeps four fp32 running sums in registers
// across the tiles of a BatchNorm group -- instead of two 31-shuffle transposes of the accumulator per warp and tile
// (ncu: the epilogue warps executed ~770 instructions per tile at 8.5 cycles each and were the critical path of the
// 128->64 decoder layer).  The statistics are then those of the bf16 values that BatchNorm will normalise.
template <int BN, int STAGES, bool TS, bool SS>
__global__ void __launch_bounds__(TC_THREADS, 2) tc_conv_kernel(const __grid_constant__ CUtensorMap map_a,
                                                         const __grid_constant__ CUtensorMap map_b,
                                                         const __grid_constant__ TcOutMaps omaps,
                                                         const __grid_constant__ TcParams p,
                                                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                         double* __restrict__ stats) {
    static_assert(!SS || TS, "statistics from the staged tile need the staged epilogue");
    using S = TcSmem<BN, STAGES, TS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* obuf = smem + STAGES * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;        // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* sstat = reinterpret_cast<float*>(smem + S::BAR_OFF + 256);      // [BN][2] per-tile column sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.ntaps * p.kchunks;
    // every CTA walks one contiguous range of work items: the parity classes / channel tiles of a pixel tile run back to
    // back on the same SM (input tile re-read from L2), and a CTA sees few BatchNorm groups (statistics are flushed
    // once per group and CTA, not once per tile)
    const int t_begin = (int)((long long)blockIdx.x * p.total_tiles / gridDim.x);
    const int t_end = (int)((long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x);

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) sstat[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item -> (class, output-channel tile, pixel tile); the class index is fastest so that the CTAs running
    // at the same time read the same input pixels (L2 hits), then the channel tile, then the pixel tile
#define VS_TC_DECODE(idx)                                                   \
    const int cls = (idx) % p.classes;                                      \
    const int rest_ = (idx) / p.classes;                                    \
    const int n0 = (rest_ % p.n_tiles) * BN;                                \
    int t_ = rest_ / p.n_tiles;                                             \
    const int tw = t_ % p.tiles_w; t_ /= p.tiles_w;                         \
    const int th = t_ % p.tiles_h;                                          \
    const int tn = t_ / p.tiles_h;                                          \
    const int j0 = tw * p.WT, i0 = th * p.HT, b0 = tn * p.NT;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int idx = t_begin; idx < t_end; ++idx) {
                VS_TC_DECODE(idx)
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
                    uint8_t* a_dst = smem + s * S::STAGE_BYTES;
                    uint8_t* b_dst = a_dst + S::A_BYTES;
                    mbar_expect_tx(&full[s], S::STAGE_BYTES);
                    tma_load_4d(a_dst, &map_a, &full[s], kc * TC_BK, j0 * p.in_sw + p.dw[cls * p.ntaps + tap], i0 * p.in_sh + p.dh[cls * p.ntaps + tap], b0);
                    tma_load_2d(b_dst, &map_b, &full[s], p.wtap[cls * p.ntaps + tap] * p.IC + kc * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp stays converged in the loop, so that barrier addresses and descriptors are
        // computed on the uniform datapath (a lane-0-only loop re-derived them per instruction through R2UR/ELECT
        // sequences, ~17 instructions per MMA: more than a 64-column MMA lasts); one elected lane issues =====
        {
            constexpr uint32_t idesc = idesc_bf16_f32(BN);
            const bool elected = elect_one();
            int s = 0, lt = 0;
            uint32_t ph = 0;
            for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
                const int acc = lt & 1;
                mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_lo = kmajor_sw128_desc_lo(smem_u32(smem + s * S::STAGE_BYTES));
                    const uint32_t b_lo = a_lo + S::A_BYTES / 16;
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t ad = kmajor_sw128_desc_from_lo(a_lo + 2 * k), bd = kmajor_sw128_desc_from_lo(b_lo + 2 * k);
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                        if (elected) umma_bf16(tmem_d, ad, bd, idesc, accum);
                    }
                    if (elected) umma_commit(&empty[s]);          // smem slot reusable once these MMAs have read it
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (elected) umma_commit(&tmem_full[acc]);        // accumulator of this tile complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: warps 2..9.  A warp may only touch the TMEM lane quarter (warp % 4); the two warps that share a
        // quarter split the accumulator columns =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = BN / (TC_EPI_WARPS / 4);
        const int m = q * 32 + lane;             // tile-local pixel (TMEM lane)
        const int w = m % p.WT, h = (m / p.WT) % p.HT, n = m / (p.WT * p.HT);
        int lt = 0;
        int stat_key = -1;
        double stat_acc = 0.0;
        constexpr int HALVES = BN / 64;              // 64-column half-tiles of the staged tile
        float ss[HALVES][4] = {};                    // SS: sum / sum of squares of this thread's two columns per half-tile
        // SS: flush the register sums of the finished BatchNorm group through the shared-memory column table
        auto ss_flush = [&](int key) {
            const int t = threadIdx.x - 64, cp = t & 31;
#pragma unroll
            for (int hf = 0; hf < HALVES; ++hf) {
                const int c = hf * 64 + 2 * cp;
                atomicAdd(&sstat[c * 2], ss[hf][0]);       atomicAdd(&sstat[c * 2 + 1], ss[hf][1]);
                atomicAdd(&sstat[(c + 1) * 2], ss[hf][2]); atomicAdd(&sstat[(c + 1) * 2 + 1], ss[hf][3]);
                ss[hf][0] = ss[hf][1] = ss[hf][2] = ss[hf][3] = 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            if (t < 2 * BN) {
                const int col = (key % p.n_tiles) * BN + (t >> 1);
                if (col < p.OC) atomicAdd(&stats[((long long)(key / p.n_tiles) * p.OC + col) * 2 + (t & 1)], (double)sstat[t]);
                sstat[t] = 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        };
        for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
            VS_TC_DECODE(idx)
            const int acc = lt & 1;
            const int i = i0 + h, j = j0 + w, nn = b0 + n;
            const bool ok = i < p.OHc && j < p.OWc && nn < p.N;
            const bool full_tile = i0 + p.HT <= p.OHc && j0 + p.WT <= p.OWc && b0 + p.NT <= p.N;
            const long long pix = ((long long)nn * p.OH + (long long)i * p.ost + p.ca[cls]) * p.OW + (long long)j * p.ost + p.cb[cls];
            __nv_bfloat16* dst = out + pix * p.OC + n0;
            if (SS && stats != nullptr) {
                const int key = (b0 / p.n_per_group) * p.n_tiles + n0 / BN;      // a tile never straddles groups (host check)
                if (key != stat_key) {                                           // uniform over the epilogue warps
                    if (stat_key >= 0) ss_flush(stat_key);
                    stat_key = key;
                }
            }
            mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = half * COLS_PER_WARP; c0 < (half + 1) * COLS_PER_WARP; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                // one coalesced bias load per chunk, broadcast by shuffle (instead of 32 loads per thread); the
                // shuffles are executed by all lanes before any per-lane predicate
                float xs[32];
                if (p.has_bias) {          // uniform: dgrad launches carry no bias and skip the 32 shuffles
                    const float bias_l = (p.has_bias && n0 + c0 + lane < p.OC) ? __ldg(bias + n0 + c0 + lane) : 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]) + __shfl_sync(0xffffffffu, bias_l, c);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]);
                }
                if (TS) {
                    if (SS && !full_tile && !ok) {          // rows beyond the tensor edge: clipped by the TMA store, must not count
#pragma unroll
                        for (int c = 0; c < 32; ++c) xs[c] = 0.f;
                    }
                    stage_row32(obuf + (lt & 1) * S::TILE_BYTES + (c0 >> 6) * (TC_BM * 128), m, c0 & 63, xs, p.act);
                } else if (p.partial || n0 + BN > p.OC) {
                    // tail tile in OC, or rows not 16-byte aligned: predicated scalar stores
                    if (ok) {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (n0 + c0 + c < p.OC) dst[c0 + c] = __float2bfloat16_rn(act_fwd(xs[c], p.act));
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
                if (!SS && stats != nullptr) {
                    // BatchNorm batch statistics of the fp32 accumulator (+bias), fused: 32 rows x 32 columns per warp
                    // (columns past OC hold exact zeros: zero-filled weight rows, no bias)
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
            // this warp's quarter of the accumulator has been read: hand the stage back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (TS) {
                // staged tile complete -> one TMA store.  The buffer written now was the source of the store issued two
                // tiles ago; the store of the previous tile (other buffer) is drained here, before anyone refills it
                fence_proxy_async();
                if (threadIdx.x == 64) tma_store_wait_read();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                if (threadIdx.x == 64) {
#pragma unroll
                    for (int hf = 0; hf < BN / 64; ++hf)
                        if (n0 + hf * 64 < p.OC)
                            tma_store_4d(&omaps.m[cls], obuf + (lt & 1) * S::TILE_BYTES + hf * (TC_BM * 128), n0 + hf * 64, j0, i0, b0);
                    tma_store_commit();
                }
            }
            if (SS && stats != nullptr) {
                // columns 2cp, 2cp+1 of rows 16rg .. 16rg+15 of the tile staged above (complete since the barrier); the
                // buffer is refilled two tiles from now, behind the next tile's barrier
                const int t = threadIdx.x - 64, cp = t & 31, rg = t >> 5;
#pragma unroll
                for (int hf = 0; hf < HALVES; ++hf) {
                    const uint8_t* tile = obuf + (lt & 1) * S::TILE_BYTES + hf * (TC_BM * 128) + (cp & 3) * 4;
#pragma unroll
                    for (int rr = 0; rr < TC_BM / TC_EPI_WARPS; ++rr) {
                        const int row = rg * (TC_BM / TC_EPI_WARPS) + rr;
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(tile + row * 128 + (((cp >> 2) ^ (row & 7)) << 4));
                        const float a = __uint_as_float(v << 16), b = __uint_as_float(v & 0xffff0000u);
                        ss[hf][0] += a; ss[hf][1] = fmaf(a, a, ss[hf][1]);
                        ss[hf][2] += b; ss[hf][3] = fmaf(b, b, ss[hf][3]);
                    }
                }
            }
            if (!SS && stats != nullptr) {
                // the epilogue warps have added their 32-row partials of this tile; thread i keeps the running fp64 sum of
                // entry i (column n0 + i/2, sum or sum of squares) and flushes it when the (group, channel tile) changes
                if (!TS) asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                const int t = threadIdx.x - 64;              // 0 .. 32*TC_EPI_WARPS-1
                if (t < 2 * BN) {
                    const int key = (b0 / p.n_per_group) * p.n_tiles + n0 / BN;      // a tile never straddles groups (host check)
                    if (key != stat_key) {
                        if (stat_key >= 0) {
                            const int col = (stat_key % p.n_tiles) * BN + (t >> 1);
                            if (col < p.OC) atomicAdd(&stats[((long long)(stat_key / p.n_tiles) * p.OC + col) * 2 + (t & 1)], stat_acc);
                        }
                        stat_key = key; stat_acc = 0.0;
                    }
                    stat_acc += (double)sstat[t];
                    sstat[t] = 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            }
        }
        if (SS) {
            if (stats != nullptr && stat_key >= 0) ss_flush(stat_key);
        } else if (stats != nullptr && stat_key >= 0) {
            const int t = threadIdx.x - 64;
            const int col = (stat_key % p.n_tiles) * BN + (t >> 1);
            if (t < 2 * BN && col < p.OC) atomicAdd(&stats[((long long)(stat_key / p.n_tiles) * p.OC + col) * 2 + (t & 1)], stat_acc);
        }
        if (TS && threadIdx.x == 64) tma_store_wait_all();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
#undef VS_TC_DECODE
}

// ------------------------------------------------------------------------------------------ fused parity classes
// ConvTranspose k4 s2 p1 with 64 output channels per tile: the four output-parity classes of a 128-pixel tile of the class
// grid are computed TOGETHER.  They read the same 3x3 neighbourhood of input shifts (each class uses 2x2 of them), so one
// work item stages 9 input boxes instead of 16, and a shift shared by 2 or 4 classes is multiplied by the classes'
// weight slabs in ONE tcgen05.mma of N = 128 or 256: the slabs sit back to back in shared memory in the order of the
// classes' accumulator column blocks [(0,0) (0,1) (1,1) (1,0)], in which three of the four class pairs that share a shift
// are adjacent.  10 MMAs per 16 input channels instead of 16 N = 64 ones (ncu: N = 64 instructions keep the tensor pipe
// 65 % busy at 36 % of its FLOP rate).  The centre shift (all four classes) is issued first, so it alone initialises
// the accumulators.  Accumulators: 2 stages x 256 columns = all of TMEM, one CTA per SM.  Epilogue: the four class
// tiles are staged as bf16 in shared memory, written by four TMA stores, and the BatchNorm sums are read back from the
// staged tiles (see SS above).
constexpr int TCF_SHIFTS = 9;
// canonical schedule: shift (dh, dw), then per MMA of the shift: first slab, slabs (x 64 columns), first column block.
// Column blocks hold the classes (0,0) (0,1) (1,1) (1,0); equal instruction shapes are issued back to back.
__device__ constexpr signed char TCF_DH[TCF_SHIFTS] = {0, -1, 0, 1, 0, -1, -1, 1, 1};
__device__ constexpr signed char TCF_DW[TCF_SHIFTS] = {0, 0, 1, 0, -1, -1, 1, 1, -1};
__device__ constexpr unsigned char TCF_NG[TCF_SHIFTS] = {1, 1, 1, 1, 2, 1, 1, 1, 1};
__device__ constexpr unsigned char TCF_SLAB[TCF_SHIFTS][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 1}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
__device__ constexpr unsigned char TCF_N[TCF_SHIFTS][2] = {{4, 0}, {2, 0}, {2, 0}, {2, 0}, {1, 1}, {1, 0}, {1, 0}, {1, 0}, {1, 0}};
__device__ constexpr unsigned char TCF_POS[TCF_SHIFTS][2] = {{0, 0}, {0, 0}, {1, 0}, {2, 0}, {0, 3}, {0, 0}, {1, 0}, {2, 0}, {3, 0}};
struct TcFusedTab {
    int nshift;
    signed char dh[TCF_SHIFTS], dw[TCF_SHIFTS];
    unsigned char nslab[TCF_SHIFTS];              // classes that use the shift
    unsigned char slab_wtap[TCF_SHIFTS][4];       // packed-weight tap of slab j (slabs ordered by column block)
    unsigned char ngrp[TCF_SHIFTS];               // MMAs per shift: runs of adjacent column blocks
    unsigned char grp_slab[TCF_SHIFTS][4], grp_n[TCF_SHIFTS][4], grp_pos[TCF_SHIFTS][4];
    unsigned char pos_cls[4];                     // class whose outputs column block `pos` holds
};

template <int STAGES>
struct TcFusedSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2, SLAB = 64 * TC_BK * 2, STAGE_BYTES = A_BYTES + 4 * SLAB;
    static constexpr int OUT_BYTES = 4 * TC_BM * 128;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES + OUT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 1024 /*align slack*/ + 256 /*barriers*/ + 64 * 2 * 4 /*column sums*/;
    static_assert((2 * STAGES + 5) * 8 <= 256, "barrier area");
};

template <int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_convT4_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b,
                                                                  const __grid_constant__ TcOutMaps omaps,
                                                                  const __grid_constant__ TcParams p,
                                                                  const __grid_constant__ TcFusedTab tab,
                                                                  const float* __restrict__ bias, double* __restrict__ stats) {
    using S = TcFusedSmem<STAGES>;
    constexpr int BN = 64, ACC = 4 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* obuf = smem + STAGES * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;        // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* sstat = reinterpret_cast<float*>(smem + S::BAR_OFF + 256);      // [BN][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_begin = (int)((long long)blockIdx.x * p.total_tiles / gridDim.x);
    const int t_end = (int)((long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x);

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<2 * ACC>(tmem_slot);
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) sstat[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item -> (output-channel tile, pixel tile of the class grid); all four classes belong to the item
#define VS_TCF_DECODE(idx)                                                  \
    const int n0 = ((idx) % p.n_tiles) * BN;                                \
    int t_ = (idx) / p.n_tiles;                                             \
    const int tw = t_ % p.tiles_w; t_ /= p.tiles_w;                         \
    const int th = t_ % p.tiles_h;                                          \
    const int tn = t_ / p.tiles_h;                                          \
    const int j0 = tw * p.WT, i0 = th * p.HT, b0 = tn * p.NT;

    if (warp == 0) {
        // ===== TMA producer: per (shift, 64-channel chunk) one input box + the weight slabs of the classes that use it =====
        if (lane == 0) {
            int it = 0;
            for (int idx = t_begin; idx < t_end; ++idx) {
                VS_TCF_DECODE(idx)
                for (int sh = 0; sh < tab.nshift; ++sh) {
                    const int ns = tab.nslab[sh];
                    for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
                        mbar_expect_tx(&full[s], (uint32_t)(S::A_BYTES + ns * S::SLAB));
                        tma_load_4d(a_dst, &map_a, &full[s], kc * TC_BK, j0 + tab.dw[sh], i0 + tab.dh[sh], b0);
                        for (int j = 0; j < ns; ++j)
                            tma_load_2d(a_dst + S::A_BYTES + j * S::SLAB, &map_b, &full[s], tab.slab_wtap[sh][j] * p.IC + kc * TC_BK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread).  The MMA schedule of a work item is fixed (TCF_* tables below, canonical shift
        // order enforced by the host), so the loops are fully unrolled with compile-time column blocks, slab offsets and
        // instruction descriptors: a table-driven issuer spent ~250 cycles of dependent scalar work per MMA and left the
        // tensor pipe 25 % busy =====
        if (lane == 0) {
            int it = 0, lt = 0;
            for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
                const int acc = lt & 1;
                mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC);
#pragma unroll
                for (int sh = 0; sh < TCF_SHIFTS; ++sh) {
#pragma unroll 1
                    for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&full[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
                        const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {
#pragma unroll
                            for (int g = 0; g < TCF_NG[sh]; ++g) {
                                // the first MMA of a work item is the centre shift: N = 256, it overwrites all four blocks
                                umma_bf16(tmem_d + (uint32_t)(TCF_POS[sh][g] * BN), kmajor_sw128_desc(a_addr + k * 32),
                                          kmajor_sw128_desc(b_addr + TCF_SLAB[sh][g] * S::SLAB + k * 32), idesc_bf16_f32(64 * TCF_N[sh][g]),
                                          (sh != 0 || (kc | k) != 0) ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty[s]);
                    }
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ===== epilogue: warps 2..9; lane quarter q, column blocks {2*half, 2*half + 1} =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int m = q * 32 + lane;
        const int w = m % p.WT, h = (m / p.WT) % p.HT, n = m / (p.WT * p.HT);
        const int t = threadIdx.x - 64, cp = t & 31, rg = t >> 5;
        int lt = 0;
        int stat_key = -1;
        float ss[4] = {0.f, 0.f, 0.f, 0.f};
        auto ss_flush = [&](int key) {
            atomicAdd(&sstat[(2 * cp) * 2], ss[0]);     atomicAdd(&sstat[(2 * cp) * 2 + 1], ss[1]);
            atomicAdd(&sstat[(2 * cp + 1) * 2], ss[2]); atomicAdd(&sstat[(2 * cp + 1) * 2 + 1], ss[3]);
            ss[0] = ss[1] = ss[2] = ss[3] = 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            if (t < 2 * BN) {
                const int col = (key % p.n_tiles) * BN + (t >> 1);
                if (col < p.OC) atomicAdd(&stats[((long long)(key / p.n_tiles) * p.OC + col) * 2 + (t & 1)], (double)sstat[t]);
                sstat[t] = 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        };
        for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
            VS_TCF_DECODE(idx)
            const int acc = lt & 1;
            const int i = i0 + h, j = j0 + w, nn = b0 + n;
            const bool ok = i < p.OHc && j < p.OWc && nn < p.N;
            const bool full_tile = i0 + p.HT <= p.OHc && j0 + p.WT <= p.OWc && b0 + p.NT <= p.N;
            if (stats != nullptr) {
                const int key = (b0 / p.n_per_group) * p.n_tiles + n0 / BN;
                if (key != stat_key) {
                    if (stat_key >= 0) ss_flush(stat_key);
                    stat_key = key;
                }
            }
            mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
            tc_fence_after();
            // the previous item's TMA stores (and everybody's statistics reads) are done with the staging tiles
            if (threadIdx.x == 64) tma_store_wait_read();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int pb = 2 * half + (cc >> 1), c0 = (cc & 1) * 32;
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC + pb * BN + c0), r);
                float xs[32];
                if (p.has_bias) {          // uniform: dgrad launches carry no bias and skip the 32 shuffles
                    const float bias_l = (p.has_bias && n0 + c0 + lane < p.OC) ? __ldg(bias + n0 + c0 + lane) : 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]) + __shfl_sync(0xffffffffu, bias_l, c);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]);
                }
                if (!full_tile && !ok) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = 0.f;
                }
                stage_row32(obuf + pb * (TC_BM * 128), m, c0, xs, p.act);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            if (threadIdx.x == 64) {
#pragma unroll
                for (int pb = 0; pb < 4; ++pb) tma_store_4d(&omaps.m[tab.pos_cls[pb]], obuf + pb * (TC_BM * 128), n0, j0, i0, b0);
                tma_store_commit();
            }
            if (stats != nullptr) {
#pragma unroll 1
                for (int pb = 0; pb < 4; ++pb) {
                    const uint8_t* tile = obuf + pb * (TC_BM * 128) + (cp & 3) * 4;
#pragma unroll
                    for (int rr = 0; rr < TC_BM / TC_EPI_WARPS; ++rr) {
                        const int row = rg * (TC_BM / TC_EPI_WARPS) + rr;
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(tile + row * 128 + (((cp >> 2) ^ (row & 7)) << 4));
                        const float a = __uint_as_float(v << 16), b = __uint_as_float(v & 0xffff0000u);
                        ss[0] += a; ss[1] = fmaf(a, a, ss[1]);
                        ss[2] += b; ss[3] = fmaf(b, b, ss[3]);
                    }
                }
            }
        }
        if (stats != nullptr && stat_key >= 0) ss_flush(stat_key);
        if (threadIdx.x == 64) tma_store_wait_all();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * ACC>(tmem_base);
    }
#undef VS_TCF_DECODE
}

// ------------------------------------------------------------------------------------------ CTA pairs
// cta_group::2 variant for the wide layers (OC a multiple of 128): the two CTAs of a cluster (one TPC) work on two
// adjacent 128-pixel tiles and the SAME BN output channels.  Each CTA stages its own A box and only HALF of the weight
// slab; one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256 over the pair, N = BN), which reads both
// shared memories and writes both tensor memories.  Per 64-channel chunk a CTA stages 16 + BN/8 KB instead of
// 16 + BN/4 KB for the same 128 x BN x 64 MACs: the L2 -> shared-memory fill that bounds the single-CTA kernel drops by
// 25 % (BN = 128) to 33 % (BN = 256, which one CTA cannot hold with double-buffered accumulators).
// One CTA per SM (both TMEM accumulator stages of BN = 256 take all 512 columns), deeper smem ring instead.
// RESB: the weight slabs of ALL (tap, 64-channel chunk) steps stay resident in shared memory for the whole kernel (each CTA
// of the pair holds its half: R*S*kchunks slabs of BN/2 rows x 128 B, at most 128 KB) and the ring stages carry only the
// input boxes.  The narrow layers are bound by what an SM can ingest from L2: this removes the weight
// bytes from every step — 16 KB instead of 24 KB (BN = 64) / 24 KB (BN = 128, already halved by the pair) per 128 x BN x 64 block.
constexpr int TC_RES_BYTES = 128 * 1024;
template <int BN, int STAGES, bool RESB = false>
struct TcPairSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2, B_BYTES = (BN / 2) * TC_BK * 2;
    static constexpr int STAGE_BYTES = RESB ? A_BYTES : A_BYTES + B_BYTES;
    static constexpr int RES_OFF = STAGES * STAGE_BYTES;
    static constexpr int OUT_OFF = RES_OFF + (RESB ? TC_RES_BYTES : 0);      // RESB: one staged bf16 output tile (TMA store)
    static constexpr int OUT_BYTES = RESB ? (BN / 64) * TC_BM * 128 : 0;
    static constexpr int BAR_OFF = OUT_OFF + OUT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 1024 /*align slack*/ + 256 /*barriers*/ + BN * 2 * 4 /*tile statistics*/;
    static_assert((2 * STAGES + 6) * 8 <= 256, "barrier area");
    static_assert(TOTAL <= 227 * 1024, "shared memory");
};

template <int BN, int STAGES, bool RESB = false>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_pair_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_b,
                                                                     const __grid_constant__ TcOutMaps omaps,
                                                                     const __grid_constant__ TcParams p,
                                                                     const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                                     double* __restrict__ stats) {
    using S = TcPairSmem<BN, STAGES, RESB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bres = smem + S::RES_OFF;              // RESB: resident weight slabs [R*S*kchunks][BN/2 rows][128 B]
    uint8_t* obuf = smem + S::OUT_OFF;              // RESB: BN/64 half-tiles of [128 pixels][128 bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;                          // leader's: both CTAs' TMA bytes of a stage have landed
    uint64_t* empty = bars + STAGES;                // each CTA's: the MMAs that read this stage have completed
    uint64_t* tmem_full = bars + 2 * STAGES;        // [2] each CTA's: accumulator stage complete
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2] leader's: both CTAs' epilogues have drained the stage
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint64_t* bres_full = bars + 2 * STAGES + 5;    // leader's: both CTAs' resident slabs have landed
    float* sstat = reinterpret_cast<float*>(smem + S::BAR_OFF + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cl_rank();
    const bool leader = rank == 0;
    const int nkb = p.ntaps * p.kchunks;
    // work items of the PAIR: (class, channel tile, pair of pixel tiles); contiguous range per pair
    const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int t_begin = (int)((long long)pair * p.total_tiles / npairs);
    const int t_end = (int)((long long)(pair + 1) * p.total_tiles / npairs);

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 2 * TC_EPI_WARPS); }
        mbar_init(bres_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) sstat[i] = 0.f;
    __syncthreads();
    if (warp == 1) tmem_alloc_pair<2 * BN>(tmem_slot);
    tc_fence_before();
    cl_sync();                 // barriers of both CTAs initialised, both tensor memories allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's pixel tile = 2 * pair index + rank (an odd tile count leaves the last rank-1 tile beyond the batch:
    // its TMA boxes are zero-filled and its epilogue rows are masked)
#define VS_TCP_DECODE(idx)                                                  \
    const int cls = (idx) % p.classes;                                      \
    const int rest_ = (idx) / p.classes;                                    \
    const int n0 = (rest_ % p.n_tiles) * BN;                                \
    int t_ = (rest_ / p.n_tiles) * 2 + rank;                                \
    const int tw = t_ % p.tiles_w; t_ /= p.tiles_w;                         \
    const int th = t_ % p.tiles_h;                                          \
    const int tn = t_ / p.tiles_h;                                          \
    const int j0 = tw * p.WT, i0 = th * p.HT, b0 = tn * p.NT;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A box + own half of the weight slab, completion on the LEADER's barrier =====
        if (lane == 0) {
            if (RESB) {
                // all weight slabs of the layer, once: slab (filter tap, channel chunk) -> this CTA's BN/2 rows of it
                const int nslab = p.res_taps * p.kchunks;
                if (leader) mbar_expect_tx(bres_full, (uint32_t)(2 * nslab * S::B_BYTES));
                const uint32_t rbar = cl_map(bres_full, 0);
                for (int sl = 0; sl < nslab; ++sl)
                    tma_load_2d_pair(bres + sl * S::B_BYTES, &map_b, rbar, (sl / p.kchunks) * p.IC + (sl % p.kchunks) * TC_BK, rank * (BN / 2));
            }
            int it = 0;
            for (int idx = t_begin; idx < t_end; ++idx) {
                VS_TCP_DECODE(idx)
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
                    uint8_t* a_dst = smem + s * S::STAGE_BYTES;
                    uint8_t* b_dst = a_dst + S::A_BYTES;
                    if (leader) mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);
                    const uint32_t bar = cl_map(&full[s], 0);
                    tma_load_4d_pair(a_dst, &map_a, bar, kc * TC_BK, j0 * p.in_sw + p.dw[cls * p.ntaps + tap],
                                     i0 * p.in_sh + p.dh[cls * p.ntaps + tap], b0);
                    if (!RESB)
                        tma_load_2d_pair(b_dst, &map_b, bar, p.wtap[cls * p.ntaps + tap] * p.IC + kc * TC_BK, n0 + rank * (BN / 2));
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader CTA's warp 1, converged (see tc_conv_kernel); one elected lane issues =====
        if (leader) {
            constexpr uint32_t idesc = idesc_bf16_f32_pair(BN);
            const bool elected = elect_one();
            if (RESB) mbar_wait(bres_full, 0);
            const uint32_t bres_lo = kmajor_sw128_desc_lo(smem_u32(bres));
            int s = 0, lt = 0;
            uint32_t ph = 0;
            for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
                const int acc = lt & 1;
                const int cls_m = idx % p.classes;
                mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_lo = kmajor_sw128_desc_lo(smem_u32(smem + s * S::STAGE_BYTES));
                    const int tap_m = kb / p.kchunks;
                    const uint32_t b_lo = RESB ? bres_lo + (uint32_t)((p.wtap[cls_m * p.ntaps + tap_m] * p.kchunks + (kb - tap_m * p.kchunks)) * (S::B_BYTES / 16))
                                               : a_lo + S::A_BYTES / 16;
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t ad = kmajor_sw128_desc_from_lo(a_lo + 2 * k), bd = kmajor_sw128_desc_from_lo(b_lo + 2 * k);
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                        if (elected) umma_bf16_pair(tmem_d, ad, bd, idesc, accum);
                    }
                    if (elected) umma_commit_pair(&empty[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (elected) umma_commit_pair(&tmem_full[acc]);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue (both CTAs): own 128 pixel rows of the pair tile =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = BN / (TC_EPI_WARPS / 4);
        const int m = q * 32 + lane;
        const int w = m % p.WT, h = (m / p.WT) % p.HT, n = m / (p.WT * p.HT);
        int lt = 0;
        int stat_key = -1;
        if constexpr (RESB) {
            // staged epilogue (as tc_conv_kernel<.., TS, SS>): the tile goes to shared memory as bf16, one thread writes it
            // with TMA stores, and the BatchNorm sums are read back from the staged tile.  One staging tile: the previous
            // item's store must have read it (and everybody's statistics reads be done) before it is refilled.
            constexpr int HALVES = BN / 64;
            const int t = threadIdx.x - 64, cp = t & 31, rg = t >> 5;
            float ss[HALVES][4] = {};
            auto ss_flush = [&](int key) {
#pragma unroll
                for (int hf = 0; hf < HALVES; ++hf) {
                    const int c = hf * 64 + 2 * cp;
                    atomicAdd(&sstat[c * 2], ss[hf][0]);       atomicAdd(&sstat[c * 2 + 1], ss[hf][1]);
                    atomicAdd(&sstat[(c + 1) * 2], ss[hf][2]); atomicAdd(&sstat[(c + 1) * 2 + 1], ss[hf][3]);
                    ss[hf][0] = ss[hf][1] = ss[hf][2] = ss[hf][3] = 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                if (t < 2 * BN) {
                    atomicAdd(&stats[((long long)key * p.OC + (t >> 1)) * 2 + (t & 1)], (double)sstat[t]);      // OC == BN
                    sstat[t] = 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            };
            for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
                VS_TCP_DECODE(idx)
                const int acc = lt & 1;
                const int i = i0 + h, j = j0 + w, nn = b0 + n;
                const bool ok = i < p.OHc && j < p.OWc && nn < p.N;
                const bool full_tile = i0 + p.HT <= p.OHc && j0 + p.WT <= p.OWc && b0 + p.NT <= p.N;
                if (stats != nullptr && b0 < p.N) {
                    const int key = b0 / p.n_per_group;
                    if (key != stat_key) {
                        if (stat_key >= 0) ss_flush(stat_key);
                        stat_key = key;
                    }
                }
                mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
                tc_fence_after();
                if (threadIdx.x == 64) tma_store_wait_read();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
#pragma unroll 1
                for (int c0 = half * COLS_PER_WARP; c0 < (half + 1) * COLS_PER_WARP; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                    float xs[32];
                    if (p.has_bias) {
                        const float bias_l = __ldg(bias + n0 + c0 + lane);
#pragma unroll
                        for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]) + __shfl_sync(0xffffffffu, bias_l, c);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]);
                    }
                    if (!full_tile && !ok) {          // rows beyond the tensor edge: clipped by the TMA store, must not count
#pragma unroll
                        for (int c = 0; c < 32; ++c) xs[c] = 0.f;
                    }
                    stage_row32(obuf + (c0 >> 6) * (TC_BM * 128), m, c0 & 63, xs, p.act);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(cl_map(&tmem_empty[acc], 0));      // ordered by the tcgen05 fences
                fence_proxy_async();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                if (threadIdx.x == 64 && b0 < p.N) {
#pragma unroll
                    for (int hf = 0; hf < HALVES; ++hf) tma_store_4d(&omaps.m[cls], obuf + hf * (TC_BM * 128), n0 + hf * 64, j0, i0, b0);
                    tma_store_commit();
                }
                if (stats != nullptr && b0 < p.N) {
#pragma unroll
                    for (int hf = 0; hf < HALVES; ++hf) {
                        const uint8_t* tile = obuf + hf * (TC_BM * 128) + (cp & 3) * 4;
#pragma unroll
                        for (int rr = 0; rr < TC_BM / TC_EPI_WARPS; ++rr) {
                            const int row = rg * (TC_BM / TC_EPI_WARPS) + rr;
                            const uint32_t v = *reinterpret_cast<const uint32_t*>(tile + row * 128 + (((cp >> 2) ^ (row & 7)) << 4));
                            const float a = __uint_as_float(v << 16), b = __uint_as_float(v & 0xffff0000u);
                            ss[hf][0] += a; ss[hf][1] = fmaf(a, a, ss[hf][1]);
                            ss[hf][2] += b; ss[hf][3] = fmaf(b, b, ss[hf][3]);
                        }
                    }
                }
            }
            if (stats != nullptr && stat_key >= 0) ss_flush(stat_key);
            if (threadIdx.x == 64) tma_store_wait_all();
        } else {
        constexpr int NE = (2 * BN + 32 * TC_EPI_WARPS - 1) / (32 * TC_EPI_WARPS);      // statistics entries per epilogue thread
        double stat_acc[NE] = {};
        for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
            VS_TCP_DECODE(idx)
            const int acc = lt & 1;
            const int i = i0 + h, j = j0 + w, nn = b0 + n;
            const bool ok = i < p.OHc && j < p.OWc && nn < p.N;
            const long long pix = ((long long)nn * p.OH + (long long)i * p.ost + p.ca[cls]) * p.OW + (long long)j * p.ost + p.cb[cls];
            __nv_bfloat16* dst = out + pix * p.OC + n0;
            mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = half * COLS_PER_WARP; c0 < (half + 1) * COLS_PER_WARP; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                float xs[32];
                if (p.has_bias) {          // uniform: dgrad launches carry no bias and skip the 32 shuffles
                    const float bias_l = p.has_bias ? __ldg(bias + n0 + c0 + lane) : 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]) + __shfl_sync(0xffffffffu, bias_l, c);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]);
                }
                if (ok) {
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
                if (stats != nullptr) {
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
            // hand the accumulator stage back to the leader's MMA issuer (its barrier counts the warps of both CTAs)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(cl_map(&tmem_empty[acc], 0));      // ordered by the tcgen05 fences
            if (stats != nullptr) {
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                const int t = threadIdx.x - 64;
                const int key = (b0 / p.n_per_group) * p.n_tiles + n0 / BN;
                if (key != stat_key) {
                    if (stat_key >= 0) {
#pragma unroll
                        for (int e = 0; e < NE; ++e) {
                            const int ii = t + e * 32 * TC_EPI_WARPS, col = (stat_key % p.n_tiles) * BN + (ii >> 1);
                            if (ii < 2 * BN) atomicAdd(&stats[((long long)(stat_key / p.n_tiles) * p.OC + col) * 2 + (ii & 1)], stat_acc[e]);
                            stat_acc[e] = 0.0;
                        }
                    }
                    stat_key = key;
                }
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int ii = t + e * 32 * TC_EPI_WARPS;
                    if (ii < 2 * BN) { stat_acc[e] += (double)sstat[ii]; sstat[ii] = 0.f; }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            }
        }
        if (stats != nullptr && stat_key >= 0) {
            const int t = threadIdx.x - 64;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const int ii = t + e * 32 * TC_EPI_WARPS, col = (stat_key % p.n_tiles) * BN + (ii >> 1);
                if (ii < 2 * BN) atomicAdd(&stats[((long long)(stat_key / p.n_tiles) * p.OC + col) * 2 + (ii & 1)], stat_acc[e]);
            }
        }
        }
    }
    // the leader's MMAs read the peer's shared memory and write its tensor memory: nobody leaves before everybody is done
    tc_fence_before();
    cl_sync();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair<2 * BN>(tmem_base);
    }
#undef VS_TCP_DECODE
}

// ------------------------------------------------------------------------------------------ shifted windows
// The 64/128-channel k4 s2 p1 layers at 16x16 / 32x32 (decoder 128->64 and its input gradient, encoder 64->128 and its
// input gradient) are bound by what the tensor core's operands cost on the way in: per 128 x 64 x 64 block the per-class
// kernels above stage 16 KB of input + 8 KB of weights from L2, 125-190 B/clk/SM at the tensor pipe's rate, and the L2
// sits at 62 % of its throughput (profiles/r02_ncu_per_launch_tc_conv.txt).  This kernel removes most of those bytes:
//   * CTA pairs (cta_group::2, M = 256) with ALL weight slabs resident in shared memory (each CTA its half, 128 KB),
//   * a pixel tile of 16 rows x 8 columns of ONE image, so that a pixel row of the tile is exactly one 1024-byte swizzle
//     atom: an input box of 17-18 rows is loaded once and the taps that differ only by a row shift read it through
//     descriptors offset by 1024 bytes per row (no re-load),
//   * transposed mode: the four output-parity classes of the tile are accumulated together (4 x 64 TMEM columns) and a
//     shift shared by 2 or 4 classes is ONE N = 128 / 256 instruction (schedule of the fused-class kernel above).
// Input bytes per work item: 3 column shifts x 18 KB per 64-channel chunk (transposed) or 8 x 17 KB (direct, stride 2:
// two row parities x four filter columns) instead of 16 x 16 KB, no weight bytes: 26-33 B/clk/SM at full tensor rate.
// The MMA schedule comes from the host as a flat table (box -> groups of {row shift, resident offset, descriptor,
// accumulator column}); the issuer does four loads from the constant bank per group of four MMAs.
constexpr int TCS_MAX_BOX = 8, TCS_MAX_GRP = 24, TCS_PIECES = 32, TCS_HT = 16, TCS_WT = 8, TCS_ACC = 256;
struct TcShiftTab {
    int nbox, ngrp, nblk, hb;                  // boxes / MMA groups per work item, 64-column accumulator blocks, rows per box
    int npieces;                               // resident 32-row weight pieces per CTA (4 KB each)
    short box_dx[TCS_MAX_BOX], box_dy[TCS_MAX_BOX], box_c0[TCS_MAX_BOX];     // input offset of the box, first channel
    unsigned char box_g0[TCS_MAX_BOX + 1];     // groups [g0[b], g0[b+1]) read box b
    // per group: x = offset of its first pixel row inside the box (1024 bytes x row shift), y = offset of its weight rows in
    // the resident area, both in 16-byte descriptor units; z = instruction descriptor (N = 64 x merged classes, or OC);
    // w = first accumulator column
    uint4 grp[TCS_MAX_GRP];
    int piece_x[2][TCS_PIECES], piece_y[2][TCS_PIECES];      // [CTA rank][piece]: coordinates in the packed weights
    unsigned char blk_cls[4], blk_ch0[4];      // accumulator block -> output map (parity class), first output channel
};

template <int STAGES>
struct TcShiftSmem {
    static constexpr int STAGE_BYTES = 18 * 1024;
    static constexpr int RES_OFF = STAGES * STAGE_BYTES;
    static constexpr int OUT_OFF = RES_OFF + TC_RES_BYTES;
    static constexpr int BAR_OFF = OUT_OFF + 2 * TC_BM * 128;      // two staged 128-pixel x 64-channel bf16 tiles
    static constexpr int TOTAL = BAR_OFF + 1024 /*align slack*/ + 256 /*barriers*/ + 128 * 2 * 4 /*column sums*/;
    static_assert((2 * STAGES + 6) * 8 <= 256, "barrier area");
    static_assert(TOTAL <= 227 * 1024, "shared memory");
};

template <int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_shift_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                      const __grid_constant__ CUtensorMap map_b,
                                                                      const __grid_constant__ TcOutMaps omaps,
                                                                      const __grid_constant__ TcParams p,
                                                                      const __grid_constant__ TcShiftTab tab,
                                                                      const float* __restrict__ bias, double* __restrict__ stats) {
    using S = TcShiftSmem<STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bres = smem + S::RES_OFF;
    uint8_t* obuf = smem + S::OUT_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;                          // leader's: both CTAs' box of a stage has landed
    uint64_t* empty = bars + STAGES;                // each CTA's: the MMAs that read this stage have completed
    uint64_t* tmem_full = bars + 2 * STAGES;        // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2] leader's
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint64_t* bres_full = bars + 2 * STAGES + 5;    // leader's: both CTAs' resident weights have landed
    float* sstat = reinterpret_cast<float*>(smem + S::BAR_OFF + 256);      // [OC][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cl_rank();
    const bool leader = rank == 0;
    const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int t_begin = (int)((long long)pair * p.total_tiles / npairs);
    const int t_end = (int)((long long)(pair + 1) * p.total_tiles / npairs);

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        // accumulator hand-back: one arrival per epilogue group that owns blocks, of both CTAs
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], tab.nblk > 1 ? 4 : 2); }
        mbar_init(bres_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * 128; i += blockDim.x) sstat[i] = 0.f;
    __syncthreads();
    if (warp == 1) tmem_alloc_pair<2 * TCS_ACC>(tmem_slot);
    tc_fence_before();
    cl_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item of the pair -> this CTA's tile (16 rows x 8 columns of one image's class grid); an odd tile count leaves
    // the last rank-1 tile beyond the batch: its boxes are zero-filled, its stores and statistics skipped
#define VS_TCS_DECODE(idx)                                                  \
    int t_ = (idx) * 2 + rank;                                              \
    const int j0 = (t_ % p.tiles_w) * TCS_WT; t_ /= p.tiles_w;              \
    const int i0 = (t_ % p.tiles_h) * TCS_HT;                               \
    const int nn = t_ / p.tiles_h;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): the resident weight pieces once, then one input box per stage =====
        if (lane == 0) {
            if (leader) mbar_expect_tx(bres_full, (uint32_t)(2 * tab.npieces * 4096));
            const uint32_t rbar = cl_map(bres_full, 0);
            for (int i = 0; i < tab.npieces; ++i)
                tma_load_2d_pair(bres + i * 4096, &map_b, rbar, tab.piece_x[rank][i], tab.piece_y[rank][i]);
            int it = 0;
            for (int idx = t_begin; idx < t_end; ++idx) {
                // the ring holds three boxes (~54 KB in flight per SM, too little to cover DRAM latency at full bandwidth):
                // the NEXT item's boxes are requested into L2 now, one whole item ahead of their shared-memory loads
                if (idx + 1 < t_end) {
                    int t2 = (idx + 1) * 2 + rank;
                    const int pj0 = (t2 % p.tiles_w) * TCS_WT; t2 /= p.tiles_w;
                    const int pi0 = (t2 % p.tiles_h) * TCS_HT, pn = t2 / p.tiles_h;
                    if (pn < p.N)
                        for (int b = 0; b < tab.nbox; ++b)
                            tma_prefetch_4d(&map_a, tab.box_c0[b], pj0 * p.in_sw + tab.box_dx[b], pi0 * p.in_sh + tab.box_dy[b], pn);
                }
                VS_TCS_DECODE(idx)
                for (int b = 0; b < tab.nbox; ++b, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    if (leader) mbar_expect_tx(&full[s], (uint32_t)(2 * tab.hb * 1024));
                    tma_load_4d_pair(smem + s * S::STAGE_BYTES, &map_a, cl_map(&full[s], 0), tab.box_c0[b],
                                     j0 * p.in_sw + tab.box_dx[b], i0 * p.in_sh + tab.box_dy[b], nn);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader CTA's warp 1, converged; one elected lane issues =====
        if (leader) {
            const bool elected = elect_one();
            mbar_wait(bres_full, 0);
            const uint32_t b_lo0 = kmajor_sw128_desc_lo(smem_u32(bres));
            int s = 0, lt = 0;
            uint32_t ph = 0;
            for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
                const int acc = lt & 1;
                mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TCS_ACC);
                for (int b = 0; b < tab.nbox; ++b) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_lo0 = kmajor_sw128_desc_lo(smem_u32(smem + s * S::STAGE_BYTES));
                    const int g1 = tab.box_g0[b + 1];
                    for (int g = tab.box_g0[b]; g < g1; ++g) {
                        const uint4 e = tab.grp[g];
                        const uint32_t a_lo = a_lo0 + e.x, b_lo = b_lo0 + e.y, d = tmem_d + e.w;
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {      // the first group of an item covers every accumulator column
                            const uint64_t ad = kmajor_sw128_desc_from_lo(a_lo + 2 * k), bd = kmajor_sw128_desc_from_lo(b_lo + 2 * k);
                            const uint32_t accum = (g | k) != 0 ? 1u : 0u;
                            if (elected) umma_bf16_pair(d, ad, bd, e.z, accum);
                        }
                    }
                    if (elected) umma_commit_pair(&empty[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (elected) umma_commit_pair(&tmem_full[acc]);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue (both CTAs): own 128 pixels.  Two groups of four warps (one warp per TMEM lane quarter), each with
        // its own staging tile, named barrier and store-issuing thread; group gi takes the 64-column accumulator blocks
        // gi, gi + 2: TMEM load -> bf16 tile in shared memory -> one TMA store, sums read back from the staged tile.
        // The two groups' barrier / TMEM-load / store latencies overlap =====
        const int q = warp & 3;
        const int gi = (warp - 2) >> 2;
        const int m = q * 32 + lane;                  // pixel (m / 8, m % 8) of the tile
        const int T = threadIdx.x - 64, t = T & 127, cp = t & 31, rg = t >> 5;
        uint8_t* tile_w = obuf + gi * (TC_BM * 128);
        int lt = 0;
        int stat_key = -1;
        float ss0[4] = {0.f, 0.f, 0.f, 0.f}, ss1[4] = {0.f, 0.f, 0.f, 0.f};      // channels 2cp, 2cp+1 of [0,64) / [64,128)
        auto group_sync = [&]() {
            if (gi == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else         asm volatile("bar.sync 2, 128;" ::: "memory");
        };
        auto ss_flush = [&](int key) {
            atomicAdd(&sstat[(2 * cp) * 2], ss0[0]);     atomicAdd(&sstat[(2 * cp) * 2 + 1], ss0[1]);
            atomicAdd(&sstat[(2 * cp + 1) * 2], ss0[2]); atomicAdd(&sstat[(2 * cp + 1) * 2 + 1], ss0[3]);
            if (p.OC > 64) {
                atomicAdd(&sstat[(64 + 2 * cp) * 2], ss1[0]);     atomicAdd(&sstat[(64 + 2 * cp) * 2 + 1], ss1[1]);
                atomicAdd(&sstat[(64 + 2 * cp + 1) * 2], ss1[2]); atomicAdd(&sstat[(64 + 2 * cp + 1) * 2 + 1], ss1[3]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) ss0[e] = ss1[e] = 0.f;
            asm volatile("bar.sync 3, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            if (T < 2 * p.OC) {
                atomicAdd(&stats[((long long)key * p.OC + (T >> 1)) * 2 + (T & 1)], (double)sstat[T]);
                sstat[T] = 0.f;
            }
            asm volatile("bar.sync 3, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        };
        for (int idx = t_begin; idx < t_end; ++idx, ++lt) {
            VS_TCS_DECODE(idx)
            const int acc = lt & 1;
            const bool live = nn < p.N;
            if (stats != nullptr && live) {
                const int key = nn / p.n_per_group;
                if (key != stat_key) {                       // uniform over the CTA's epilogue warps
                    if (stat_key >= 0) ss_flush(stat_key);
                    stat_key = key;
                }
            }
            mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int blk = gi; blk < tab.nblk; blk += 2) {
                const int ch0 = tab.blk_ch0[blk];
                // the group's previous store has read the staging tile, and the group's statistics reads of it are done
                if (t == 0) tma_store_wait_read();
                group_sync();
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TCS_ACC + blk * 64 + half * 32), r);
                    float xs[32];
                    if (p.has_bias) {
                        const float bias_l = __ldg(bias + ch0 + half * 32 + lane);
#pragma unroll
                        for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]) + __shfl_sync(0xffffffffu, bias_l, c);
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) xs[c] = __uint_as_float(r[c]);
                    }
                    stage_row32(tile_w, m, half * 32, xs, p.act);
                }
                tc_fence_before();
                fence_proxy_async();
                group_sync();
                if (t == 0) {
                    if (live) {
                        tma_store_4d(&omaps.m[tab.blk_cls[blk]], tile_w, ch0, j0, i0, nn);
                        tma_store_commit();
                    }
                    // the group's warps have read its last block of the accumulator stage (barrier above): one arrival per
                    // group on the leader's barrier; no memory ordering needed, the tcgen05 fences order the TMEM accesses
                    if (blk + 2 >= tab.nblk) { tc_fence_after(); mbar_arrive_cluster_relaxed(cl_map(&tmem_empty[acc], 0)); }
                }
                if (stats != nullptr && live) {
                    const uint8_t* tile = tile_w + (cp & 3) * 4;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
                    for (int rr = 0; rr < TC_BM / 4; ++rr) {
                        const int row = rg * (TC_BM / 4) + rr;
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(tile + row * 128 + (((cp >> 2) ^ (row & 7)) << 4));
                        const float a = __uint_as_float(v << 16), b = __uint_as_float(v & 0xffff0000u);
                        a0 += a; a1 = fmaf(a, a, a1);
                        a2 += b; a3 = fmaf(b, b, a3);
                    }
                    if (ch0 == 0) { ss0[0] += a0; ss0[1] += a1; ss0[2] += a2; ss0[3] += a3; }
                    else          { ss1[0] += a0; ss1[1] += a1; ss1[2] += a2; ss1[3] += a3; }
                }
            }
        }
        if (stats != nullptr && stat_key >= 0) ss_flush(stat_key);
        if (t == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cl_sync();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair<2 * TCS_ACC>(tmem_base);
    }
#undef VS_TCS_DECODE
}

// ------------------------------------------------------------------------------------------ host
template <int BN, int STAGES, bool TS, bool SS = false>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcOutMaps& om, const TcParams& p, const float* bias,
                     void* out, int classes, double* stats, cudaStream_t stream) {
    using S = TcSmem<BN, STAGES, TS>;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<BN, STAGES, TS, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail("tc_conv_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    TcParams q = p;
    q.classes = classes;
    q.n_tiles = (int)cdiv(p.OC, BN);
    q.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * q.n_tiles * classes;
    // __launch_bounds__(TC_THREADS, 2); the deep-ring staged 128-column variant fills an SM's shared memory alone
    const int resident = (S::TOTAL > 113 * 1024 ? 1 : 2) * num_sms();
    const int grid = q.total_tiles < resident ? q.total_tiles : resident;      // every CTA gets at least one work item
    tc_conv_kernel<BN, STAGES, TS, SS><<<grid, TC_THREADS, S::TOTAL, stream>>>(ma, mb, om, q, bias, (__nv_bfloat16*)out, stats);
    return launched("tc_conv_kernel");
}

static bool staged_epilogue_disabled();
static bool pair_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_DISABLE_PAIR"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <int BN, int STAGES, bool RESB = false>
static int launch_tc_pair(const CUtensorMap& ma, const CUtensorMap& mb, const TcOutMaps& om, const TcParams& p, const float* bias, void* out,
                          int classes, double* stats, cudaStream_t stream) {
    using S = TcPairSmem<BN, STAGES, RESB>;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_pair_kernel<BN, STAGES, RESB>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail("tc_conv_pair_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    TcParams q = p;
    q.classes = classes;
    q.n_tiles = p.OC / BN;
    const int ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
    q.total_tiles = ((ptiles + 1) / 2) * q.n_tiles * classes;          // work items of a PAIR
    int pairs = num_sms() / 2;                                         // one CTA per SM
    if (pairs > q.total_tiles) pairs = q.total_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_conv_pair_kernel<BN, STAGES, RESB>, ma, mb, om, q, bias, (__nv_bfloat16*)out, stats);
    if (e != cudaSuccess) return fail("tc_conv_pair_kernel launch: %s", cudaGetErrorString(e));
    return launched("tc_conv_pair_kernel");
}

static bool staged_stats_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_DISABLE_STAGED_STATS"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// The fused-class kernel is correct (tests/test_kernels_gpu.py runs it) but measured SLOWER than the per-class kernel
// on the 128->64 decoder layer: 488 us with a table-driven MMA issuer (~250 cycles of dependent scalar work per MMA,
// tensor pipe 25 % busy), 275 us with the compile-time schedule, against 223 us for four per-class launches' worth of
// work in tc_conv_kernel.  With all of TMEM and a 64 KB staging area it runs ONE CTA per SM with three stages in
// flight and a single epilogue, where the per-class kernel overlaps two CTAs (six stages) per SM.  Opt-in
// (VARSEP_ENABLE_FUSED_CLASSES=1) until it gets a deeper ring / resident weight slabs.
static bool fused_classes_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_ENABLE_FUSED_CLASSES"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

// shift table of the fused-class kernel from the per-class tap tables of a k4 s2 p1 transposed convolution;
// false when the geometry does not have the 3x3-shifts / 2x2-per-class structure
static bool build_fused_tab(const TcParams& p, int classes, TcFusedTab& tab) {
    if (classes != 4 || p.ntaps != 4 || p.ost != 2) return false;
    memset(&tab, 0, sizeof(tab));
    static const int cls_pos[4] = {0, 1, 3, 2};          // class a*2+b -> column block: (0,0) (0,1) (1,1) (1,0)
    for (int cls = 0; cls < 4; ++cls) {
        if (p.ca[cls] * 2 + p.cb[cls] != cls) return false;
        tab.pos_cls[cls_pos[cls]] = (unsigned char)cls;
    }
    int uses[TCF_SHIFTS][4];                              // [shift][pos] = packed tap, or -1
    int sdh[TCF_SHIFTS], sdw[TCF_SHIFTS], ns = 0;
    for (int cls = 0; cls < 4; ++cls)
        for (int t = 0; t < 4; ++t) {
            const int dh = p.dh[cls * 4 + t], dw = p.dw[cls * 4 + t];
            int sh = -1;
            for (int i = 0; i < ns; ++i) if (sdh[i] == dh && sdw[i] == dw) sh = i;
            if (sh < 0) {
                if (ns == TCF_SHIFTS) return false;
                sh = ns++; sdh[sh] = dh; sdw[sh] = dw;
                for (int q = 0; q < 4; ++q) uses[sh][q] = -1;
            }
            if (uses[sh][cls_pos[cls]] >= 0) return false;
            uses[sh][cls_pos[cls]] = p.wtap[cls * 4 + t];
        }
    // canonical order of the device-side schedule (TCF_* tables); every shift must have exactly the expected MMAs
    static const signed char cdh[TCF_SHIFTS] = {0, -1, 0, 1, 0, -1, -1, 1, 1}, cdw[TCF_SHIFTS] = {0, 0, 1, 0, -1, -1, 1, 1, -1};
    static const unsigned char cng[TCF_SHIFTS] = {1, 1, 1, 1, 2, 1, 1, 1, 1};
    static const unsigned char cn[TCF_SHIFTS][2] = {{4, 0}, {2, 0}, {2, 0}, {2, 0}, {1, 1}, {1, 0}, {1, 0}, {1, 0}, {1, 0}};
    static const unsigned char cpos[TCF_SHIFTS][2] = {{0, 0}, {0, 0}, {1, 0}, {2, 0}, {0, 3}, {0, 0}, {1, 0}, {2, 0}, {3, 0}};
    static const unsigned char cslab[TCF_SHIFTS][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 1}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
    if (ns != TCF_SHIFTS) return false;
    tab.nshift = ns;
    for (int o = 0; o < ns; ++o) {
        int i = -1;
        for (int j = 0; j < ns; ++j) if (sdh[j] == cdh[o] && sdw[j] == cdw[o]) i = j;
        if (i < 0) return false;
        tab.dh[o] = (signed char)sdh[i]; tab.dw[o] = (signed char)sdw[i];
        int nslab = 0, ng = 0;
        for (int q = 0; q < 4; ++q) {
            if (uses[i][q] < 0) continue;
            tab.slab_wtap[o][nslab] = (unsigned char)uses[i][q];
            if (q > 0 && uses[i][q - 1] >= 0) {           // adjacent to the previous slab's column block: same MMA
                tab.grp_n[o][ng - 1]++;
            } else {
                tab.grp_slab[o][ng] = (unsigned char)nslab; tab.grp_pos[o][ng] = (unsigned char)q; tab.grp_n[o][ng] = 1;
                ++ng;
            }
            ++nslab;
        }
        if (ng != cng[o]) return false;
        for (int g = 0; g < ng; ++g)
            if (tab.grp_n[o][g] != cn[o][g] || tab.grp_pos[o][g] != cpos[o][g] || tab.grp_slab[o][g] != cslab[o][g]) return false;
        tab.nslab[o] = (unsigned char)nslab; tab.ngrp[o] = (unsigned char)ng;
    }
    return true;
}

template <int STAGES>
static int launch_tc_fused(const CUtensorMap& ma, const CUtensorMap& mb, const TcOutMaps& om, const TcParams& p, const TcFusedTab& tab,
                           const float* bias, double* stats, cudaStream_t stream) {
    using S = TcFusedSmem<STAGES>;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(tc_convT4_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail("tc_convT4_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    TcParams q = p;
    q.classes = 4;
    q.n_tiles = (int)cdiv(p.OC, 64);
    q.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * q.n_tiles;
    const int grid = q.total_tiles < num_sms() ? q.total_tiles : num_sms();      // one CTA per SM (all 512 TMEM columns)
    tc_convT4_kernel<STAGES><<<grid, TC_THREADS, S::TOTAL, stream>>>(ma, mb, om, q, tab, bias, stats);
    return launched("tc_convT4_kernel");
}

static bool staged_epilogue_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_DISABLE_STAGED_EPILOGUE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// VARSEP_DISABLE_SHIFT=1 switches the shifted-window kernel off.  Eligibility is a property of the layer, never of the
// batch size (per-sample results must not depend on what else is in the batch); VARSEP_SHIFT_MIN_ITEMS (pair
// work items below which the per-class kernels are used instead) exists for experiments only.  Measured on the mnist step:
// 3.68 ms with a threshold of two items per CTA pair, 3.57 ms with one or none (the encoders' 256-image launches included).
static bool shift_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_DISABLE_SHIFT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// k4 s2 p1 layers with 64 / 128 output channels on 16-row x 8-column tiles (see tc_conv_shift_kernel).
// returns 0 = launched, -1 = not this kernel's geometry, >0 = error
static int conv_forward_shift(EncodeTiledFn enc, const vs_conv_geom* g, bool tr, const TcParams& p0, int classes, const void* in,
                              const void* wp, const float* bias, void* out, double* stats, cudaStream_t stream, int IH, int IW, int IC,
                              int OH, int OW, int OC) {
    constexpr int STAGES = 3;
    using S = TcShiftSmem<STAGES>;
    if (shift_disabled() || pair_disabled() || staged_epilogue_disabled()) return -1;
    if (g->R != 4 || g->S != 4 || g->stride != 2 || g->pad != 1) return -1;
    if ((OC != 64 && OC != 128) || IC % 64 != 0 || p0.partial) return -1;
    if (p0.OHc % TCS_HT != 0 || p0.OWc % TCS_WT != 0) return -1;
    if (stats != nullptr && p0.act != VS_ACT_NONE) return -1;          // the sums are taken from the staged (activated) tile
    const int kchunks = IC / 64;
    if (16 * kchunks * (OC / 64) > TCS_PIECES) return -1;              // resident weights: 128 KB per CTA
    TcParams p = p0;
    p.WT = TCS_WT; p.HT = TCS_HT; p.NT = 1;
    p.tiles_w = p.OWc / TCS_WT; p.tiles_h = p.OHc / TCS_HT; p.tiles_n = g->N;
    p.classes = classes; p.n_tiles = 1;
    const long long ptiles = (long long)p.tiles_w * p.tiles_h * g->N;
    p.total_tiles = (int)((ptiles + 1) / 2);
    int pairs = num_sms() / 2;
    static long long min_items = -1;
    if (min_items < 0) { const char* e = getenv("VARSEP_SHIFT_MIN_ITEMS"); min_items = e ? atoll(e) : 0; }
    if (p.total_tiles < min_items) return -1;
    if (pairs > p.total_tiles) pairs = p.total_tiles;
    p.n_per_group = g->N / g->groups;

    TcShiftTab tab;
    memset(&tab, 0, sizeof(tab));
    int ng = 0, nb = 0, pieces = 0;
    if (tr) {
        TcFusedTab ft;
        if (OC != 64 || 3 * kchunks > TCS_MAX_BOX || !build_fused_tab(p, classes, ft)) return -1;
        tab.hb = TCS_HT + 2;
        static const int dw_order[3] = {0, 1, -1};          // the centre shift (all four classes, N = 256) comes first
        for (int wi = 0; wi < 3; ++wi)
            for (int kc = 0; kc < kchunks; ++kc, ++nb) {
                tab.box_dx[nb] = (short)dw_order[wi]; tab.box_dy[nb] = -1; tab.box_c0[nb] = (short)(kc * 64);
                tab.box_g0[nb] = (unsigned char)ng;
                for (int o = 0; o < ft.nshift; ++o) {
                    if (ft.dw[o] != dw_order[wi]) continue;
                    for (int gi = 0; gi < ft.ngrp[o]; ++gi, ++ng) {
                        if (ng >= TCS_MAX_GRP) return -1;
                        const int n = ft.grp_n[o][gi];
                        tab.grp[ng] = make_uint4((unsigned)((ft.dh[o] + 1) * 1024 / 16), (unsigned)(pieces * 4096 / 16),
                                                 idesc_bf16_f32_pair(64 * n), (unsigned)(ft.grp_pos[o][gi] * 64));
                        // the n class slabs side by side are the N = 64n weight rows of the instruction; CTA r stages rows
                        // [32n r, 32n (r+1)) of them as 32-row pieces
                        for (int r = 0; r < 2; ++r)
                            for (int i = 0; i < n; ++i) {
                                const int pi = r * n + i, slab = ft.grp_slab[o][gi] + pi / 2;
                                tab.piece_x[r][pieces + i] = ft.slab_wtap[o][slab] * IC + kc * 64;
                                tab.piece_y[r][pieces + i] = (pi % 2) * 32;
                            }
                        pieces += n;
                    }
                }
            }
        tab.nblk = 4;
        for (int b = 0; b < 4; ++b) { tab.blk_cls[b] = ft.pos_cls[b]; tab.blk_ch0[b] = 0; }
    } else {
        if (classes != 1 || kchunks != 1) return -1;
        tab.hb = TCS_HT + 1;
        for (int dyb = -1; dyb <= 0; ++dyb)                 // odd input rows (filter rows 0, 2), even input rows (1, 3)
            for (int sx = 0; sx < 4; ++sx, ++nb) {
                tab.box_dx[nb] = (short)(sx - 1); tab.box_dy[nb] = (short)dyb; tab.box_c0[nb] = 0;
                tab.box_g0[nb] = (unsigned char)ng;
                for (int r = dyb + 1; r < 4; r += 2, ++ng) {
                    tab.grp[ng] = make_uint4((unsigned)(((r - 1 - dyb) / 2) * 1024 / 16), (unsigned)(pieces * 4096 / 16),
                                             idesc_bf16_f32_pair(OC), 0u);
                    for (int rk = 0; rk < 2; ++rk)
                        for (int i = 0; i < OC / 64; ++i) {
                            tab.piece_x[rk][pieces + i] = (r * 4 + sx) * IC;
                            tab.piece_y[rk][pieces + i] = rk * (OC / 2) + 32 * i;
                        }
                    pieces += OC / 64;
                }
            }
        tab.nblk = OC / 64;
        for (int b = 0; b < tab.nblk; ++b) { tab.blk_cls[b] = 0; tab.blk_ch0[b] = (unsigned char)(64 * b); }
    }
    tab.nbox = nb; tab.ngrp = ng; tab.npieces = pieces; tab.box_g0[nb] = (unsigned char)ng;
    if (pieces > TCS_PIECES || nb > TCS_MAX_BOX) return -1;

    CUtensorMap ma, mb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)IC, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)IC * 2, (cuuint64_t)IC * IW * 2, (cuuint64_t)IC * IW * IH * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(TCS_WT * p.in_sw), (cuuint32_t)(tab.hb * p.in_sh), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)p.in_sw, (cuuint32_t)p.in_sh, 1};
        CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(shift A) failed: %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)g->R * g->S * IC, (cuuint64_t)OC};
        cuuint64_t strides[1] = {(cuuint64_t)g->R * g->S * IC * 2};
        cuuint32_t box[2] = {64, 32};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(shift B) failed: %d", (int)r);
    }
    TcOutMaps om;
    memset(&om, 0, sizeof(om));
    const int ost = p.ost;
    for (int cls = 0; cls < classes; ++cls) {
        const char* base = reinterpret_cast<const char*>(out) + ((size_t)p.ca[cls] * OW + p.cb[cls]) * OC * 2;
        cuuint64_t dims[4] = {(cuuint64_t)OC, (cuuint64_t)p.OWc, (cuuint64_t)p.OHc, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)ost * OC * 2, (cuuint64_t)ost * OW * OC * 2, (cuuint64_t)OH * OW * OC * 2};
        cuuint32_t box[4] = {64, TCS_WT, TCS_HT, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&om.m[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(shift out) failed: %d", (int)r);
    }
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_shift_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail("tc_conv_shift_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_conv_shift_kernel<STAGES>, ma, mb, om, p, tab, bias, stats);
    if (e != cudaSuccess) return fail("tc_conv_shift_kernel launch: %s", cudaGetErrorString(e));
    return launched("tc_conv_shift_kernel");
}

int conv_forward_tc_eligible(const vs_conv_geom* g, int mode) {
    if (g->dtype != VS_BF16 || (g->flags & VS_FLAG_FORCE_SIMT) || tc_disabled()) return 0;
    const bool tr = mode == VS_CONV_TRANSPOSED;
    const int IC = tr ? g->K : g->C, OH = tr ? g->H : g->P, OW = tr ? g->W : g->Q;
    const int st = g->stride, ost = tr ? st : 1;
    if (IC % 8 != 0 || IC < 32 || st > 2 || g->R * g->S > TC_MAX_TAPS) return 0;      // TMA needs a 16-byte channel pitch
    if (OH % ost != 0 || OW % ost != 0) return 0;
    if (tr && st == 2 && ((g->R % 2) || (g->S % 2))) return 0;      // parity classes with unequal tap counts
    return 1;
}

// Which of the tap-GEMM kernels conv_forward_tc launches for a geometry it accepts, from the geometry alone (default
// switches, 16-byte aligned tensors): 1 = per-class single-CTA kernel, 2 = CTA-pair kernel, 3 = shifted-window kernel.
// Host-only mirror of the selection below for vs_conv_forward_variant; the batch size g->N does not enter.
int conv_forward_tc_variant(const vs_conv_geom* g, int mode) {
    if (!conv_forward_tc_eligible(g, mode)) return 0;
    const bool tr = mode == VS_CONV_TRANSPOSED;
    const int IC = tr ? g->K : g->C, OC = tr ? g->C : g->K;
    const int OH = tr ? g->H : g->P, OW = tr ? g->W : g->Q;
    const int st = g->stride;
    const bool one_px = tr && g->P == 1 && g->Q == 1 && g->pad == 0 && st == 1 && g->R == g->S && g->R * g->S <= TC_MAX_CLASSES;
    const int cst = one_px ? g->R : st, ost = tr ? cst : 1;
    const int kchunks = (IC + 63) / 64;
    const bool aligned_rows = OC % 8 == 0;
    if (!one_px && g->R == 4 && g->S == 4 && st == 2 && g->pad == 1 && (OC == 64 || OC == 128) && IC % 64 == 0 &&
        (OH / ost) % TCS_HT == 0 && (OW / ost) % TCS_WT == 0 && 16 * kchunks * (OC / 64) <= TCS_PIECES &&
        (tr ? (OC == 64 && 3 * kchunks <= TCS_MAX_BOX) : kchunks == 1))
        return 3;
    const int ntaps = !tr ? g->R * g->S : one_px ? 1 : (g->R / cst) * (g->S / cst);
    const int PBN = OC % 256 == 0 ? 256 : 128;
    if (OC % 128 == 0 && IC >= 64 && aligned_rows && (PBN == 256 || ntaps * kchunks >= 64)) return 2;
    return 1;
}

// returns 0 = done, -1 = geometry not eligible (caller falls back), >0 = error
int conv_forward_tc(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                    double* stats, cudaStream_t stream) {
    if (g->dtype != VS_BF16 || (g->flags & VS_FLAG_FORCE_SIMT) || tc_disabled()) return -1;
    const bool tr = mode == VS_CONV_TRANSPOSED;
    const int IH = tr ? g->P : g->H, IW = tr ? g->Q : g->W, IC = tr ? g->K : g->C;
    const int OH = tr ? g->H : g->P, OW = tr ? g->W : g->Q, OC = tr ? g->C : g->K;
    const int st = g->stride;
    if (IC % 8 != 0 || IC < 32 || st > 2 || g->R * g->S > TC_MAX_TAPS) return -1;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(wp)) & 15) return -1;
    // A transposed convolution of a 1x1 input (the decoder's first up-convolution) has exactly one tap per output pixel:
    // every output position is its own "parity class" (class stride R), so no structurally-zero tap is multiplied.
    const bool one_px = tr && g->P == 1 && g->Q == 1 && g->pad == 0 && st == 1 && g->R == g->S && g->R * g->S <= TC_MAX_CLASSES;
    const int cst = one_px ? g->R : st;          // stride between the pixels of one class
    const int ost = tr ? cst : 1;
    if (OH % ost != 0 || OW % ost != 0) return -1;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1;

    TcParams p;
    memset(&p, 0, sizeof(p));
    p.N = g->N; p.OH = OH; p.OW = OW; p.OC = OC; p.ost = ost;
    p.OHc = OH / ost; p.OWc = OW / ost;
    p.WT = pow2ceil(p.OWc) < 128 ? pow2ceil(p.OWc) : 128;
    p.HT = pow2ceil(p.OHc) < 128 / p.WT ? pow2ceil(p.OHc) : 128 / p.WT;
    p.NT = 128 / (p.WT * p.HT);
    p.tiles_w = (int)cdiv(p.OWc, p.WT); p.tiles_h = (int)cdiv(p.OHc, p.HT); p.tiles_n = (int)cdiv(g->N, p.NT);
    // a trailing partial 64-channel chunk reads zeros for the input (TMA out-of-bounds fill), so whatever the weight
    // box picks up beyond this tap's IC columns (the next tap's weights, or zeros past the matrix) is multiplied by 0
    p.IC = IC; p.kchunks = (int)cdiv(IC, 64);
    p.in_sh = p.in_sw = tr ? 1 : st;
    p.act = g->act; p.has_bias = bias != nullptr;
    int classes = ost * ost, ntaps = -1;
    if (classes > TC_MAX_CLASSES) return -1;
    for (int cls = 0; cls < classes; ++cls) {
        const int a = cls / ost, b = cls % ost;
        p.ca[cls] = (unsigned char)a; p.cb[cls] = (unsigned char)b;
        // first pass counts the taps of this class, second pass fills the flat tables at [cls * ntaps + tap]
        for (int pass = 0; pass < 2; ++pass) {
            int n = 0;
            if (!tr) {
                for (int r = 0; r < g->R; ++r)
                    for (int s = 0; s < g->S; ++s, ++n)
                        if (pass) { p.dh[cls * ntaps + n] = (signed char)(r - g->pad); p.dw[cls * ntaps + n] = (signed char)(s - g->pad); p.wtap[cls * ntaps + n] = (unsigned char)(r * g->S + s); }
            } else {
                for (int r = (a + g->pad) % cst; r < g->R; r += cst)
                    for (int s = (b + g->pad) % cst; s < g->S; s += cst, ++n)
                        if (pass) {
                            p.dh[cls * ntaps + n] = (signed char)((a + g->pad - r) / cst); p.dw[cls * ntaps + n] = (signed char)((b + g->pad - s) / cst);
                            p.wtap[cls * ntaps + n] = (unsigned char)(r * g->S + s);
                        }
            }
            if (!pass) {
                if (ntaps >= 0 && n != ntaps) return -1;     // classes with unequal tap counts (odd filters, stride 2)
                ntaps = n;
                if (ntaps <= 0 || classes * ntaps > TC_TABLE) return -1;
            }
        }
    }
    p.ntaps = ntaps;
    if (p.WT * p.in_sw > 256 || p.HT * p.in_sh > 256) return -1;

    // A: NHWC input [N, IH, IW, IC] viewed as (IC, IW, IH, N); traversal stride = convolution stride
    CUtensorMap ma, mb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)IC, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)IC * 2, (cuuint64_t)IC * IW * 2, (cuuint64_t)IC * IW * IH * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(p.WT * p.in_sw), (cuuint32_t)(p.HT * p.in_sh), (cuuint32_t)p.NT};
        cuuint32_t estr[4] = {1, (cuuint32_t)p.in_sw, (cuuint32_t)p.in_sh, 1};
        CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    }
    const int BN = (OC % 128 == 0 || OC >= 512) ? 128 : 64;
    p.partial = (OC % 8 != 0) || (reinterpret_cast<uintptr_t>(out) & 15) ? 1 : 0;      // rows not 16-byte aligned
    if (!one_px) {
        const int rc = conv_forward_shift(enc, g, tr, p, classes, in, wp, bias, out, stats, stream, IH, IW, IC, OH, OW, OC);
        if (rc >= 0) return rc;
    }
    // wide layers: CTA pairs (cta_group::2), BN = 256 when OC allows.  A property of the layer, never of the batch size.
    // (measured: 256-column pair tiles reach 1.2-1.4 PFLOP/s on the decoder layers; 128-column pair tiles lose to two
    // single CTAs per SM except for long reduction loops, where the deeper ring of the pair kernel hides the latency)
    const int PBN = OC % 256 == 0 ? 256 : 128;
    static int pair128_min = -1;       // (tap, 64-channel chunk) steps from which 128-column layers use CTA pairs
    if (pair128_min < 0) { const char* e = getenv("VARSEP_PAIR128_MIN_STEPS"); pair128_min = e ? atoi(e) : 64; }
    const bool pair = OC % 128 == 0 && IC >= 64 && !p.partial && !pair_disabled() && (PBN == 256 || p.ntaps * p.kchunks >= pair128_min);
    // narrow layers with little weight data: CTA pairs with ALL weight slabs resident in shared memory (see TcPairSmem)
    static long long res_min_items = -1;      // work items (pixel tiles x parity classes) from which it pays: 8 per SM
    // OPT-IN (VARSEP_RESIDENT_OC: bit 0 = 64-channel layers, bit 1 = 128-channel layers).  Measured on the mnist step
    // (bf16, batch 128): 3.76 ms without, 3.80 ms with the 128-channel layers, 3.99 ms with the 64-channel layers resident —
    // one pair tile per SM with 64-column MMAs loses more tensor-pipe rate than the removed weight traffic gives back.
    static int res_oc = -1;
    if (res_oc < 0) { const char* e = getenv("VARSEP_RESIDENT_OC"); res_oc = e ? atoi(e) : 0; }
    if (res_min_items < 0) { const char* e = getenv("VARSEP_RESIDENT_MIN_ITEMS"); res_min_items = e ? atoll(e) : 8LL * num_sms(); }
    const long long work_items = (long long)p.tiles_w * p.tiles_h * p.tiles_n * classes;
    const bool resb = !pair_disabled() && ((OC == 64 && (res_oc & 1)) || (OC == 128 && (res_oc & 2))) && IC % 64 == 0 && !p.partial && !one_px &&
                      classes <= TC_MAX_OUT_MAPS && !staged_epilogue_disabled() && (stats == nullptr || p.act == VS_ACT_NONE) &&
                      (long long)g->R * g->S * p.kchunks * (OC / 2) * 128 <= TC_RES_BYTES && work_items >= res_min_items;
    p.res_taps = g->R * g->S;
    {
        cuuint64_t dims[2] = {(cuuint64_t)g->R * g->S * IC, (cuuint64_t)OC};
        cuuint64_t strides[1] = {(cuuint64_t)g->R * g->S * IC * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(resb ? OC / 2 : pair ? PBN / 2 : BN)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    // statistics are fused into the epilogue when every 128-pixel tile lies inside one BatchNorm group
    p.n_per_group = g->N / g->groups;
    const bool fuse_stats = stats != nullptr && (p.n_per_group % p.NT) == 0;
    // staged epilogue + TMA stores for the 64-column tiles (16-byte aligned rows): one [OC, OW/ost, OH/ost, N] view per
    // output-parity class, starting at that class's first pixel
    TcOutMaps om;
    memset(&om, 0, sizeof(om));
    // (BN = 128: only the statistics-carrying forward launches, see below; one CTA per SM with a five-stage ring)
    const bool staged128 = BN == 128 && OC % 128 == 0 && fuse_stats && p.act == VS_ACT_NONE && !staged_stats_disabled();
    const bool staged = (resb || BN == 64 || staged128) && !p.partial && classes <= TC_MAX_OUT_MAPS && !staged_epilogue_disabled();
    if (staged) {
        for (int cls = 0; cls < classes; ++cls) {
            const char* base = reinterpret_cast<const char*>(out) + ((size_t)p.ca[cls] * OW + p.cb[cls]) * OC * 2;
            cuuint64_t dims[4] = {(cuuint64_t)OC, (cuuint64_t)p.OWc, (cuuint64_t)p.OHc, (cuuint64_t)g->N};
            cuuint64_t strides[3] = {(cuuint64_t)ost * OC * 2, (cuuint64_t)ost * OW * OC * 2, (cuuint64_t)OH * OW * OC * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)p.WT, (cuuint32_t)p.HT, (cuuint32_t)p.NT};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&om.m[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(out) failed: %d", (int)r);
        }
    }
    double* stp = fuse_stats ? stats : nullptr;
    if (resb) {
        int rc = OC == 64 ? launch_tc_pair<64, 5, true>(ma, mb, om, p, bias, out, classes, stp, stream)
                          : launch_tc_pair<128, 4, true>(ma, mb, om, p, bias, out, classes, stp, stream);
        if (rc) return rc;
        if (stats != nullptr && !fuse_stats) return stats_of_output(g, g->dtype, out, (long long)g->N * OH * OW, OC, stats, stream);
        return rc;
    }
    TcFusedTab ftab;
    if (tr && staged && BN == 64 && !pair && (stp == nullptr || p.act == VS_ACT_NONE) && !fused_classes_disabled() && build_fused_tab(p, classes, ftab)) {
        int rc = launch_tc_fused<3>(ma, mb, om, p, ftab, bias, stp, stream);
        if (rc) return rc;
        if (stats != nullptr && !fuse_stats) return stats_of_output(g, g->dtype, out, (long long)g->N * OH * OW, OC, stats, stream);
        return rc;
    }
    int rc = pair ? (PBN == 256 ? launch_tc_pair<256, 6>(ma, mb, om, p, bias, out, classes, stp, stream)
                                : launch_tc_pair<128, 8>(ma, mb, om, p, bias, out, classes, stp, stream))
             : BN == 128 ? (staged ? launch_tc<128, 5, true, true>(ma, mb, om, p, bias, out, classes, stp, stream)
                                   : launch_tc<128, 3, false>(ma, mb, om, p, bias, out, classes, stp, stream))
             : staged && stp != nullptr && p.act == VS_ACT_NONE && !staged_stats_disabled()
                       ? launch_tc<64, 3, true, true>(ma, mb, om, p, bias, out, classes, stp, stream)
             : staged  ? launch_tc<64, 3, true>(ma, mb, om, p, bias, out, classes, stp, stream)
                       : launch_tc<64, 4, false>(ma, mb, om, p, bias, out, classes, stp, stream);
    if (rc) return rc;
    if (stats != nullptr && !fuse_stats) return stats_of_output(g, g->dtype, out, (long long)g->N * OH * OW, OC, stats, stream);
    return rc;
}

}  // namespace vs

// =================================================================================================
// Weight gradient on the tensor cores.
//   dw[k][c][tap] += sum_pix small[pix][k] * big[src(pix, tap)][c]
// is a GEMM whose REDUCTION dimension is the pixel index, i.e. both operands are "MN-major" in
// shared memory (channels contiguous, pixels along K): exactly what a TMA box of an NHWC tensor is.
// A = 128 channels of `small` (two 64-channel boxes), B = BN channels of `big` gathered at the tap
// offset (strided box for stride-2 convolutions, zero-filled outside the image), 64 pixels per
// pipeline stage, fp32 accumulation in TMEM, split over pixel ranges across CTAs, and an epilogue of
// fp32 reductions (red.global.add) straight into the torch-layout gradient.
// =================================================================================================
namespace vs {

struct TcWgradParams {
    int N, P, Q, K, C, R, S, stride, pad;
    int WT, HT, NT, tiles_w, tiles_h, tiles_n;   // 64-pixel boxes over the small grid
    int c_tiles, k_tiles;
    int total_ptiles, ptiles_per_split;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 2) tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_small,
                                                          const __grid_constant__ CUtensorMap map_big,
                                                          const __grid_constant__ TcWgradParams p,
                                                          float* __restrict__ dw) {
    constexpr int PIX = 64;
    constexpr int A_BYTES = 128 * PIX * 2, B_BYTES = BN * PIX * 2, STAGE_BYTES = A_BYTES + B_BYTES, CHUNK = 64 * PIX * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x % p.k_tiles, ct = blockIdx.x / p.k_tiles;
    const int tap = blockIdx.y;
    const int r = tap / p.S, s = tap % p.S;
    const int pt0 = blockIdx.z * p.ptiles_per_split;
    int pt1 = pt0 + p.ptiles_per_split;
    if (pt1 > p.total_ptiles) pt1 = p.total_ptiles;
    const int nkb = pt1 - pt0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (nkb <= 0) {          // uniform: nothing to reduce for this split
        __syncthreads();
        if (warp == 1) tmem_dealloc<BN>(tmem_base);
        return;
    }

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[st], ph ^ 1);
                int t = pt0 + kb;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h;
                const int tn = t / p.tiles_h;
                const int q0 = tw * p.WT, p0 = th * p.HT, b0 = tn * p.NT;
                uint8_t* a_dst = smem + st * STAGE_BYTES;
                uint8_t* b_dst = a_dst + A_BYTES;
                mbar_expect_tx(&full[st], STAGE_BYTES);
                tma_load_4d(a_dst, &map_small, &full[st], kt * 128, q0, p0, b0);
                tma_load_4d(a_dst + CHUNK, &map_small, &full[st], kt * 128 + 64, q0, p0, b0);
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                    tma_load_4d(b_dst + j * CHUNK, &map_big, &full[st], ct * BN + j * 64, q0 * p.stride - p.pad + s,
                                p0 * p.stride - p.pad + r, b0);
            }
        }
    } else if (warp == 1) {
        {   // converged warp, one elected lane issues (see tc_conv_kernel)
            // D=f32, A=B=bf16, both MN-major (bits 15, 16), M=128, N=BN
            constexpr uint32_t idesc = idesc_bf16_f32(BN) | (1u << 15) | (1u << 16);
            const bool elected = elect_one();
            int st = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t a_lo = mnmajor_sw128_desc_lo(smem_u32(smem + st * STAGE_BYTES), CHUNK);
                const uint32_t b_lo = a_lo + A_BYTES / 16;
#pragma unroll
                for (int k = 0; k < PIX / 16; ++k) {
                    // 16 pixels = two 8-pixel swizzle groups = 2048 bytes along K
                    const uint64_t ad = kmajor_sw128_desc_from_lo(a_lo + k * 128), bd = kmajor_sw128_desc_from_lo(b_lo + k * 128);
                    const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                    if (elected) umma_bf16(tmem_base, ad, bd, idesc, accum);
                }
                if (elected) umma_commit(&empty[st]);
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
            if (elected) umma_commit(tmem_full);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int k = kt * 128 + q * 32 + lane;          // output row = channel of `small`
        const int RS = p.R * p.S;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            float* dst = dw + ((long long)k * p.C + ct * BN + c0) * RS + tap;
            if (k < p.K) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (ct * BN + c0 + c < p.C) atomicAdd(dst + (long long)c * RS, __uint_as_float(v[c]));
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

// Several taps per CTA: the box of `small` is staged ONCE per 64-pixel block and multiplied with the TAPS shifted boxes
// of CB channels of `big`, which sit back to back in shared memory and so form ONE MN-major B operand of N = TAPS * CB
// = 256 columns (accumulator column = tap * CB + channel).  Per block a CTA stages 16 + 32 KB for 128 x 256 x 64 MACs:
// 87 FLOP per staged byte instead of 44 (CB = 64) / 64 (CB = 128) -- the one-tap kernel sits on the ~64 B/clk/SM
// L2 -> shared-memory fill (measured 62 B/clk/SM, tensor pipe 72 % busy with N = 64 instructions).  The taps of a CTA
// are consecutive in s, so for one (k, c) their gradients are contiguous in the torch layout: the epilogue issues one
// vector reduction (red.global.add.v4/v2.f32) per (k, c) instead of TAPS scalar ones.
template <int CB, int TAPS, int STAGES>
__global__ void __launch_bounds__(192, 2) tc_wgrad_taps_kernel(const __grid_constant__ CUtensorMap map_small,
                                                               const __grid_constant__ CUtensorMap map_big,
                                                               const __grid_constant__ TcWgradParams p,
                                                               float* __restrict__ dw, int vec_ok) {
    constexpr int PIX = 64, BN = CB * TAPS;
    constexpr int A_BYTES = 128 * PIX * 2, B_BYTES = BN * PIX * 2, STAGE_BYTES = A_BYTES + B_BYTES, CHUNK = 64 * PIX * 2;
    static_assert(BN == 256 && (TAPS == 2 || TAPS == 4), "one 256-column accumulator");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kt = blockIdx.x % p.k_tiles, ct = blockIdx.x / p.k_tiles;
    const int tap0 = blockIdx.y * TAPS;                  // TAPS consecutive taps (same r: S % TAPS == 0, host check)
    const int r = tap0 / p.S, s0 = tap0 % p.S;
    const int pt0 = blockIdx.z * p.ptiles_per_split;
    int pt1 = pt0 + p.ptiles_per_split;
    if (pt1 > p.total_ptiles) pt1 = p.total_ptiles;
    const int nkb = pt1 - pt0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_small);
        prefetch_tmap(&map_big);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (nkb <= 0) {          // uniform: nothing to reduce for this split
        __syncthreads();
        if (warp == 1) tmem_dealloc<BN>(tmem_base);
        return;
    }

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[st], ph ^ 1);
                int t = pt0 + kb;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h;
                const int tn = t / p.tiles_h;
                const int q0 = tw * p.WT, p0 = th * p.HT, b0 = tn * p.NT;
                uint8_t* a_dst = smem + st * STAGE_BYTES;
                uint8_t* b_dst = a_dst + A_BYTES;
                mbar_expect_tx(&full[st], STAGE_BYTES);
                tma_load_4d(a_dst, &map_small, &full[st], kt * 128, q0, p0, b0);
                tma_load_4d(a_dst + CHUNK, &map_small, &full[st], kt * 128 + 64, q0, p0, b0);
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) {      // chunk j = tap j / (CB/64), 64-channel slice j % (CB/64)
                    const int tj = j / (CB / 64), cj = j % (CB / 64);
                    tma_load_4d(b_dst + j * CHUNK, &map_big, &full[st], ct * CB + cj * 64, q0 * p.stride - p.pad + s0 + tj,
                                p0 * p.stride - p.pad + r, b0);
                }
            }
        }
    } else if (warp == 1) {
        {   // converged warp, one elected lane issues (see tc_conv_kernel)
            constexpr uint32_t idesc = idesc_bf16_f32(BN) | (1u << 15) | (1u << 16);      // both operands MN-major
            const bool elected = elect_one();
            int st = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t a_lo = mnmajor_sw128_desc_lo(smem_u32(smem + st * STAGE_BYTES), CHUNK);
                const uint32_t b_lo = a_lo + A_BYTES / 16;
#pragma unroll
                for (int k = 0; k < PIX / 16; ++k) {
                    const uint64_t ad = kmajor_sw128_desc_from_lo(a_lo + k * 128), bd = kmajor_sw128_desc_from_lo(b_lo + k * 128);
                    const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                    if (elected) umma_bf16(tmem_base, ad, bd, idesc, accum);
                }
                if (elected) umma_commit(&empty[st]);
                if (++st == STAGES) { st = 0; ph ^= 1; }
            }
            if (elected) umma_commit(tmem_full);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int k = kt * 128 + q * 32 + lane;          // output row = channel of `small`
        const int RS = p.R * p.S;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < CB; c0 += 8) {
            uint32_t v[TAPS][8];
#pragma unroll
            for (int j = 0; j < TAPS; ++j) tmem_ld8_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * CB + c0), v[j]);
            tmem_ld_wait();
            if (k < p.K) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = ct * CB + c0 + e;
                    if (c >= p.C) break;
                    float* dst = dw + ((long long)k * p.C + c) * RS + tap0;
                    if (vec_ok) {
                        if (TAPS == 4)
                            red_add_v4(dst, __uint_as_float(v[0][e]), __uint_as_float(v[1][e]), __uint_as_float(v[2 % TAPS][e]),
                                       __uint_as_float(v[3 % TAPS][e]));
                        else
                            red_add_v2(dst, __uint_as_float(v[0][e]), __uint_as_float(v[1][e]));
                    } else {
#pragma unroll
                        for (int j = 0; j < TAPS; ++j) atomicAdd(dst + j, __uint_as_float(v[j][e]));
                    }
                }
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

// CTA-pair variant (cta_group::2): the pair owns 256 channels of `small` (128 per CTA) x BN channels of `big` (BN/2 per
// CTA) for one tap and one pixel range; per 64-pixel block a CTA stages 16 + BN/16 KB for 128 x BN x 64 MACs
// (128 FLOP per staged byte at BN = 256 instead of 64).
template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 2) tc_wgrad_pair_kernel(const __grid_constant__ CUtensorMap map_small,
                                                               const __grid_constant__ CUtensorMap map_big,
                                                               const __grid_constant__ TcWgradParams p,
                                                               float* __restrict__ dw) {
    constexpr int PIX = 64, HB = BN / 2;
    constexpr int A_BYTES = 128 * PIX * 2, B_BYTES = HB * PIX * 2, STAGE_BYTES = A_BYTES + B_BYTES, CHUNK = 64 * PIX * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                      // leader's
    uint64_t* empty = bars + STAGES;            // each CTA's
    uint64_t* tmem_full = bars + 2 * STAGES;    // each CTA's
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cl_rank();
    const bool leader = rank == 0;
    const int pairx = blockIdx.x >> 1;
    const int kt = pairx % p.k_tiles, ct = pairx / p.k_tiles;          // k_tiles counts 256-channel tiles here
    const int tap = blockIdx.y;
    const int r = tap / p.S, s = tap % p.S;
    const int pt0 = blockIdx.z * p.ptiles_per_split;
    int pt1 = pt0 + p.ptiles_per_split;
    if (pt1 > p.total_ptiles) pt1 = p.total_ptiles;
    const int nkb = pt1 - pt0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_small);
        prefetch_tmap(&map_big);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) tmem_alloc_pair<BN>(tmem_slot);
    tc_fence_before();
    cl_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (nkb > 0) {                 // uniform over the pair
        if (warp == 0) {
            if (lane == 0) {
                const uint32_t k_base = (uint32_t)(kt * 256 + rank * 128), c_base = (uint32_t)(ct * BN + rank * HB);
                for (int kb = 0; kb < nkb; ++kb) {
                    const int st = kb % STAGES;
                    mbar_wait(&empty[st], ((kb / STAGES) & 1) ^ 1);
                    int t = pt0 + kb;
                    const int tw = t % p.tiles_w; t /= p.tiles_w;
                    const int th = t % p.tiles_h;
                    const int tn = t / p.tiles_h;
                    const int q0 = tw * p.WT, p0 = th * p.HT, b0 = tn * p.NT;
                    uint8_t* a_dst = smem + st * STAGE_BYTES;
                    uint8_t* b_dst = a_dst + A_BYTES;
                    if (leader) mbar_expect_tx(&full[st], 2 * STAGE_BYTES);
                    const uint32_t bar = cl_map(&full[st], 0);
                    tma_load_4d_pair(a_dst, &map_small, bar, (int)k_base, q0, p0, b0);
                    tma_load_4d_pair(a_dst + CHUNK, &map_small, bar, (int)k_base + 64, q0, p0, b0);
#pragma unroll
                    for (int j = 0; j < HB / 64; ++j)
                        tma_load_4d_pair(b_dst + j * CHUNK, &map_big, bar, (int)c_base + j * 64, q0 * p.stride - p.pad + s,
                                         p0 * p.stride - p.pad + r, b0);
                }
            }
        } else if (warp == 1) {
            if (leader) {   // converged warp, one elected lane issues (see tc_conv_kernel)
                constexpr uint32_t idesc = idesc_bf16_f32_pair(BN) | (1u << 15) | (1u << 16);      // both operands MN-major
                const bool elected = elect_one();
                int st = 0;
                uint32_t ph = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t a_lo = mnmajor_sw128_desc_lo(smem_u32(smem + st * STAGE_BYTES), CHUNK);
                    const uint32_t b_lo = a_lo + A_BYTES / 16;
#pragma unroll
                    for (int k = 0; k < PIX / 16; ++k) {
                        const uint64_t ad = kmajor_sw128_desc_from_lo(a_lo + k * 128), bd = kmajor_sw128_desc_from_lo(b_lo + k * 128);
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                        if (elected) umma_bf16_pair(tmem_base, ad, bd, idesc, accum);
                    }
                    if (elected) umma_commit_pair(&empty[st]);
                    if (++st == STAGES) { st = 0; ph ^= 1; }
                }
                if (elected) umma_commit_pair(tmem_full);
                __syncwarp();
            }
        } else {
            const int q = warp & 3;
            const int k = kt * 256 + rank * 128 + q * 32 + lane;          // output row = channel of `small`
            const int RS = p.R * p.S;
            mbar_wait(tmem_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                float* dst = dw + ((long long)k * p.C + ct * BN + c0) * RS + tap;
                if (k < p.K) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (ct * BN + c0 + c < p.C) atomicAdd(dst + (long long)c * RS, __uint_as_float(v[c]));
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    cl_sync();               // the leader's MMAs read the peer's shared memory and write its tensor memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair<BN>(tmem_base);
    }
}

template <int BN, int STAGES>
static int launch_tc_wgrad_pair(const CUtensorMap& ms, const CUtensorMap& mb, const TcWgradParams& p, float* dw, int splits,
                                cudaStream_t stream) {
    constexpr int SMEM = STAGES * (128 * 64 * 2 + (BN / 2) * 64 * 2) + 1024 + 256;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_pair_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return fail("tc_wgrad_pair_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * p.k_tiles * p.c_tiles), (unsigned)(p.R * p.S), (unsigned)splits);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_wgrad_pair_kernel<BN, STAGES>, ms, mb, p, dw);
    if (e != cudaSuccess) return fail("tc_wgrad_pair_kernel launch: %s", cudaGetErrorString(e));
    return launched("tc_wgrad_pair_kernel");
}

template <int BN, int STAGES>
static int launch_tc_wgrad(const CUtensorMap& ms, const CUtensorMap& mb, const TcWgradParams& p, float* dw, int splits,
                           cudaStream_t stream) {
    constexpr int SMEM = STAGES * (128 * 64 * 2 + BN * 64 * 2) + 1024 + 256;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return fail("tc_wgrad_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    dim3 grid((unsigned)(p.k_tiles * p.c_tiles), (unsigned)(p.R * p.S), (unsigned)splits);
    tc_wgrad_kernel<BN, STAGES><<<grid, 192, SMEM, stream>>>(ms, mb, p, dw);
    return launched("tc_wgrad_kernel");
}

template <int CB, int TAPS, int STAGES>
static int launch_tc_wgrad_taps(const CUtensorMap& ms, const CUtensorMap& mb, const TcWgradParams& p, float* dw, int splits,
                                cudaStream_t stream) {
    constexpr int SMEM = STAGES * (128 * 64 * 2 + CB * TAPS * 64 * 2) + 1024 + 256;
    static DeviceOnce configured;
    if (!configured.flag()) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_taps_kernel<CB, TAPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return fail("tc_wgrad_taps_kernel smem attribute: %s", cudaGetErrorString(e));
        configured.flag() = true;
    }
    // vector reductions need the TAPS gradients of one (k, c) aligned to their total size
    const int vec_ok = ((p.R * p.S) % TAPS == 0 && (reinterpret_cast<uintptr_t>(dw) % (TAPS * 4)) == 0) ? 1 : 0;
    dim3 grid((unsigned)(p.k_tiles * p.c_tiles), (unsigned)(p.R * p.S / TAPS), (unsigned)splits);
    tc_wgrad_taps_kernel<CB, TAPS, STAGES><<<grid, 192, SMEM, stream>>>(ms, mb, p, dw, vec_ok);
    return launched("tc_wgrad_taps_kernel");
}

static bool wgrad_taps_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_DISABLE_WGRAD_TAPS"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// The CTA-pair weight gradient is correct (tests/test_kernels_gpu.py runs it) but measured slower than two single CTAs
// per SM on this workload (210 vs 189 us on the 512x256 decoder layer: three 32 KB stages per CTA do not cover the TMA
// latency); it stays opt-in until it gets a deeper ring.
static bool wgrad_pair_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VARSEP_ENABLE_WGRAD_PAIR"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

int conv_wgrad_tc(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream) {
    if (g->dtype != VS_BF16 || (g->flags & VS_FLAG_FORCE_SIMT) || tc_disabled()) return -1;
    if (g->K % 8 != 0 || g->C % 8 != 0 || g->K < 64 || g->stride > 2) return -1;   // TMA: 16-byte channel pitch
    if ((reinterpret_cast<uintptr_t>(small_) | reinterpret_cast<uintptr_t>(big)) & 15) return -1;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1;
    TcWgradParams p;
    memset(&p, 0, sizeof(p));
    p.N = g->N; p.P = g->P; p.Q = g->Q; p.K = g->K; p.C = g->C; p.R = g->R; p.S = g->S; p.stride = g->stride; p.pad = g->pad;
    p.WT = pow2ceil(g->Q) < 64 ? pow2ceil(g->Q) : 64;
    p.HT = pow2ceil(g->P) < 64 / p.WT ? pow2ceil(g->P) : 64 / p.WT;
    p.NT = 64 / (p.WT * p.HT);
    // partial boxes would add zero-filled pixels of `small` (harmless) but the boxes must tile the grid exactly
    // in W/H for the pixel <-> pixel correspondence between the two operands
    if (g->Q % p.WT != 0 || g->P % p.HT != 0) return -1;
    p.tiles_w = g->Q / p.WT; p.tiles_h = g->P / p.HT; p.tiles_n = (int)cdiv(g->N, p.NT);
    p.total_ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
    // wide layers: CTA pairs, 256 channels of `small` x 256 (or 128) channels of `big` per pair
    const bool pair = g->K % 256 == 0 && g->C % 128 == 0 && !wgrad_pair_disabled();
    const int BN = pair ? (g->C % 256 == 0 ? 256 : 128) : ((g->C % 128 == 0 || g->C >= 512) ? 128 : 64);
    // several taps per CTA (one 256-column accumulator = TAPS x BN channels) when the filter rows split evenly
    const int TAPS = pair || wgrad_taps_disabled() ? 1 : 256 / BN;
    const bool taps = TAPS > 1 && g->S % TAPS == 0;
    p.k_tiles = (int)cdiv(g->K, pair ? 256 : 128); p.c_tiles = (int)cdiv(g->C, BN);
    const long long base = (long long)p.k_tiles * p.c_tiles * (g->R * g->S / (taps ? TAPS : 1)) * (pair ? 2 : 1);
    // Split of the pixel range over CTAs: the grid runs in waves of (2 CTAs per SM) and every CTA ends with an epilogue
    // of 128*BN fp32 reductions (worth about 24 pixel blocks, fitted to the measured launches), so pick the split count that minimises
    //   waves(base * splits) * (blocks per split + 24)
    // (e.g. 640 CTAs = 2.2 waves cost 3 waves; 576 cost 2).
    const long long slots = 2LL * num_sms();
    long long splits = 1, best = -1;
    // (with TAPS taps per CTA a pixel block carries TAPS times the MMA work and the epilogue TAPS times the data)
    const long long epi = taps ? 12 : 24;
    for (long long sp = 1; sp <= (taps ? 128 : 64) && sp <= p.total_ptiles; ++sp) {
        const long long cost = cdiv(base * sp, slots) * (cdiv(p.total_ptiles, sp) + epi);
        if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
    p.ptiles_per_split = (int)cdiv(p.total_ptiles, splits);
    splits = cdiv(p.total_ptiles, p.ptiles_per_split);
    if (p.WT * g->stride > 256 || p.HT * g->stride > 256) return -1;

    CUtensorMap ms, mb;
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
    {
        cuuint64_t dims[4] = {(cuuint64_t)g->C, (cuuint64_t)g->W, (cuuint64_t)g->H, (cuuint64_t)g->N};
        cuuint64_t strides[3] = {(cuuint64_t)g->C * 2, (cuuint64_t)g->C * g->W * 2, (cuuint64_t)g->C * g->W * g->H * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(p.WT * g->stride), (cuuint32_t)(p.HT * g->stride), (cuuint32_t)p.NT};
        cuuint32_t estr[4] = {1, (cuuint32_t)g->stride, (cuuint32_t)g->stride, 1};
        CUresult rc = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(big), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(big) failed: %d", (int)rc);
    }
    if (taps)
        return BN == 128 ? launch_tc_wgrad_taps<128, 2, 2>(ms, mb, p, dw, (int)splits, stream)
                         : launch_tc_wgrad_taps<64, 4, 2>(ms, mb, p, dw, (int)splits, stream);
    if (pair)
        return BN == 256 ? launch_tc_wgrad_pair<256, 3>(ms, mb, p, dw, (int)splits, stream)
                         : launch_tc_wgrad_pair<128, 4>(ms, mb, p, dw, (int)splits, stream);
    return BN == 128 ? launch_tc_wgrad<128, 3>(ms, mb, p, dw, (int)splits, stream)
                     : launch_tc_wgrad<64, 4>(ms, mb, p, dw, (int)splits, stream);
}

}  // namespace vs

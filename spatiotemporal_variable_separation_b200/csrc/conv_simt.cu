// Generic fp32-accumulate gather-GEMM convolution (CUDA cores).  This is the exact-arithmetic path
// (fp32 storage: parity mode) and the fallback geometry coverage for the bf16 mode; the tensor-core
// path for the hot layers lives in conv_tc.cu and is selected by vs_conv_forward/vs_conv_wgrad.
//
// Replaces aten::convolution / convolution_backward / addmm issued by
// /root/reference/var_sep/networks/conv.py:41-60 (make_conv_block) and mlp.py:40.
#include "common.cuh"

namespace vs {

struct GatherArgs {
    int N, IH, IW, IC;   // tensor that is read through the filter window
    int OH, OW, OC;      // tensor that is produced
    int R, S, stride, pad;
    int transposed;      // 0: ih = oh*stride - pad + r ; 1: ih = (oh + pad - r)/stride when divisible
    int groups, act;
    int n_per_group;     // samples per BatchNorm group
};

constexpr int BM = 64, BN = 64, BK = 16, LDS = BM + 4;

// out[m][oc] = act(bias[oc] + sum_{tap,c} in[src(m,tap)][c] * wp[oc][tap][c])
// grid: (row tiles, oc tiles, stride*stride output-parity classes when transposed)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) gather_gemm_kernel(GatherArgs a, const T* __restrict__ in,
                                                          const T* __restrict__ wp, const float* __restrict__ bias,
                                                          T* __restrict__ out, double* __restrict__ stats) {
    __shared__ __align__(16) float smem[2 * BK * LDS];
    float(*As)[LDS] = reinterpret_cast<float(*)[LDS]>(smem);
    float(*Bs)[LDS] = reinterpret_cast<float(*)[LDS]>(smem + BK * LDS);

    const int tid = threadIdx.x;
    const int ost = a.transposed ? a.stride : 1;          // output sub-sampling of this class
    const int ca = blockIdx.z / ost, cb = blockIdx.z % ost;  // parity class (transposed only)
    const int OHc = (a.OH - ca + ost - 1) / ost, OWc = (a.OW - cb + ost - 1) / ost;
    const long long Mc = (long long)a.N * OHc * OWc;
    const long long m0 = (long long)blockIdx.x * BM;
    if (m0 >= Mc) return;
    const int n0 = blockIdx.y * BN;

    // ---- loader role: one row (pixel) and 4 consecutive k per thread
    const int lr = tid >> 2, cq = (tid & 3) * 4;
    const long long lm = m0 + lr;
    const bool lrow_ok = lm < Mc;
    int ln = 0, loh = 0, low = 0;
    if (lrow_ok) {
        ln = (int)(lm / ((long long)OHc * OWc));
        int rem = (int)(lm - (long long)ln * OHc * OWc);
        loh = (rem / OWc) * ost + ca;
        low = (rem % OWc) * ost + cb;
    }
    const int woc = n0 + lr;   // weight row loaded by this thread
    const bool wrow_ok = woc < a.OC;
    const long long wrow = (long long)woc * a.R * a.S * a.IC;

    // ---- compute role
    const int tx = tid & 15, ty = tid >> 4;
    constexpr bool EXACT = sizeof(T) == 4;
    float acc[4][4];
    double tot[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.0; }

    int r_begin = 0, r_step = 1, s_begin = 0, s_step = 1;
    if (a.transposed) {
        r_begin = (ca + a.pad) % a.stride; r_step = a.stride;
        s_begin = (cb + a.pad) % a.stride; s_step = a.stride;
    }
    for (int r = r_begin; r < a.R; r += r_step) {
        int ih; bool rok;
        if (a.transposed) { int t = loh + a.pad - r; ih = t / a.stride; rok = t >= 0 && ih < a.IH; }
        else { ih = loh * a.stride - a.pad + r; rok = ih >= 0 && ih < a.IH; }
        for (int s = s_begin; s < a.S; s += s_step) {
            int iw; bool sok;
            if (a.transposed) { int t = low + a.pad - s; iw = t / a.stride; sok = t >= 0 && iw < a.IW; }
            else { iw = low * a.stride - a.pad + s; sok = iw >= 0 && iw < a.IW; }
            const bool pix_ok = lrow_ok && rok && sok;
            const T* src = in + (((long long)ln * a.IH + ih) * a.IW + iw) * a.IC;
            const T* wsrc = wp + wrow + (long long)(r * a.S + s) * a.IC;
            for (int c0 = 0; c0 < a.IC; c0 += BK) {
                const int c = c0 + cq;
                float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
                if (VEC) {
                    if (pix_ok && c < a.IC) av = ld4<T>(src + c);
                    if (wrow_ok && c < a.IC) bv = ld4<T>(wsrc + c);
                } else {
                    if (pix_ok) {
                        if (c + 0 < a.IC) av.x = ld<T>(src + c + 0);
                        if (c + 1 < a.IC) av.y = ld<T>(src + c + 1);
                        if (c + 2 < a.IC) av.z = ld<T>(src + c + 2);
                        if (c + 3 < a.IC) av.w = ld<T>(src + c + 3);
                    }
                    if (wrow_ok) {
                        if (c + 0 < a.IC) bv.x = ld<T>(wsrc + c + 0);
                        if (c + 1 < a.IC) bv.y = ld<T>(wsrc + c + 1);
                        if (c + 2 < a.IC) bv.z = ld<T>(wsrc + c + 2);
                        if (c + 3 < a.IC) bv.w = ld<T>(wsrc + c + 3);
                    }
                }
                __syncthreads();   // previous slab fully consumed
                As[cq + 0][lr] = av.x; As[cq + 1][lr] = av.y; As[cq + 2][lr] = av.z; As[cq + 3][lr] = av.w;
                Bs[cq + 0][lr] = bv.x; Bs[cq + 1][lr] = bv.y; Bs[cq + 2][lr] = bv.z; Bs[cq + 3][lr] = bv.w;
                __syncthreads();
#pragma unroll
                for (int k = 0; k < BK; ++k) {
                    const float4 x = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                    const float4 w = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                    const float xs[4] = {x.x, x.y, x.z, x.w}, ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xs[i], ws[j], acc[i][j]);
                }
                if (EXACT) {
                    // parity mode: the 16-term fp32 partial sums of this slab are added up in fp64, so the rounding
                    // error of a K = 4608 reduction stays at the level of ONE short fp32 sum (a sequential fp32
                    // accumulation over K terms loses ~sqrt(K) ulp, more than the blocked sums of the CPU libraries)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) { tot[i][j] += (double)acc[i][j]; acc[i][j] = 0.f; }
                }
            }
        }
    }
    if (EXACT) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = (float)tot[i][j];
    }

    // ---- epilogue: bias, statistics of the pre-activation, activation, store
    float bj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int oc = n0 + tx * 4 + j;
        bj[j] = (bias != nullptr && oc < a.OC) ? bias[oc] : 0.f;
    }
    int rg[4];   // BatchNorm group of each of this thread's rows (-1: row out of range)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        rg[i] = -1;
        if (m >= Mc) continue;
        const int n = (int)(m / ((long long)OHc * OWc));
        const int rem = (int)(m - (long long)n * OHc * OWc);
        const int oh = (rem / OWc) * ost + ca, ow = (rem % OWc) * ost + cb;
        rg[i] = n / a.n_per_group;
        T* dst = out + (((long long)n * a.OH + oh) * a.OW + ow) * a.OC + n0 + tx * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] += bj[j]; v[j] = act_fwd(acc[i][j], a.act); }
        if (VEC && (a.OC & 3) == 0 && n0 + tx * 4 + 3 < a.OC) {
            st4<T>(dst, make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + tx * 4 + j < a.OC) st<T>(dst + j, v[j]);
        }
    }
    if (stats == nullptr) return;

    const long long m_last = (m0 + BM - 1 < Mc ? m0 + BM - 1 : Mc - 1);
    const int g_first = (int)(m0 / ((long long)OHc * OWc)) / a.n_per_group;
    const int g_last = (int)(m_last / ((long long)OHc * OWc)) / a.n_per_group;
    if (g_first == g_last) {
        // whole tile in one group: reduce the 16 row-slices through shared memory, 1 atomic per column
        float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (rg[i] >= 0)
#pragma unroll
                for (int j = 0; j < 4; ++j) { s1[j] += acc[i][j]; s2[j] += acc[i][j] * acc[i][j]; }
        __syncthreads();
        float* red = smem;   // [16][64][2]
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[(ty * 64 + tx * 4 + j) * 2 + 0] = s1[j];
            red[(ty * 64 + tx * 4 + j) * 2 + 1] = s2[j];
        }
        __syncthreads();
        if (tid < 128) {
            const int col = tid >> 1, which = tid & 1;
            double t = 0.0;
#pragma unroll
            for (int y = 0; y < 16; ++y) t += (double)red[(y * 64 + col) * 2 + which];
            if (n0 + col < a.OC) atomicAdd(&stats[((long long)g_first * a.OC + n0 + col) * 2 + which], t);
        }
    } else {
        // tile straddles groups (tiny batches): per-thread atomics
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (rg[i] >= 0)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + tx * 4 + j < a.OC) {
                        double* p = &stats[((long long)rg[i] * a.OC + n0 + tx * 4 + j) * 2];
                        atomicAdd(p, (double)acc[i][j]);
                        atomicAdd(p + 1, (double)acc[i][j] * (double)acc[i][j]);
                    }
    }
}

// dw[k][c][tap] += sum_m small[m][k] * big[src(m, tap)][c]        (direct-form gather, fp32 atomics)
// grid: (k tiles, taps * c tiles, splits over m)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) wgrad_kernel(vs_conv_geom g, const T* __restrict__ small_,
                                                    const T* __restrict__ big, float* __restrict__ dw,
                                                    int c_tiles, long long rows_per_split) {
    __shared__ __align__(16) float smem[2 * BK * LDS];
    float(*Ds)[LDS] = reinterpret_cast<float(*)[LDS]>(smem);             // [m][k]
    float(*Gs)[LDS] = reinterpret_cast<float(*)[LDS]>(smem + BK * LDS);  // [m][c]
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * BN;
    const int tap = blockIdx.y / c_tiles, c0 = (blockIdx.y % c_tiles) * BN;
    const int r = tap / g.S, s = tap % g.S;
    const long long M = (long long)g.N * g.P * g.Q;
    const long long m_begin = (long long)blockIdx.z * rows_per_split;
    const long long m_end = m_begin + rows_per_split < M ? m_begin + rows_per_split : M;

    const int lrow = tid >> 4, lcol = (tid & 15) * 4;   // loader: 16 rows x 64 columns, 4 per thread
    const int tx = tid & 15, ty = tid >> 4;
    constexpr bool EXACT = sizeof(T) == 4;
    float acc[4][4];
    double tot[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.0; }

    for (long long mb = m_begin; mb < m_end; mb += BK) {
        const long long m = mb + lrow;
        float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), gv = dv;
        if (m < m_end) {
            const T* dp = small_ + m * g.K + k0 + lcol;
            const int n = (int)(m / ((long long)g.P * g.Q));
            const int rem = (int)(m - (long long)n * g.P * g.Q);
            const int ih = (rem / g.Q) * g.stride - g.pad + r, iw = (rem % g.Q) * g.stride - g.pad + s;
            const bool ok = ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
            const T* gp = big + (((long long)n * g.H + ih) * g.W + iw) * g.C + c0 + lcol;
            if (VEC) {
                if (k0 + lcol < g.K) dv = ld4<T>(dp);
                if (ok && c0 + lcol < g.C) gv = ld4<T>(gp);
            } else {
                if (k0 + lcol + 0 < g.K) dv.x = ld<T>(dp + 0);
                if (k0 + lcol + 1 < g.K) dv.y = ld<T>(dp + 1);
                if (k0 + lcol + 2 < g.K) dv.z = ld<T>(dp + 2);
                if (k0 + lcol + 3 < g.K) dv.w = ld<T>(dp + 3);
                if (ok) {
                    if (c0 + lcol + 0 < g.C) gv.x = ld<T>(gp + 0);
                    if (c0 + lcol + 1 < g.C) gv.y = ld<T>(gp + 1);
                    if (c0 + lcol + 2 < g.C) gv.z = ld<T>(gp + 2);
                    if (c0 + lcol + 3 < g.C) gv.w = ld<T>(gp + 3);
                }
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&Ds[lrow][lcol]) = dv;
        *reinterpret_cast<float4*>(&Gs[lrow][lcol]) = gv;
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < BK; ++mm) {
            const float4 x = *reinterpret_cast<const float4*>(&Ds[mm][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Gs[mm][tx * 4]);
            const float xs[4] = {x.x, x.y, x.z, x.w}, ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xs[i], ws[j], acc[i][j]);
        }
        if (EXACT) {          // parity mode: fp64 sum of the 16-row fp32 partial sums (see gather_gemm_kernel)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { tot[i][j] += (double)acc[i][j]; acc[i][j] = 0.f; }
        }
    }
    if (EXACT) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = (float)tot[i][j];
    }
    const int RS = g.R * g.S;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= g.K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c < g.C) atomicAdd(&dw[((long long)k * g.C + c) * RS + tap], acc[i][j]);
        }
    }
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int K, int C, int RS, int swap) {
    const long long total = (long long)K * C * RS;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT (coalesced writes): [A][RS][B]
        const int B = swap ? K : C;
        const int b = (int)(i % B);
        const int tap = (int)((i / B) % RS);
        const int a = (int)(i / ((long long)B * RS));
        const int k = swap ? b : a, c = swap ? a : b;
        st<T>(out + i, w[((long long)k * C + c) * RS + tap]);
    }
}

// all packed weight copies of a model in ONE launch: table rows = {src, dst, K, C, RS, swap, dtype, first_block}
struct PackEntry { const float* src; void* dst; int K, C, RS, swap, dtype, first_block; };
constexpr int PACK_BT = 64, PACK_MAX_RS = 25;
// Persistent CTAs walk the table's virtual blocks (block vb belongs to the entry with the largest first_block <= vb);
// the table sits in shared memory, so the binary search costs a few shared-memory reads instead of ~6 dependent
// global loads per 1024 elements (which, not bandwidth, bounded the one-block-per-1024-elements form at 0.6 TB/s).
__global__ void pack_multi_kernel(const PackEntry* __restrict__ gtable, int n, int total_blocks) {
    extern __shared__ __align__(16) unsigned char pack_smem[];
    PackEntry* table = reinterpret_cast<PackEntry*>(pack_smem);
    for (int i = threadIdx.x; i < n * (int)(sizeof(PackEntry) / 8); i += blockDim.x)
        reinterpret_cast<unsigned long long*>(table)[i] = reinterpret_cast<const unsigned long long*>(gtable)[i];
    // PACK_U tiles per iteration of the fast paths: four 16-byte loads per thread in flight before the first barrier (one
    // tile per iteration left the kernel latency-bound: 93 us for 22 M elements, 0.7 TB/s)
    constexpr int PACK_U = 4, TILE_F = PACK_BT * (PACK_MAX_RS + 1);
    __shared__ float tile[PACK_U * TILE_F];
    __syncthreads();
    for (int vb = blockIdx.x; vb < total_blocks; vb += gridDim.x) {
        int lo = 0, hi = n - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid].first_block <= vb) lo = mid; else hi = mid - 1;
        }
        const PackEntry e = table[lo];
        const long long total = (long long)e.K * e.C * e.RS;
        if (e.RS >= 4 && e.RS <= PACK_MAX_RS) {
            // Filters: for one output row a (= k, or c when swapped) the copy is a [b][tap] -> [tap][b] transpose of rows of
            // RS contiguous floats (row pitch RS, or C*RS when swapped).  The entry's blocks walk tiles of (a, 64 values of
            // b): coalesced reads of whole tap rows into a padded shared-memory tile, coalesced writes of b-contiguous runs.
            const int nblk = (lo + 1 < n ? table[lo + 1].first_block : total_blocks) - e.first_block;
            if (e.swap && e.RS == 16 && e.K % 32 == 0 && e.C % 2 == 0 && (reinterpret_cast<uintptr_t>(e.src) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(e.dst) & 7) == 0) {
                // swapped 4x4 filters: out[c][tap][k] = w[k][c][tap].  Tiles of 32 k x 2 c x 16 taps: every k row contributes one
                // full 128-byte line (the 64-value form read 64-byte pieces 8+ KB apart: DRAM-page bound at 0.5 TB/s), every
                // (c, tap) row of the output receives one 64-byte run of k.
                const int tiles_k = e.K / 32, ntl = (e.C / 2) * tiles_k;
                for (int tl0 = vb - e.first_block; tl0 < ntl; tl0 += PACK_U * nblk) {
                    const int kl = threadIdx.x >> 3, f = threadIdx.x & 7;
                    float4 v[PACK_U];
#pragma unroll
                    for (int u = 0; u < PACK_U; ++u) {
                        const int tl = tl0 + u * nblk;
                        if (tl < ntl) {
                            const int c0 = (tl / tiles_k) * 2, k0 = (tl % tiles_k) * 32;
                            v[u] = __ldg(reinterpret_cast<const float4*>(e.src + ((long long)(k0 + kl) * e.C + c0) * 16) + f);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PACK_U; ++u) {
                        float* d = tile + u * TILE_F + kl * 33 + f * 4;
                        if (tl0 + u * nblk < ntl) { d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w; }
                    }
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < PACK_U; ++u) {
                        const int tl = tl0 + u * nblk;
                        if (tl >= ntl) break;
                        const int c0 = (tl / tiles_k) * 2, k0 = (tl % tiles_k) * 32;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int idx2 = threadIdx.x + h * 256, kp = idx2 & 15, ct = idx2 >> 4;       // ct = c_local * 16 + tap
                            const float v0 = tile[u * TILE_F + (2 * kp) * 33 + ct], v1 = tile[u * TILE_F + (2 * kp + 1) * 33 + ct];
                            const long long o = ((long long)c0 * 16 + ct) * e.K + k0 + 2 * kp;
                            if (e.dtype == VS_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(e.dst) + o) = make_float2(v0, v1);
                            else *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(e.dst) + o) = __floats2bfloat162_rn(v0, v1);
                        }
                    }
                    __syncthreads();
                }
                continue;
            }
            const int A = e.swap ? e.C : e.K, B = e.swap ? e.K : e.C, RS = e.RS;
            const int pitch = RS | 1;
            const int tiles_b = (B + PACK_BT - 1) / PACK_BT, ntiles = A * tiles_b;
            const long long row_pitch = e.swap ? (long long)e.C * RS : RS;
            if (RS == 16 && B % PACK_BT == 0 && ((reinterpret_cast<uintptr_t>(e.src) | (uintptr_t)(row_pitch * 4)) & 15) == 0 &&
                (((uintptr_t)e.dst | (uintptr_t)(B * 2)) & 7) == 0) {
                // 4x4 filters, whole tiles only: PACK_U tiles per iteration (see above)
                for (int tl0 = vb - e.first_block; tl0 < ntiles; tl0 += PACK_U * nblk) {
                    const int row = threadIdx.x >> 2, quad = threadIdx.x & 3;
                    float4 v[PACK_U];
#pragma unroll
                    for (int u = 0; u < PACK_U; ++u) {
                        const int tl = tl0 + u * nblk;
                        if (tl < ntiles) {
                            const int a = tl / tiles_b, b0 = (tl - a * tiles_b) * PACK_BT;
                            const float* src = e.src + (e.swap ? ((long long)b0 * e.C + a) * RS : ((long long)a * e.C + b0) * RS);
                            v[u] = __ldg(reinterpret_cast<const float4*>(src + row * row_pitch) + quad);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PACK_U; ++u) {
                        float* d = tile + u * TILE_F + row * pitch + quad * 4;
                        if (tl0 + u * nblk < ntiles) { d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w; }
                    }
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < PACK_U; ++u) {
                        const int tl = tl0 + u * nblk;
                        if (tl >= ntiles) break;
                        const int a = tl / tiles_b, b0 = (tl - a * tiles_b) * PACK_BT;
                        const long long obase = (long long)a * RS * B + b0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int idx2 = threadIdx.x + h * 256, tap = idx2 >> 5, bp = idx2 & 31;
                            const float v0 = tile[u * TILE_F + (2 * bp) * pitch + tap], v1 = tile[u * TILE_F + (2 * bp + 1) * pitch + tap];
                            const long long o = obase + (long long)tap * B + 2 * bp;
                            if (e.dtype == VS_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(e.dst) + o) = make_float2(v0, v1);
                            else *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(e.dst) + o) = __floats2bfloat162_rn(v0, v1);
                        }
                    }
                    __syncthreads();
                }
                continue;
            }
            for (int tl = vb - e.first_block; tl < ntiles; tl += nblk) {
                const int a = tl / tiles_b, b0 = (tl - a * tiles_b) * PACK_BT;
                const int nb = B - b0 < PACK_BT ? B - b0 : PACK_BT;
                const float* src = e.src + (e.swap ? ((long long)b0 * e.C + a) * RS : ((long long)a * e.C + b0) * RS);
                const bool fast = RS == 16 && nb == PACK_BT && ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(row_pitch * 4)) & 15) == 0 &&
                                  (((uintptr_t)e.dst | (uintptr_t)(B * 2)) & 7) == 0;
                const long long obase = (long long)a * RS * B + b0;
                if (fast) {
                    // 4x4 filters, full tile: one float4 (four taps of one row) per thread in, two packed pairs of b per
                    // thread out; shifts instead of divisions (the element-wise loops below spend ~80 instructions per element)
                    {
                        const int row = threadIdx.x >> 2, quad = threadIdx.x & 3;
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src + row * row_pitch) + quad);
                        float* d = tile + row * pitch + quad * 4;
                        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                    }
                    __syncthreads();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int idx2 = threadIdx.x + h * 256, tap = idx2 >> 5, bp = idx2 & 31;
                        const float v0 = tile[(2 * bp) * pitch + tap], v1 = tile[(2 * bp + 1) * pitch + tap];
                        const long long o = obase + (long long)tap * B + 2 * bp;
                        if (e.dtype == VS_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(e.dst) + o) = make_float2(v0, v1);
                        else *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(e.dst) + o) = __floats2bfloat162_rn(v0, v1);
                    }
                    __syncthreads();
                    continue;
                }
                for (int idx = threadIdx.x; idx < nb * RS; idx += blockDim.x) {
                    const int row = idx / RS, tap = idx - row * RS;
                    tile[row * pitch + tap] = src[row * row_pitch + tap];
                }
                __syncthreads();
                for (int idx = threadIdx.x; idx < nb * RS; idx += blockDim.x) {
                    const int tap = idx / nb, bb = idx - tap * nb;
                    const float v = tile[bb * pitch + tap];
                    const long long o = obase + (long long)tap * B + bb;
                    if (e.dtype == VS_F32) reinterpret_cast<float*>(e.dst)[o] = v;
                    else reinterpret_cast<__nv_bfloat16*>(e.dst)[o] = __float2bfloat16_rn(v);
                }
                __syncthreads();
            }
            continue;
        }
        const long long base = (long long)(vb - e.first_block) * VS_PACK_BLOCK_ELEMS;
        for (int t = threadIdx.x; t < VS_PACK_BLOCK_ELEMS; t += blockDim.x) {
            const long long i = base + t;
            if (i >= total) break;
            const int B = e.swap ? e.K : e.C;
            const int bb = (int)(i % B);
            const int tap = (int)((i / B) % e.RS);
            const int a = (int)(i / ((long long)B * e.RS));
            const int k = e.swap ? bb : a, c = e.swap ? a : bb;
            const float v = e.src[((long long)k * e.C + c) * e.RS + tap];
            if (e.dtype == VS_F32) reinterpret_cast<float*>(e.dst)[i] = v;
            else reinterpret_cast<__nv_bfloat16*>(e.dst)[i] = __float2bfloat16_rn(v);
        }
    }
}

template <typename T>
__global__ void colsum_kernel(const T* __restrict__ a, long long rows, int C, float* __restrict__ db,
                              long long rows_per_block) {
    // block = 32 columns x 8 row lanes
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float s = 0.f;
    if (c < C)
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += ld<T>(a + r * C + c);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
        atomicAdd(&db[c], t);
    }
}

// Column sums of a wide, long matrix (bias gradient of a layer without BatchNorm: 33 MB for the first encoder layer):
// every thread owns 16 bytes of channels, four rows in flight, fp32 partial sums, one shared-memory reduction and one
// atomic per column and block.  (The 32-column form above issues one 2-byte load per thread at a time: 1.3 TB/s.)
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ a, long long rows, int C, float* __restrict__ db) {
    constexpr int V = 16 / (int)sizeof(T);
    const int tpr = C / V;                               // threads per row (C % V == 0, tpr divides 256)
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr, rpi = 256 / tpr;
    float s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = 0.f;
    const long long step = (long long)gridDim.x * rpi;
    for (long long r = (long long)blockIdx.x * rpi + rl; r < rows; r += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long rr = r + u * step;
            v[u] = rr < rows ? __ldg(reinterpret_cast<const uint4*>(a + rr * C) + cg) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T* e = reinterpret_cast<const T*>(&v[u]);
#pragma unroll
            for (int j = 0; j < V; ++j) s[j] += ld<T>(e + j);
        }
    }
    __shared__ float red[256 * 8];
#pragma unroll
    for (int j = 0; j < V; ++j) red[(rl * tpr + cg) * V + j] = s[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.f;
        for (int y = 0; y < rpi; ++y) t += red[y * C + c];
        atomicAdd(&db[c], t);
    }
}

// column sums when C <= 4: every thread strides over rows, block reduction, one atomic per column and block
template <typename T>
__global__ void __launch_bounds__(256) colsum_narrow_kernel(const T* __restrict__ a, long long rows, int C, float* __restrict__ db) {
    __shared__ float red[8][4];
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
        for (int c = 0; c < C; ++c) s[c] += ld<T>(a + r * C + c);
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] = warp_sum(s[c]);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int c = 0; c < 4; ++c) red[threadIdx.x >> 5][c] = s[c];
    __syncthreads();
    if (threadIdx.x < C) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(&db[threadIdx.x], t);
    }
}

// C == 1: a flat sum.  16-byte loads, four per thread in flight (the per-row form above issues one 2-byte load at a time)
template <typename T>
__global__ void __launch_bounds__(256) sum_flat_kernel(const T* __restrict__ a, long long n, float* __restrict__ db) {
    constexpr int V = 16 / (int)sizeof(T);
    __shared__ float red[8];
    const long long nv = n / V, stride = (long long)gridDim.x * blockDim.x;
    float s = 0.f;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < nv; i0 += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            v[u] = i < nv ? __ldg(reinterpret_cast<const uint4*>(a) + i) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T* e = reinterpret_cast<const T*>(&v[u]);
#pragma unroll
            for (int j = 0; j < V; ++j) s += ld<T>(e + j);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - nv * V)) s += ld<T>(a + nv * V + threadIdx.x);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(db, t);
    }
}

// ------------------------------------------------------------------------------------------ host
int conv_forward_simt(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                      double* stats, cudaStream_t stream) {
    GatherArgs a;
    a.N = g->N; a.R = g->R; a.S = g->S; a.stride = g->stride; a.pad = g->pad;
    a.groups = g->groups; a.act = g->act; a.transposed = mode == VS_CONV_TRANSPOSED;
    if (mode == VS_CONV_DIRECT) { a.IH = g->H; a.IW = g->W; a.IC = g->C; a.OH = g->P; a.OW = g->Q; a.OC = g->K; }
    else { a.IH = g->P; a.IW = g->Q; a.IC = g->K; a.OH = g->H; a.OW = g->W; a.OC = g->C; }
    a.n_per_group = g->N / g->groups;
    // transposed convolution of a 1x1 input (first_upconv, conv.py:258): every output pixel (h,w) has exactly
    // one contributing tap (r,s) = (h,w).  Declaring the class stride to be the filter size makes each output
    // pixel position its own parity class with a single tap, i.e. R*S plain GEMMs instead of R*S-fold padding.
    if (a.transposed && a.IH == 1 && a.IW == 1 && g->stride == 1 && g->pad == 0 && g->R == g->S && a.OH == g->R && a.OW == g->S)
        a.stride = g->R;
    const int st = a.transposed ? a.stride : 1;
    const long long Mc0 = (long long)a.N * cdiv(a.OH, st) * cdiv(a.OW, st);
    dim3 grid((unsigned)cdiv(Mc0, BM), (unsigned)cdiv(a.OC, BN), (unsigned)(st * st));
    const bool vec = (a.IC % 4) == 0;
    VS_DISPATCH_DTYPE(g->dtype, T, {
        if (vec) gather_gemm_kernel<T, true><<<grid, 256, 0, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out, stats);
        else gather_gemm_kernel<T, false><<<grid, 256, 0, stream>>>(a, (const T*)in, (const T*)wp, bias, (T*)out, stats);
    });
    return launched("gather_gemm_kernel");
}

int conv_wgrad_simt(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream) {
    const int c_tiles = (int)cdiv(g->C, BN), k_tiles = (int)cdiv(g->K, BN);
    const long long M = (long long)g->N * g->P * g->Q;
    const long long base = (long long)k_tiles * c_tiles * g->R * g->S;
    long long splits = cdiv(4LL * num_sms(), base);
    // few output tiles (the stepper's 512 x 20 linears: 8 tiles): the reduction over the rows is one dependent chain of
    // load -> barrier -> FMA -> barrier per 16 rows, so short ranges (64 rows) on many CTAs beat 256-row ranges on 56
    const long long max_splits = cdiv(M, base < num_sms() ? 64 : 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long rps = cdiv(cdiv(M, splits), BK) * BK;
    splits = cdiv(M, rps);
    VS_REQUIRE((long long)c_tiles * g->R * g->S <= 65535 && splits <= 65535, "wgrad grid too large");
    dim3 grid((unsigned)k_tiles, (unsigned)(c_tiles * g->R * g->S), (unsigned)splits);
    const bool vec = (g->C % 4) == 0 && (g->K % 4) == 0;
    VS_DISPATCH_DTYPE(g->dtype, T, {
        if (vec) wgrad_kernel<T, true><<<grid, 256, 0, stream>>>(*g, (const T*)small_, (const T*)big, dw, c_tiles, rps);
        else wgrad_kernel<T, false><<<grid, 256, 0, stream>>>(*g, (const T*)small_, (const T*)big, dw, c_tiles, rps);
    });
    return launched("wgrad_kernel");
}

}  // namespace vs

using namespace vs;

extern "C" int vs_pack_weight(const float* w, void* out, int32_t dtype, int32_t K, int32_t C, int32_t RS,
                              int32_t swap, void* stream) {
    const long long total = (long long)K * C * RS;
    if (total == 0) return 0;
    const int blocks = (int)(cdiv(total, 256) < 4096 ? cdiv(total, 256) : 4096);
    VS_DISPATCH_DTYPE(dtype, T, (pack_weight_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(w, (T*)out, K, C, RS, swap)));
    return launched("pack_weight_kernel");
}

extern "C" int vs_pack_weights_multi(const void* table, int32_t n, int32_t total_blocks, void* stream) {
    if (n <= 0 || total_blocks <= 0) return 0;
    VS_REQUIRE(n <= 1024, "pack_weights_multi: at most 1024 table rows (got %d)", n);
    VS_REQUIRE((reinterpret_cast<uintptr_t>(table) & 7) == 0, "pack_weights_multi: table must be 8-byte aligned");
    const int grid = total_blocks < 8 * num_sms() ? total_blocks : 8 * num_sms();
    pack_multi_kernel<<<(unsigned)grid, 256, (size_t)n * sizeof(PackEntry), as_stream(stream)>>>(reinterpret_cast<const PackEntry*>(table), n, total_blocks);
    return launched("pack_multi_kernel");
}

extern "C" int vs_colsum(const void* a, int32_t dtype, int64_t rows, int32_t C, float* db, void* stream) {
    if (rows == 0 || C == 0) return 0;
    if (C == 1 && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
        long long blocks = cdiv(rows, 256 * 16 * 4);
        if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
        if (blocks < 1) blocks = 1;
        VS_DISPATCH_DTYPE(dtype, T, (sum_flat_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const T*)a, rows, db)));
        return launched("sum_flat_kernel");
    }
    if (C <= 4) {
        long long blocks = cdiv(rows, 1024);
        if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
        VS_DISPATCH_DTYPE(dtype, T, (colsum_narrow_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const T*)a, rows, C, db)));
        return launched("colsum_narrow_kernel");
    }
    {
        const int V = dtype == VS_F32 ? 4 : 8;
        if (C % V == 0 && 256 % (C / V) == 0 && C / V <= 256 && rows >= 4096 && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
            long long blocks = cdiv(rows, (256 / (C / V)) * 8);
            if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
            VS_DISPATCH_DTYPE(dtype, T, (colsum_vec_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const T*)a, rows, C, db)));
            return launched("colsum_vec_kernel");
        }
    }
    const int cx = (int)cdiv(C, 32);
    long long by = cdiv(2LL * num_sms(), cx);
    if (by > cdiv(rows, 64)) by = cdiv(rows, 64);
    if (by < 1) by = 1;
    if (by > 65535) by = 65535;
    const long long rpb = cdiv(rows, by);
    dim3 grid(cx, (unsigned)cdiv(rows, rpb)), block(32, 8);
    VS_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<grid, block, 0, as_stream(stream)>>>((const T*)a, rows, C, db, rpb)));
    return launched("colsum_kernel");
}

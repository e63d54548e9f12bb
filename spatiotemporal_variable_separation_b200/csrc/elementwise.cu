// Streaming (HBM-bound) kernels: activations, residual adds, channel concat / S-code broadcast,
// pooling, nearest upsampling, layout boundary, loss reductions and the fused Adam update.
// Citations: /root/reference/var_sep/networks/utils.py:50-72, conv.py:151-169,221-228,296-314,388-394,
// resnet.py:29,69, train.py:38-42,86,139-149, main.py:145.
#include "common.cuh"

namespace vs {

static int ew_grid(long long work) {
    long long b = cdiv(work, 256);
    const long long cap = 8LL * num_sms();
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}
#define GRID_STRIDE(i, total)                                                             \
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (total);     \
         i += (long long)gridDim.x * blockDim.x)

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ out, T* __restrict__ dx, long long n,
                               int act) {
    GRID_STRIDE(i, n) st<T>(dx + i, ld<T>(dout + i) * act_grad_from_out(ld<T>(out + i), act));
}

template <typename T>
__global__ void add_act_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n,
                               int act) {
    GRID_STRIDE(i, n) st<T>(out + i, act_fwd(ld<T>(a + i) + ld<T>(b + i), act));
}

template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ src, int src_C, long long src_rows, T* __restrict__ dst,
                                     int dst_C, int dst_off, long long rows) {
    const long long total = rows * src_C;
    GRID_STRIDE(i, total) {
        const long long r = i / src_C;
        const int c = (int)(i - r * src_C);
        dst[r * dst_C + dst_off + c] = src[(r % src_rows) * src_C + c];
    }
}

template <typename T>
__global__ void slice_reduce_kernel(const T* __restrict__ ddst, int dst_C, int dst_off, long long rows,
                                    T* __restrict__ dsrc, int src_C, long long src_rows) {
    const long long total = src_rows * src_C;
    const long long reps = rows / src_rows;
    GRID_STRIDE(i, total) {
        const long long r = i / src_C;
        const int c = (int)(i - r * src_C);
        float s = 0.f;
        for (long long j = 0; j < reps; ++j) s += ld<T>(ddst + (r + j * src_rows) * dst_C + dst_off + c);
        st<T>(dsrc + i, s);
    }
}

template <typename T>
__global__ void mul_bcast_kernel(const T* __restrict__ s, long long s_rows, const T* __restrict__ t,
                                 T* __restrict__ out, long long rows, int C) {
    const long long total = rows * C;
    GRID_STRIDE(i, total) {
        const long long r = i / C;
        const int c = (int)(i - r * C);
        st<T>(out + i, ld<T>(s + (r % s_rows) * C + c) * ld<T>(t + i));
    }
}

template <typename T>
__global__ void mul_bcast_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ s, long long s_rows,
                                     const T* __restrict__ t, T* __restrict__ ds, T* __restrict__ dt, long long rows,
                                     int C) {
    const long long total = s_rows * C;
    const long long reps = rows / s_rows;
    GRID_STRIDE(i, total) {
        const float sv = ld<T>(s + i);
        float acc = 0.f;
        for (long long j = 0; j < reps; ++j) {
            const long long e = i + j * s_rows * C;
            const float d = ld<T>(dout + e);
            acc += d * ld<T>(t + e);
            if (dt) st<T>(dt + e, d * sv);
        }
        if (ds) st<T>(ds + i, acc);
    }
}

template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int P,
                                   int Q, int k, int stride, int pad) {
    const long long total = (long long)N * P * Q * C;
    GRID_STRIDE(i, total) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int q = (int)(t % Q); t /= Q;
        const int p = (int)(t % P);
        const int n = (int)(t / P);
        float best = -INFINITY;
        for (int r = 0; r < k; ++r) {
            const int h = p * stride - pad + r;
            if (h < 0 || h >= H) continue;
            for (int s = 0; s < k; ++s) {
                const int w = q * stride - pad + s;
                if (w < 0 || w >= W) continue;
                const float v = ld<T>(x + (((long long)n * H + h) * W + w) * C + c);
                if (v > best) best = v;
            }
        }
        st<T>(y + i, best);
    }
}

// gather form: each input element sums dy of the windows whose (first-in-scan-order) maximum it is
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int N,
                                   int H, int W, int C, int P, int Q, int k, int stride, int pad) {
    const long long total = (long long)N * H * W * C;
    GRID_STRIDE(i, total) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        const float xv = ld<T>(x + i);
        float acc = 0.f;
        // windows p with p*stride - pad <= h <= p*stride - pad + k - 1
        int p_lo = (h + pad - k + 1 + stride - 1) / stride; if (h + pad - k + 1 < 0) p_lo = 0;
        int q_lo = (w + pad - k + 1 + stride - 1) / stride; if (w + pad - k + 1 < 0) q_lo = 0;
        const int p_hi = min((h + pad) / stride, P - 1), q_hi = min((w + pad) / stride, Q - 1);
        for (int p = p_lo; p <= p_hi; ++p)
            for (int q = q_lo; q <= q_hi; ++q) {
                // is (h,w) the first maximum of window (p,q)?
                float best = -INFINITY; int bh = -1, bw = -1;
                for (int r = 0; r < k; ++r) {
                    const int hh = p * stride - pad + r;
                    if (hh < 0 || hh >= H) continue;
                    for (int s = 0; s < k; ++s) {
                        const int ww = q * stride - pad + s;
                        if (ww < 0 || ww >= W) continue;
                        const float v = (hh == h && ww == w) ? xv : ld<T>(x + (((long long)n * H + hh) * W + ww) * C + c);
                        if (v > best) { best = v; bh = hh; bw = ww; }
                    }
                }
                if (bh == h && bw == w) acc += ld<T>(dy + (((long long)n * P + p) * Q + q) * C + c);
            }
        st<T>(dx + i, acc);
    }
}

template <typename T>
__global__ void upsample2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
    const long long total = (long long)N * 2 * H * 2 * W * C;
    GRID_STRIDE(i, total) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w2 = (int)(t % (2 * W)); t /= 2 * W;
        const int h2 = (int)(t % (2 * H));
        const int n = (int)(t / (2 * H));
        y[i] = x[(((long long)n * H + (h2 >> 1)) * W + (w2 >> 1)) * C + c];
    }
}

template <typename T>
__global__ void upsample2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C) {
    const long long total = (long long)N * H * W * C;
    GRID_STRIDE(i, total) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        const long long base = (((long long)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
        const long long rs = (long long)2 * W * C;
        st<T>(dx + i, ld<T>(dy + base) + ld<T>(dy + base + C) + ld<T>(dy + base + rs) + ld<T>(dy + base + rs + C));
    }
}

template <typename T>
__global__ void frames_to_nhwc_kernel(const float* __restrict__ frames, int B, int T_, int Cf, int H, int W, int t0,
                                      int nt, T* __restrict__ out) {
    const int Cn = nt * Cf;
    const long long total = (long long)B * H * W * Cn;
    GRID_STRIDE(i, total) {
        // iterate in INPUT order over (b, t, cf, h, w) so reads are coalesced; writes are strided by Cn
        const int w = (int)(i % W);
        long long t = i / W;
        const int h = (int)(t % H); t /= H;
        const int cf = (int)(t % Cf); t /= Cf;
        const int tt = (int)(t % nt);
        const int b = (int)(t / nt);
        const float v = frames[((((long long)b * T_ + t0 + tt) * Cf + cf) * H + h) * W + w];
        st<T>(out + (((long long)b * H + h) * W + w) * Cn + tt * Cf + cf, v);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int N, int C, int H, int W) {
    const long long total = (long long)N * C * H * W;
    GRID_STRIDE(i, total) {   // i over the NCHW output
        const int w = (int)(i % W);
        long long t = i / W;
        const int h = (int)(t % H); t /= H;
        const int c = (int)(t % C);
        const int n = (int)(t / C);
        out[i] = ld<T>(in + (((long long)n * H + h) * W + w) * C + c);
    }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int N, int C, int H, int W) {
    const long long total = (long long)N * C * H * W;
    GRID_STRIDE(i, total) {   // i over the NCHW input
        const int w = (int)(i % W);
        long long t = i / W;
        const int h = (int)(t % H); t /= H;
        const int c = (int)(t % C);
        const int n = (int)(t / C);
        st<T>(out + (((long long)n * H + h) * W + w) * C + c, in[i]);
    }
}

// C == 1 or H*W == 1: NCHW and NHWC coincide, the layout change is a plain cast (four elements per thread and iteration)
template <typename TI, typename TO>
__global__ void cast4_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n4, long long n) {
    GRID_STRIDE(i, n4) st4<TO>(out + 4 * i, ld4<TI>(in + 4 * i));
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) st<TO>(out + 4 * n4 + threadIdx.x, ld<TI>(in + 4 * n4 + threadIdx.x));
}

// frames_to_nhwc for windows of at most 16 folded channels: a block owns 256 consecutive pixels, every thread gathers the
// Cn values of its pixel (each a coalesced read of one (frame, channel) plane) into shared memory and the block writes
// its contiguous 256*Cn-element output range with 16-byte stores (the element-wise form wrote 2 bytes at a 2*Cn-byte
// stride and spent a chain of 64-bit divisions per element).
template <typename T>
__global__ void frames_to_nhwc_pix_kernel(const float* __restrict__ frames, int T_, int Cf, int HW, int t0, int Cn,
                                          long long npix, T* __restrict__ out) {
    __shared__ __align__(16) T tile[256 * 16];
    for (long long p0 = (long long)blockIdx.x * 256; p0 < npix; p0 += (long long)gridDim.x * 256) {
        const long long pp = p0 + threadIdx.x;
        if (pp < npix) {
            const long long b = pp / HW;
            const int hw = (int)(pp - b * HW);
            const float* src = frames + ((b * T_ + t0) * Cf) * HW + hw;          // window planes are consecutive: (t0 + tt, cf)
            for (int ch = 0; ch < Cn; ++ch) st<T>(&tile[threadIdx.x * Cn + ch], __ldg(src + (long long)ch * HW));
        }
        __syncthreads();
        const long long left = npix - p0;
        const int elems = (int)(left < 256 ? left : 256) * Cn;
        T* dst = out + p0 * Cn;
        constexpr int V = 16 / sizeof(T);
        if (elems % V == 0) {          // block base is 16-byte aligned: 256 * Cn * sizeof(T) per block
            for (int v = threadIdx.x; v < elems / V; v += 256)
                reinterpret_cast<uint4*>(dst)[v] = reinterpret_cast<const uint4*>(tile)[v];
        } else {
            for (int v = threadIdx.x; v < elems; v += 256) dst[v] = tile[v];
        }
        __syncthreads();
    }
}

__global__ void sqdiff_sum_kernel(const float* __restrict__ a, long long a_sb, long long a_st,
                                  const float* __restrict__ b, long long b_sb, long long b_st, long long B,
                                  long long T_, long long L, double* __restrict__ acc) {
    __shared__ double red[8];
    const long long total = B * T_ * L;
    double s = 0.0;
    GRID_STRIDE(i, total) {
        const long long l = i % L;
        const long long t = (i / L) % T_;
        const long long bb = i / (L * T_);
        const float av = a[bb * a_sb + t * a_st + l];
        const float d = b ? av - b[bb * b_sb + t * b_st + l] : av;
        s += (double)(d * d);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        atomicAdd(acc, t);
    }
}

__global__ void sqdiff_bwd_kernel(const float* __restrict__ a, long long a_sb, long long a_st,
                                  const float* __restrict__ b, long long b_sb, long long b_st, long long B,
                                  long long T_, long long L, float scale, const float* __restrict__ g_term,
                                  const float* __restrict__ g_total, float lamb, float* __restrict__ da,
                                  int accumulate) {
    const long long total = B * T_ * L;
    const float sc = scale * ((g_term ? g_term[0] : 0.f) + (g_total ? lamb * g_total[0] : 0.f));
    GRID_STRIDE(i, total) {
        const long long l = i % L;
        const long long t = (i / L) % T_;
        const long long bb = i / (L * T_);
        const long long ia = bb * a_sb + t * a_st + l;
        const float av = a[ia];
        const float d = b ? av - b[bb * b_sb + t * b_st + l] : av;
        if (accumulate) da[ia] += sc * d; else da[ia] = sc * d;
    }
}

// Vectorised forms for rows of L contiguous floats (L % 4 == 0, 16-byte aligned rows, < 2^31 float4 groups): one float4
// per operand and iteration, 32-bit index arithmetic (the scalar kernels spend three 64-bit divisions per element and ran
// at 0.7 TB/s on the 63 MB forecast term), four squares summed in fp32 and added to the thread's fp64 sum.
__global__ void sqdiff_sum_vec_kernel(const float* __restrict__ a, long long a_sb, long long a_st,
                                      const float* __restrict__ b, long long b_sb, long long b_st, unsigned T_,
                                      unsigned L4, unsigned total4, double* __restrict__ acc) {
    __shared__ double red[8];
    double s = 0.0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const unsigned row = i / L4, l4 = i - row * L4;
        const unsigned bb = row / T_, t = row - bb * T_;
        const float4 av = __ldg(reinterpret_cast<const float4*>(a + bb * a_sb + t * a_st) + l4);
        float4 d = av;
        if (b) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(b + bb * b_sb + t * b_st) + l4);
            d.x -= bv.x; d.y -= bv.y; d.z -= bv.z; d.w -= bv.w;
        }
        s += (double)(fmaf(d.x, d.x, d.y * d.y) + fmaf(d.z, d.z, d.w * d.w));
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        atomicAdd(acc, t);
    }
}

__global__ void sqdiff_bwd_vec_kernel(const float* __restrict__ a, long long a_sb, long long a_st,
                                      const float* __restrict__ b, long long b_sb, long long b_st, unsigned T_,
                                      unsigned L4, unsigned total4, float scale, const float* __restrict__ g_term,
                                      const float* __restrict__ g_total, float lamb, float* __restrict__ da,
                                      int accumulate) {
    const float sc = scale * ((g_term ? g_term[0] : 0.f) + (g_total ? lamb * g_total[0] : 0.f));
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const unsigned row = i / L4, l4 = i - row * L4;
        const unsigned bb = row / T_, t = row - bb * T_;
        const long long ia = bb * a_sb + t * a_st;
        float4 d = __ldg(reinterpret_cast<const float4*>(a + ia) + l4);
        if (b) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(b + bb * b_sb + t * b_st) + l4);
            d.x -= bv.x; d.y -= bv.y; d.z -= bv.z; d.w -= bv.w;
        }
        float4* dst = reinterpret_cast<float4*>(da + ia) + l4;
        float4 o = make_float4(sc * d.x, sc * d.y, sc * d.z, sc * d.w);
        if (accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
        *dst = o;
    }
}

static bool sqdiff_vec_ok(const float* a, long long a_sb, long long a_st, const float* b, long long b_sb, long long b_st,
                          long long B, long long T_, long long L, const float* da) {
    if (L % 4 || a_sb % 4 || a_st % 4 || b_sb % 4 || b_st % 4) return false;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(da)) & 15) return false;
    return B * T_ * (L / 4) < (1LL << 31) && T_ < (1LL << 31);
}

struct CombineArgs { double coef[8]; double lamb[8]; int n; };
__global__ void loss_combine_kernel(const double* __restrict__ acc, CombineArgs c, float* __restrict__ terms) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float total = 0.f;
    for (int i = 0; i < c.n; ++i) {
        const float t = (float)(acc[i] * c.coef[i]);
        terms[i] = t;
        total += (float)c.lamb[i] * t;
    }
    terms[c.n] = total;
}

// torch.optim.Adam single-tensor math (eps outside the bias-corrected sqrt)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float b1, float b2, float eps, float lr,
                            float grad_scale, int step_host, const int* __restrict__ step_dev,
                            const float* __restrict__ lr_dev) {
    // bias corrections in fp64 from the (host or device-resident) step count; a device counter keeps a
    // captured CUDA graph valid across replays
    const int step = step_dev ? *step_dev : step_host;
    if (lr_dev) lr = *lr_dev;               // device-resident learning rate: a schedule does not invalidate a captured graph
    const float step_size = (float)((double)lr / (1.0 - pow((double)b1, (double)step)));
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)step)));
    const long long n4 = n >> 2;
    GRID_STRIDE(i, n4) {
        float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i],
               vv = reinterpret_cast<float4*>(v)[i];
        float4 gv = reinterpret_cast<const float4*>(g)[i];
#define VS_ADAM1(P, G, M, V)                                 \
        G *= grad_scale;                                     \
        M = b1 * M + (1.f - b1) * G;                         \
        V = b2 * V + (1.f - b2) * G * G;                     \
        P -= step_size * (M / (sqrtf(V) * inv_sqrt_bc2 + eps));
        VS_ADAM1(pv.x, gv.x, mv.x, vv.x) VS_ADAM1(pv.y, gv.y, mv.y, vv.y)
        VS_ADAM1(pv.z, gv.z, mv.z, vv.z) VS_ADAM1(pv.w, gv.w, mv.w, vv.w)
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail
    const long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        float pv = p[i], mv = m[i], vv = v[i];
        float gv = g[i];
        VS_ADAM1(pv, gv, mv, vv)
        p[i] = pv; m[i] = mv; v[i] = vv;
    }
#undef VS_ADAM1
}

}  // namespace vs

using namespace vs;
#define S_ as_stream(stream)

extern "C" int vs_act_backward(const void* dout, const void* out, void* dx, int32_t dtype, int64_t n, int32_t act,
                               void* stream) {
    if (n == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (act_bwd_kernel<T><<<ew_grid(n), 256, 0, S_>>>((const T*)dout, (const T*)out, (T*)dx, n, act)));
    return launched("act_bwd_kernel");
}

extern "C" int vs_add_act(const void* a, const void* b, void* out, int32_t dtype, int64_t n, int32_t act, void* stream) {
    if (n == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (add_act_kernel<T><<<ew_grid(n), 256, 0, S_>>>((const T*)a, (const T*)b, (T*)out, n, act)));
    return launched("add_act_kernel");
}

extern "C" int vs_copy_channels(const void* src, int32_t src_C, int64_t src_rows_mod, void* dst, int32_t dst_C,
                                int32_t dst_off, int64_t rows, int32_t dtype, void* stream) {
    VS_REQUIRE(dst_off >= 0 && dst_off + src_C <= dst_C && src_rows_mod > 0, "copy_channels: bad channel window");
    if (rows == 0 || src_C == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (copy_channels_kernel<T><<<ew_grid(rows * src_C), 256, 0, S_>>>(
        (const T*)src, src_C, src_rows_mod, (T*)dst, dst_C, dst_off, rows)));
    return launched("copy_channels_kernel");
}

extern "C" int vs_slice_channels_reduce(const void* ddst, int32_t dst_C, int32_t dst_off, int64_t rows, void* dsrc,
                                        int32_t src_C, int64_t src_rows, int32_t dtype, void* stream) {
    VS_REQUIRE(src_rows > 0 && rows % src_rows == 0, "slice_channels_reduce: rows must be a multiple of src_rows");
    if (rows == 0 || src_C == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (slice_reduce_kernel<T><<<ew_grid(src_rows * src_C), 256, 0, S_>>>(
        (const T*)ddst, dst_C, dst_off, rows, (T*)dsrc, src_C, src_rows)));
    return launched("slice_reduce_kernel");
}

extern "C" int vs_mul_bcast(const void* s, int64_t s_rows, const void* t, void* out, int64_t rows, int32_t C,
                            int32_t dtype, void* stream) {
    VS_REQUIRE(s_rows > 0 && rows % s_rows == 0, "mul_bcast: rows must be a multiple of s_rows");
    if (rows == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (mul_bcast_kernel<T><<<ew_grid(rows * C), 256, 0, S_>>>((const T*)s, s_rows, (const T*)t, (T*)out, rows, C)));
    return launched("mul_bcast_kernel");
}

extern "C" int vs_mul_bcast_backward(const void* dout, const void* s, int64_t s_rows, const void* t, void* ds,
                                     void* dt, int64_t rows, int32_t C, int32_t dtype, void* stream) {
    VS_REQUIRE(s_rows > 0 && rows % s_rows == 0, "mul_bcast_backward: rows must be a multiple of s_rows");
    if (rows == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (mul_bcast_bwd_kernel<T><<<ew_grid(s_rows * C), 256, 0, S_>>>(
        (const T*)dout, (const T*)s, s_rows, (const T*)t, (T*)ds, (T*)dt, rows, C)));
    return launched("mul_bcast_bwd_kernel");
}

static inline int pool_out(int H, int k, int stride, int pad) { return (H + 2 * pad - k) / stride + 1; }

extern "C" int vs_maxpool_forward(const void* x, void* y, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                                  int32_t k, int32_t stride, int32_t pad, void* stream) {
    const int P = pool_out(H, k, stride, pad), Q = pool_out(W, k, stride, pad);
    const long long total = (long long)N * P * Q * C;
    if (total == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (maxpool_fwd_kernel<T><<<ew_grid(total), 256, 0, S_>>>((const T*)x, (T*)y, N, H, W, C, P, Q, k, stride, pad)));
    return launched("maxpool_fwd_kernel");
}

extern "C" int vs_maxpool_backward(const void* x, const void* dy, void* dx, int32_t dtype, int32_t N, int32_t H,
                                   int32_t W, int32_t C, int32_t k, int32_t stride, int32_t pad, void* stream) {
    const int P = pool_out(H, k, stride, pad), Q = pool_out(W, k, stride, pad);
    const long long total = (long long)N * H * W * C;
    if (total == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (maxpool_bwd_kernel<T><<<ew_grid(total), 256, 0, S_>>>((const T*)x, (const T*)dy, (T*)dx, N, H, W, C, P, Q, k, stride, pad)));
    return launched("maxpool_bwd_kernel");
}

extern "C" int vs_upsample2_forward(const void* x, void* y, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                                    void* stream) {
    const long long total = (long long)N * 4 * H * W * C;
    if (total == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (upsample2_fwd_kernel<T><<<ew_grid(total), 256, 0, S_>>>((const T*)x, (T*)y, N, H, W, C)));
    return launched("upsample2_fwd_kernel");
}

extern "C" int vs_upsample2_backward(const void* dy, void* dx, int32_t dtype, int32_t N, int32_t H, int32_t W,
                                     int32_t C, void* stream) {
    const long long total = (long long)N * H * W * C;
    if (total == 0) return 0;
    VS_DISPATCH_DTYPE(dtype, T, (upsample2_bwd_kernel<T><<<ew_grid(total), 256, 0, S_>>>((const T*)dy, (T*)dx, N, H, W, C)));
    return launched("upsample2_bwd_kernel");
}

extern "C" int vs_frames_to_nhwc(const float* frames, int32_t B, int32_t T_, int32_t Cf, int32_t H, int32_t W,
                                 int32_t t0, int32_t nt, void* out, int32_t dtype, void* stream) {
    VS_REQUIRE(t0 >= 0 && nt >= 1 && t0 + nt <= T_, "frames_to_nhwc: window [%d,%d) outside %d frames", t0, t0 + nt, T_);
    const long long total = (long long)B * nt * Cf * H * W;
    if (total == 0) return 0;
    if (nt * Cf <= 16 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const long long npix = (long long)B * H * W;
        long long blocks = cdiv(npix, 256);
        if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
        VS_DISPATCH_DTYPE(dtype, T, (frames_to_nhwc_pix_kernel<T><<<(unsigned)blocks, 256, 0, S_>>>(frames, T_, Cf, H * W, t0, nt * Cf, npix, (T*)out)));
        return launched("frames_to_nhwc_pix_kernel");
    }
    VS_DISPATCH_DTYPE(dtype, T, (frames_to_nhwc_kernel<T><<<ew_grid(total), 256, 0, S_>>>(frames, B, T_, Cf, H, W, t0, nt, (T*)out)));
    return launched("frames_to_nhwc_kernel");
}

extern "C" int vs_nhwc_to_nchw(const void* in, int32_t dtype, float* out, int32_t N, int32_t C, int32_t H, int32_t W,
                               void* stream) {
    const long long total = (long long)N * C * H * W;
    if (total == 0) return 0;
    if ((C == 1 || H * W == 1) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        VS_DISPATCH_DTYPE(dtype, T, (cast4_kernel<T, float><<<ew_grid(cdiv(total, 4)), 256, 0, S_>>>((const T*)in, out, total / 4, total)));
        return launched("cast4_kernel");
    }
    VS_DISPATCH_DTYPE(dtype, T, (nhwc_to_nchw_kernel<T><<<ew_grid(total), 256, 0, S_>>>((const T*)in, out, N, C, H, W)));
    return launched("nhwc_to_nchw_kernel");
}

extern "C" int vs_nchw_to_nhwc(const float* in, void* out, int32_t dtype, int32_t N, int32_t C, int32_t H, int32_t W,
                               void* stream) {
    const long long total = (long long)N * C * H * W;
    if (total == 0) return 0;
    if ((C == 1 || H * W == 1) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        VS_DISPATCH_DTYPE(dtype, T, (cast4_kernel<float, T><<<ew_grid(cdiv(total, 4)), 256, 0, S_>>>(in, (T*)out, total / 4, total)));
        return launched("cast4_kernel");
    }
    VS_DISPATCH_DTYPE(dtype, T, (nchw_to_nhwc_kernel<T><<<ew_grid(total), 256, 0, S_>>>(in, (T*)out, N, C, H, W)));
    return launched("nchw_to_nhwc_kernel");
}

extern "C" int vs_sqdiff_sum(const float* a, int64_t a_sb, int64_t a_st, const float* b, int64_t b_sb, int64_t b_st,
                             int64_t B, int64_t T_, int64_t L, double* acc, void* stream) {
    const long long total = B * T_ * L;
    if (total == 0) return 0;
    if (sqdiff_vec_ok(a, a_sb, a_st, b, b_sb, b_st, B, T_, L, nullptr)) {
        const long long total4 = total / 4;
        int blocks = (int)cdiv(total4, 256 * 4);           // four float4 per thread and operand in flight
        if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
        sqdiff_sum_vec_kernel<<<blocks, 256, 0, S_>>>(a, a_sb, a_st, b, b_sb, b_st, (unsigned)T_, (unsigned)(L / 4), (unsigned)total4, acc);
        return launched("sqdiff_sum_vec_kernel");
    }
    int blocks = ew_grid(total);
    if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
    sqdiff_sum_kernel<<<blocks, 256, 0, S_>>>(a, a_sb, a_st, b, b_sb, b_st, B, T_, L, acc);
    return launched("sqdiff_sum_kernel");
}

extern "C" int vs_sqdiff_backward(const float* a, int64_t a_sb, int64_t a_st, const float* b, int64_t b_sb,
                                  int64_t b_st, int64_t B, int64_t T_, int64_t L, float scale, const float* g_term,
                                  const float* g_total, float lamb, float* da, int32_t accumulate, void* stream) {
    const long long total = B * T_ * L;
    if (total == 0) return 0;
    if (sqdiff_vec_ok(a, a_sb, a_st, b, b_sb, b_st, B, T_, L, da)) {
        const long long total4 = total / 4;
        sqdiff_bwd_vec_kernel<<<ew_grid(total4), 256, 0, S_>>>(a, a_sb, a_st, b, b_sb, b_st, (unsigned)T_, (unsigned)(L / 4), (unsigned)total4,
                                                               scale, g_term, g_total, lamb, da, accumulate);
        return launched("sqdiff_bwd_vec_kernel");
    }
    sqdiff_bwd_kernel<<<ew_grid(total), 256, 0, S_>>>(a, a_sb, a_st, b, b_sb, b_st, B, T_, L, scale, g_term, g_total, lamb, da, accumulate);
    return launched("sqdiff_bwd_kernel");
}

extern "C" int vs_loss_combine(const double* acc, const double* coef_host, const double* lamb_host, int32_t n,
                               float* terms, void* stream) {
    VS_REQUIRE(n >= 1 && n <= 8, "loss_combine: n must be in [1,8]");
    CombineArgs c;
    c.n = n;
    for (int i = 0; i < n; ++i) { c.coef[i] = coef_host[i]; c.lamb[i] = lamb_host[i]; }
    loss_combine_kernel<<<1, 32, 0, S_>>>(acc, c, terms);
    return launched("loss_combine_kernel");
}

extern "C" int vs_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                            float beta2, float eps, float grad_scale, int32_t step_host, const int32_t* step_dev,
                            const float* lr_dev, void* stream) {
    VS_REQUIRE(step_dev != nullptr || step_host >= 1, "adam_step: step must be >= 1");
    VS_REQUIRE((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) % 16 == 0, "adam_step: arenas must be 16-byte aligned");
    if (n == 0) return 0;
    int blocks = ew_grid(cdiv(n, 4));
    adam_kernel<<<blocks, 256, 0, S_>>>(param, grad, m, v, n, beta1, beta2, eps, lr, grad_scale, step_host, step_dev, lr_dev);
    return launched("adam_kernel");
}

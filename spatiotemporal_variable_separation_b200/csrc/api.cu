// C-ABI plumbing: error text, launch accounting, and the convolution dispatch (tcgen05 path for the
// eligible bf16 layers, CUDA-core gather-GEMM for everything else).
#include <stdarg.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace vs {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};

int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return 1;
}

int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
    return 0;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev;
}

int num_sms() {
    static int cache[64] = {};
    const int dev = current_device();
    int& n = cache[(dev >= 0 && dev < 64) ? dev : 0];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 148;   // B200
        }
    }
    return n;
}

int conv_forward_simt(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                      double* stats, cudaStream_t stream);
int conv_wgrad_tc(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream);
// streaming kernels for layers with a handful of channels on one side (same return convention)
int conv_forward_thin(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                      double* stats, cudaStream_t stream);
int conv_wgrad_thin(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream);
int conv_wgrad_simt(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream);
// tensor-core paths: 0 = done, -1 = geometry not eligible (fall back to the CUDA-core kernels), >0 = error
int conv_forward_im2col(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                        double* stats, cudaStream_t stream);
int conv_forward_col2im(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                        double* stats, cudaStream_t stream, const BnApplyArgs* bn = nullptr);
int conv_wgrad_im2col(const vs_conv_geom* g, const void* small_, const void* big, float* dw, cudaStream_t stream,
                      const BnApplyArgs* bn = nullptr);
int conv_forward_col2im_eligible(const vs_conv_geom* g, int mode);
int tail_eligible(const vs_conv_geom* g);
int tail_bn_backward(const vs_conv_geom* g, const BnBwdArgs& bb, const void* dout, const void* wp, int phase, double* sums, void* dy,
                     cudaStream_t stream);
void bn_param_grad(const double* sums, int G, int C, float* dgamma, float* dbeta, cudaStream_t stream);   // bn.cu
int conv_forward_tc(const vs_conv_geom* g, int mode, const void* in, const void* wp, const float* bias, void* out,
                    double* stats, cudaStream_t stream);

static int check_geom(const vs_conv_geom* g) {
    VS_REQUIRE(g != nullptr, "null geometry");
    VS_REQUIRE(g->dtype == VS_F32 || g->dtype == VS_BF16, "unsupported dtype %d", g->dtype);
    VS_REQUIRE(g->N >= 1 && g->H >= 1 && g->W >= 1 && g->C >= 1 && g->K >= 1 && g->R >= 1 && g->S >= 1 && g->stride >= 1 && g->pad >= 0,
               "bad convolution geometry");
    VS_REQUIRE(g->P == (g->H + 2 * g->pad - g->R) / g->stride + 1 && g->Q == (g->W + 2 * g->pad - g->S) / g->stride + 1,
               "P/Q inconsistent with H/W, filter, stride and pad");
    VS_REQUIRE(g->groups >= 1 && g->N % g->groups == 0, "N=%d not divisible by groups=%d", g->N, g->groups);
    return 0;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_abi_version(void) { return 1; }
extern "C" const char* vs_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t vs_launch_count(void) { return g_launches.load(); }

extern "C" int vs_conv_forward(const vs_conv_geom* g, int32_t mode, const void* in, const void* wp, const float* bias,
                               void* out, double* stats, void* stream) {
    if (int rc = check_geom(g)) return rc;
    VS_REQUIRE(mode == VS_CONV_DIRECT || mode == VS_CONV_TRANSPOSED, "bad mode %d", mode);
    VS_REQUIRE(stats == nullptr || g->act == VS_ACT_NONE, "statistics are taken of the pre-activation: act must be NONE");
    // last decoder up-convolution (k4 s2 p1 to one or two channels): GEMM + col2im, the input is read once
    int rc = conv_forward_col2im(g, mode, in, wp, bias, out, stats, as_stream(stream));
    if (rc >= 0) return rc;
    // 64 -> 1 channel transposed convolution: the dedicated streaming kernel reads the input once (the tensor-core
    // tap GEMM would fetch it 16 times to fill 1 of 64 accumulator columns)
    if (mode == VS_CONV_TRANSPOSED && g->C == 1 && g->K == 64 && g->R == 4 && g->stride == 2)
        rc = conv_forward_thin(g, mode, in, wp, bias, out, stats, as_stream(stream));
    if (rc >= 0) return rc;
    rc = conv_forward_im2col(g, mode, in, wp, bias, out, stats, as_stream(stream));
    if (rc >= 0) return rc;
    rc = conv_forward_tc(g, mode, in, wp, bias, out, stats, as_stream(stream));
    if (rc >= 0) return rc;
    rc = conv_forward_thin(g, mode, in, wp, bias, out, stats, as_stream(stream));
    if (rc >= 0) return rc;
    return conv_forward_simt(g, mode, in, wp, bias, out, stats, as_stream(stream));
}

// which kernel family vs_conv_forward dispatches this geometry to: 0 = CUDA-core gather GEMM, 1 = tcgen05 tap GEMM,
// 2 = thin streaming kernel, 3 = tcgen05 GEMM over a CTA-built im2col tile, 4 = tcgen05 GEMM + col2im.  Pure host logic
// (no launch, assumes no fused statistics); mirrors the dispatch order above.
namespace vs {
int conv_forward_col2im_eligible(const vs_conv_geom* g, int mode);
int conv_forward_im2col_eligible(const vs_conv_geom* g, int mode);
int conv_forward_tc_eligible(const vs_conv_geom* g, int mode);
int conv_forward_thin_eligible(const vs_conv_geom* g, int mode);
int conv_forward_tc_variant(const vs_conv_geom* g, int mode);
}
extern "C" int vs_conv_forward_path(const vs_conv_geom* g, int32_t mode) {
    if (check_geom(g)) return -1;
    if (conv_forward_col2im_eligible(g, mode)) return 4;
    if (mode == VS_CONV_TRANSPOSED && g->C == 1 && g->K == 64 && g->R == 4 && g->stride == 2 && conv_forward_thin_eligible(g, mode)) return 2;
    if (conv_forward_im2col_eligible(g, mode)) return 3;
    if (conv_forward_tc_eligible(g, mode)) return 1;
    if (conv_forward_thin_eligible(g, mode)) return 2;
    return 0;
}

extern "C" int vs_conv_forward_variant(const vs_conv_geom* g, int32_t mode) {
    if (vs_conv_forward_path(g, mode) != 1) return 0;
    return conv_forward_tc_variant(g, mode);
}

extern "C" int vs_conv_wgrad(const vs_conv_geom* g, const void* small_, const void* big, float* dw, void* stream) {
    if (int rc = check_geom(g)) return rc;
    int rc = conv_wgrad_im2col(g, small_, big, dw, as_stream(stream));
    if (rc >= 0) return rc;
    rc = conv_wgrad_thin(g, small_, big, dw, as_stream(stream));
    if (rc >= 0) return rc;
    rc = conv_wgrad_tc(g, small_, big, dw, as_stream(stream));
    if (rc >= 0) return rc;
    return conv_wgrad_simt(g, small_, big, dw, as_stream(stream));
}


// ---------------------------------------------------------------------------------------------- fused decoder tail
static int tail_check(const vs_conv_geom* g, int32_t bn_groups, int32_t bn_act) {
    if (int rc = check_geom(g)) return rc;
    VS_REQUIRE(bn_act_supported(bn_act), "tail: the BatchNorm block's activation must be none, ReLU or LeakyReLU (got %d)", bn_act);
    VS_REQUIRE(bn_groups >= 1 && g->N % bn_groups == 0, "tail: N=%d not divisible by the BatchNorm groups=%d", g->N, bn_groups);
    VS_REQUIRE(vs_tail_eligible(g) == 1, "tail: geometry not eligible for the fused kernels (query vs_tail_eligible first)");
    return 0;
}

extern "C" int vs_tail_eligible(const vs_conv_geom* g) {
    if (check_geom(g)) return -1;
    return (conv_forward_col2im_eligible(g, VS_CONV_TRANSPOSED) && tail_eligible(g)) ? 1 : 0;
}

extern "C" int vs_tail_forward(const vs_conv_geom* g, const void* y, const float* mean, const float* invstd, const float* gamma,
                               const float* beta, int32_t bn_groups, int32_t bn_act, const void* wp, const float* bias, void* out,
                               void* stream) {
    if (int rc = tail_check(g, bn_groups, bn_act)) return rc;
    BnApplyArgs bn = {mean, invstd, gamma, beta, g->N / bn_groups, bn_act_neg_slope(bn_act)};
    int rc = conv_forward_col2im(g, VS_CONV_TRANSPOSED, y, wp, bias, out, nullptr, as_stream(stream), &bn);
    VS_REQUIRE(rc >= 0, "tail_forward: operands not 16-byte aligned");
    return rc;
}

extern "C" int vs_tail_wgrad(const vs_conv_geom* g, const void* y, const float* mean, const float* invstd, const float* gamma,
                             const float* beta, int32_t bn_groups, int32_t bn_act, const void* dout, float* dw, void* stream) {
    if (int rc = tail_check(g, bn_groups, bn_act)) return rc;
    BnApplyArgs bn = {mean, invstd, gamma, beta, g->N / bn_groups, bn_act_neg_slope(bn_act)};
    int rc = conv_wgrad_im2col(g, y, dout, dw, as_stream(stream), &bn);
    VS_REQUIRE(rc >= 0, "tail_wgrad: operands not 16-byte aligned");
    return rc;
}

extern "C" int vs_tail_bn_backward(const vs_conv_geom* g, const void* y, const float* mean, const float* invstd, const float* gamma,
                                   const float* beta, int32_t bn_groups, int32_t bn_act, const void* dout, const void* wp_direct,
                                   int32_t phase, int32_t train, double* sums, void* dy, float* dgamma, float* dbeta, void* stream) {
    if (int rc = tail_check(g, bn_groups, bn_act)) return rc;
    VS_REQUIRE(phase == 0 || phase == 1, "tail_bn_backward: phase must be 0 (reduce) or 1 (apply)");
    VS_REQUIRE(sums != nullptr && (phase == 0 || dy != nullptr), "tail_bn_backward: null output");
    BnBwdArgs bb;
    bb.y = reinterpret_cast<const __nv_bfloat16*>(y);
    bb.mean = mean; bb.invstd = invstd; bb.gamma = gamma; bb.beta = beta; bb.sums = sums;
    bb.n_per_group = g->N / bn_groups; bb.neg_slope = bn_act_neg_slope(bn_act); bb.train = train;
    bb.inv_count = 1.f / (float)((long long)bb.n_per_group * g->P * g->Q);
    int rc = tail_bn_backward(g, bb, dout, wp_direct, phase, sums, dy, as_stream(stream));
    VS_REQUIRE(rc >= 0, "tail_bn_backward: operands not 16-byte aligned");
    if (rc) return rc;
    if (phase == 1 && (dgamma != nullptr || dbeta != nullptr)) {
        bn_param_grad(sums, bn_groups, g->K, dgamma, dbeta, as_stream(stream));
        rc = launched("bn_param_grad_kernel");
    }
    return rc;
}

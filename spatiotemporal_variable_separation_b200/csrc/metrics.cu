// Evaluation metrics on the device: per (image, channel) plane the mean SSIM (11x11 Gaussian window by default) and the
// mean squared error, from which the callers of get_forecast in the reference's test/* scripts derive MSE / PSNR / SSIM
// (/root/reference/var_sep/test/mnist/test.py:136-141, test/utils.py:19-24, utils/ssim.py:81-149).
// One CTA per plane: both planes and the window weights are staged in shared memory once, every thread evaluates
// output pixels with a direct fs x fs window (the planes are at most 64 x 64: 54 x 54 outputs x 121 taps), block
// reduction of the two sums.  HBM traffic = the two planes read once (+ the optional SSIM map written once).
#include "common.cuh"
#include "metrics_core.h"

namespace vs {

constexpr int SSIM_MAX_PIX = 4096, SSIM_MAX_FS = 15;

__global__ void __launch_bounds__(256) ssim_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H,
                                                       int W, const float* __restrict__ kern, int fs, float c1, float c2,
                                                       float* __restrict__ ssim_map, float* __restrict__ ssim_mean,
                                                       float* __restrict__ mse_mean) {
    __shared__ float X[SSIM_MAX_PIX], Y[SSIM_MAX_PIX], Kw[SSIM_MAX_FS * SSIM_MAX_FS];
    __shared__ float red[8][2];
    const long long plane = blockIdx.x;
    const int HW = H * W;
    const float* px = pred + plane * HW;
    const float* py = target + plane * HW;
    float se = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float x = px[i], y = py[i];
        X[i] = x;
        Y[i] = y;
        const float d = x - y;
        se += d * d;
    }
    for (int i = threadIdx.x; i < fs * fs; i += blockDim.x) Kw[i] = kern[i];
    __syncthreads();
    const int OH = H - fs + 1, OW = W - fs + 1, n_out = OH * OW;
    float ss = 0.f;
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
        const int oy = o / OW, ox = o - oy * OW;
        const float v = vs_ssim_at(X, Y, Kw, W, fs, oy, ox, c1, c2);
        if (ssim_map != nullptr) ssim_map[plane * n_out + o] = v;
        ss += v;
    }
    se = warp_sum(se);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = se; red[threadIdx.x >> 5][1] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[w][0]; b += red[w][1]; }
        mse_mean[plane] = a / (float)HW;
        ssim_mean[plane] = b / (float)n_out;
    }
}

}  // namespace vs

using namespace vs;

extern "C" int vs_ssim_mse_planes(const float* pred, const float* target, int64_t planes, int32_t H, int32_t W, const float* kernel,
                                  int32_t fs, float c1, float c2, float* ssim_map, float* ssim_mean, float* mse_mean,
                                  void* stream) {
    VS_REQUIRE(H >= 1 && W >= 1 && (int64_t)H * W <= SSIM_MAX_PIX, "ssim_mse_planes: planes of at most %d pixels (got %d x %d)",
               SSIM_MAX_PIX, H, W);
    VS_REQUIRE(fs >= 1 && fs <= SSIM_MAX_FS && fs <= H && fs <= W, "ssim_mse_planes: window %d does not fit (max %d, plane %d x %d)", fs,
               SSIM_MAX_FS, H, W);
    VS_REQUIRE(planes >= 0 && planes < (1LL << 31), "ssim_mse_planes: plane count out of range");
    if (planes == 0) return 0;
    ssim_mse_kernel<<<(unsigned)planes, 256, 0, as_stream(stream)>>>(pred, target, H, W, kernel, fs, c1, c2, ssim_map, ssim_mean, mse_mean);
    return launched("ssim_mse_kernel");
}

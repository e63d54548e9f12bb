// SSIM at one output pixel: the arithmetic shared by the CUDA kernel (metrics.cu) and the host harness that checks it
// on the CPU (tests/test_metrics.py compiles this header with g++).
// Follows /root/reference/var_sep/utils/ssim.py:95-116 (_ssim): five "valid" Gaussian-weighted window sums, then
//   ssim = ((2 mu1 mu2 + c1) (2 s12 + c2)) / ((mu1^2 + mu2^2 + c1) (s1 + s2 + c2)).
#pragma once

#if defined(__CUDACC__)
#define VS_HD __host__ __device__ __forceinline__
#else
#define VS_HD inline
#endif

// X, Y: one H x W plane each (row pitch W); Kw: fs x fs window weights; (oy, ox): output pixel = top-left of the window
VS_HD float vs_ssim_at(const float* X, const float* Y, const float* Kw, int W, int fs, int oy, int ox, float c1, float c2) {
    float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int r = 0; r < fs; ++r) {
        const float* xr = X + (oy + r) * W + ox;
        const float* yr = Y + (oy + r) * W + ox;
        const float* kr = Kw + r * fs;
        for (int s = 0; s < fs; ++s) {
            const float w = kr[s], x = xr[s], y = yr[s];
            const float wx = w * x, wy = w * y;
            mu1 += wx;
            mu2 += wy;
            s11 += wx * x;
            s22 += wy * y;
            s12 += wx * y;
        }
    }
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float v1 = 2.f * (s12 - mu12) + c2;
    const float v2 = (s11 - mu1_sq) + (s22 - mu2_sq) + c2;
    return ((2.f * mu12 + c1) * v1) / ((mu1_sq + mu2_sq + c1) * v2);
}

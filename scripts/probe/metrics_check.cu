// Stand-alone check of vs_ssim_mse_planes on the GPU (no Python: runs in about a second).  The CPU side evaluates the same
// per-pixel function (csrc/metrics_core.h) sequentially; prints the largest differences and exits non-zero on failure.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/probe/metrics_check scripts/probe/metrics_check.cu \
//        -Lspatiotemporal_variable_separation_b200 -lvarsep_sm100a -Xlinker -rpath -Xlinker '$ORIGIN/../../spatiotemporal_variable_separation_b200'
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/varsep.h"
#include "../../spatiotemporal_variable_separation_b200/csrc/metrics_core.h"

static unsigned lcg_state = 12345u;
static float frand() { lcg_state = lcg_state * 1664525u + 1013904223u; return (lcg_state >> 8) * (1.0f / 16777216.0f); }

static int run(int planes, int H, int W, int fs, float sigma, bool with_map) {
    const int HW = H * W, OH = H - fs + 1, OW = W - fs + 1, n_out = OH * OW;
    std::vector<float> x((size_t)planes * HW), y((size_t)planes * HW), k(fs * fs);
    for (auto& v : x) v = frand();
    for (size_t i = 0; i < y.size(); ++i) y[i] = 0.6f * x[i] + 0.4f * frand();
    double ksum = 0;
    for (int r = 0; r < fs; ++r)
        for (int s = 0; s < fs; ++s) {
            const double a = r - (fs - 1) / 2.0, b = s - (fs - 1) / 2.0;
            k[r * fs + s] = (float)std::exp(-(a * a + b * b) / (2.0 * sigma * sigma));
            ksum += k[r * fs + s];
        }
    for (auto& v : k) v = (float)(v / ksum);
    float *dx, *dy, *dk, *dmap = nullptr, *dsm, *dmm;
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dy, y.size() * 4); cudaMalloc(&dk, k.size() * 4);
    cudaMalloc(&dsm, planes * 4); cudaMalloc(&dmm, planes * 4);
    if (with_map) cudaMalloc(&dmap, (size_t)planes * n_out * 4);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, y.data(), y.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dk, k.data(), k.size() * 4, cudaMemcpyHostToDevice);
    const float c1 = 1e-4f, c2 = 9e-4f;
    int rc = vs_ssim_mse_planes(dx, dy, planes, H, W, dk, fs, c1, c2, dmap, dsm, dmm, nullptr);
    if (rc != 0) { printf("call failed: %s\n", vs_last_error()); return 1; }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    std::vector<float> sm(planes), mm(planes), map(with_map ? (size_t)planes * n_out : 0);
    cudaMemcpy(sm.data(), dsm, planes * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(mm.data(), dmm, planes * 4, cudaMemcpyDeviceToHost);
    if (with_map) cudaMemcpy(map.data(), dmap, map.size() * 4, cudaMemcpyDeviceToHost);
    double e_map = 0, e_sm = 0, e_mm = 0;
    for (int p = 0; p < planes; ++p) {
        const float* X = &x[(size_t)p * HW];
        const float* Y = &y[(size_t)p * HW];
        double ss = 0, se = 0;
        for (int o = 0; o < n_out; ++o) {
            const float v = vs_ssim_at(X, Y, k.data(), W, fs, o / OW, o % OW, c1, c2);
            ss += v;
            if (with_map) e_map = std::fmax(e_map, std::fabs(v - map[(size_t)p * n_out + o]));
        }
        for (int i = 0; i < HW; ++i) se += (double)(X[i] - Y[i]) * (X[i] - Y[i]);
        e_sm = std::fmax(e_sm, std::fabs(ss / n_out - sm[p]));
        e_mm = std::fmax(e_mm, std::fabs(se / HW - mm[p]) / (se / HW));
    }
    const bool ok = e_map < 2e-4 && e_sm < 2e-5 && e_mm < 1e-5;
    printf("planes %d  %dx%d  window %d  map %d:  max |map| diff %.2e, |mean ssim| diff %.2e, rel mse diff %.2e  %s\n", planes, H, W, fs,
           (int)with_map, e_map, e_sm, e_mm, ok ? "ok" : "FAIL");
    cudaFree(dx); cudaFree(dy); cudaFree(dk); cudaFree(dsm); cudaFree(dmm); if (dmap) cudaFree(dmap);
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run(37, 64, 64, 11, 1.5f, true);
    bad += run(300, 32, 32, 11, 1.5f, false);
    bad += run(5, 40, 24, 7, 1.0f, true);
    bad += run(3, 15, 15, 15, 2.0f, true);
    printf(bad ? "metrics_check: FAILED\n" : "metrics_check: all ok\n");
    return bad;
}

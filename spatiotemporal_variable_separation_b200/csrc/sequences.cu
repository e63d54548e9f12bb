// Device-side Moving-MNIST sequence generator (deterministic variant).
//
// Replaces the per-sample Python loops of /root/reference/var_sep/data/moving_mnist.py:112-130 (blit of num_digits
// glyphs along their trajectories, clip at 255, scale to [0, 1]) and :131-170 + :172-253 (_compute_trajectory /
// _process_collision).  With integer start positions and speeds and the deterministic setting (speed kept at a bounce,
// moving_mnist.py:229-231 skipped) the collision loop is an elastic reflection on each axis independently: the exact
// position at frame t is the triangle wave of period 2 * x_max through  s + d * t, which the reference then rounds
// (int(round(.)), moving_mnist.py:166).  The host draws (glyph, sx, sy, dx, dy) per object in the reference's order
// (data.py); this kernel does everything else, one thread per output pixel quad, frames written once with 16-byte stores.
#include "common.cuh"

namespace vs {

// position of an object at frame t: reflection of s + d*t into [0, lim]
__device__ __forceinline__ int bounce(int s, int d, int t, int lim) {
    if (lim <= 0) return 0;
    const int period = 2 * lim;
    int m = (s + d * t) % period;
    if (m < 0) m += period;
    return m > lim ? period - m : m;
}

// frames [B][T][1][F][F] fp32; glyphs [G][gh][gw] uint8; objs [B][n_obj][5] int32 = {glyph, sx, sy, dx, dy}
// (sx indexes rows, sy columns, as x[t, 0, sx:sx+h, sy:sy+w] += img in moving_mnist.py:124)
__global__ void __launch_bounds__(256) moving_sequences_kernel(const unsigned char* __restrict__ glyphs, int gh, int gw,
                                                               const int* __restrict__ objs, int n_obj, int B, int T, int F,
                                                               float* __restrict__ frames) {
    const long long quads = (long long)B * T * F * (F / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < quads; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % (F / 4));
        long long r = i / (F / 4);
        const int row = (int)(r % F); r /= F;
        const int t = (int)(r % T);
        const int b = (int)(r / T);
        int acc[4] = {0, 0, 0, 0};
        for (int o = 0; o < n_obj; ++o) {
            const int* ob = objs + ((long long)b * n_obj + o) * 5;
            const int px = bounce(ob[1], ob[3], t, F - gh), py = bounce(ob[2], ob[4], t, F - gw);
            const int gr = row - px;
            if (gr < 0 || gr >= gh) continue;
            const unsigned char* g = glyphs + ((long long)ob[0] * gh + gr) * gw;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gc = cq * 4 + e - py;
                if (gc >= 0 && gc < gw) acc[e] += g[gc];
            }
        }
        float4 v;
        v.x = (float)(acc[0] > 255 ? 255 : acc[0]) / 255.f; v.y = (float)(acc[1] > 255 ? 255 : acc[1]) / 255.f;
        v.z = (float)(acc[2] > 255 ? 255 : acc[2]) / 255.f; v.w = (float)(acc[3] > 255 ? 255 : acc[3]) / 255.f;
        reinterpret_cast<float4*>(frames)[i] = v;
    }
}

}  // namespace vs

using namespace vs;

extern "C" int vs_moving_sequences(const uint8_t* glyphs, int32_t n_glyphs, int32_t gh, int32_t gw, const int32_t* objs,
                                   int32_t n_obj, int32_t B, int32_t T, int32_t F, float* frames, void* stream) {
    VS_REQUIRE(gh >= 1 && gw >= 1 && gh <= F && gw <= F && F % 4 == 0, "moving_sequences: glyph %dx%d does not fit a %d-pixel frame (F %% 4 == 0)", gh, gw, F);
    VS_REQUIRE(n_glyphs >= 1 && n_obj >= 1 && B >= 0 && T >= 0, "moving_sequences: bad sizes");
    VS_REQUIRE((reinterpret_cast<uintptr_t>(frames) & 15) == 0, "moving_sequences: frames must be 16-byte aligned");
    const long long quads = (long long)B * T * F * (F / 4);
    if (quads == 0) return 0;
    long long blocks = cdiv(quads, 256);
    if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
    moving_sequences_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(glyphs, gh, gw, objs, n_obj, B, T, F, frames);
    return launched("moving_sequences_kernel");
}

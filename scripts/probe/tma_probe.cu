// Standalone probe: which no-swizzle TMA box shapes / coordinates are accepted by the hardware.
// usage: tma_probe W H N box0 box1 c0 c1 c2 swizzle(0|128)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, unsigned short* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar_a) : "memory");
    }
    for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<unsigned short*>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    if (argc < 10) return 2;
    int W = atoi(argv[1]), H = atoi(argv[2]), N = atoi(argv[3]), b0 = atoi(argv[4]), b1 = atoi(argv[5]);
    int c0 = atoi(argv[6]), c1 = atoi(argv[7]), c2 = atoi(argv[8]), sw = atoi(argv[9]);
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    std::vector<unsigned short> h((size_t)W * H * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned short)(i % 60000 + 1);
    unsigned short *d, *o;
    cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    int bytes = b0 * b1 * 2;
    cudaMalloc(&o, bytes); cudaMemset(o, 0xff, bytes);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<<<1, 128, 65536>>>(m, c0, c1, c2, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned short> res(bytes / 2);
    cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost);
    long bad = 0;
    if (sw == 0)
        for (int y = 0; y < b1; ++y)
            for (int x = 0; x < b0; ++x) {
                int gx = c0 + x, gy = c1 + y;
                unsigned short want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)c2 * H + gy) * W + gx] : 0;
                if (res[(size_t)y * b0 + x] != want) ++bad;
            }
    printf("ok, mismatches vs dense row-major expectation: %ld of %d\n", bad, b0 * b1);
    return 0;
}

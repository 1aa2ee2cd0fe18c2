// Diagnostic (not part of the product): which cp.async.bulk.tensor configurations complete on this GPU.
// usage: tma_matrix <variant>; prints "VARIANT n: OK" or the CUDA error.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, uint32_t bytes, uint32_t words, uint32_t *out, uint32_t *flag,
                  int proxy_fence)
{
    extern __shared__ __align__(128) uint32_t buf[];
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (proxy_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar)), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_addr(buf)), "l"(&tmap), "r"(c0), "r"(c1), "r"(smem_addr(&mbar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_addr(buf)), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(&mbar)) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spins = 0; !done && spins < (1u << 22); ++spins)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(&mbar)) : "memory");
    if (!done) { if (threadIdx.x == 0) *flag = 1; return; }
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 160, H = 120, K = 6;           // "tiny" scene: 480 bytes = 120 words per row
    const int roww = 3 * W / 4;
    std::vector<uint32_t> img((size_t)roww * H * K);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint32_t)(i * 2654435761u);
    uint32_t *d_img, *d_out, *d_flag;
    cudaMalloc(&d_img, img.size() * 4); cudaMalloc(&d_out, 1 << 20); cudaMalloc(&d_flag, 4);
    cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice); cudaMemset(d_flag, 0, 4); cudaMemset(d_out, 0xFF, 1 << 20);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    EncodeFn encode = (EncodeFn)fn;
    // variant table: rank, box0, box1, c0, c1, c2, proxy fence, data type (0 = u32, 1 = u8 with box0 in bytes)
    struct V { int rank, b0, b1, c0, c1, c2, pf, u8; } tab[] = {
        {2, 64, 16, 0, 0, 0, 0, 0},      // 0: plain 2-D
        {2, 64, 16, -1, -1, 0, 0, 0},    // 1: 2-D, negative start
        {3, 64, 16, 0, 0, 1, 0, 0},      // 2: 3-D
        {3, 100, 34, 0, 0, 1, 0, 0},     // 3: 3-D, the product's box
        {3, 100, 34, -1, -1, 1, 1, 0},   // 4: the product's call
        {3, 100, 34, 95, 31, 1, 1, 0},   // 5: box hanging over the right / bottom edge
        {3, 96, 34, -1, -1, 1, 0, 0},    // 6: 384-byte rows
        {3, 64, 34, -1, -1, 1, 0, 0},    // 7: 256-byte rows
        {3, 256, 16, -16, -1, 1, 0, 1},  // 8: u8 elements, 256-byte rows
        {2, 32, 34, -1, -1, 0, 0, 0},    // 9: 128-byte rows
    };
    const V v = tab[variant];
    alignas(64) CUtensorMap tmap;
    cuuint64_t dims[3] = {(cuuint64_t)(v.u8 ? 3 * W : roww), (cuuint64_t)H, (cuuint64_t)K};
    cuuint64_t strides[2] = {(cuuint64_t)3 * W, (cuuint64_t)3 * W * H};
    cuuint32_t box[3] = {(cuuint32_t)v.b0, (cuuint32_t)v.b1, 1u}, es[3] = {1, 1, 1};
    CUresult r = encode(&tmap, v.u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT32, v.rank, d_img, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("VARIANT %d: encode failed (%d)\n", variant, (int)r); return 0; }
    const uint32_t bytes = (uint32_t)v.b0 * (v.u8 ? 1 : 4) * v.b1, words = bytes / 4;
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (v.rank == 2) k<2><<<1, 128, bytes>>>(tmap, v.c0, v.c1, v.c2, bytes, words, d_out, d_flag, v.pf);
    else k<3><<<1, 128, bytes>>>(tmap, v.c0, v.c1, v.c2, bytes, words, d_out, d_flag, v.pf);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("VARIANT %d: %s\n", variant, cudaGetErrorString(e)); return 0; }
    uint32_t flag = 0; cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost);
    if (flag) { printf("VARIANT %d: copy never arrived (timeout)\n", variant); return 0; }
    std::vector<uint32_t> out(words); cudaMemcpy(out.data(), d_out, words * 4, cudaMemcpyDeviceToHost);
    // reference
    size_t wrong = 0;
    const int b0w = v.u8 ? v.b0 / 4 : v.b0, c0w = v.u8 ? v.c0 / 4 : v.c0;
    for (int rr = 0; rr < v.b1; ++rr)
        for (int x = 0; x < b0w; ++x) {
            const int gy = v.c1 + rr, gx = c0w + x;
            uint32_t want = 0;
            if (gy >= 0 && gy < H && gx >= 0 && gx < roww) want = img[((size_t)(v.rank == 3 ? v.c2 : 0) * H + gy) * roww + gx];
            if (out[(size_t)rr * b0w + x] != want) ++wrong;
        }
    printf("VARIANT %d: OK, %zu of %u words differ from the expected tile\n", variant, wrong, words);
    return 0;
}

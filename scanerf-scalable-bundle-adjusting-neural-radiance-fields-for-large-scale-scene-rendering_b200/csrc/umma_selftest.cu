// Self-test of the tcgen05 conventions in umma.cuh: one CTA computes, from bf16-rounded
// inputs X[128,64], W[64,64], G[128,64],
//   Y  = X W^T      (A = X  K-major,  B = W K-major,  M=128, N=64, K=64)
//   DX = G W        (A = G  K-major,  B = W MN-major, M=128, N=64, K=64)
//   DW = 2 G^T X    (A = G  MN-major, B = X MN-major, M=64,  N=64, K=128; issued twice to
//                    exercise accumulation across separately committed batches)
//   DWo = G^T X[:, 32:48]   (as DW with N=16 and the B operand starting at column 32 of its tile)
//   YS  = X[:, 32:64] W[0:16, 0:32]^T  (M=128, N=16, K=32 with the A operand starting at column 32)
// i.e. the three GEMM shapes of a linear layer's forward, input gradient and weight gradient,
// all reading the SAME shared-memory images.  tests/test_decoder_gpu.py compares with torch.
#include "common.cuh"
#include "umma.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ G,
                     float* __restrict__ Y, float* __restrict__ DX, float* __restrict__ DW,
                     float* __restrict__ DWo, float* __restrict__ YS)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* tX = smem;               // 128 x 64 bf16 = 16 KB
    unsigned char* tG = smem + 16384;       // 16 KB
    unsigned char* tW = smem + 32768;       // 64 x 64 bf16 = 8 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    // stage operands: thread t owns row t of X and G, threads 0..63 own a row of W
    for (int c = 0; c < 8; ++c) {
        umma::tile_store8(tX, tid, c, X + tid * 64 + c * 8);
        umma::tile_store8(tG, tid, c, G + tid * 64 + c * 8);
        if (tid < 64) umma::tile_store8(tW, tid, c, W + tid * 64 + c * 8);
    }
    if (warp == 0) umma::tmem_alloc<256>(&tmem_base_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_async_smem();               // generic-proxy stores -> visible to the tensor core
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_base_slot;

    if (tid == 0) {
        const uint32_t aX = umma::smem_u32(tX), aG = umma::smem_u32(tG), aW = umma::smem_u32(tW);
        const uint32_t id_y = umma::idesc_bf16(128, 64, 0, 0);
        const uint32_t id_dx = umma::idesc_bf16(128, 64, 0, 1);
        const uint32_t id_dw = umma::idesc_bf16(64, 64, 1, 1);
        for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + 0, umma::desc_kmajor(aX, k), umma::desc_kmajor(aW, k), id_y, k > 0);
        for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + 64, umma::desc_kmajor(aG, k), umma::desc_mnmajor(aW, k), id_dx, k > 0);
        for (int k = 0; k < 8; ++k) umma::mma_bf16(tmem + 128, umma::desc_mnmajor(aG, k), umma::desc_mnmajor(aX, k), id_dw, k > 0);
        const uint32_t id_dwo = umma::idesc_bf16(64, 16, 1, 1);
        for (int k = 0; k < 8; ++k) umma::mma_bf16(tmem + 192, umma::desc_mnmajor(aG, k), umma::desc_mnmajor(aX + 64, k), id_dwo, k > 0);
        const uint32_t id_ys = umma::idesc_bf16(128, 16, 0, 0);
        for (int k = 0; k < 2; ++k) umma::mma_bf16(tmem + 208, umma::desc_kmajor(aX, 2 + k), umma::desc_kmajor(aW, k), id_ys, k > 0);
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    if (tid == 0) {
        const uint32_t aX = umma::smem_u32(tX), aG = umma::smem_u32(tG);
        const uint32_t id_dw = umma::idesc_bf16(64, 64, 1, 1);
        for (int k = 0; k < 8; ++k) umma::mma_bf16(tmem + 128, umma::desc_mnmajor(aG, k), umma::desc_mnmajor(aX, k), id_dw, 1);
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 1);
    umma::tc_fence_after();

    float v[32];
    const int lane_base = 32 * (warp & 3);
    for (int h = 0; h < 2; ++h) {           // Y and DX: thread t holds row t
        umma::tmem_ld32(umma::tmem_addr(tmem, lane_base, 0 + 32 * h), v);
        umma::tc_wait_ld();
        for (int j = 0; j < 32; ++j) Y[tid * 64 + 32 * h + j] = v[j];
        umma::tmem_ld32(umma::tmem_addr(tmem, lane_base, 64 + 32 * h), v);
        umma::tc_wait_ld();
        for (int j = 0; j < 32; ++j) DX[tid * 64 + 32 * h + j] = v[j];
    }
    for (int h = 0; h < 2; ++h) {           // DW (M=64): row m lives in lane 32*(m/16) + m%16
        umma::tmem_ld32(umma::tmem_addr(tmem, lane_base, 128 + 32 * h), v);
        umma::tc_wait_ld();
        const int l = tid & 31;
        if (l < 16) {
            const int m = 16 * (warp & 3) + l;
            for (int j = 0; j < 32; ++j) DW[m * 64 + 32 * h + j] = v[j];
        }
    }
    {
        float w[16];
        umma::tmem_ld16(umma::tmem_addr(tmem, lane_base, 192), w);
        umma::tc_wait_ld();
        const int l = tid & 31;
        if (l < 16) {
            const int m = 16 * (warp & 3) + l;
            for (int j = 0; j < 16; ++j) DWo[m * 16 + j] = w[j];
        }
        umma::tmem_ld16(umma::tmem_addr(tmem, lane_base, 208), w);
        umma::tc_wait_ld();
        for (int j = 0; j < 16; ++j) YS[tid * 16 + j] = w[j];
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free<256>(tmem);
}

}  // namespace

// X[128,64], W[64,64], G[128,64] f32 (rounded to bf16 inside) -> Y[128,64], DX[128,64], DW[64,64], DWo[64,16], YS[128,16] f32
SNRF_API int snrf_umma_selftest(const float* X, const float* W, const float* G, float* Y, float* DX, float* DW,
                                float* DWo, float* YS, void* stream)
{
    const int smem = 16384 * 2 + 8192 + 1024;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { snrf_set_error("snrf_umma_selftest: %s", cudaGetErrorString(e)); return (int)e; }
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(X, W, G, Y, DX, DW, DWo, YS);
    SNRF_RETURN_LAUNCH("snrf_umma_selftest");
}

// Density / colour decoder MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators
// in TMEM), forward and backward, sm_100a.
//
// Replaces (behaviour, not code) network.ShallowMLP.forward (network.py:151-190) and its
// autograd backward, which the reference runs as 8 cuBLAS SGEMMs + ~25 elementwise torch
// kernels per direction:
//   x   = feat(32) * level_mask(32)
//   h1  = g(W1 x + b1)            32 -> 64      g(v) = exp(-v^2 / 0.02)
//   H   = W2 h1 + b2              64 -> 64      (no activation)
//   sigma   = softplus(ws  . H[0:32] + bs)
//   diffuse = sigmoid (Wd    H[0:32] + bd)      3
//   tint    = sigmoid (Wt    H[0:32] + bt)      3
//   x2  = [H[32:64], SH16(d / (|d| + 1e-8))]    48
//   specular = sigmoid(W5 g(W4 g(W3 x2 + b3) + b4) + b5)    48 -> 64 -> 64 -> 3
//
// Design (B200).  A CTA owns tiles of 128 consecutive samples (M = 128, one TMEM lane and one
// thread per sample).  Every operand -- activations and weights -- is a bf16 tile with 128-byte
// rows, 128B-swizzled (umma.cuh); a layer is 1-4 tcgen05.mma (K = 16 each) issued by one
// thread, completion is signalled through an mbarrier (tcgen05.commit), the epilogue pulls the
// fp32 accumulator row out of TMEM (tcgen05.ld), applies bias + activation in registers and
// writes the next layer's operand tile.  Weights are converted to bf16 tiles once per CTA
// (persistent grid).  Nothing but the 40-byte head row per sample is written to HBM in the
// forward; the backward recomputes the forward per tile, keeps every intermediate in shared
// memory, accumulates all weight / bias gradients in TMEM across the CTA's tiles (M = 64
// accumulators) and flushes them once at the end.
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int kRows = 128;                 // samples per tile
constexpr int kTile = kRows * 128;         // bytes of one operand tile (128 rows x 64 bf16)
constexpr float kGaussLog2 = -50.0f * 1.4426950408889634f;   // exp(-v^2/0.02) = exp2(v^2 * kGaussLog2)

// parameter tensors in network.ShallowMLP state_dict order (weight, bias per Linear)
struct DecoderParams {
    const float *W1, *b1, *W2, *b2, *Ws, *bs, *Wd, *bd, *Wt, *bt, *W3, *b3, *W4, *b4, *W5, *b5;
};

// ---- shared-memory plan of the forward kernel (offsets from the 1024-aligned base)
//   weights : W1 [64 x 64 (32 used)]  W2 [64 x 64]  Wh [16 x 64 (7 rows, 32 cols used)]
//             W3 [64 x 64 (48 used: 32 H + 16 SH)]  W4 [64 x 64]  W5 [16 x 64 (3 rows used)]
constexpr int oW1 = 0, oW2 = 8192, oW3 = 16384, oW4 = 24576, oWh = 32768, oW5 = 34816, oWend = 36864;
//   biases (fp32): b1[64] b2[64] bh[16] b3[64] b4[64] b5[16]
constexpr int oB = oWend, nB = 64 + 64 + 16 + 64 + 64 + 16;
constexpr int oB1 = 0, oB2 = 64, oBh = 128, oB3 = 144, oB4 = 208, oB5 = 272;
constexpr int oMask = oB + nB * 4;                  // 32 floats
constexpr int oTiles = ((oMask + 128 + 1023) / 1024) * 1024;
constexpr int kFwdSmem = oTiles + 3 * kTile + 1024; // T0 (x | SH), T1, T2  (+ alignment slack)

__device__ __forceinline__ float gauss_act(float v) { return exp2f(v * v * kGaussLog2); }
__device__ __forceinline__ float sigmoidf(float v) { return 1.0f / (1.0f + __expf(-v)); }
// torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf(float v) { return v > 20.0f ? v : log1pf(__expf(v)); }

// Degree-3 real spherical harmonics of a unit vector in the order of network.py:38-77.
__device__ __forceinline__ void sh16(float x, float y, float z, float* o)
{
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    o[0] = 0.28209479177387814f;
    o[1] = 0.4886025119029199f * y; o[2] = 0.4886025119029199f * z; o[3] = 0.4886025119029199f * x;
    o[4] = 1.0925484305920792f * xy; o[5] = -1.0925484305920792f * yz;
    o[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    o[7] = -1.0925484305920792f * xz; o[8] = 0.5462742152960396f * (xx - yy);
    o[9] = -0.5900435899266435f * y * (3.0f * xx - yy); o[10] = 2.890611442640554f * xy * z;
    o[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    o[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    o[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    o[14] = 1.445305721320277f * z * (xx - yy); o[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

// Convert one fp32 weight matrix W[out, in] (row-major, nn.Linear layout) into a swizzled bf16
// tile of `rows` rows; input column j lands in tile column col0 + j; everything else is zero.
__device__ void stage_weight(unsigned char* tile, int rows, const float* __restrict__ W, int out, int in, int col0,
                             int tid, int nthreads)
{
    for (int t = tid; t < rows * 8; t += nthreads) {
        const int r = t >> 3, c = t & 7;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = c * 8 + j - col0;
            v[j] = (r < out && col >= 0 && col < in) ? W[r * in + col] : 0.0f;
        }
        umma::tile_store8(tile, r, c, v);
    }
}
// rows [row0, row0+out) of a tile from W[out, in] (used to stack the three 32-input heads)
__device__ void stage_weight_rows(unsigned char* tile, int row0, const float* __restrict__ W, int out, int in,
                                  int tid, int nthreads)
{
    for (int t = tid; t < out * 8; t += nthreads) {
        const int r = t >> 3, c = t & 7;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int col = c * 8 + j; v[j] = col < in ? W[r * in + col] : 0.0f; }
        umma::tile_store8(tile, row0 + r, c, v);
    }
}

__device__ void stage_all_weights(unsigned char* smem, const DecoderParams& p, const float* __restrict__ mask32, int tid, int nthreads)
{
    stage_weight(smem + oW1, 64, p.W1, 64, 32, 0, tid, nthreads);
    stage_weight(smem + oW2, 64, p.W2, 64, 64, 0, tid, nthreads);
    stage_weight(smem + oW3, 64, p.W3, 64, 48, 0, tid, nthreads);
    stage_weight(smem + oW4, 64, p.W4, 64, 64, 0, tid, nthreads);
    stage_weight(smem + oWh, 16, p.Ws, 0, 32, 0, tid, nthreads);          // zero the 16 x 64 tile
    stage_weight(smem + oW5, 16, p.W5, 3, 64, 0, tid, nthreads);
    __syncthreads();
    stage_weight_rows(smem + oWh, 0, p.Ws, 1, 32, tid, nthreads);
    stage_weight_rows(smem + oWh, 1, p.Wd, 3, 32, tid, nthreads);
    stage_weight_rows(smem + oWh, 4, p.Wt, 3, 32, tid, nthreads);
    float* b = reinterpret_cast<float*>(smem + oB);
    for (int i = tid; i < nB; i += nthreads) {
        float v = 0.0f;
        if (i < 64) v = p.b1[i];
        else if (i < 128) v = p.b2[i - 64];
        else if (i < 144) { const int j = i - 128; v = j == 0 ? p.bs[0] : (j < 4 ? p.bd[j - 1] : (j < 7 ? p.bt[j - 4] : 0.0f)); }
        else if (i < 208) v = p.b3[i - 144];
        else if (i < 272) v = p.b4[i - 208];
        else { const int j = i - 272; v = j < 3 ? p.b5[j] : 0.0f; }
        b[i] = v;
    }
    float* m = reinterpret_cast<float*>(smem + oMask);
    for (int i = tid; i < 32; i += nthreads) m[i] = mask32 ? mask32[i] : 1.0f;
}

// One MMA "stage": issued by a single thread; `n_ops` (A tile, k-step, B tile, k-step) pairs into one accumulator.
struct MmaOp { uint32_t a_addr; int a_k; uint32_t b_addr; int b_k; };

// ------------------------------- forward ------------------------------------
// feats [N,32] f32, rays_d [R,3] (sample n belongs to ray n / S), out [N,10] f32 =
// (sigma, tint3, diffuse3, specular3).
__global__ void __launch_bounds__(kRows, 1)
decoder_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ mask32, const float* __restrict__ rays_d,
                   DecoderParams p, float* __restrict__ out, int N, int S, int num_tiles)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* T0 = smem + oTiles;
    unsigned char* T1 = T0 + kTile;
    unsigned char* T2 = T1 + kTile;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    stage_all_weights(smem, p, mask32, tid, kRows);
    if (warp == 0) umma::tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t Da = tmem, Db = tmem + 64, Dh = tmem + 128;          // accumulators: 64, 64, 16 columns
    const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    const float* bias = reinterpret_cast<const float*>(smem + oB);
    const float* mask = reinterpret_cast<const float*>(smem + oMask);
    const uint32_t aT0 = umma::smem_u32(T0), aT1 = umma::smem_u32(T1), aT2 = umma::smem_u32(T2);
    const uint32_t aW1 = umma::smem_u32(smem + oW1), aW2 = umma::smem_u32(smem + oW2), aW3 = umma::smem_u32(smem + oW3),
                   aW4 = umma::smem_u32(smem + oW4), aWh = umma::smem_u32(smem + oWh), aW5 = umma::smem_u32(smem + oW5);
    constexpr uint32_t id64 = umma::idesc_bf16(128, 64, 0, 0), id16 = umma::idesc_bf16(128, 16, 0, 0);
    uint32_t phase = 0;

    auto sync_operands = [&]() {       // my smem stores + TMEM reads are done -> MMA may run
        umma::fence_async_smem();
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
    };
    auto wait_mma = [&]() { umma::mbar_wait(&bar, phase); phase ^= 1u; umma::tc_fence_after(); };

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n = tile * kRows + tid;
        const bool live = n < N;
        // ---- stage the input row: x = feat * mask (cols 0..31), SH16(view dir) (cols 32..47), zeros (48..63)
        {
            float v[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (live) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(feats + (size_t)n * 32 + c * 8));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(feats + (size_t)n * 32 + c * 8 + 4));
                    v[0] = a.x * mask[c * 8 + 0]; v[1] = a.y * mask[c * 8 + 1]; v[2] = a.z * mask[c * 8 + 2]; v[3] = a.w * mask[c * 8 + 3];
                    v[4] = b.x * mask[c * 8 + 4]; v[5] = b.y * mask[c * 8 + 5]; v[6] = b.z * mask[c * 8 + 6]; v[7] = b.w * mask[c * 8 + 7];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = 0.0f;
                }
                umma::tile_store8(T0, tid, c, v);
            }
            float sh[16];
            if (live) {
                const f3 d = ld3(rays_d + 3 * (size_t)(n / S));
                const float inv = 1.0f / (sqrtf(d.x * d.x + d.y * d.y + d.z * d.z) + 1e-8f);
                sh16(d.x * inv, d.y * inv, d.z * inv, sh);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) sh[j] = 0.0f;
            }
            umma::tile_store8(T0, tid, 4, sh);
            umma::tile_store8(T0, tid, 5, sh + 8);
            umma::tile_zero8(T0, tid, 6);
            umma::tile_zero8(T0, tid, 7);
        }
        sync_operands();
        // ---- L1: Da = x W1^T   (K = 32)
        if (tid == 0) {
            for (int k = 0; k < 2; ++k) umma::mma_bf16(Da, umma::desc_kmajor(aT0, k), umma::desc_kmajor(aW1, k), id64, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
        float v[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            umma::tmem_ld32(Da + lane_addr + 32 * h, v);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gauss_act(v[j] + bias[oB1 + 32 * h + j]);
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::tile_store8(T1, tid, 4 * h + c, v + 8 * c);
        }
        sync_operands();
        // ---- L2: Db = h1 W2^T  (K = 64)
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(Db, umma::desc_kmajor(aT1, k), umma::desc_kmajor(aW2, k), id64, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            umma::tmem_ld32(Db + lane_addr + 32 * h, v);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bias[oB2 + 32 * h + j];
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::tile_store8(T2, tid, 4 * h + c, v + 8 * c);
        }
        sync_operands();
        // ---- heads: Dh = H[0:32] Wh^T (K = 32, N = 16);  L3: Da = [H[32:64], SH] W3^T (K = 48)
        if (tid == 0) {
            for (int k = 0; k < 2; ++k) umma::mma_bf16(Dh, umma::desc_kmajor(aT2, k), umma::desc_kmajor(aWh, k), id16, k > 0);
            umma::mma_bf16(Da, umma::desc_kmajor(aT2, 2), umma::desc_kmajor(aW3, 0), id64, 0);
            umma::mma_bf16(Da, umma::desc_kmajor(aT2, 3), umma::desc_kmajor(aW3, 1), id64, 1);
            umma::mma_bf16(Da, umma::desc_kmajor(aT0, 2), umma::desc_kmajor(aW3, 2), id64, 1);
            umma::mma_commit(&bar);
        }
        wait_mma();
        float head[10];
        {
            float z[16];
            umma::tmem_ld16(Dh + lane_addr, z);
            umma::tc_wait_ld();
            head[0] = softplusf(z[0] + bias[oBh + 0]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                head[4 + j] = sigmoidf(z[1 + j] + bias[oBh + 1 + j]);     // diffuse
                head[1 + j] = sigmoidf(z[4 + j] + bias[oBh + 4 + j]);     // tint
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            umma::tmem_ld32(Da + lane_addr + 32 * h, v);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gauss_act(v[j] + bias[oB3 + 32 * h + j]);
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::tile_store8(T1, tid, 4 * h + c, v + 8 * c);
        }
        sync_operands();
        // ---- L4: Db = a3 W4^T (K = 64)
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(Db, umma::desc_kmajor(aT1, k), umma::desc_kmajor(aW4, k), id64, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            umma::tmem_ld32(Db + lane_addr + 32 * h, v);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gauss_act(v[j] + bias[oB4 + 32 * h + j]);
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::tile_store8(T2, tid, 4 * h + c, v + 8 * c);
        }
        sync_operands();
        // ---- L5: Dh = a4 W5^T (K = 64, N = 16)
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(Dh, umma::desc_kmajor(aT2, k), umma::desc_kmajor(aW5, k), id16, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
        {
            float z[16];
            umma::tmem_ld16(Dh + lane_addr, z);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 3; ++j) head[7 + j] = sigmoidf(z[j] + bias[oB5 + j]);
        }
        if (live) {
            float2* o = reinterpret_cast<float2*>(out + (size_t)n * 10);
#pragma unroll
            for (int j = 0; j < 5; ++j) o[j] = make_float2(head[2 * j], head[2 * j + 1]);
        }
        // the next tile's input staging overwrites T0, last read by the L3 MMA that has completed;
        // its first accumulator write (Da) is ordered after this tile's TMEM loads by sync_operands()
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free<256>(tmem);
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
// params: 16 device pointers in network.ShallowMLP state_dict order
//   Spatial_MLP.mlp.0.{weight[64,32],bias}, Spatial_MLP.mlp.2.{weight[64,64],bias},
//   sigma_layer.mlp.0.{[1,32]}, diffuse_layer.mlp.0.{[3,32]}, tint_layer.mlp.0.{[3,32]},
//   Directional_MLP.mlp.0.{[64,48]}, .2.{[64,64]}, .4.{[3,64]}
SNRF_API int snrf_decoder_fwd(const float* feats, const float* mask32, const float* rays_d, const float* const* params,
                              float* heads_out, int N, int S, void* stream)
{
    SNRF_CHECK_ARG(N >= 0 && S > 0, "snrf_decoder_fwd: need N >= 0, S > 0 (N=%d S=%d)", N, S);
    SNRF_CHECK_ARG(params != nullptr, "snrf_decoder_fwd: params is required");
    if (N == 0) return 0;
    DecoderParams p{params[0], params[1], params[2], params[3], params[4], params[5], params[6], params[7],
                    params[8], params[9], params[10], params[11], params[12], params[13], params[14], params[15]};
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(decoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
        if (e != cudaSuccess) { snrf_set_error("snrf_decoder_fwd: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int num_tiles = snrf_div_up(N, kRows);
    const int ctas_per_sm = 2;     // 2 x (256 TMEM columns, ~86 KB smem)
    int grid = snrf_sm_count() * ctas_per_sm;
    if (grid > num_tiles) grid = num_tiles;
    decoder_fwd_kernel<<<grid, kRows, kFwdSmem, (cudaStream_t)stream>>>(feats, mask32, rays_d, p, heads_out, N, S, num_tiles);
    SNRF_RETURN_LAUNCH("snrf_decoder_fwd");
}

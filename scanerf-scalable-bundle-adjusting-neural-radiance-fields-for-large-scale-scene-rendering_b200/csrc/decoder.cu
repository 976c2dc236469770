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


// ------------------------------- backward -----------------------------------
// Per tile: recompute the forward keeping every intermediate in shared memory, then walk the
// layers backwards.  Each backward stage is one commit group holding the input-gradient GEMM
// (dA_{l-1} = dz_l W_l, weights read MN-major from the very tile the forward used K-major) and
// the weight-gradient GEMMs (dW_l += dz_l^T A_{l-1}, both operands read MN-major, M = 64
// accumulators that stay in TMEM for the whole kernel).  Bias gradients of the 64-wide layers
// are one more GEMM against a column of ones kept in the spare columns of the input tile.
//
// shared-memory tiles (16 KB each): A0 = [x | SH | 1 | 0], a1, g1 -> dz1, H, dz2, a3, g3 -> dz4,
// a4, g4 -> dz5, dzs = [dz_heads(16) | dz_spec(16)]   with g = dg/dz of the Gaussian activation.
constexpr int kBwdTiles = 10;
constexpr int oSmallGrad = oMask + 128;                       // 16 floats: heads(7) + spec(3) bias gradients
constexpr int oTilesB = ((oSmallGrad + 64 + 1023) / 1024) * 1024;
constexpr int kBwdSmem = oTilesB + kBwdTiles * kTile + 1024;
// TMEM columns
constexpr int cDa = 0, cDb = 64, cDh = 128;
constexpr int cGW1 = 144, cGW2 = 176, cGW3a = 240, cGW3b = 272, cGW4 = 288, cGWhT = 352, cGW5T = 368,
              cGb1 = 384, cGb2 = 392, cGb3 = 400, cGb4 = 408;   // < 416

struct DecoderGrads {
    float *W1, *b1, *W2, *b2, *Ws, *bs, *Wd, *bd, *Wt, *bt, *W3, *b3, *W4, *b4, *W5, *b5;
};

__global__ void __launch_bounds__(kRows, 1)
decoder_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ mask32, const float* __restrict__ rays_d,
                   DecoderParams p, const float* __restrict__ grad_heads, float* __restrict__ grad_feats,
                   float* __restrict__ grad_rays_d, DecoderGrads gp, int N, int S, int num_tiles)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* TA0 = smem + oTilesB;
    unsigned char* Ta1 = TA0 + kTile;
    unsigned char* Tg1 = Ta1 + kTile;
    unsigned char* TH = Tg1 + kTile;
    unsigned char* Tdz2 = TH + kTile;
    unsigned char* Ta3 = Tdz2 + kTile;
    unsigned char* Tg3 = Ta3 + kTile;
    unsigned char* Ta4 = Tg3 + kTile;
    unsigned char* Tg4 = Ta4 + kTile;
    unsigned char* Tdzs = Tg4 + kTile;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_all_weights(smem, p, mask32, tid, kRows);
    float* small_grad = reinterpret_cast<float*>(smem + oSmallGrad);
    if (tid < 16) small_grad[tid] = 0.0f;
    if (warp == 0) umma::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    const float* bias = reinterpret_cast<const float*>(smem + oB);
    const float* mask = reinterpret_cast<const float*>(smem + oMask);
    const uint32_t aA0 = umma::smem_u32(TA0), aa1 = umma::smem_u32(Ta1), ag1 = umma::smem_u32(Tg1), aH = umma::smem_u32(TH),
                   adz2 = umma::smem_u32(Tdz2), aa3 = umma::smem_u32(Ta3), ag3 = umma::smem_u32(Tg3),
                   aa4 = umma::smem_u32(Ta4), ag4 = umma::smem_u32(Tg4), adzs = umma::smem_u32(Tdzs);
    const uint32_t aW1 = umma::smem_u32(smem + oW1), aW2 = umma::smem_u32(smem + oW2), aW3 = umma::smem_u32(smem + oW3),
                   aW4 = umma::smem_u32(smem + oW4), aWh = umma::smem_u32(smem + oWh), aW5 = umma::smem_u32(smem + oW5);
    constexpr uint32_t id64 = umma::idesc_bf16(128, 64, 0, 0), id16 = umma::idesc_bf16(128, 16, 0, 0);
    // input-gradient GEMMs: A K-major (dz rows), B MN-major (weight tile: rows = K = out, cols = N = in)
    constexpr uint32_t idg64 = umma::idesc_bf16(128, 64, 0, 1), idg32 = umma::idesc_bf16(128, 32, 0, 1);
    // weight-gradient GEMMs: both MN-major, M = 64
    constexpr uint32_t idw64 = umma::idesc_bf16(64, 64, 1, 1), idw32 = umma::idesc_bf16(64, 32, 1, 1),
                       idw16 = umma::idesc_bf16(64, 16, 1, 1), idw8 = umma::idesc_bf16(64, 8, 1, 1);
    uint32_t phase = 0;
    auto sync_operands = [&]() { umma::fence_async_smem(); umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after(); };
    auto wait_mma = [&]() { umma::mbar_wait(&bar, phase); phase ^= 1u; umma::tc_fence_after(); };
    // dW += A^T B over the 128 rows of the tile (8 k-steps); `first` = first tile of this CTA
    auto wgrad = [&](uint32_t col, uint32_t a_tile, uint32_t b_tile_plus_off, uint32_t idesc, bool first) {
        for (int k = 0; k < 8; ++k)
            umma::mma_bf16(tmem + col, umma::desc_mnmajor(a_tile, k), umma::desc_mnmajor(b_tile_plus_off, k), idesc, (!first) || k > 0);
    };

    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, first = false) {
        const int n = tile * kRows + tid;
        const bool live = n < N;
        f3 d = mk3(0.f, 0.f, 1.f);
        float dn = 1.0f;
        // ================= forward recompute =================
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
                umma::tile_store8(TA0, tid, c, v);
            }
            float sh[16];
            if (live) {
                d = ld3(rays_d + 3 * (size_t)(n / S));
                dn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
                const float inv = 1.0f / (dn + 1e-8f);
                sh16(d.x * inv, d.y * inv, d.z * inv, sh);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) sh[j] = 0.0f;
            }
            umma::tile_store8(TA0, tid, 4, sh);
            umma::tile_store8(TA0, tid, 5, sh + 8);
            const float ones[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // column 48 = 1: bias-gradient GEMMs
            umma::tile_store8(TA0, tid, 6, ones);
            umma::tile_zero8(TA0, tid, 7);
        }
        sync_operands();
        if (tid == 0) {
            for (int k = 0; k < 2; ++k) umma::mma_bf16(tmem + cDa, umma::desc_kmajor(aA0, k), umma::desc_kmajor(aW1, k), id64, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
        float v[32], g[32];
        // activation a = exp(-50 z^2) and its derivative g = -100 z a, both as operand tiles
        auto gauss_epilogue = [&](uint32_t col, int boff, unsigned char* Ta, unsigned char* Tg) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                umma::tmem_ld32(tmem + col + lane_addr + 32 * h, v);
                umma::tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float z = v[j] + bias[boff + 32 * h + j];
                    const float a = gauss_act(z);
                    v[j] = a;
                    g[j] = -100.0f * z * a;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    umma::tile_store8(Ta, tid, 4 * h + c, v + 8 * c);
                    umma::tile_store8(Tg, tid, 4 * h + c, g + 8 * c);
                }
            }
        };
        gauss_epilogue(cDa, oB1, Ta1, Tg1);
        sync_operands();
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + cDb, umma::desc_kmajor(aa1, k), umma::desc_kmajor(aW2, k), id64, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            umma::tmem_ld32(tmem + cDb + lane_addr + 32 * h, v);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bias[oB2 + 32 * h + j];
#pragma unroll
            for (int c = 0; c < 4; ++c) umma::tile_store8(TH, tid, 4 * h + c, v + 8 * c);
        }
        sync_operands();
        if (tid == 0) {
            for (int k = 0; k < 2; ++k) umma::mma_bf16(tmem + cDh, umma::desc_kmajor(aH, k), umma::desc_kmajor(aWh, k), id16, k > 0);
            umma::mma_bf16(tmem + cDa, umma::desc_kmajor(aH, 2), umma::desc_kmajor(aW3, 0), id64, 0);
            umma::mma_bf16(tmem + cDa, umma::desc_kmajor(aH, 3), umma::desc_kmajor(aW3, 1), id64, 1);
            umma::mma_bf16(tmem + cDa, umma::desc_kmajor(aA0, 2), umma::desc_kmajor(aW3, 2), id64, 1);
            umma::mma_commit(&bar);
        }
        wait_mma();
        float dzh[16];                       // d(loss)/d(pre-activation) of the 7 heads, then of the 3 specular outputs
        float gh[10];
        {
            if (live) {
                const float2* gsrc = reinterpret_cast<const float2*>(grad_heads + (size_t)n * 10);
#pragma unroll
                for (int j = 0; j < 5; ++j) { const float2 t = __ldg(gsrc + j); gh[2 * j] = t.x; gh[2 * j + 1] = t.y; }
            } else {
#pragma unroll
                for (int j = 0; j < 10; ++j) gh[j] = 0.0f;
            }
            float z[16];
            umma::tmem_ld16(tmem + cDh + lane_addr, z);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) dzh[j] = 0.0f;
            // softplus'(z) = sigmoid(z) (0 contribution beyond torch's threshold handled by sigmoid -> 1)
            const float z0 = z[0] + bias[oBh];
            dzh[0] = gh[0] * (z0 > 20.0f ? 1.0f : sigmoidf(z0));
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float sd = sigmoidf(z[1 + j] + bias[oBh + 1 + j]);        // diffuse
                const float st = sigmoidf(z[4 + j] + bias[oBh + 4 + j]);        // tint
                dzh[1 + j] = gh[4 + j] * sd * (1.0f - sd);
                dzh[4 + j] = gh[1 + j] * st * (1.0f - st);
            }
        }
        gauss_epilogue(cDa, oB3, Ta3, Tg3);
        sync_operands();
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + cDb, umma::desc_kmajor(aa3, k), umma::desc_kmajor(aW4, k), id64, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
        gauss_epilogue(cDb, oB4, Ta4, Tg4);
        sync_operands();
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + cDh, umma::desc_kmajor(aa4, k), umma::desc_kmajor(aW5, k), id16, k > 0);
            umma::mma_commit(&bar);
        }
        wait_mma();
        // ================= backward =================
        float dzs[16];
        {
            float z[16];
            umma::tmem_ld16(tmem + cDh + lane_addr, z);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) dzs[j] = 0.0f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float s = sigmoidf(z[j] + bias[oB5 + j]);
                dzs[j] = gh[7 + j] * s * (1.0f - s);
            }
            umma::tile_store8(Tdzs, tid, 0, dzh);
            umma::tile_store8(Tdzs, tid, 1, dzh + 8);
            umma::tile_store8(Tdzs, tid, 2, dzs);
            umma::tile_store8(Tdzs, tid, 3, dzs + 8);
#pragma unroll
            for (int c = 4; c < 8; ++c) umma::tile_zero8(Tdzs, tid, c);
            // bias gradients of the narrow layers: warp reduction, one shared atomic per warp
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                float t = j < 7 ? dzh[j] : dzs[j - 7];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                if (lane == 0) atomicAdd(small_grad + j, t);
            }
        }
        sync_operands();
        // ---- stage B1: dA4 = dz_spec W5 ; dH[0:32] = dz_heads Wh ; dW5^T += a4^T dz_spec ; dWh^T += H^T dz_heads
        if (tid == 0) {
            umma::mma_bf16(tmem + cDa, umma::desc_kmajor(adzs, 1), umma::desc_mnmajor(aW5, 0), idg64, 0);
            umma::mma_bf16(tmem + cDb, umma::desc_kmajor(adzs, 0), umma::desc_mnmajor(aWh, 0), idg32, 0);
            wgrad(cGW5T, aa4, adzs + 32, idw16, first);
            wgrad(cGWhT, aH, adzs, idw16, first);
            umma::mma_commit(&bar);
        }
        wait_mma();
        // dz5 = dA4 * g4 (in place over g4);  dH[0:32] -> dz2 tile columns 0..31
        auto mul_inplace = [&](uint32_t col, unsigned char* Tg) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                umma::tmem_ld32(tmem + col + lane_addr + 32 * h, v);
                umma::tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 q = *reinterpret_cast<const uint4*>(Tg + umma::tile_chunk_off(tid, 4 * h + c));
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
                        o[2 * e] = v[8 * c + 2 * e] * __low2float(b2);
                        o[2 * e + 1] = v[8 * c + 2 * e + 1] * __high2float(b2);
                    }
                    umma::tile_store8(Tg, tid, 4 * h + c, o);
                }
            }
        };
        mul_inplace(cDa, Tg4);
        umma::tmem_ld32(tmem + cDb + lane_addr, v);
        umma::tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 4; ++c) umma::tile_store8(Tdz2, tid, c, v + 8 * c);
        sync_operands();
        // ---- stage B2: dA3 = dz5 W4 ; dW4 += dz5^T a3 ; db4
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + cDb, umma::desc_kmajor(ag4, k), umma::desc_mnmajor(aW4, k), idg64, k > 0);
            wgrad(cGW4, ag4, aa3, idw64, first);
            wgrad(cGb4, ag4, aA0 + 96, idw8, first);
            umma::mma_commit(&bar);
        }
        wait_mma();
        mul_inplace(cDb, Tg3);
        sync_operands();
        // ---- stage B3: d[x2] = dz4 W3 ; dW3 += dz4^T [H[32:64] | SH] ; db3
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + cDa, umma::desc_kmajor(ag3, k), umma::desc_mnmajor(aW3, k), idg64, k > 0);
            wgrad(cGW3a, ag3, aH + 64, idw32, first);
            wgrad(cGW3b, ag3, aA0 + 64, idw16, first);
            wgrad(cGb3, ag3, aA0 + 96, idw8, first);
            umma::mma_commit(&bar);
        }
        wait_mma();
        umma::tmem_ld32(tmem + cDa + lane_addr, v);          // dH[32:64]
        umma::tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 4; ++c) umma::tile_store8(Tdz2, tid, 4 + c, v + 8 * c);
        if (grad_rays_d != nullptr) {                        // d/d(ray direction) through the SH encoding
            float dsh[16];
            umma::tmem_ld16(tmem + cDa + lane_addr + 32, dsh);
            umma::tc_wait_ld();
            float gx = 0.f, gy = 0.f, gz = 0.f;
            if (live) {
                const float inv = 1.0f / (dn + 1e-8f);
                const float x = d.x * inv, y = d.y * inv, z = d.z * inv;
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                const float C1 = 0.4886025119029199f, C2a = 1.0925484305920792f, C2c = 0.31539156525252005f,
                            C2e = 0.5462742152960396f, C3a = 0.5900435899266435f, C3b = 2.890611442640554f,
                            C3c = 0.4570457994644658f, C3d = 0.3731763325901154f, C3e = 1.445305721320277f;
                float vx = 0.f, vy = 0.f, vz = 0.f;
                vy += dsh[1] * C1; vz += dsh[2] * C1; vx += dsh[3] * C1;
                vx += dsh[4] * C2a * y; vy += dsh[4] * C2a * x;
                vy += dsh[5] * -C2a * z; vz += dsh[5] * -C2a * y;
                vx += dsh[6] * C2c * -2.f * x; vy += dsh[6] * C2c * -2.f * y; vz += dsh[6] * C2c * 4.f * z;
                vx += dsh[7] * -C2a * z; vz += dsh[7] * -C2a * x;
                vx += dsh[8] * C2e * 2.f * x; vy += dsh[8] * C2e * -2.f * y;
                vx += dsh[9] * -C3a * 6.f * xy; vy += dsh[9] * -C3a * (3.f * xx - 3.f * yy);
                vx += dsh[10] * C3b * yz; vy += dsh[10] * C3b * xz; vz += dsh[10] * C3b * xy;
                vx += dsh[11] * -C3c * -2.f * xy; vy += dsh[11] * -C3c * (4.f * zz - xx - 3.f * yy); vz += dsh[11] * -C3c * 8.f * yz;
                vx += dsh[12] * C3d * -6.f * xz; vy += dsh[12] * C3d * -6.f * yz; vz += dsh[12] * C3d * (6.f * zz - 3.f * xx - 3.f * yy);
                vx += dsh[13] * -C3c * (4.f * zz - 3.f * xx - yy); vy += dsh[13] * -C3c * -2.f * xy; vz += dsh[13] * -C3c * 8.f * xz;
                vx += dsh[14] * C3e * 2.f * xz; vy += dsh[14] * C3e * -2.f * yz; vz += dsh[14] * C3e * (xx - yy);
                vx += dsh[15] * -C3a * (3.f * xx - 3.f * yy); vy += dsh[15] * -C3a * -6.f * xy;
                // v = d / (|d| + eps):  dL/dd = dv / (n+eps) - d (d . dv) / (n (n+eps)^2)
                const float ddv = d.x * vx + d.y * vy + d.z * vz;
                const float k2 = dn > 0.f ? ddv * inv * inv / dn : 0.f;
                gx = vx * inv - d.x * k2; gy = vy * inv - d.y * k2; gz = vz * inv - d.z * k2;
            }
            const int ray = live ? n / S : -1;
            const int ray0 = __shfl_sync(0xffffffffu, ray, 0);
            if (__all_sync(0xffffffffu, ray == ray0 || ray < 0)) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    gx += __shfl_xor_sync(0xffffffffu, gx, off);
                    gy += __shfl_xor_sync(0xffffffffu, gy, off);
                    gz += __shfl_xor_sync(0xffffffffu, gz, off);
                }
                if (lane == 0 && ray0 >= 0) {
                    atomicAdd(grad_rays_d + 3 * (size_t)ray0 + 0, gx);
                    atomicAdd(grad_rays_d + 3 * (size_t)ray0 + 1, gy);
                    atomicAdd(grad_rays_d + 3 * (size_t)ray0 + 2, gz);
                }
            } else if (live) {
                atomicAdd(grad_rays_d + 3 * (size_t)ray + 0, gx);
                atomicAdd(grad_rays_d + 3 * (size_t)ray + 1, gy);
                atomicAdd(grad_rays_d + 3 * (size_t)ray + 2, gz);
            }
        }
        sync_operands();
        // ---- stage B4: dA1 = dH W2 ; dW2 += dH^T a1 ; db2
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + cDb, umma::desc_kmajor(adz2, k), umma::desc_mnmajor(aW2, k), idg64, k > 0);
            wgrad(cGW2, adz2, aa1, idw64, first);
            wgrad(cGb2, adz2, aA0 + 96, idw8, first);
            umma::mma_commit(&bar);
        }
        wait_mma();
        mul_inplace(cDb, Tg1);
        sync_operands();
        // ---- stage B5: dx = dz1 W1 (32 columns) ; dW1 += dz1^T x ; db1
        if (tid == 0) {
            for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem + cDa, umma::desc_kmajor(ag1, k), umma::desc_mnmajor(aW1, k), idg32, k > 0);
            wgrad(cGW1, ag1, aA0, idw32, first);
            wgrad(cGb1, ag1, aA0 + 96, idw8, first);
            umma::mma_commit(&bar);
        }
        wait_mma();
        umma::tmem_ld32(tmem + cDa + lane_addr, v);
        umma::tc_wait_ld();
        if (live) {
            float4* dst = reinterpret_cast<float4*>(grad_feats + (size_t)n * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                dst[c] = make_float4(v[4 * c] * mask[4 * c], v[4 * c + 1] * mask[4 * c + 1], v[4 * c + 2] * mask[4 * c + 2],
                                     v[4 * c + 3] * mask[4 * c + 3]);
        }
        // every MMA of this tile has completed (the last commit covers all earlier ones), so the
        // next tile may overwrite the operand tiles; TMEM reads are ordered by its first sync_operands()
    }

    // ================= flush the weight / bias gradients accumulated in TMEM =================
    umma::tc_fence_after();
    {
        // M = 64 accumulators: row m lives in TMEM lane 32*(m/16) + m%16 -> warp q, lanes 0..15 hold rows 16q..16q+15
        const int m = 16 * (warp & 3) + lane;
        const bool own = lane < 16;
        float w[32];
        auto flush = [&](int col, int ncols, float* dst, int ld, int col0, int max_rows) {
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                if (ncols - c0 >= 32) umma::tmem_ld32(tmem + col + c0 + lane_addr, w);
                else umma::tmem_ld16(tmem + col + c0 + lane_addr, w);
                umma::tc_wait_ld();
                const int nc = ncols - c0 >= 32 ? 32 : ncols - c0;
                if (own && m < max_rows)
                    for (int j = 0; j < nc; ++j) atomicAdd(dst + (size_t)m * ld + col0 + c0 + j, w[j]);
            }
        };
        flush(cGW1, 32, gp.W1, 32, 0, 64);
        flush(cGW2, 64, gp.W2, 64, 0, 64);
        flush(cGW3a, 32, gp.W3, 48, 0, 64);
        flush(cGW3b, 16, gp.W3, 48, 32, 64);
        flush(cGW4, 64, gp.W4, 64, 0, 64);
        // bias gradients: column 0 of the [64 x 8] accumulators
        float b8[16];
        auto flush_bias = [&](int col, float* dst) {
            umma::tmem_ld16(tmem + col + lane_addr, b8);     // reads 16 columns; only column 0 of this accumulator is used
            umma::tc_wait_ld();
            if (own) atomicAdd(dst + m, b8[0]);
        };
        flush_bias(cGb1, gp.b1); flush_bias(cGb2, gp.b2); flush_bias(cGb3, gp.b3); flush_bias(cGb4, gp.b4);
        // transposed narrow layers: accumulator row = input feature k, column = output o
        umma::tmem_ld16(tmem + cGWhT + lane_addr, b8);
        umma::tc_wait_ld();
        if (own && m < 32) {
            atomicAdd(gp.Ws + m, b8[0]);
            for (int o = 0; o < 3; ++o) { atomicAdd(gp.Wd + o * 32 + m, b8[1 + o]); atomicAdd(gp.Wt + o * 32 + m, b8[4 + o]); }
        }
        umma::tmem_ld16(tmem + cGW5T + lane_addr, b8);
        umma::tc_wait_ld();
        if (own)
            for (int o = 0; o < 3; ++o) atomicAdd(gp.W5 + o * 64 + m, b8[o]);
    }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(gp.bs, small_grad[0]);
        for (int o = 0; o < 3; ++o) { atomicAdd(gp.bd + o, small_grad[1 + o]); atomicAdd(gp.bt + o, small_grad[4 + o]); atomicAdd(gp.b5 + o, small_grad[7 + o]); }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free<512>(tmem);
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

// Backward of snrf_decoder_fwd.  grad_heads[N,10] (same column order as heads) ->
// grad_feats[N,32] (WRITTEN), grad_rays_d[R,3] (ACCUMULATED; may be NULL: the view direction
// only enters through the SH encoding), grad_params = HOST array of 16 DEVICE pointers,
// same order and shapes as params (ACCUMULATED).
SNRF_API int snrf_decoder_bwd(const float* feats, const float* mask32, const float* rays_d, const float* const* params,
                              const float* grad_heads, float* grad_feats, float* grad_rays_d, float* const* grad_params,
                              int N, int S, void* stream)
{
    SNRF_CHECK_ARG(N >= 0 && S > 0, "snrf_decoder_bwd: need N >= 0, S > 0 (N=%d S=%d)", N, S);
    SNRF_CHECK_ARG(params != nullptr && grad_params != nullptr, "snrf_decoder_bwd: params and grad_params are required");
    if (N == 0) return 0;
    DecoderParams p{params[0], params[1], params[2], params[3], params[4], params[5], params[6], params[7],
                    params[8], params[9], params[10], params[11], params[12], params[13], params[14], params[15]};
    DecoderGrads g{grad_params[0], grad_params[1], grad_params[2], grad_params[3], grad_params[4], grad_params[5],
                   grad_params[6], grad_params[7], grad_params[8], grad_params[9], grad_params[10], grad_params[11],
                   grad_params[12], grad_params[13], grad_params[14], grad_params[15]};
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(decoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
        if (e != cudaSuccess) { snrf_set_error("snrf_decoder_bwd: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int num_tiles = snrf_div_up(N, kRows);
    int grid = snrf_sm_count();            // one CTA per SM: all 512 TMEM columns, ~200 KB of shared memory
    if (grid > num_tiles) grid = num_tiles;
    decoder_bwd_kernel<<<grid, kRows, kBwdSmem, (cudaStream_t)stream>>>(feats, mask32, rays_d, p, grad_heads, grad_feats,
                                                                        grad_rays_d, g, N, S, num_tiles);
    SNRF_RETURN_LAUNCH("snrf_decoder_bwd");
}

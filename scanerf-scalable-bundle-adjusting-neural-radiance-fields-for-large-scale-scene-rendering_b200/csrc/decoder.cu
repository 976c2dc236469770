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
// rows, 128B-swizzled (umma.cuh); a layer is a handful of tcgen05.mma (K = 16 each) issued by one
// thread, completion is signalled through an mbarrier (tcgen05.commit), the epilogue pulls the
// fp32 accumulator row out of TMEM (tcgen05.ld), applies bias + activation in registers and
// writes the next layer's operand tile.  Weights are converted to bf16 tiles once per CTA
// (persistent grid).  Nothing but the 40-byte head row per sample is written to HBM in the
// forward; the backward recomputes the forward per tile, keeps every intermediate in shared
// memory, accumulates all weight / bias gradients in TMEM across the CTA's tiles (M = 64
// accumulators) and flushes them once at the end.
//
// Precision.  The Gaussian activation amplifies operand rounding (d a / a = -100 z dz), so plain
// bf16 operands cost ~1 % in the directional branch.  SPLIT mode (the default) therefore feeds the
// tensor cores error-compensated operands in every forward GEMM: v = hi + lo with hi = bf16(v),
// lo = bf16(v - hi), and  A W^T ~= A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T  (three MMAs into the
// same fp32 accumulator, ~16 mantissa bits).  In the backward the input-gradient GEMMs (the chain
// that ends in d/d features -> table and pose gradients) are compensated the same way, with
// dz = dz_hi + dz_lo; the weight-gradient GEMMs use the hi parts only (their rounding errors are
// independent per sample and average out over the batch).
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int kRows = 128;                 // samples per tile
constexpr int kTile = kRows * 128;         // bytes of one operand tile (128 rows x 64 bf16)
constexpr float kGaussLog2 = -50.0f * 1.4426950408889634f;   // exp(-v^2/0.02) = exp2(v^2 * kGaussLog2)
constexpr float kInvSH0 = 1.0f / 0.28125f;                   // 1 / bf16(0.28209479): the SH_0 column doubles as the ones column

// parameter tensors in network.ShallowMLP state_dict order (weight, bias per Linear)
struct DecoderParams {
    const float *W1, *b1, *W2, *b2, *Ws, *bs, *Wd, *bd, *Wt, *bt, *W3, *b3, *W4, *b4, *W5, *b5;
};
struct DecoderGrads {
    float *W1, *b1, *W2, *b2, *Ws, *bs, *Wd, *bd, *Wt, *bt, *W3, *b3, *W4, *b4, *W5, *b5;
};

// ---- shared-memory plan (byte offsets from the 1024-aligned base)
// weight tiles: W1 [64 rows: cols 0..31 hi | cols 32..63 lo], W2, W3 (48 cols used: 32 H + 16 SH),
// W4 [64 x 64], Wh [16 rows = sigma, diffuse3, tint3: cols 0..31 hi | 32..63 lo], W5 [16 rows, 3 used];
// in SPLIT mode W2, W3, W4, W5 have a second tile holding their lo parts.
constexpr int oW1 = 0, oW2 = 8192, oW3 = 16384, oW4 = 24576, oWh = 32768, oW5 = 34816;
constexpr int oW2l = 36864, oW3l = 45056, oW4l = 53248, oW5l = 61440;
template <bool SPLIT> constexpr int weights_end() { return SPLIT ? 63488 : 36864; }
// biases (fp32): b1[64] b2[64] bh[16] b3[64] b4[64] b5[16]; then mask[32], small bias grads[16]
constexpr int nB = 64 + 64 + 16 + 64 + 64 + 16;
constexpr int oB1 = 0, oB2 = 64, oBh = 128, oB3 = 144, oB4 = 208, oB5 = 272;
template <bool SPLIT> constexpr int off_bias() { return weights_end<SPLIT>(); }
template <bool SPLIT> constexpr int off_mask() { return off_bias<SPLIT>() + nB * 4; }
template <bool SPLIT> constexpr int off_small() { return off_mask<SPLIT>() + 128; }
template <bool SPLIT> constexpr int off_tiles() { return ((off_small<SPLIT>() + 64 + 1023) / 1024) * 1024; }
template <bool SPLIT> constexpr int fwd_smem() { return off_tiles<SPLIT>() + (SPLIT ? 5 : 3) * kTile + 1024; }
template <bool SPLIT> constexpr int bwd_smem() { return off_tiles<SPLIT>() + 10 * kTile + 1024; }

// TMEM columns: working accumulators, then (backward only) the persistent gradient accumulators
constexpr int cDa = 0, cDb = 64, cDh = 128;
constexpr int cGW1 = 144, cGW2 = 176, cGW3a = 240, cGW3b = 272, cGW4 = 288, cGWhT = 352, cGW5T = 368,
              cGb1 = 384, cGb2 = 392, cGb3 = 400, cGb4 = 408;   // last one ends at 416 (+8 slack read)

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

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// store 8 values as the hi tile chunk and (SPLIT) their bf16 residuals as the lo tile chunk
template <bool SPLIT>
__device__ __forceinline__ void store8_hl(unsigned char* Thi, int chi, unsigned char* Tlo, int clo, int row, const float* v)
{
    umma::tile_store8(Thi, row, chi, v);
    if (SPLIT) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = v[j] - bf16_round(v[j]);
        umma::tile_store8(Tlo, row, clo, r);
    }
}

// W[out, in] (row-major f32, nn.Linear layout) -> rows [row0, row0+out) of a bf16 tile: hi part in
// columns [0, in), and in SPLIT mode the lo part either in columns [32, 32+in) of the same tile
// (lo_tile == tile, needs in <= 32) or in columns [0, in) of `lo_tile`.
template <bool SPLIT>
__device__ void stage_weight(unsigned char* tile, unsigned char* lo_tile, int row0, const float* __restrict__ W, int out,
                             int in, int tid, int nthreads)
{
    const bool packed = (lo_tile == tile);
    for (int t = tid; t < out * 8; t += nthreads) {
        const int r = t >> 3, c = t & 7;
        float hi[8], lo[8];
        const int src_c = (SPLIT && packed && c >= 4) ? c - 4 : c;     // packed lo chunks mirror chunks 0..3
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = src_c * 8 + j;
            const float w = col < in ? W[r * in + col] : 0.0f;
            hi[j] = w;
            lo[j] = w - bf16_round(w);
        }
        if (SPLIT && packed) {
            umma::tile_store8(tile, row0 + r, c, c >= 4 ? lo : hi);
        } else {
            umma::tile_store8(tile, row0 + r, c, hi);
            if (SPLIT) umma::tile_store8(lo_tile, row0 + r, c, lo);
        }
    }
}
__device__ void zero_tile_rows(unsigned char* tile, int rows, int tid, int nthreads)
{
    for (int t = tid; t < rows * 8; t += nthreads) umma::tile_zero8(tile, t >> 3, t & 7);
}

template <bool SPLIT>
__device__ void stage_all_weights(unsigned char* smem, const DecoderParams& p, const float* __restrict__ mask32, int tid, int nthreads)
{
    zero_tile_rows(smem + oWh, 16, tid, nthreads);
    zero_tile_rows(smem + oW5, 16, tid, nthreads);
    if (SPLIT) zero_tile_rows(smem + oW5l, 16, tid, nthreads);
    __syncthreads();
    stage_weight<SPLIT>(smem + oW1, smem + oW1, 0, p.W1, 64, 32, tid, nthreads);
    stage_weight<SPLIT>(smem + oW2, smem + oW2l, 0, p.W2, 64, 64, tid, nthreads);
    stage_weight<SPLIT>(smem + oW3, smem + oW3l, 0, p.W3, 64, 48, tid, nthreads);
    stage_weight<SPLIT>(smem + oW4, smem + oW4l, 0, p.W4, 64, 64, tid, nthreads);
    stage_weight<SPLIT>(smem + oW5, smem + oW5l, 0, p.W5, 3, 64, tid, nthreads);
    stage_weight<SPLIT>(smem + oWh, smem + oWh, 0, p.Ws, 1, 32, tid, nthreads);
    stage_weight<SPLIT>(smem + oWh, smem + oWh, 1, p.Wd, 3, 32, tid, nthreads);
    stage_weight<SPLIT>(smem + oWh, smem + oWh, 4, p.Wt, 3, 32, tid, nthreads);
    float* b = reinterpret_cast<float*>(smem + off_bias<SPLIT>());
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
    float* m = reinterpret_cast<float*>(smem + off_mask<SPLIT>());
    for (int i = tid; i < 32; i += nthreads) m[i] = mask32 ? mask32[i] : 1.0f;
    float* sg = reinterpret_cast<float*>(smem + off_small<SPLIT>());
    for (int i = tid; i < 16; i += nthreads) sg[i] = 0.0f;
}

// Forward GEMM D (+)= A W^T over `nk` k-steps, issued by one thread.  A = (a_hi tile from k-step
// a_hi_k, a_lo tile from a_lo_k), W likewise; in SPLIT mode three MMAs per k-step.
template <bool SPLIT>
__device__ __forceinline__ void fwd_gemm(uint32_t d, uint32_t a_hi, int a_hi_k, uint32_t a_lo, int a_lo_k, uint32_t w_hi,
                                         int w_hi_k, uint32_t w_lo, int w_lo_k, int nk, uint32_t idesc, bool accumulate)
{
    for (int k = 0; k < nk; ++k)
        umma::mma_bf16(d, umma::desc_kmajor(a_hi, a_hi_k + k), umma::desc_kmajor(w_hi, w_hi_k + k), idesc, accumulate || k > 0);
    if (SPLIT) {
        for (int k = 0; k < nk; ++k)
            umma::mma_bf16(d, umma::desc_kmajor(a_lo, a_lo_k + k), umma::desc_kmajor(w_hi, w_hi_k + k), idesc, 1);
        for (int k = 0; k < nk; ++k)
            umma::mma_bf16(d, umma::desc_kmajor(a_hi, a_hi_k + k), umma::desc_kmajor(w_lo, w_lo_k + k), idesc, 1);
    }
}

// The tiles one forward pass writes.  Forward kernel: a1 == a3, H == a4, g* unused.
struct Tiles {
    unsigned char *A0, *a1, *g1, *H, *a3, *g3, *a4, *g4, *LOa, *LOb;
};

// Per-thread context shared by the forward stages
template <bool SPLIT>
struct Ctx {
    unsigned char* smem;
    uint64_t* bar;
    uint32_t tmem, lane_addr, phase;
    int tid;
    const float* bias;
    const float* mask;
    __device__ __forceinline__ void sync_operands()
    {   // my shared-memory stores and TMEM loads are done -> the next MMAs may run
        umma::fence_async_smem();
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
    }
    __device__ __forceinline__ void wait_mma()
    {
        umma::mbar_wait(bar, phase);
        phase ^= 1u;
        umma::tc_fence_after();
    }
};

// Forward of one tile up to (and including) the L5 GEMM.  Writes the operand tiles, leaves the
// head pre-activations readable: returns sigma/diffuse/tint in head[0..6] (activated) and their
// pre-activations in zh[0..6]; the specular pre-activations are in TMEM columns cDh..cDh+2.
template <bool SPLIT, bool TRAIN>
__device__ __forceinline__ void forward_tile(Ctx<SPLIT>& c, const Tiles& T, const float* __restrict__ feats,
                                             const float* __restrict__ rays_d, int n, bool live, int S, float* head, float* zh,
                                             f3& d, float& dn)
{
    unsigned char* smem = c.smem;
    const int tid = c.tid;
    const uint32_t tmem = c.tmem, lane_addr = c.lane_addr;
    const float* bias = c.bias;
    const float* mask = c.mask;
    const uint32_t aA0 = umma::smem_u32(T.A0), aa1 = umma::smem_u32(T.a1), aH = umma::smem_u32(T.H), aa3 = umma::smem_u32(T.a3),
                   aa4 = umma::smem_u32(T.a4), aLOa = umma::smem_u32(T.LOa), aLOb = umma::smem_u32(T.LOb);
    const uint32_t aW1 = umma::smem_u32(smem + oW1), aW2 = umma::smem_u32(smem + oW2), aW3 = umma::smem_u32(smem + oW3),
                   aW4 = umma::smem_u32(smem + oW4), aWh = umma::smem_u32(smem + oWh), aW5 = umma::smem_u32(smem + oW5),
                   aW2l = umma::smem_u32(smem + oW2l), aW3l = umma::smem_u32(smem + oW3l), aW4l = umma::smem_u32(smem + oW4l),
                   aW5l = umma::smem_u32(smem + oW5l);
    constexpr uint32_t id64 = umma::idesc_bf16(128, 64, 0, 0), id16 = umma::idesc_bf16(128, 16, 0, 0);

    // ---- input row: A0 = [x_hi (0..31) | SH_hi (32..47) | SH_lo (48..63)], LOb = [x_lo (0..31)]
    {
        float v[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (live) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(feats + (size_t)n * 32 + q * 8));
                const float4 b = __ldg(reinterpret_cast<const float4*>(feats + (size_t)n * 32 + q * 8 + 4));
                v[0] = a.x * mask[q * 8 + 0]; v[1] = a.y * mask[q * 8 + 1]; v[2] = a.z * mask[q * 8 + 2]; v[3] = a.w * mask[q * 8 + 3];
                v[4] = b.x * mask[q * 8 + 4]; v[5] = b.y * mask[q * 8 + 5]; v[6] = b.z * mask[q * 8 + 6]; v[7] = b.w * mask[q * 8 + 7];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.0f;
            }
            store8_hl<SPLIT>(T.A0, q, T.LOb, q, tid, v);
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
        store8_hl<SPLIT>(T.A0, 4, T.A0, 6, tid, sh);
        store8_hl<SPLIT>(T.A0, 5, T.A0, 7, tid, sh + 8);
        if (!SPLIT) { umma::tile_zero8(T.A0, tid, 6); umma::tile_zero8(T.A0, tid, 7); }
    }
    c.sync_operands();
    // ---- L1: Da = x W1^T (K = 32)
    if (tid == 0) {
        fwd_gemm<SPLIT>(tmem + cDa, aA0, 0, aLOb, 0, aW1, 0, aW1, 2, 2, id64, false);
        umma::mma_commit(c.bar);
    }
    c.wait_mma();
    float v[32], g[32];
    // Gaussian layer epilogue: a = exp(-50 z^2) -> (Ta, lo in Tlo);  TRAIN: g = da/dz = -100 z a -> Tg
    auto gauss_epilogue = [&](int col, int boff, unsigned char* Ta, unsigned char* Tlo, unsigned char* Tg) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            umma::tmem_ld32(tmem + col + lane_addr + 32 * h, v);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float z = v[j] + bias[boff + 32 * h + j];
                const float a = gauss_act(z);
                v[j] = a;
                if (TRAIN) g[j] = -100.0f * z * a;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                store8_hl<SPLIT>(Ta, 4 * h + q, Tlo, 4 * h + q, tid, v + 8 * q);
                if (TRAIN) umma::tile_store8_f16(Tg, tid, 4 * h + q, g + 8 * q);   // fp16: |g| <= 6.1, read back element-wise only
            }
        }
    };
    gauss_epilogue(cDa, oB1, T.a1, T.LOa, T.g1);
    c.sync_operands();
    // ---- L2: Db = h1 W2^T (K = 64)
    if (tid == 0) {
        fwd_gemm<SPLIT>(tmem + cDb, aa1, 0, aLOa, 0, aW2, 0, aW2l, 0, 4, id64, false);
        umma::mma_commit(c.bar);
    }
    c.wait_mma();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        umma::tmem_ld32(tmem + cDb + lane_addr + 32 * h, v);
        umma::tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += bias[oB2 + 32 * h + j];
#pragma unroll
        for (int q = 0; q < 4; ++q) store8_hl<SPLIT>(T.H, 4 * h + q, T.LOb, 4 * h + q, tid, v + 8 * q);
    }
    c.sync_operands();
    // ---- heads: Dh = H[0:32] Wh^T (K = 32, N = 16);   L3: Da = [H[32:64] | SH] W3^T (K = 48)
    if (tid == 0) {
        fwd_gemm<SPLIT>(tmem + cDh, aH, 0, aLOb, 0, aWh, 0, aWh, 2, 2, id16, false);
        fwd_gemm<SPLIT>(tmem + cDa, aH, 2, aLOb, 2, aW3, 0, aW3l, 0, 2, id64, false);
        fwd_gemm<SPLIT>(tmem + cDa, aA0, 2, aA0, 3, aW3, 2, aW3l, 2, 1, id64, true);
        umma::mma_commit(c.bar);
    }
    c.wait_mma();
    {
        float z[16];
        umma::tmem_ld16(tmem + cDh + lane_addr, z);
        umma::tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 7; ++j) zh[j] = z[j] + bias[oBh + j];
        head[0] = softplusf(zh[0]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            head[4 + j] = sigmoidf(zh[1 + j]);     // diffuse
            head[1 + j] = sigmoidf(zh[4 + j]);     // tint
        }
    }
    gauss_epilogue(cDa, oB3, T.a3, T.LOa, T.g3);
    c.sync_operands();
    // ---- L4: Db = a3 W4^T (K = 64)
    if (tid == 0) {
        fwd_gemm<SPLIT>(tmem + cDb, aa3, 0, aLOa, 0, aW4, 0, aW4l, 0, 4, id64, false);
        umma::mma_commit(c.bar);
    }
    c.wait_mma();
    gauss_epilogue(cDb, oB4, T.a4, T.LOb, T.g4);
    c.sync_operands();
    // ---- L5: Dh = a4 W5^T (K = 64, N = 16)
    if (tid == 0) {
        fwd_gemm<SPLIT>(tmem + cDh, aa4, 0, aLOb, 0, aW5, 0, aW5l, 0, 4, id16, false);
        umma::mma_commit(c.bar);
    }
    c.wait_mma();
}

// ------------------------------- forward ------------------------------------
// feats [N,32] f32, rays_d [R,3] (sample n belongs to ray n / S), out [N,10] f32 =
// (sigma, tint3, diffuse3, specular3).
template <bool SPLIT>
__global__ void __launch_bounds__(kRows, 1)
decoder_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ mask32, const float* __restrict__ rays_d,
                   DecoderParams p, float* __restrict__ out, int N, int S, int num_tiles)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* T0 = smem + off_tiles<SPLIT>();
    unsigned char* T1 = T0 + kTile;
    unsigned char* T2 = T1 + kTile;
    unsigned char* LOa = SPLIT ? T2 + kTile : T1;      // never written when !SPLIT
    unsigned char* LOb = SPLIT ? LOa + kTile : T2;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    stage_all_weights<SPLIT>(smem, p, mask32, tid, kRows);
    if (warp == 0) umma::tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    Ctx<SPLIT> c;
    c.smem = smem; c.bar = &bar; c.tmem = tmem_slot; c.lane_addr = (uint32_t)(32 * (warp & 3)) << 16; c.phase = 0; c.tid = tid;
    c.bias = reinterpret_cast<const float*>(smem + off_bias<SPLIT>());
    c.mask = reinterpret_cast<const float*>(smem + off_mask<SPLIT>());
    const Tiles T{T0, T1, nullptr, T2, T1, nullptr, T2, nullptr, LOa, LOb};

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n = tile * kRows + tid;
        const bool live = n < N;
        float head[10], zh[7];
        f3 d = mk3(0.f, 0.f, 1.f);
        float dn = 1.0f;
        forward_tile<SPLIT, false>(c, T, feats, rays_d, n, live, S, head, zh, d, dn);
        {
            float z[16];
            umma::tmem_ld16(c.tmem + cDh + c.lane_addr, z);
            umma::tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 3; ++j) head[7 + j] = sigmoidf(z[j] + c.bias[oB5 + j]);
        }
        if (live) {
            float2* o = reinterpret_cast<float2*>(out + (size_t)n * 10);
#pragma unroll
            for (int j = 0; j < 5; ++j) o[j] = make_float2(head[2 * j], head[2 * j + 1]);
        }
        // every MMA of this tile has completed; the next tile's first sync_operands() orders its
        // operand stores and this tile's TMEM loads before the next accumulator writes
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free<256>(c.tmem);
}

// ------------------------------- backward -----------------------------------
// Per tile: recompute the forward keeping every intermediate in shared memory, then walk the
// layers backwards.  Each backward stage is one commit group holding the input-gradient GEMM
// (dA_{l-1} = dz_l W_l, weights read MN-major from the very tile the forward used K-major) and
// the weight-gradient GEMMs (dW_l += dz_l^T A_{l-1}, both operands read MN-major, M = 64
// accumulators that stay in TMEM for the whole kernel).  Bias gradients of the 64-wide layers
// are one more GEMM against a column of ones.
//
// shared-memory tiles (16 KB each): A0 = [x | SH | SH_lo], a1, g1 -> dz1, H, a3, g3 -> dz4, a4,
// g4 -> dz5, LOa (a1_lo / a3_lo in the forward, then the lo part of the current dz), LOb (x_lo / H_lo /
// a4_lo in the forward, then [dz_heads | dz_spec | lo parts] and finally dH = dz2); a4 holds dH_lo after
// B1; g = d(activation)/dz.
template <bool SPLIT>
__global__ void __launch_bounds__(kRows, 1)
decoder_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ mask32, const float* __restrict__ rays_d,
                   DecoderParams p, const float* __restrict__ grad_heads, float* __restrict__ grad_feats,
                   float* __restrict__ grad_rays_d, DecoderGrads gp, int N, int S, int num_tiles)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* base = smem + off_tiles<SPLIT>();
    const Tiles T{base, base + kTile, base + 2 * kTile, base + 3 * kTile, base + 4 * kTile, base + 5 * kTile,
                  base + 6 * kTile, base + 7 * kTile, base + 8 * kTile, base + 9 * kTile};
    // backward-phase aliases of tiles the forward no longer needs
    unsigned char* Tdz = T.LOb;          // [dz_heads | dz_spec | their lo parts], then dH = dz2 (hi)
    unsigned char* Tdzlo = T.LOa;        // lo part of the current layer's dz (dz5, dz4, dz1 in turn)
    unsigned char* Tdhlo = T.a4;         // lo part of dH (a4 is dead once B1 has read it)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_all_weights<SPLIT>(smem, p, mask32, tid, kRows);
    float* small_grad = reinterpret_cast<float*>(smem + off_small<SPLIT>());
    if (warp == 0) umma::tmem_alloc<512>(&tmem_slot);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    Ctx<SPLIT> c;
    c.smem = smem; c.bar = &bar; c.tmem = tmem_slot; c.lane_addr = (uint32_t)(32 * (warp & 3)) << 16; c.phase = 0; c.tid = tid;
    c.bias = reinterpret_cast<const float*>(smem + off_bias<SPLIT>());
    c.mask = reinterpret_cast<const float*>(smem + off_mask<SPLIT>());
    const uint32_t tmem = c.tmem, lane_addr = c.lane_addr;
    const float* bias = c.bias;
    const float* mask = c.mask;
    const uint32_t aA0 = umma::smem_u32(T.A0), aa1 = umma::smem_u32(T.a1), ag1 = umma::smem_u32(T.g1), aH = umma::smem_u32(T.H),
                   aa3 = umma::smem_u32(T.a3), ag3 = umma::smem_u32(T.g3), aa4 = umma::smem_u32(T.a4), ag4 = umma::smem_u32(T.g4),
                   adz = umma::smem_u32(Tdz), adzlo = umma::smem_u32(Tdzlo), adhlo = umma::smem_u32(Tdhlo);
    // bias gradients of the 64-wide layers ride on the weight-gradient GEMMs: column 32 of A0 holds
    // SH_0 = 0.28209479 (a constant, exactly representable products), so dz^T A0[:, 32:40] column 0
    // is c0 * sum_n dz_n; the flush divides by c0.  Rows past N have SH = 0 and dz = 0.
    const uint32_t aones = aA0 + 64;
    const uint32_t aW1 = umma::smem_u32(smem + oW1), aW2 = umma::smem_u32(smem + oW2), aW3 = umma::smem_u32(smem + oW3),
                   aW4 = umma::smem_u32(smem + oW4), aWh = umma::smem_u32(smem + oWh), aW5 = umma::smem_u32(smem + oW5);
    // input-gradient GEMMs: A K-major (dz rows), B MN-major (weight tile: rows = K = out, cols = N = in)
    constexpr uint32_t idg64 = umma::idesc_bf16(128, 64, 0, 1), idg32 = umma::idesc_bf16(128, 32, 0, 1);
    // weight-gradient GEMMs: both operands MN-major, M = 64
    constexpr uint32_t idw64 = umma::idesc_bf16(64, 64, 1, 1), idw32 = umma::idesc_bf16(64, 32, 1, 1),
                       idw16 = umma::idesc_bf16(64, 16, 1, 1), idw8 = umma::idesc_bf16(64, 8, 1, 1);
    const uint32_t aW2l = umma::smem_u32(smem + oW2l), aW3l = umma::smem_u32(smem + oW3l), aW4l = umma::smem_u32(smem + oW4l),
                   aW5l = umma::smem_u32(smem + oW5l);
    // dA = dz W over `nk` k-steps; in SPLIT mode dz = dz_hi + dz_lo and W = W_hi + W_lo (the lo tile,
    // or the lo half of a packed tile at a 64-byte column offset): dz_hi W_hi + dz_hi W_lo + dz_lo W_hi
    auto dgrad = [&](int col, uint32_t dz_tile, int dz_k, uint32_t dzlo_tile, int dzlo_k, uint32_t w_hi, uint32_t w_lo, int nk,
                     uint32_t idesc) {
        for (int k = 0; k < nk; ++k)
            umma::mma_bf16(tmem + col, umma::desc_kmajor(dz_tile, dz_k + k), umma::desc_mnmajor(w_hi, k), idesc, k > 0);
        if (SPLIT) {
            for (int k = 0; k < nk; ++k)
                umma::mma_bf16(tmem + col, umma::desc_kmajor(dz_tile, dz_k + k), umma::desc_mnmajor(w_lo, k), idesc, 1);
            for (int k = 0; k < nk; ++k)
                umma::mma_bf16(tmem + col, umma::desc_kmajor(dzlo_tile, dzlo_k + k), umma::desc_mnmajor(w_hi, k), idesc, 1);
        }
    };
    // dW += dz^T B over the 128 rows of the tile (8 k-steps); `first` = first tile of this CTA.  In SPLIT
    // mode the dz operand is compensated (dz_hi + dz_lo: `a_lo` = the lo tile, or the packed lo columns);
    // the activation operand B stays bf16 (its lo part is gone by now).
    auto wgrad = [&](int col, uint32_t a_tile, uint32_t a_lo, uint32_t b_tile_plus_off, uint32_t idesc, bool first) {
        for (int k = 0; k < 8; ++k)
            umma::mma_bf16(tmem + col, umma::desc_mnmajor(a_tile, k), umma::desc_mnmajor(b_tile_plus_off, k), idesc, (!first) || k > 0);
        if (SPLIT)
            for (int k = 0; k < 8; ++k)
                umma::mma_bf16(tmem + col, umma::desc_mnmajor(a_lo, k), umma::desc_mnmajor(b_tile_plus_off, k), idesc, 1);
    };
    // transposed narrow layers (dW^T += A^T dz): the compensated operand is B
    auto wgrad_t = [&](int col, uint32_t a_tile, uint32_t b_hi, uint32_t b_lo, uint32_t idesc, bool first) {
        for (int k = 0; k < 8; ++k)
            umma::mma_bf16(tmem + col, umma::desc_mnmajor(a_tile, k), umma::desc_mnmajor(b_hi, k), idesc, (!first) || k > 0);
        if (SPLIT)
            for (int k = 0; k < 8; ++k)
                umma::mma_bf16(tmem + col, umma::desc_mnmajor(a_tile, k), umma::desc_mnmajor(b_lo, k), idesc, 1);
    };
    float v[32];
    // dz = dA * g: hi part in place over the g tile, lo part (SPLIT) into Tlo (this thread's row only)
    auto mul_inplace = [&](int col, unsigned char* Tg, unsigned char* Tlo) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            umma::tmem_ld32(tmem + col + lane_addr + 32 * h, v);
            umma::tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(Tg + umma::tile_chunk_off(tid, 4 * h + q));
                const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
                float o[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __half2 h2 = *reinterpret_cast<const __half2*>(&w[e]);
                    o[2 * e] = v[8 * q + 2 * e] * __low2float(h2);
                    o[2 * e + 1] = v[8 * q + 2 * e + 1] * __high2float(h2);
                }
                store8_hl<SPLIT>(Tg, 4 * h + q, Tlo, 4 * h + q, tid, o);
            }
        }
    };

    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, first = false) {
        const int n = tile * kRows + tid;
        const bool live = n < N;
        float head[10], zh[7];
        f3 d = mk3(0.f, 0.f, 1.f);
        float dn = 1.0f;
        forward_tile<SPLIT, true>(c, T, feats, rays_d, n, live, S, head, zh, d, dn);

        // ---- d(loss)/d(pre-activations) of the 7 heads and the 3 specular outputs
        float dzh[16], dzs[16];
        {
            float gh[10];
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
            for (int j = 0; j < 16; ++j) { dzh[j] = 0.0f; dzs[j] = 0.0f; }
            dzh[0] = gh[0] * (zh[0] > 20.0f ? 1.0f : sigmoidf(zh[0]));          // softplus' = sigmoid
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                dzh[1 + j] = gh[4 + j] * head[4 + j] * (1.0f - head[4 + j]);    // diffuse
                dzh[4 + j] = gh[1 + j] * head[1 + j] * (1.0f - head[1 + j]);    // tint
                const float s = sigmoidf(z[j] + bias[oB5 + j]);
                dzs[j] = gh[7 + j] * s * (1.0f - s);                            // specular
            }
            // Tdz = [dz_heads 0..15 | dz_spec 16..31 | dz_heads_lo 32..47 | dz_spec_lo 48..63]
            store8_hl<SPLIT>(Tdz, 0, Tdz, 4, tid, dzh);
            store8_hl<SPLIT>(Tdz, 1, Tdz, 5, tid, dzh + 8);
            store8_hl<SPLIT>(Tdz, 2, Tdz, 6, tid, dzs);
            store8_hl<SPLIT>(Tdz, 3, Tdz, 7, tid, dzs + 8);
            // bias gradients of the narrow layers: warp reduction, one shared atomic per warp
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                float t = j < 7 ? dzh[j] : dzs[j - 7];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                if (lane == 0) atomicAdd(small_grad + j, t);
            }
        }
        c.sync_operands();
        // ---- B1: dA4 = dz_spec W5 ; dH[0:32] = dz_heads Wh ; dW5^T += a4^T dz_spec ; dWh^T += H^T dz_heads
        if (tid == 0) {
            dgrad(cDa, adz, 1, adz, 3, aW5, aW5l, 1, idg64);
            dgrad(cDb, adz, 0, adz, 2, aWh, aWh + 64, 1, idg32);
            wgrad_t(cGW5T, aa4, adz + 32, adz + 96, idw16, first);
            wgrad_t(cGWhT, aH, adz, adz + 64, idw16, first);
            umma::mma_commit(&bar);
        }
        c.wait_mma();
        mul_inplace(cDa, T.g4, Tdzlo);                           // dz5
        umma::tmem_ld32(tmem + cDb + lane_addr, v);              // dH[0:32] -> dz2 tile columns 0..31 (over the consumed dz_heads/spec)
        umma::tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) store8_hl<SPLIT>(Tdz, q, Tdhlo, q, tid, v + 8 * q);
        c.sync_operands();
        // ---- B2: dA3 = dz5 W4 ; dW4 += dz5^T a3 ; db4
        if (tid == 0) {
            dgrad(cDb, ag4, 0, adzlo, 0, aW4, aW4l, 4, idg64);
            wgrad(cGW4, ag4, adzlo, aa3, idw64, first);
            wgrad(cGb4, ag4, adzlo, aones, idw8, first);
            umma::mma_commit(&bar);
        }
        c.wait_mma();
        mul_inplace(cDb, T.g3, Tdzlo);                           // dz4
        c.sync_operands();
        // ---- B3: d[x2] = dz4 W3 ; dW3 += dz4^T [H[32:64] | SH] ; db3
        if (tid == 0) {
            dgrad(cDa, ag3, 0, adzlo, 0, aW3, aW3l, 4, idg64);
            wgrad(cGW3a, ag3, adzlo, aH + 64, idw32, first);
            wgrad(cGW3b, ag3, adzlo, aA0 + 64, idw16, first);
            wgrad(cGb3, ag3, adzlo, aones, idw8, first);
            umma::mma_commit(&bar);
        }
        c.wait_mma();
        umma::tmem_ld32(tmem + cDa + lane_addr, v);              // dH[32:64]
        umma::tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) store8_hl<SPLIT>(Tdz, 4 + q, Tdhlo, 4 + q, tid, v + 8 * q);
        if (grad_rays_d != nullptr) {                            // d/d(ray direction) through the SH encoding
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
        c.sync_operands();
        // ---- B4: dA1 = dH W2 ; dW2 += dH^T a1 ; db2
        if (tid == 0) {
            dgrad(cDb, adz, 0, adhlo, 0, aW2, aW2l, 4, idg64);
            wgrad(cGW2, adz, adhlo, aa1, idw64, first);
            wgrad(cGb2, adz, adhlo, aones, idw8, first);
            umma::mma_commit(&bar);
        }
        c.wait_mma();
        mul_inplace(cDb, T.g1, Tdzlo);                           // dz1
        c.sync_operands();
        // ---- B5: dx = dz1 W1 (32 columns) ; dW1 += dz1^T x ; db1
        if (tid == 0) {
            dgrad(cDa, ag1, 0, adzlo, 0, aW1, aW1 + 64, 4, idg32);
            wgrad(cGW1, ag1, adzlo, aA0, idw32, first);
            wgrad(cGb1, ag1, adzlo, aones, idw8, first);
            umma::mma_commit(&bar);
        }
        c.wait_mma();
        umma::tmem_ld32(tmem + cDa + lane_addr, v);
        umma::tc_wait_ld();
        if (live) {
            float4* dst = reinterpret_cast<float4*>(grad_feats + (size_t)n * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(v[4 * q] * mask[4 * q], v[4 * q + 1] * mask[4 * q + 1], v[4 * q + 2] * mask[4 * q + 2],
                                     v[4 * q + 3] * mask[4 * q + 3]);
        }
        // every MMA of this tile has completed (the last commit covers all earlier ones), so the
        // next tile may overwrite the operand tiles
    }

    // ================= flush the weight / bias gradients accumulated in TMEM =================
    umma::tc_fence_after();
    {
        // M = 64 accumulators: row m lives in TMEM lane 32*(m/16) + m%16 -> warp q, lanes 0..15 hold rows 16q..16q+15
        const int m = 16 * (warp & 3) + lane;
        const bool own = lane < 16;
        float w[32];
        auto flush = [&](int col, int ncols, float* dst, int ld, int col0) {
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                const int nc = ncols - c0 >= 32 ? 32 : ncols - c0;
                if (nc == 32) umma::tmem_ld32(tmem + col + c0 + lane_addr, w);
                else umma::tmem_ld16(tmem + col + c0 + lane_addr, w);
                umma::tc_wait_ld();
                if (own)
                    for (int j = 0; j < nc; ++j) atomicAdd(dst + (size_t)m * ld + col0 + c0 + j, w[j]);
            }
        };
        flush(cGW1, 32, gp.W1, 32, 0);
        flush(cGW2, 64, gp.W2, 64, 0);
        flush(cGW3a, 32, gp.W3, 48, 0);
        flush(cGW3b, 16, gp.W3, 48, 32);
        flush(cGW4, 64, gp.W4, 64, 0);
        float b8[16];
        auto flush_bias = [&](int col, float* dst) {            // column 0 of a [64 x 8] accumulator
            umma::tmem_ld16(tmem + col + lane_addr, b8);
            umma::tc_wait_ld();
            if (own) atomicAdd(dst + m, b8[0] * kInvSH0);
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

int g_split = 1;       // 1 = error-compensated bf16x3 operands in the forward GEMMs (default), 0 = plain bf16

template <typename K>
int set_smem(K kernel, int bytes, const char* name)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { snrf_set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
// 0 = plain bf16 operands (fastest), 1 = bf16x3 split operands in the forward GEMMs (default)
SNRF_API void snrf_decoder_set_precision(int split) { g_split = split ? 1 : 0; }

// params: HOST array of 16 DEVICE pointers in network.ShallowMLP state_dict order
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
        int rc = set_smem(decoder_fwd_kernel<true>, fwd_smem<true>(), "snrf_decoder_fwd");
        if (rc == 0) rc = set_smem(decoder_fwd_kernel<false>, fwd_smem<false>(), "snrf_decoder_fwd");
        if (rc) return rc;
        configured = true;
    }
    const int num_tiles = snrf_div_up(N, kRows);
    cudaStream_t s = (cudaStream_t)stream;
    if (g_split) {
        int grid = snrf_sm_count();                 // ~145 KB of shared memory: one CTA per SM
        if (grid > num_tiles) grid = num_tiles;
        decoder_fwd_kernel<true><<<grid, kRows, fwd_smem<true>(), s>>>(feats, mask32, rays_d, p, heads_out, N, S, num_tiles);
    } else {
        int grid = snrf_sm_count() * 2;             // 2 x (256 TMEM columns, ~87 KB)
        if (grid > num_tiles) grid = num_tiles;
        decoder_fwd_kernel<false><<<grid, kRows, fwd_smem<false>(), s>>>(feats, mask32, rays_d, p, heads_out, N, S, num_tiles);
    }
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
        int rc = set_smem(decoder_bwd_kernel<true>, bwd_smem<true>(), "snrf_decoder_bwd");
        if (rc == 0) rc = set_smem(decoder_bwd_kernel<false>, bwd_smem<false>(), "snrf_decoder_bwd");
        if (rc) return rc;
        configured = true;
    }
    const int num_tiles = snrf_div_up(N, kRows);
    int grid = snrf_sm_count();            // one CTA per SM: all 512 TMEM columns, 200-225 KB of shared memory
    if (grid > num_tiles) grid = num_tiles;
    cudaStream_t s = (cudaStream_t)stream;
    if (g_split)
        decoder_bwd_kernel<true><<<grid, kRows, bwd_smem<true>(), s>>>(feats, mask32, rays_d, p, grad_heads, grad_feats, grad_rays_d, g, N, S, num_tiles);
    else
        decoder_bwd_kernel<false><<<grid, kRows, bwd_smem<false>(), s>>>(feats, mask32, rays_d, p, grad_heads, grad_feats, grad_rays_d, g, N, S, num_tiles);
    SNRF_RETURN_LAUNCH("snrf_decoder_bwd");
}

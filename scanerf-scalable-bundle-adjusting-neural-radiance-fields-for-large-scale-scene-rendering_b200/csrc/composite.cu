// Front-to-back alpha compositing of per-sample decoder outputs, forward and
// backward, sm_100a.
//
// Replaces (behaviour, not code) the torch chain in the reference's
// HashGrid.cal_integrate_weight / accumulate / render_batch_rays tail
// (hashgrid/__init__.py:344-366, 564-596):
//   delta_k = dist_k * |d|  (last = 1e10 when `infinity`)
//   alpha_k = 1 - exp(-sigma_k delta_k)
//   T_k     = prod_{j<k} (1 - alpha_j + 1e-6)        (torch.cumprod with the 1e-6 fudge)
//   w_k     = alpha_k T_k ;  T_left = T_{S-1}        (T[:, -1] of the exclusive product)
//   depth = sum w z ; tint = sum w tint ; diffuse = sum w c_d ; specular = sum w (tint * c_s)
//   l2    = sum stopgrad(w) c_s^2                    (per ray, per channel)
//
// One WARP per ray: lane l owns samples l, l+32, ... so every global access is
// coalesced; the transmittance is a multiplicative warp scan per 32-sample chunk
// with the chunk product carried forward (warp-segmented scan); the backward runs the
// mirrored suffix-sum scan.  HBM-bound: 48 B read + 4 B written per sample forward.
// Per-sample inputs are addressed as base + k*stride so both the separate tensors of
// the torch decoder (strides 1,3,3,3) and the packed heads of the fused MLP kernel work.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kOutStride = 16;   // per-ray output row: depth, tint3, diffuse3, specular3, l2_3, T_left, 2 pad

struct Heads {
    const float* sigma; const float* tint; const float* diffuse; const float* specular;
    int s_sigma, s_tint, s_diffuse, s_specular;   // strides in floats between samples
};
struct HeadGrads {
    float* sigma; float* tint; float* diffuse; float* specular;
    int s_sigma, s_tint, s_diffuse, s_specular;
};

__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v *= o;
    }
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// inclusive suffix sum: lane l receives the sum over lanes >= l
__device__ __forceinline__ float warp_incl_suffix_sum(float v, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float o = __shfl_down_sync(0xffffffffu, v, off);
        if (lane + off < 32) v += o;
    }
    return v;
}

// PACKED: the four heads are the columns (sigma | tint3 | diffuse3 | specular3) of ONE [R*S, 10] array (the tensor-core decoder's
// output): a warp's 32 rows are 1280 contiguous bytes, fetched with coalesced 8-byte loads into shared memory and read from
// there -- the per-column loads at a 40-byte stride cost ten L1 wavefronts each and bound these kernels, not HBM.
template <bool PACKED>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(Heads in, const float* __restrict__ z_vals, const float* __restrict__ dists,
                     const float* __restrict__ rays_d, const unsigned char* __restrict__ ray_valid, int R, int S, int infinity,
                     int inf_start, float* __restrict__ weights, float* __restrict__ trans, float* __restrict__ out,
                     float ert_eps, unsigned char* __restrict__ sample_live)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int chunks = (S + 31) >> 5;
    __shared__ __align__(16) float stage_all[PACKED ? kWarpsPerBlock * 320 : 1];
    float* sw = stage_all + (PACKED ? (threadIdx.x >> 5) * 320 : 0);
    // composites ray r; `front` = transmittance in front of the ray's first sample for the termination test only (1, or the
    // T_left of the foreground chain when r is the background chain of the same pixel); returns the ray's T_left
    auto composite_ray = [&](int r, float front) -> float {
        if (ray_valid && !ray_valid[r]) {   // masked-out ray: nothing composited, full transmittance (HashGrid._scatter_back defaults)
            for (int k = lane; k < S; k += 32) {
                weights[(size_t)r * S + k] = 0.0f;
                if (sample_live) sample_live[(size_t)r * S + k] = 0;
            }
            if (lane < kOutStride) out[(size_t)r * kOutStride + lane] = lane == 13 ? 1.0f : 0.0f;
            return 1.0f;
        }
        const f3 d = ld3(rays_d + 3 * (size_t)r);
        const float dn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        const bool inf = infinity != 0 && r >= inf_start;      // last step of this ray reaches to infinity (1e10)
        float carry = 1.0f;                 // product of beta over all previous chunks
        float acc[13];
#pragma unroll
        for (int i = 0; i < 13; ++i) acc[i] = 0.f;
        float t_left = 0.0f;
        for (int c = 0; c < chunks; ++c) {
            const int k = c * 32 + lane;
            const bool live = k < S;
            const size_t n = (size_t)r * S + k;
            float beta = 1.0f, alpha = 0.0f;
            if (PACKED) {
                const int rows = min(32, S - c * 32);
                const float2* src = reinterpret_cast<const float2*>(in.sigma + ((size_t)r * S + (size_t)c * 32) * 10);
                __syncwarp();
                for (int i = lane; i < rows * 5; i += 32) reinterpret_cast<float2*>(sw)[i] = __ldg(src + i);
                __syncwarp();
            }
            if (live) {
                float delta = dists[n] * dn;
                if (inf && k == S - 1) delta = 1e10f;
                alpha = 1.0f - expf(-(PACKED ? sw[lane * 10] : in.sigma[n * in.s_sigma]) * delta);
                beta = 1.0f - alpha + 1e-6f;
            }
            const float incl = warp_incl_scan_mul(beta, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float T = carry * excl;
            if (live) {
                // early ray termination (opt-in, ert_eps > 0): a sample behind transmittance < ert_eps is composited with
                // weight 0 and flagged dead -- the backward kernels (composite, decoder, encode scatter) skip it
                const bool dead = front * T < ert_eps;
                const float w = dead ? 0.0f : alpha * T;
                if (sample_live) sample_live[n] = dead ? 0 : 1;
                weights[n] = w;
                if (trans) trans[n] = T;
                const float z = z_vals[n];
                const f3 ti = PACKED ? ld3(sw + lane * 10 + 1) : ld3(in.tint + n * in.s_tint);
                const f3 di = PACKED ? ld3(sw + lane * 10 + 4) : ld3(in.diffuse + n * in.s_diffuse);
                const f3 sp = PACKED ? ld3(sw + lane * 10 + 7) : ld3(in.specular + n * in.s_specular);
                acc[0] += w * z;
                acc[1] += w * ti.x; acc[2] += w * ti.y; acc[3] += w * ti.z;
                acc[4] += w * di.x; acc[5] += w * di.y; acc[6] += w * di.z;
                acc[7] += w * (ti.x * sp.x); acc[8] += w * (ti.y * sp.y); acc[9] += w * (ti.z * sp.z);
                acc[10] += w * (sp.x * sp.x); acc[11] += w * (sp.y * sp.y); acc[12] += w * (sp.z * sp.z);
                if (k == S - 1) t_left = T;
            }
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
#pragma unroll
        for (int i = 0; i < 13; ++i) acc[i] = warp_sum(acc[i]);
        t_left = warp_sum(t_left);          // exactly one lane holds it
        float* o = out + (size_t)r * kOutStride;
        if (lane < 13) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < 13; ++i) if (lane == i) v = acc[i];
            o[lane] = v;
        } else if (lane == 13) {
            o[13] = t_left;
        }
        return t_left;
    };
    if (ert_eps > 0.0f && infinity != 0 && inf_start > 0 && R == 2 * inf_start) {
        // joint batch with termination: rays [0, R/2) are the foreground chains, ray r + R/2 the background chain of the same
        // pixel, whose colour is later weighted by the foreground's T_left -- one warp takes both, the second with that factor
        for (int r = warp; r < inf_start; r += nwarps) {
            const float t_fore = composite_ray(r, 1.0f);
            composite_ray(r + inf_start, t_fore);
        }
    } else {
        for (int r = warp; r < R; r += nwarps) composite_ray(r, 1.0f);
    }
}

// Backward.  g_out[R,16] uses the forward's row layout (depth, tint3, diffuse3, specular3,
// l2_3, T_left); g_weights[R,S] optional.  Writes (overwrites) per-sample gradients and
// accumulates nothing: grad_rays_d[R,3] is written (via |d| in delta).
template <bool PACKED>      // (see composite_fwd_kernel; here the ten gradient columns of a row go back the same way)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(Heads in, const float* __restrict__ z_vals, const float* __restrict__ dists,
                     const float* __restrict__ rays_d, const float* __restrict__ trans,
                     const float* __restrict__ g_out, const float* __restrict__ g_weights, const unsigned char* __restrict__ ray_valid,
                     int R, int S, int infinity, int inf_start, HeadGrads g, float* __restrict__ grad_rays_d,
                     const unsigned char* __restrict__ sample_live)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int chunks = (S + 31) >> 5;
    __shared__ __align__(16) float stage_all[PACKED ? kWarpsPerBlock * 320 : 1];
    float* sw = stage_all + (PACKED ? (threadIdx.x >> 5) * 320 : 0);
    for (int r = warp; r < R; r += nwarps) {
        if (ray_valid && !ray_valid[r]) {   // masked-out ray: its head gradients are never read downstream
            if (lane == 0 && grad_rays_d) st3(grad_rays_d + 3 * (size_t)r, mk3(0.f, 0.f, 0.f));
            continue;
        }
        const f3 d = ld3(rays_d + 3 * (size_t)r);
        const float dn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        const bool inf = infinity != 0 && r >= inf_start;
        const float* go = g_out + (size_t)r * kOutStride;
        const float g_depth = go[0];
        const f3 g_ti = mk3(go[1], go[2], go[3]), g_di = mk3(go[4], go[5], go[6]), g_sp = mk3(go[7], go[8], go[9]);
        const f3 g_l2 = mk3(go[10], go[11], go[12]);
        const float g_tl = go[13];
        // suffix accumulates sum_{m>k} G_m w_m (+ g_Tleft * T_left, attached to the last sample)
        float suffix = 0.0f;
        float g_dn = 0.0f;
        for (int c = chunks - 1; c >= 0; --c) {
            const int k = c * 32 + lane;
            const bool live = k < S;
            const size_t n = (size_t)r * S + k;
            float Gw = 0.0f, w = 0.0f, e = 1.0f, delta = 0.0f, sig = 0.0f, G = 0.0f, T = 0.0f;
            const bool dead = live && sample_live != nullptr && sample_live[n] == 0;    // terminated in the forward: weight 0, no gradient
            const int rows = min(32, S - c * 32);
            if (PACKED) {
                const float2* src = reinterpret_cast<const float2*>(in.sigma + ((size_t)r * S + (size_t)c * 32) * 10);
                __syncwarp();
                for (int i = lane; i < rows * 5; i += 32) reinterpret_cast<float2*>(sw)[i] = __ldg(src + i);
                __syncwarp();
            }
            f3 o_ti = mk3(0.f, 0.f, 0.f), o_di = o_ti, o_sp = o_ti;
            if (live) {
                T = trans[n];
                sig = PACKED ? sw[lane * 10] : in.sigma[n * in.s_sigma];
                delta = dists[n] * dn;
                if (inf && k == S - 1) delta = 1e10f;
                e = expf(-sig * delta);                     // 1 - alpha
                w = dead ? 0.0f : (1.0f - e) * T;
                const f3 ti = PACKED ? ld3(sw + lane * 10 + 1) : ld3(in.tint + n * in.s_tint);
                const f3 di = PACKED ? ld3(sw + lane * 10 + 4) : ld3(in.diffuse + n * in.s_diffuse);
                const f3 sp = PACKED ? ld3(sw + lane * 10 + 7) : ld3(in.specular + n * in.s_specular);
                G = g_depth * z_vals[n] + dot3(g_ti, ti) + dot3(g_di, di) +
                    (g_sp.x * ti.x * sp.x + g_sp.y * ti.y * sp.y + g_sp.z * ti.z * sp.z);
                if (g_weights) G += g_weights[n];
                Gw = G * w;
                // per-sample attribute gradients
                o_ti = mk3(w * (g_ti.x + g_sp.x * sp.x), w * (g_ti.y + g_sp.y * sp.y), w * (g_ti.z + g_sp.z * sp.z));
                o_di = w * g_di;
                o_sp = mk3(w * (g_sp.x * ti.x + 2.0f * g_l2.x * sp.x), w * (g_sp.y * ti.y + 2.0f * g_l2.y * sp.y),
                           w * (g_sp.z * ti.z + 2.0f * g_l2.z * sp.z));
                if (!PACKED) {
                    st3(g.tint + n * g.s_tint, o_ti);
                    st3(g.diffuse + n * g.s_diffuse, o_di);
                    st3(g.specular + n * g.s_specular, o_sp);
                }
            }
            // T_left = T_{S-1}: contributes to d/d alpha_k for k < S-1, i.e. it rides with sample S-1
            float tail = Gw;
            if (live && k == S - 1) tail += g_tl * T;
            const float incl = warp_incl_suffix_sum(tail, lane);
            const float after = suffix + (incl - tail);      // sum over samples strictly after k
            if (live) {
                const float beta = e + 1e-6f;
                const float g_alpha = dead ? 0.0f : G * T - after / beta;
                const float o_sig = g_alpha * delta * e;
                if (PACKED) {                                // (every lane has read its row: the shuffles above synchronised the warp)
                    sw[lane * 10] = o_sig;
                    st3(sw + lane * 10 + 1, o_ti); st3(sw + lane * 10 + 4, o_di); st3(sw + lane * 10 + 7, o_sp);
                } else {
                    g.sigma[n * g.s_sigma] = o_sig;
                }
                if (!(inf && k == S - 1)) g_dn += g_alpha * sig * e * dists[n];
            }
            if (PACKED) {
                float2* dst = reinterpret_cast<float2*>(g.sigma + ((size_t)r * S + (size_t)c * 32) * 10);
                __syncwarp();
                for (int i = lane; i < rows * 5; i += 32) dst[i] = reinterpret_cast<const float2*>(sw)[i];
            }
            suffix += __shfl_sync(0xffffffffu, incl, 0);
        }
        g_dn = warp_sum(g_dn);
        if (lane == 0 && grad_rays_d) {
            const float inv = dn > 0.f ? g_dn / dn : 0.f;
            st3(grad_rays_d + 3 * (size_t)r, mk3(inv * d.x, inv * d.y, inv * d.z));
        }
    }
}

int g_fwd_packed = 0;      // snrf_composite_set_fwd_packed

inline int grid_rays(int R)
{
    const int want = snrf_div_up(R, kWarpsPerBlock);
    const int cap = snrf_sm_count() * 8;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
// measurement hook: 1 = the forward stages packed head rows through shared memory like the backward does (default 0: slower)
SNRF_API void snrf_composite_set_fwd_packed(int on) { g_fwd_packed = on ? 1 : 0; }
// ert_eps > 0: early ray termination -- samples behind transmittance < ert_eps get weight 0; sample_live [R*S] (may be NULL)
// receives 1 / 0 per sample (0 also for the samples of masked-out rays) for the backward kernels.
SNRF_API int snrf_composite_fwd_ert(const float* sigma, const float* tint, const float* diffuse, const float* specular,
                                    int s_sigma, int s_tint, int s_diffuse, int s_specular,
                                    const float* z_vals, const float* dists, const float* rays_d, const unsigned char* ray_valid,
                                    int R, int S, int infinity, int inf_start, float* weights, float* trans, float* out,
                                    float ert_eps, unsigned char* sample_live, void* stream)
{
    SNRF_CHECK_ARG(S > 0, "snrf_composite_fwd: S must be positive");
    SNRF_CHECK_ARG(ert_eps >= 0.0f && ert_eps < 1.0f, "snrf_composite_fwd_ert: ert_eps must lie in [0, 1) (got %g)", (double)ert_eps);
    if (R <= 0) return 0;
    Heads in{sigma, tint, diffuse, specular, s_sigma, s_tint, s_diffuse, s_specular};
    // one [R*S, 10] array behind the four pointers (the decoder's packed head rows)?  Staging pays in the backward (ten strided
    // stores per row on top of the loads: 0.152 -> 0.099 ms at C2) but not in the forward (0.063 -> 0.088 ms: its strided loads hit
    // L1 after the first column): off here, kept selectable for measurements.
    const bool packed = g_fwd_packed && s_sigma == 10 && s_tint == 10 && s_diffuse == 10 && s_specular == 10 && tint == sigma + 1 &&
                        diffuse == sigma + 4 && specular == sigma + 7 && ((uintptr_t)sigma & 7) == 0;
    if (packed) composite_fwd_kernel<true><<<grid_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(in, z_vals, dists, rays_d, ray_valid, R, S, infinity, inf_start, weights, trans, out, ert_eps, sample_live);
    else composite_fwd_kernel<false><<<grid_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(in, z_vals, dists, rays_d, ray_valid, R, S, infinity, inf_start, weights, trans, out, ert_eps, sample_live);
    SNRF_RETURN_LAUNCH("snrf_composite_fwd");
}
SNRF_API int snrf_composite_fwd(const float* sigma, const float* tint, const float* diffuse, const float* specular,
                                int s_sigma, int s_tint, int s_diffuse, int s_specular,
                                const float* z_vals, const float* dists, const float* rays_d, const unsigned char* ray_valid,
                                int R, int S, int infinity, int inf_start, float* weights, float* trans, float* out, void* stream)
{
    return snrf_composite_fwd_ert(sigma, tint, diffuse, specular, s_sigma, s_tint, s_diffuse, s_specular, z_vals, dists, rays_d, ray_valid,
                                  R, S, infinity, inf_start, weights, trans, out, 0.0f, nullptr, stream);
}

SNRF_API int snrf_composite_bwd_ert(const float* sigma, const float* tint, const float* diffuse, const float* specular,
                                    int s_sigma, int s_tint, int s_diffuse, int s_specular,
                                    const float* z_vals, const float* dists, const float* rays_d, const float* trans,
                                    const float* g_out, const float* g_weights, const unsigned char* ray_valid, int R, int S, int infinity,
                                    int inf_start, float* g_sigma, float* g_tint, float* g_diffuse, float* g_specular,
                                    int gs_sigma, int gs_tint, int gs_diffuse, int gs_specular,
                                    float* grad_rays_d, const unsigned char* sample_live, void* stream);
SNRF_API int snrf_composite_bwd(const float* sigma, const float* tint, const float* diffuse, const float* specular,
                                int s_sigma, int s_tint, int s_diffuse, int s_specular,
                                const float* z_vals, const float* dists, const float* rays_d, const float* trans,
                                const float* g_out, const float* g_weights, const unsigned char* ray_valid, int R, int S, int infinity,
                                int inf_start, float* g_sigma, float* g_tint, float* g_diffuse, float* g_specular,
                                int gs_sigma, int gs_tint, int gs_diffuse, int gs_specular,
                                float* grad_rays_d, void* stream)
{
    return snrf_composite_bwd_ert(sigma, tint, diffuse, specular, s_sigma, s_tint, s_diffuse, s_specular, z_vals, dists, rays_d, trans,
                                  g_out, g_weights, ray_valid, R, S, infinity, inf_start, g_sigma, g_tint, g_diffuse, g_specular,
                                  gs_sigma, gs_tint, gs_diffuse, gs_specular, grad_rays_d, nullptr, stream);
}
// sample_live (may be NULL): the flags snrf_composite_fwd_ert wrote -- dead samples get zero gradients
SNRF_API int snrf_composite_bwd_ert(const float* sigma, const float* tint, const float* diffuse, const float* specular,
                                    int s_sigma, int s_tint, int s_diffuse, int s_specular,
                                    const float* z_vals, const float* dists, const float* rays_d, const float* trans,
                                    const float* g_out, const float* g_weights, const unsigned char* ray_valid, int R, int S, int infinity,
                                    int inf_start, float* g_sigma, float* g_tint, float* g_diffuse, float* g_specular,
                                    int gs_sigma, int gs_tint, int gs_diffuse, int gs_specular,
                                    float* grad_rays_d, const unsigned char* sample_live, void* stream)
{
    SNRF_CHECK_ARG(S > 0, "snrf_composite_bwd: S must be positive");
    if (R <= 0) return 0;
    Heads in{sigma, tint, diffuse, specular, s_sigma, s_tint, s_diffuse, s_specular};
    HeadGrads g{g_sigma, g_tint, g_diffuse, g_specular, gs_sigma, gs_tint, gs_diffuse, gs_specular};
    const bool packed = s_sigma == 10 && s_tint == 10 && s_diffuse == 10 && s_specular == 10 && tint == sigma + 1 && diffuse == sigma + 4 &&
                        specular == sigma + 7 && ((uintptr_t)sigma & 7) == 0 &&
                        gs_sigma == 10 && gs_tint == 10 && gs_diffuse == 10 && gs_specular == 10 && g_tint == g_sigma + 1 &&
                        g_diffuse == g_sigma + 4 && g_specular == g_sigma + 7 && ((uintptr_t)g_sigma & 7) == 0;
    if (packed) composite_bwd_kernel<true><<<grid_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(in, z_vals, dists, rays_d, trans, g_out, g_weights, ray_valid, R, S, infinity, inf_start, g, grad_rays_d, sample_live);
    else composite_bwd_kernel<false><<<grid_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(in, z_vals, dists, rays_d, trans, g_out, g_weights, ray_valid, R, S, infinity, inf_start, g, grad_rays_d, sample_live);
    SNRF_RETURN_LAUNCH("snrf_composite_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
// Colour loss of the joint foreground + background batch, forward and gradient in one pass.
//
// Replaces (behaviour, not code) the torch chain between the compositing rows and loss.backward() in the training step:
// TILE.render_rays' merge of the two chains (tile.py:661-681), the clamp of HashGrid.render_batch_rays (hashgrid/__init__.py:
// 589), the masked MSE (criterions.py:126-147) and the specular L2 regulariser (tile.py:999) -- ~30 small launches forward and
// ~30 backward per step.  row [2R,16] = the rows of snrf_composite_fwd, rays [0,R) the foreground chains, [R,2R) the background
// chains of the same pixels:
//   rgb_f = clamp(dif_f + spec_f, 0, 1), rgb_b likewise;  pred = rgb_f + T_left_f * rgb_b
//   loss  = sum_{valid} |pred - gt|^2 / (3 n_valid) + l2_weight * (sum row_f[10:13] / (3 n_f) + sum row_b[10:13] / (3 n_b))
// with valid = valid_f | valid_b, n_* = max(count, 1).  g_row [2R,16] = d loss / d row (WRITTEN, every entry).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256)
joint_loss_count_kernel(const unsigned char* __restrict__ vf, const unsigned char* __restrict__ vb, int R, int* __restrict__ counts)
{
    int cf = 0, cb = 0, cv = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        const int f = vf[r] != 0, b = vb[r] != 0;
        cf += f; cb += b; cv += f | b;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        cf += __shfl_xor_sync(0xffffffffu, cf, off); cb += __shfl_xor_sync(0xffffffffu, cb, off); cv += __shfl_xor_sync(0xffffffffu, cv, off);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(counts, cf); atomicAdd(counts + 1, cb); atomicAdd(counts + 2, cv); }
}

__global__ void __launch_bounds__(256)
joint_loss_kernel(const float* __restrict__ row, const unsigned char* __restrict__ vf, const unsigned char* __restrict__ vb,
                  const float* __restrict__ gt, float l2_weight, int R, const int* __restrict__ counts, float* __restrict__ loss,
                  float* __restrict__ g_row, float* __restrict__ pred_out)
{
    const float nf = (float)max(counts[0], 1), nb = (float)max(counts[1], 1), nv = (float)max(counts[2], 1);
    const float inv_mse = 1.0f / (3.0f * nv), l2f = l2_weight / (3.0f * nf), l2b = l2_weight / (3.0f * nb);
    float acc = 0.0f;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        const float4* pf = reinterpret_cast<const float4*>(row + (size_t)r * kOutStride);
        const float4* pb = reinterpret_cast<const float4*>(row + (size_t)(R + r) * kOutStride);
        const float4 f1 = pf[1], f2 = pf[2], f3v = pf[3], b1 = pb[1], b2 = pb[2], b3v = pb[3];
        // row: [0] depth, [1..3] tint, [4..6] diffuse, [7..9] specular, [10..12] l2, [13] T_left
        const float xf[3] = {f1.x + f1.w, f1.y + f2.x, f1.z + f2.y};
        const float xb[3] = {b1.x + b1.w, b1.y + b2.x, b1.z + b2.y};
        const float T = f3v.y;
        const bool valid = (vf[r] | vb[r]) != 0;
        float gp[3], gT = 0.0f, pf_pass[3], pb_pass[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float rf = fminf(fmaxf(xf[c], 0.0f), 1.0f), rb = fminf(fmaxf(xb[c], 0.0f), 1.0f);
            const float pred = rf + T * rb;
            const float diff = pred - gt[3 * (size_t)r + c];
            if (pred_out) pred_out[3 * (size_t)r + c] = pred;
            if (valid) acc += diff * diff * inv_mse;
            gp[c] = valid ? 2.0f * diff * inv_mse : 0.0f;
            pf_pass[c] = (xf[c] >= 0.0f && xf[c] <= 1.0f) ? gp[c] : 0.0f;           // torch.clamp: the gradient passes on [min, max]
            pb_pass[c] = (xb[c] >= 0.0f && xb[c] <= 1.0f) ? gp[c] * T : 0.0f;
            gT += gp[c] * rb;
        }
        acc += l2f * (f2.z + f2.w + f3v.x) + l2b * (b2.z + b2.w + b3v.x);
        float4* gf = reinterpret_cast<float4*>(g_row + (size_t)r * kOutStride);
        float4* gb = reinterpret_cast<float4*>(g_row + (size_t)(R + r) * kOutStride);
        gf[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        gf[1] = make_float4(pf_pass[0], pf_pass[1], pf_pass[2], pf_pass[0]);
        gf[2] = make_float4(pf_pass[1], pf_pass[2], l2f, l2f);
        gf[3] = make_float4(l2f, gT, 0.f, 0.f);
        gb[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        gb[1] = make_float4(pb_pass[0], pb_pass[1], pb_pass[2], pb_pass[0]);
        gb[2] = make_float4(pb_pass[1], pb_pass[2], l2b, l2b);
        gb[3] = make_float4(l2b, 0.f, 0.f, 0.f);
    }
    acc = warp_sum(acc);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
        atomicAdd(loss, t);
    }
}

}  // namespace

// loss: one float, ACCUMULATED (zero it first); counts_scratch: 3 ints of device scratch (cleared here); pred_color [R,3] optional
SNRF_API int snrf_joint_loss(const float* row, const unsigned char* valid_f, const unsigned char* valid_b, const float* gt,
                             float l2_weight, int R, float* loss, float* g_row, float* pred_color, int* counts_scratch, void* stream)
{
    SNRF_CHECK_ARG(row && valid_f && valid_b && gt && loss && g_row && counts_scratch, "snrf_joint_loss: NULL argument");
    if (R <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(counts_scratch, 0, 3 * sizeof(int), s);
    int gx = snrf_div_up(R, 256);
    const int cap = snrf_sm_count() * 4;
    if (gx > cap) gx = cap;
    joint_loss_count_kernel<<<gx, 256, 0, s>>>(valid_f, valid_b, R, counts_scratch);
    joint_loss_kernel<<<gx, 256, 0, s>>>(row, valid_f, valid_b, gt, l2_weight, R, counts_scratch, loss, g_row, pred_color);
    SNRF_RETURN_LAUNCH("snrf_joint_loss");
}

// Sparse Adam step over the entries that received gradient, sm_100a.
//
// Replaces (behaviour, not code) the reference operators
//   adam_step_cuda / adam_step_cuda_fp16      cuda/adam_kernel.cu:72-94, 147-168
// whose semantics are: an element whose gradient is exactly 0 is skipped (its moments do
// not decay); otherwise
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
//   p -= (lr / (1-b1^t)) * m / (sqrt(v / (1-b2^t)) + eps)
// The fp16 variant keeps m, v in half scaled by 128 / 128^2 (LOSS_SCALE).
//
// Design (B200): pure HBM streaming, 28 B per touched float (p,g,m,v read; p,m,v written)
// and 4 B per untouched one.  One thread owns 4 consecutive floats (128-bit accesses when
// the row layout allows), the two bias corrections are evaluated once per CTA instead of
// two powf per element, and the gradient can be cleared in the same pass (`zero_grad`) so
// the next backward needs no 2 GiB memset of the hash-table gradient.  The reference
// addresses element (k, d) at k*8 + d whatever D is (cuda/adam_kernel.cu:43); the C ABI
// takes the row stride explicitly so both that layout and dense [.., D] tensors work.
#include "adam_core.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int kThreads = 256;
constexpr float kLossScale = 128.0f;

using adamcore::Hyper;
using adamcore::adam_elem;

// dense, contiguous: n floats, n % 4 == 0 handled by the vector body + scalar tail
__global__ void __launch_bounds__(kThreads)
adam_dense_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n, Hyper h, int zero_grad)
{
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
        s_bc[0] = 1.0f - powf(h.b1, (float)h.step);
        s_bc[1] = 1.0f - powf(h.b2, (float)h.step);
    }
    __syncthreads();
    const float bc1 = s_bc[0], bc2 = s_bc[1];
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // kUnroll float4 groups per thread and trip, in two phases: all gradient loads first, then the (conditional) loads of
    // the parameter and the two moments of every touched group, so that independent loads overlap instead of forming a
    // dependent pair.  Measured on B200 at C2 (2 GiB table, ~60 % of the groups touched): 1 group 2.35 ms, 2 groups
    // 2.14 ms, 4 groups 2.62 ms (112 registers: the occupancy lost costs more than the overlap gains); streaming
    // cache hints (ld.cs / st.cs) on top made it slower still.
    constexpr int kUnroll = 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long base = (long long)blockIdx.x * blockDim.x * kUnroll + threadIdx.x; base < n4; base += stride * kUnroll) {
        float4 gg[kUnroll], pp[kUnroll], mm[kUnroll], vv[kUnroll];
        bool act[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = base + (long long)u * blockDim.x;
            gg[u] = i < n4 ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = base + (long long)u * blockDim.x;
            act[u] = !(gg[u].x == 0.0f && gg[u].y == 0.0f && gg[u].z == 0.0f && gg[u].w == 0.0f);
            if (act[u]) { pp[u] = p4[i]; mm[u] = m4[i]; vv[u] = v4[i]; }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!act[u]) continue;
            const long long i = base + (long long)u * blockDim.x;
            if (gg[u].x != 0.0f) adam_elem(pp[u].x, gg[u].x, mm[u].x, vv[u].x, h, bc1, bc2);
            if (gg[u].y != 0.0f) adam_elem(pp[u].y, gg[u].y, mm[u].y, vv[u].y, h, bc1, bc2);
            if (gg[u].z != 0.0f) adam_elem(pp[u].z, gg[u].z, mm[u].z, vv[u].z, h, bc1, bc2);
            if (gg[u].w != 0.0f) adam_elem(pp[u].w, gg[u].w, mm[u].w, vv[u].w, h, bc1, bc2);
            p4[i] = pp[u];
            m4[i] = mm[u];
            v4[i] = vv[u];
            if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // tail
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i];
        if (gi == 0.0f) continue;
        float pi = p[i], mi = m[i], vi = v[i];
        adam_elem(pi, gi, mi, vi, h, bc1, bc2);
        p[i] = pi; m[i] = mi; v[i] = vi;
        if (zero_grad) g[i] = 0.0f;
    }
}

// rows x dim elements at k*row_stride + d (the reference's K x 8 addressing)
template <bool HALF_STATE>
__global__ void __launch_bounds__(kThreads)
adam_strided_kernel(float* __restrict__ p, float* __restrict__ g, void* __restrict__ m_v, void* __restrict__ v_v,
                    long long rows, int dim, int row_stride, Hyper h, int zero_grad)
{
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
        s_bc[0] = 1.0f - powf(h.b1, (float)h.step);
        s_bc[1] = 1.0f - powf(h.b2, (float)h.step);
    }
    __syncthreads();
    const float bc1 = s_bc[0], bc2 = s_bc[1];
    const long long n = rows * dim;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long k = t / dim;
        const int d = (int)(t - k * dim);
        const long long i = k * row_stride + d;
        if (!HALF_STATE) {
            const float gi = g[i];
            if (gi == 0.0f) continue;
            float* m = (float*)m_v; float* v = (float*)v_v;
            float pi = p[i], mi = m[i], vi = v[i];
            adam_elem(pi, gi, mi, vi, h, bc1, bc2);
            p[i] = pi; m[i] = mi; v[i] = vi;
        } else {
            // cuda/adam_kernel.cu:97-144: gradient pre-scaled by 128, state stored in half
            const float gi = g[i] * kLossScale;
            if (gi == 0.0f) continue;
            __half* m = (__half*)m_v; __half* v = (__half*)v_v;
            const float mi = h.b1 * __half2float(m[i]) + (1.0f - h.b1) * gi;
            const float vi = h.b2 * __half2float(v[i]) + (1.0f - h.b2) * gi * gi;
            const float denom = sqrtf(vi / (bc2 * kLossScale * kLossScale)) + h.eps;
            const float step_size = h.lr / bc1;
            p[i] = p[i] - step_size * mi / (denom * kLossScale);
            m[i] = __float2half(mi);
            v[i] = __float2half(vi);
        }
        if (zero_grad) g[i] = 0.0f;
    }
}

inline int grid_for(long long work)
{
    long long gx = (work + kThreads - 1) / kThreads;
    const long long cap = (long long)snrf_sm_count() * 16;
    if (gx > cap) gx = cap;
    return gx > 0 ? (int)gx : 1;
}

}  // namespace

// ------------------------------- C ABI --------------------------------------
SNRF_API int snrf_adam_step(float* params, float* grads, void* exp_avg, void* exp_avg_sq,
                            long long rows, int dim, int row_stride, int half_state,
                            float lr, float beta1, float beta2, float eps, int step, int zero_grad,
                            void* stream)
{
    SNRF_CHECK_ARG(rows >= 0 && dim > 0 && row_stride >= dim, "snrf_adam_step: need rows>=0, 0<dim<=row_stride (rows=%lld dim=%d stride=%d)", rows, dim, row_stride);
    SNRF_CHECK_ARG(step >= 1, "snrf_adam_step: step counts from 1 (got %d)", step);
    if (rows == 0) return 0;
    const Hyper h{lr, beta1, beta2, eps, step};
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = rows * dim;
    const bool aligned = ((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)exp_avg) | ((uintptr_t)exp_avg_sq)) & 15) == 0;
    if (!half_state && row_stride == dim && aligned) {
        adam_dense_kernel<<<grid_for((n + 3) / 4), kThreads, 0, s>>>(params, grads, (float*)exp_avg, (float*)exp_avg_sq, n, h, zero_grad);
    } else if (half_state) {
        adam_strided_kernel<true><<<grid_for(n), kThreads, 0, s>>>(params, grads, exp_avg, exp_avg_sq, rows, dim, row_stride, h, zero_grad);
    } else {
        adam_strided_kernel<false><<<grid_for(n), kThreads, 0, s>>>(params, grads, exp_avg, exp_avg_sq, rows, dim, row_stride, h, zero_grad);
    }
    SNRF_RETURN_LAUNCH("snrf_adam_step");
}

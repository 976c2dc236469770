// Bundle-adjustment pose chain: se(3) refinement -> SE(3) -> compose with the base world-to-camera
// pose -> invert to camera-to-world, forward and analytic backward, sm_100a.
//
// Replaces (behaviour, not code) this torch chain, which the reference runs as ~170 tiny kernels
// forward and ~350 through autograd for a [N_cam, 6] tensor:
//   CAM.get_rts            camera_utils.py:86-89   rts = compose([se3_to_SE3(se3_refine), base_rts])
//   Lie.se3_to_SE3         camera.py:84-95         R = I + A wx + B wx^2, t = (I + B wx + C wx^2) u
//   Lie.taylor_A/B/C       camera.py:118-141       11-term Taylor series of sin x / x, (1-cos x)/x^2, (x-sin x)/x^3
//   Pose.compose_pair      camera.py:53-60         R_new = R_b R_a, t_new = R_b t_a + t_b
//   Pose.invert            camera.py:37-43         c2w = [R^T | -R^T t]
// One thread per camera.  The backward evaluates the same function on dual numbers carrying the six
// partial derivatives (forward-mode AD) and contracts them with the incoming d L / d c2w -- exact
// derivatives of the very polynomial the forward evaluates, no singularity at w = 0 (the series are
// polynomials in theta^2).
#include "common.cuh"

namespace {

struct Dual {
    float v, d[6];
};
__device__ __forceinline__ Dual mkd(float v) { Dual r; r.v = v; for (int i = 0; i < 6; ++i) r.d[i] = 0.f; return r; }
__device__ __forceinline__ Dual seed(float v, int k) { Dual r = mkd(v); r.d[k] = 1.f; return r; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { Dual r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ Dual operator-(Dual a) { Dual r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { Dual r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ Dual operator*(Dual a, float s) { Dual r; r.v = a.v * s; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * s; return r; }
__device__ __forceinline__ Dual operator+(Dual a, float s) { a.v += s; return a; }
__device__ __forceinline__ float mkval(float v, float) { return v; }
__device__ __forceinline__ Dual mkval(float v, Dual) { return mkd(v); }

// Series in s = theta^2: sum_i (-1)^i s^i / denom_i with the reference's denominators
// (kind 0: sin x / x, 1: (1 - cos x) / x^2, 2: (x - sin x) / x^3), 11 terms, summed in the reference's order.
template <class T>
__device__ __forceinline__ T taylor(T s, int kind)
{
    T ans = mkval(0.0f, s), p = mkval(1.0f, s);
    double denom = 1.0;                       // python float arithmetic in the reference
    float sign = 1.0f;
    for (int i = 0; i <= 10; ++i) {
        if (kind == 0) { if (i > 0) denom *= (double)((2 * i) * (2 * i + 1)); }
        else if (kind == 1) denom *= (double)((2 * i + 1) * (2 * i + 2));
        else denom *= (double)((2 * i + 2) * (2 * i + 3));
        ans = ans + p * (sign / (float)denom);
        p = p * s;
        sign = -sign;
    }
    return ans;
}

// wu[6] (rotation generators, translation generators), base = world->camera [3x4] -> c2w [3x4] (12 values)
template <class T>
__device__ __forceinline__ void pose_chain(const T* wu, const float* base, T* c2w)
{
    const T w0 = wu[0], w1 = wu[1], w2 = wu[2];
    const T s = w0 * w0 + w1 * w1 + w2 * w2;
    const T A = taylor(s, 0), B = taylor(s, 1), C = taylor(s, 2);
    const T zero = mkval(0.0f, s);
    // wx and wx^2
    const T X[9] = {zero, -w2, w1, w2, zero, -w0, -w1, w0, zero};
    T X2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) X2[3 * i + j] = X[3 * i] * X[j] + X[3 * i + 1] * X[3 + j] + X[3 * i + 2] * X[6 + j];
    T Ra[9], V[9];
    for (int k = 0; k < 9; ++k) {
        const float eye = (k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f;
        Ra[k] = A * X[k] + B * X2[k] + eye;
        V[k] = B * X[k] + C * X2[k] + eye;
    }
    T ta[3];
    for (int i = 0; i < 3; ++i) ta[i] = V[3 * i] * wu[3] + V[3 * i + 1] * wu[4] + V[3 * i + 2] * wu[5];
    // compose: R = R_b R_a, t = R_b t_a + t_b  (b = base)
    T R[9], t[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[3 * i + j] = Ra[j] * base[4 * i] + Ra[3 + j] * base[4 * i + 1] + Ra[6 + j] * base[4 * i + 2];
        t[i] = ta[0] * base[4 * i] + ta[1] * base[4 * i + 1] + ta[2] * base[4 * i + 2] + base[4 * i + 3];
    }
    // invert: [R^T | -R^T t]
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c2w[4 * i + j] = R[3 * j + i];
        c2w[4 * i + 3] = -(R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);
    }
}

__global__ void pose_fwd_kernel(const float* __restrict__ se3, const float* __restrict__ base, float* __restrict__ c2w, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float wu[6], b[12], out[12];
    for (int k = 0; k < 6; ++k) wu[k] = se3[6 * i + k];
    for (int k = 0; k < 12; ++k) b[k] = base[12 * i + k];
    pose_chain<float>(wu, b, out);
    for (int k = 0; k < 12; ++k) c2w[12 * i + k] = out[k];
}

__global__ void pose_bwd_kernel(const float* __restrict__ se3, const float* __restrict__ base, const float* __restrict__ grad_c2w,
                                float* __restrict__ grad_se3, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Dual wu[6], out[12];
    float b[12];
    for (int k = 0; k < 6; ++k) wu[k] = seed(se3[6 * i + k], k);
    for (int k = 0; k < 12; ++k) b[k] = base[12 * i + k];
    pose_chain<Dual>(wu, b, out);
    float g[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 12; ++k) {
        const float gk = grad_c2w[12 * i + k];
        for (int j = 0; j < 6; ++j) g[j] += gk * out[k].d[j];
    }
    for (int j = 0; j < 6; ++j) grad_se3[6 * i + j] = g[j];
}

}  // namespace

// se3_refine [N,6], base_w2c [N,12] (row-major 3x4) -> c2w [N,12]
SNRF_API int snrf_pose_fwd(const float* se3_refine, const float* base_w2c, float* c2w, int n, void* stream)
{
    if (n <= 0) return 0;
    pose_fwd_kernel<<<snrf_div_up(n, 64), 64, 0, (cudaStream_t)stream>>>(se3_refine, base_w2c, c2w, n);
    SNRF_RETURN_LAUNCH("snrf_pose_fwd");
}

// grad_c2w [N,12] -> grad_se3 [N,6] (WRITTEN)
SNRF_API int snrf_pose_bwd(const float* se3_refine, const float* base_w2c, const float* grad_c2w, float* grad_se3, int n, void* stream)
{
    if (n <= 0) return 0;
    pose_bwd_kernel<<<snrf_div_up(n, 64), 64, 0, (cudaStream_t)stream>>>(se3_refine, base_w2c, grad_c2w, grad_se3, n);
    SNRF_RETURN_LAUNCH("snrf_pose_bwd");
}

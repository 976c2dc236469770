// The Adam element update shared by adam.cu (snrf_adam_step) and field_encode.cu (the scatter + update fusion).
#pragma once
#include "common.cuh"

namespace adamcore {

struct Hyper { float lr, b1, b2, eps; int step; };

// cuda/adam_kernel.cu:43-69 semantics (one element with a non-zero gradient)
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const Hyper& h, float bc1, float bc2)
{
    const float mi = h.b1 * m + (1.0f - h.b1) * g;
    const float vi = h.b2 * v + (1.0f - h.b2) * g * g;
    const float denom = sqrtf(vi / bc2) + h.eps;
    const float step_size = h.lr / bc1;
    p = p - step_size * mi / denom;
    m = mi; v = vi;
}

}  // namespace adamcore

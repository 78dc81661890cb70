// k_linsolve_rb.cuh -- red-black Gauss-Seidel fast path for lin_solve
// (fluid.rs:301-325 with the sweep order changed): each iteration updates the
// cells with (i+j) even, then the cells with (i+j) odd, then set_boundaries.
// Same formula, same iteration count; results are tolerance-checked against the
// oracle (tests/test_red_black.py), not bit-compared.
#pragma once
#include "eq_common.cuh"

// v1: one launch per colour.  Thread t of row j owns cell i = 1 + 2t + off.
__global__ void __launch_bounds__(256) k_rb_half(float *__restrict__ x, const float *__restrict__ x0, float a,
                                                 float c_recip, int colour, EqLayout L) {
    const int j = blockIdx.y + max(L.row0, 1);   // owned interior rows
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = 1 + 2 * t + ((colour ^ (j + 1)) & 1);
    if (i > L.N - 2) return;
    const size_t o = (size_t)i + (size_t)j * L.P;
    x[o] = gs_update(x0[o], x[o + 1], x[o - 1], x[o + L.P], x[o - L.P], a, c_recip);
}

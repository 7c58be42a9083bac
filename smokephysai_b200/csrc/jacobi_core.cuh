// jacobi_core.cuh -- the register-strip Jacobi sweep shared by k_jacobi (jacobi.cu) and the fused whole-step
// kernel (fused.cu).  See jacobi.cu for the design notes.
#pragma once
#include "common.cuh"

namespace smk {

// One row of four cells: m * ((((up + dn) + left) + right) - div), left/right neighbours of the strip by shuffle.
__device__ __forceinline__ float4 stencil_row(const float4 up, const float4 cur, const float4 dn, const float4 d,
                                              const float m0, const float m1, const float m2, const float m3)
{
    const float left = __shfl_up_sync(0xffffffffu, cur.w, 1);
    const float right = __shfl_down_sync(0xffffffffu, cur.x, 1);
    float4 nw;
    nw.x = m0 * ((((up.x + dn.x) + left) + cur.y) - d.x);
    nw.y = m1 * ((((up.y + dn.y) + cur.x) + cur.z) - d.y);
    nw.z = m2 * ((((up.z + dn.z) + cur.y) + cur.w) - d.z);
    nw.w = m3 * ((((up.w + dn.w) + cur.z) + right) - d.w);
    return nw;
}

// One sweep of a warp's R x 128 strip, boundary rows first: the new first and last rows are what the
// neighbouring warps need for the NEXT sweep, so they are computed and posted to shared memory before the
// R-2 interior rows; the CTA barrier that publishes them then overlaps with the interior arithmetic.
template <int R, bool FAST>
__device__ __forceinline__ void sweep_rows(float4 (&P)[R], const float4 (&D)[R], const float4 uph, const float4 dnh,
                                           const float cm0, const float cm1, const float cm2, const float cm3,
                                           const int gi0, const int h, float4* post_first, float4* post_last)
{
    auto rowok = [&](int r) { const int gi = gi0 + r; return FAST || (gi >= 1 && gi <= h - 2); };
    const float4 o0 = P[0], oL = P[R - 1];
    float4 n0, nL;
    {
        const bool ok = rowok(0);
        n0 = stencil_row(uph, o0, R > 1 ? P[1] : dnh, D[0], ok ? cm0 : 0.f, ok ? cm1 : 0.f, ok ? cm2 : 0.f, ok ? cm3 : 0.f);
    }
    if (R > 1) {
        const bool ok = rowok(R - 1);
        nL = stencil_row(P[R - 2], oL, dnh, D[R - 1], ok ? cm0 : 0.f, ok ? cm1 : 0.f, ok ? cm2 : 0.f, ok ? cm3 : 0.f);
    } else nL = n0;
    if (post_first) { *post_first = n0; *post_last = nL; }
    float4 up = o0;
#pragma unroll
    for (int r = 1; r < R - 1; ++r) {
        const float4 cur = P[r];
        const float4 dn = (r < R - 2) ? P[r + 1] : oL;
        const bool ok = rowok(r);
        P[r] = stencil_row(up, cur, dn, D[r], ok ? cm0 : 0.f, ok ? cm1 : 0.f, ok ? cm2 : 0.f, ok ? cm3 : 0.f);
        up = cur;
    }
    P[0] = n0;
    if (R > 1) P[R - 1] = nL;
}

}  // namespace smk

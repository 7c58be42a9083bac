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

// ---- packed variant ------------------------------------------------------------------------------------
// sm_100 has add.f32x2 / mul.f32x2 (__fadd2_rn / __fmul2_rn): two IEEE-rounded fp32 results per instruction,
// i.e. half the issue slots for the same arithmetic -- and the Jacobi sweep is bound by FP32 issue, not memory.
// A float2 pairs row rr of the thread's 8 x 4 strip with row rr + 4, so the up / down neighbours of a pair are
// another pair (rows rr-1, rr+3 and rr+1, rr+5); only the first add of the two seam row pairs (rows 0|4 and
// 3|7) is done with scalar adds.  x - div is x + (-div): the divergence is held negated.  A sweep reads one
// register set and writes another (ping-pong), so no copies are needed to keep the old values alive.
struct PackedStrip {
    float2 q[4][4];            // q[rr][c] = (p[rr][c], p[rr + 4][c])
};

__device__ __forceinline__ float4 packed_row(const PackedStrip& S, const int r)      // r is a compile-time constant after unrolling
{
    const int rr = r & 3;
    return (r < 4) ? make_float4(S.q[rr][0].x, S.q[rr][1].x, S.q[rr][2].x, S.q[rr][3].x)
                   : make_float4(S.q[rr][0].y, S.q[rr][1].y, S.q[rr][2].y, S.q[rr][3].y);
}
__device__ __forceinline__ void packed_set_row(PackedStrip& S, const int r, const float4 v)
{
    const int rr = r & 3;
    if (r < 4) { S.q[rr][0].x = v.x; S.q[rr][1].x = v.y; S.q[rr][2].x = v.z; S.q[rr][3].x = v.w; }
    else       { S.q[rr][0].y = v.x; S.q[rr][1].y = v.y; S.q[rr][2].y = v.z; S.q[rr][3].y = v.w; }
}

// one row pair of a sweep: Dst.q[RR] from the old strip S.  PACKED selects f32x2 arithmetic for this pair; with
// PACKED false the same operations are issued as scalar FADD / FMUL (the two forms run on different issue /
// pipe resources, so a mix of packed and scalar row pairs is faster than either alone: PMASK in sweep_packed).
template <int RR, bool PACKED, int DBG = 0>
__device__ __forceinline__ void packed_pair(const PackedStrip& S, PackedStrip& Dst, const PackedStrip& ND,
                                            const float4 uph, const float4 dnh, const float2 (&M)[4])
{
    const float uh[4] = {uph.x, uph.y, uph.z, uph.w}, dh[4] = {dnh.x, dnh.y, dnh.z, dnh.w};
    float2 left, right;
    if (DBG & 1) { left = S.q[RR][3]; right = S.q[RR][0]; }          // probe builds only (tools/micro/jacobi_probe.cu)
    else {
        left.x = __shfl_up_sync(0xffffffffu, S.q[RR][3].x, 1);
        left.y = __shfl_up_sync(0xffffffffu, S.q[RR][3].y, 1);
        right.x = __shfl_down_sync(0xffffffffu, S.q[RR][0].x, 1);
        right.y = __shfl_down_sync(0xffffffffu, S.q[RR][0].y, 1);
    }
    if (DBG & 8) {
        Dst.q[RR][0] = left; Dst.q[RR][3] = right; Dst.q[RR][1] = make_float2(uh[1], dh[1]); Dst.q[RR][2] = make_float2(uh[2], dh[2]);
        return;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2 t;
        if (RR == 0)      { t.x = uh[c] + S.q[1][c].x;        t.y = S.q[3][c].x + S.q[1][c].y; }     // rows 0 | 4: up = halo | row 3
        else if (RR == 3) { t.x = S.q[2][c].x + S.q[0][c].y;  t.y = S.q[2][c].y + dh[c]; }           // rows 3 | 7: down = row 4 | halo
        else if (PACKED) t = __fadd2_rn(S.q[RR - 1][c], S.q[RR + 1][c]);
        else { t.x = S.q[RR - 1][c].x + S.q[RR + 1][c].x; t.y = S.q[RR - 1][c].y + S.q[RR + 1][c].y; }
        const float2 l = (c == 0) ? left : S.q[RR][c - 1];
        const float2 r = (c == 3) ? right : S.q[RR][c + 1];
        if (PACKED) {
            t = __fadd2_rn(t, l);
            t = __fadd2_rn(t, r);
            t = __fadd2_rn(t, ND.q[RR][c]);
            Dst.q[RR][c] = __fmul2_rn(M[c], t);
        } else {
            t.x = t.x + l.x; t.y = t.y + l.y;
            t.x = t.x + r.x; t.y = t.y + r.y;
            t.x = t.x + ND.q[RR][c].x; t.y = t.y + ND.q[RR][c].y;
            Dst.q[RR][c] = make_float2(M[c].x * t.x, M[c].y * t.y);
        }
    }
}

// rows of the strip that are on the Dirichlet ring or outside the grid: bit r set -> row r is forced to 0
__device__ __forceinline__ void packed_zero_rows(PackedStrip& S, const unsigned ringmask)
{
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (ringmask & (1u << r)) packed_set_row(S, r, make_float4(0.f, 0.f, 0.f, 0.f));
}

// One sweep S -> Dst, boundary rows (0 and 7) first: they are posted to shared memory for the neighbouring
// warps before the interior row pairs are computed (ORDER 0), or one of the other orders below.
__device__ __forceinline__ void mbar_arrive_cta(unsigned long long* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}

// posted (optional, MBAR builds): an mbarrier every lane arrives on once its two boundary rows are in shared memory
template <int PMASK, int DBG = 0, bool MBAR = false, int ORDER = 0>
__device__ __forceinline__ void sweep_packed(const PackedStrip& S, PackedStrip& Dst, const PackedStrip& ND,
                                             const float4 uph, const float4 dnh, const float2 (&M)[4], const unsigned ringmask,
                                             float4* post_first, float4* post_last, unsigned long long* posted = nullptr)
{
    // ORDER 0: boundary pairs, post, interior pairs;  1: interior pairs, boundary pairs, post;  2: pair 1, boundary pairs,
    // post, pair 2 (the halo loads complete under pair 1, the posts under pair 2)
    if (ORDER == 1 || ORDER == 2) packed_pair<1, (PMASK & 2) != 0, DBG>(S, Dst, ND, uph, dnh, M);
    if (ORDER == 1) packed_pair<2, (PMASK & 4) != 0, DBG>(S, Dst, ND, uph, dnh, M);
    packed_pair<0, (PMASK & 1) != 0, DBG>(S, Dst, ND, uph, dnh, M);
    packed_pair<3, (PMASK & 8) != 0, DBG>(S, Dst, ND, uph, dnh, M);
    if (ringmask & 0x81u) {
        if (ringmask & 1u) packed_set_row(Dst, 0, make_float4(0.f, 0.f, 0.f, 0.f));
        if (ringmask & 0x80u) packed_set_row(Dst, 7, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (post_first) { *post_first = packed_row(Dst, 0); *post_last = packed_row(Dst, 7); }
    if (MBAR) mbar_arrive_cta(posted);
    if (ORDER == 0) packed_pair<1, (PMASK & 2) != 0, DBG>(S, Dst, ND, uph, dnh, M);
    if (ORDER == 0 || ORDER == 2) packed_pair<2, (PMASK & 4) != 0, DBG>(S, Dst, ND, uph, dnh, M);
    if (ringmask & 0x7eu) packed_zero_rows(Dst, ringmask & 0x7eu);
}

}  // namespace smk

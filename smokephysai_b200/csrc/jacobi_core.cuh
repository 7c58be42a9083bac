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
// The element type E of a strip is float2, or a 64-bit register holding the same pair (PackedStrip64): with float2 the
// register allocator is free to keep the two halves anywhere and builds an aligned pair in front of every f32x2
// instruction when it decides on the row-major (float4) layout of the loads and stores -- k_jacobi_stream compiled to
// 52 MOVs per sweep and warp that way (profiles/r01j_*); a 64-bit element pins the pair, and taking a half out of it
// (mov.b64 {lo, hi}) costs no instruction.
template <class E>
struct PackedStripT {
    E q[4][4];                 // q[rr][c] = (p[rr][c], p[rr + 4][c])
};
using PackedStrip = PackedStripT<float2>;
using PackedStrip64 = PackedStripT<unsigned long long>;

__device__ __forceinline__ float2 pe_get(const float2 e) { return e; }
__device__ __forceinline__ float2 pe_get(const unsigned long long e)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(e));
    return r;
}
__device__ __forceinline__ void pe_set(float2& e, const float2 v) { e = v; }
__device__ __forceinline__ void pe_set(unsigned long long& e, const float2 v) { asm("mov.b64 %0, {%1, %2};" : "=l"(e) : "f"(v.x), "f"(v.y)); }
__device__ __forceinline__ float2 pe_add(const float2 a, const float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ unsigned long long pe_add(const unsigned long long a, const unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float2 pe_mul(const float2 a, const float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ unsigned long long pe_mul(const unsigned long long a, const unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
template <class E> __device__ __forceinline__ E pe_make(const float2 v) { E e; pe_set(e, v); return e; }

template <class E>
__device__ __forceinline__ float4 packed_row(const PackedStripT<E>& S, const int r)      // r is a compile-time constant after unrolling
{
    const int rr = r & 3;
    const float2 a = pe_get(S.q[rr][0]), b = pe_get(S.q[rr][1]), c = pe_get(S.q[rr][2]), d = pe_get(S.q[rr][3]);
    return (r < 4) ? make_float4(a.x, b.x, c.x, d.x) : make_float4(a.y, b.y, c.y, d.y);
}
template <class E>
__device__ __forceinline__ void packed_set_row(PackedStripT<E>& S, const int r, const float4 v)
{
    const int rr = r & 3;
    const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2 e = pe_get(S.q[rr][c]);
        if (r < 4) e.x = in[c]; else e.y = in[c];
        pe_set(S.q[rr][c], e);
    }
}

// rows rr and rr + 4 at once (no read of the old element: usable on an uninitialised strip)
template <class E>
__device__ __forceinline__ void packed_set_rows(PackedStripT<E>& S, const int rr, const float4 lo, const float4 hi)
{
    pe_set(S.q[rr][0], make_float2(lo.x, hi.x)); pe_set(S.q[rr][1], make_float2(lo.y, hi.y));
    pe_set(S.q[rr][2], make_float2(lo.z, hi.z)); pe_set(S.q[rr][3], make_float2(lo.w, hi.w));
}

// one row pair of a sweep: Dst.q[RR] from the old strip S.  PACKED selects f32x2 arithmetic for this pair; with
// PACKED false the same operations are issued as scalar FADD / FMUL (the two forms run on different issue /
// pipe resources, so a mix of packed and scalar row pairs is faster than either alone: PMASK in sweep_packed).
template <int RR, bool PACKED, int DBG = 0, class E>
__device__ __forceinline__ void packed_pair(const PackedStripT<E>& S, PackedStripT<E>& Dst, const PackedStripT<E>& ND,
                                            const float4 uph, const float4 dnh, const E (&M)[4])
{
    const float uh[4] = {uph.x, uph.y, uph.z, uph.w}, dh[4] = {dnh.x, dnh.y, dnh.z, dnh.w};
    float2 left, right;
    if (DBG & 1) { left = pe_get(S.q[RR][3]); right = pe_get(S.q[RR][0]); }          // probe builds only (tools/micro/jacobi_probe.cu)
    else {
        const float2 s3 = pe_get(S.q[RR][3]), s0 = pe_get(S.q[RR][0]);
        left.x = __shfl_up_sync(0xffffffffu, s3.x, 1);
        left.y = __shfl_up_sync(0xffffffffu, s3.y, 1);
        right.x = __shfl_down_sync(0xffffffffu, s0.x, 1);
        right.y = __shfl_down_sync(0xffffffffu, s0.y, 1);
    }
    if (DBG & 8) {
        pe_set(Dst.q[RR][0], left); pe_set(Dst.q[RR][3], right);
        pe_set(Dst.q[RR][1], make_float2(uh[1], dh[1])); pe_set(Dst.q[RR][2], make_float2(uh[2], dh[2]));
        return;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2 t;
        E te;
        if (RR == 0) {                                                                                // rows 0 | 4: up = halo | row 3
            const float2 a = pe_get(S.q[1][c]), b = pe_get(S.q[3][c]);
            t.x = uh[c] + a.x; t.y = b.x + a.y;
            if (PACKED) te = pe_make<E>(t);
        } else if (RR == 3) {                                                                         // rows 3 | 7: down = row 4 | halo
            const float2 a = pe_get(S.q[2][c]), b = pe_get(S.q[0][c]);
            t.x = a.x + b.y; t.y = a.y + dh[c];
            if (PACKED) te = pe_make<E>(t);
        } else if (PACKED) te = pe_add(S.q[RR - 1][c], S.q[RR + 1][c]);
        else {
            const float2 a = pe_get(S.q[RR - 1][c]), b = pe_get(S.q[RR + 1][c]);
            t.x = a.x + b.x; t.y = a.y + b.y;
        }
        if (PACKED) {
            const E l = (c == 0) ? pe_make<E>(left) : S.q[RR][c - 1];
            const E r = (c == 3) ? pe_make<E>(right) : S.q[RR][c + 1];
            te = pe_add(te, l);
            te = pe_add(te, r);
            te = pe_add(te, ND.q[RR][c]);
            Dst.q[RR][c] = pe_mul(M[c], te);
        } else {
            const float2 l = (c == 0) ? left : pe_get(S.q[RR][c - 1]);
            const float2 r = (c == 3) ? right : pe_get(S.q[RR][c + 1]);
            const float2 nd = pe_get(ND.q[RR][c]), m = pe_get(M[c]);
            t.x = t.x + l.x; t.y = t.y + l.y;
            t.x = t.x + r.x; t.y = t.y + r.y;
            t.x = t.x + nd.x; t.y = t.y + nd.y;
            pe_set(Dst.q[RR][c], make_float2(m.x * t.x, m.y * t.y));
        }
    }
}

// rows of the strip that are on the Dirichlet ring or outside the grid: bit r set -> row r is forced to 0
template <class E>
__device__ __forceinline__ void packed_zero_rows(PackedStripT<E>& S, const unsigned ringmask)
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
template <int PMASK, int DBG = 0, bool MBAR = false, int ORDER = 0, class E>
__device__ __forceinline__ void sweep_packed(const PackedStripT<E>& S, PackedStripT<E>& Dst, const PackedStripT<E>& ND,
                                             const float4 uph, const float4 dnh, const E (&M)[4], const unsigned ringmask,
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

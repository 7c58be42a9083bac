// jacobi.cu -- a6: the Jacobi pressure sweeps of pressure_projection (navier_stokes.py:139-145).
//
//   K times:  p' = 0 ;  p'[i][j] = 0.25 * ((((p[i-1][j] + p[i+1][j]) + p[i][j-1]) + p[i][j+1]) - div[i][j])
//             for 1 <= i <= h-2, 1 <= j <= w-2 ;  p = p'
//
// Design (B200): the pressure tile lives in REGISTERS for T fused sweeps (temporal blocking).  A warp
// spans 128 columns (lane l owns columns 4l..4l+3 as a float4) and R consecutive rows; a CTA stacks NW
// warps, i.e. a (NW*R) x 128 tile.  Per sweep a thread needs, from outside its registers, only
//   - the column to its left / right: one __shfl_up / __shfl_down per row,
//   - the row above its first / below its last row: one float4 through a double-buffered shared-memory
//     halo line per warp, one __syncthreads per sweep.
// div is loop-invariant and stays in registers too.  HBM/L2 traffic per launch is one read of p and div
// and one write of p for T sweeps: 12/T B per cell-sweep (+ the overlap redundancy) instead of 12.
// Tiles overlap by T rows and HX = roundup4(T) columns (trapezoid scheme: the outer ring of a tile goes
// stale by one cell per sweep and is not stored); a grid that fits one tile needs no halo and does all K
// sweeps in a single launch.  The Dirichlet ring and everything outside the domain is handled with a
// per-cell multiplier (0.25 inside, 0 on the ring/outside), which costs no extra instruction.
//
// Bit-exactness: the association above is kept literally; 0.25*x == x*0.25 in IEEE; no FMA (-fmad=false).
#include <cstdlib>
#include "common.cuh"
#include "jacobi_core.cuh"

namespace smk {

template <int R, int NW>
__global__ void __launch_bounds__(NW * 32, (NW * 32 <= 256) ? 2 : 1)
k_jacobi(const float* __restrict__ pin, float* __restrict__ pout, const float* __restrict__ div,
         const int h, const int w, const int pitch, const long long bstride,
         const int T, const int HX, const int ox, const int oy)
{
    __shared__ float4 halo[2][2][NW][32];          // [buffer][0: first row, 1: last row][warp][lane]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * ox, y0 = blockIdx.y * oy;
    const int gj = x0 + lane * 4;
    const int gi0 = y0 + warp * R;
    const size_t boff = (size_t)blockIdx.z * (size_t)bstride;
    pin += boff; pout += boff; div += boff;

    float4 P[R], D[R];
    const bool colin = gj < pitch;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int gi = gi0 + r;
        if (colin && gi < h) {
            P[r] = *reinterpret_cast<const float4*>(pin + (size_t)gi * pitch + gj);
            D[r] = __ldg(reinterpret_cast<const float4*>(div + (size_t)gi * pitch + gj));
        } else {
            P[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            D[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float cm0 = (gj + 0 >= 1 && gj + 0 <= w - 2) ? 0.25f : 0.f;
    const float cm1 = (gj + 1 >= 1 && gj + 1 <= w - 2) ? 0.25f : 0.f;
    const float cm2 = (gj + 2 >= 1 && gj + 2 <= w - 2) ? 0.25f : 0.f;
    const float cm3 = (gj + 3 >= 1 && gj + 3 <= w - 2) ? 0.25f : 0.f;
    const bool fast = (gi0 >= 1) && (gi0 + R - 1 <= h - 2);     // warp-uniform: no ring row in this warp
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    halo[0][0][warp][lane] = P[0];
    halo[0][1][warp][lane] = P[R - 1];
    __syncthreads();
    for (int s = 0; s < T; ++s) {
        const int buf = s & 1;
        const float4 up = warp > 0 ? halo[buf][1][warp - 1][lane] : zero4;
        const float4 dn = warp < NW - 1 ? halo[buf][0][warp + 1][lane] : zero4;
        float4* pf = (s + 1 < T) ? &halo[buf ^ 1][0][warp][lane] : nullptr;
        float4* pl = (s + 1 < T) ? &halo[buf ^ 1][1][warp][lane] : nullptr;
        if (fast) sweep_rows<R, true>(P, D, up, dn, cm0, cm1, cm2, cm3, gi0, h, pf, pl);
        else      sweep_rows<R, false>(P, D, up, dn, cm0, cm1, cm2, cm3, gi0, h, pf, pl);
        if (s + 1 < T) __syncthreads();
    }

    // store the part of the tile that is still exact after T sweeps
    const int vx0 = x0 + (blockIdx.x > 0 ? HX : 0);
    const int vx1 = (blockIdx.x + 1 < gridDim.x) ? x0 + 128 - HX : pitch;
    const int vy0 = y0 + (blockIdx.y > 0 ? T : 0);
    const int vy1 = (blockIdx.y + 1 < gridDim.y) ? y0 + NW * R - T : h;
    if (gj >= vx0 && gj < vx1 && gj < pitch) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gi = gi0 + r;
            if (gi >= vy0 && gi < vy1 && gi < h)
                *reinterpret_cast<float4*>(pout + (size_t)gi * pitch + gj) = P[r];
        }
    }
}

// Packed variant (see jacobi_core.cuh): add.f32x2 / mul.f32x2 halve the FP32 issue slots of a sweep.  The strip is
// always 8 rows per thread (R == 8); sweeps ping-pong between two register sets, two sweeps per loop trip.
template <int R, int NW, int PMASK>
__global__ void __launch_bounds__(NW * 32, (NW * 32 <= 256) ? 2 : 1)
k_jacobi_packed(const float* __restrict__ pin, float* __restrict__ pout, const float* __restrict__ div,
                const int h, const int w, const int pitch, const long long bstride,
                const int T, const int HX, const int ox, const int oy)
{
    static_assert(R == 8, "the packed strip pairs row r with row r + 4");
    __shared__ float4 halo[2][2][NW][32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * ox, y0 = blockIdx.y * oy;
    const int gj = x0 + lane * 4;
    const int gi0 = y0 + warp * R;
    const size_t boff = (size_t)blockIdx.z * (size_t)bstride;
    pin += boff; pout += boff; div += boff;

    PackedStrip A, B, ND;
    const bool colin = gj < pitch;
    unsigned ringmask = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int gi = gi0 + r;
        float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = p4;
        if (colin && gi < h) {
            p4 = *reinterpret_cast<const float4*>(pin + (size_t)gi * pitch + gj);
            d4 = __ldg(reinterpret_cast<const float4*>(div + (size_t)gi * pitch + gj));
        }
        if (gi < 1 || gi > h - 2) ringmask |= 1u << r;
        packed_set_row(A, r, p4);
        packed_set_row(ND, r, make_float4(-d4.x, -d4.y, -d4.z, -d4.w));
    }
    float2 M[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float m = (gj + c >= 1 && gj + c <= w - 2) ? 0.25f : 0.f;
        M[c] = make_float2(m, m);
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    halo[0][0][warp][lane] = packed_row(A, 0);
    halo[0][1][warp][lane] = packed_row(A, 7);
    __syncthreads();
    int s = 0;
    for (; s + 1 < T; s += 2) {            // s is even here: sweep s reads halo[0], sweep s+1 reads halo[1]
        {
            const float4 up = warp > 0 ? halo[0][1][warp - 1][lane] : zero4;
            const float4 dn = warp < NW - 1 ? halo[0][0][warp + 1][lane] : zero4;
            sweep_packed<PMASK, 0, false, 1>(A, B, ND, up, dn, M, ringmask, &halo[1][0][warp][lane], &halo[1][1][warp][lane]);
            __syncthreads();
        }
        {
            const float4 up = warp > 0 ? halo[1][1][warp - 1][lane] : zero4;
            const float4 dn = warp < NW - 1 ? halo[1][0][warp + 1][lane] : zero4;
            const bool more = s + 2 < T;
            sweep_packed<PMASK, 0, false, 1>(B, A, ND, up, dn, M, ringmask, more ? &halo[0][0][warp][lane] : nullptr, more ? &halo[0][1][warp][lane] : nullptr);
            if (more) __syncthreads();
        }
    }
    if (s < T) {                           // odd T: one more sweep, then move the result back into A
        const float4 up = warp > 0 ? halo[0][1][warp - 1][lane] : zero4;
        const float4 dn = warp < NW - 1 ? halo[0][0][warp + 1][lane] : zero4;
        sweep_packed<PMASK, 0, false, 1>(A, B, ND, up, dn, M, ringmask, nullptr, nullptr);
        A = B;
    }

    const int vx0 = x0 + (blockIdx.x > 0 ? HX : 0);
    const int vx1 = (blockIdx.x + 1 < gridDim.x) ? x0 + 128 - HX : pitch;
    const int vy0 = y0 + (blockIdx.y > 0 ? T : 0);
    const int vy1 = (blockIdx.y + 1 < gridDim.y) ? y0 + NW * R - T : h;
    if (gj >= vx0 && gj < vx1 && gj < pitch) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gi = gi0 + r;
            if (gi >= vy0 && gi < vy1 && gi < h)
                *reinterpret_cast<float4*>(pout + (size_t)gi * pitch + gj) = packed_row(A, r);
        }
    }
}

static int ntiles(int n, int tile, int halo)
{
    // first tile starts at 0, tiles advance by tile-2*halo, last tile must reach n
    if (n <= tile) return 1;
    const int adv = tile - 2 * halo;
    return (n - tile + adv - 1) / adv + 1;
}

// Sweeps per launch for an overlapped-tile run, from a cost model fitted on a B200 (tools/tune_jacobi.py):
//   time = launches * 2.3 us  +  sum over launches of  waves * (3 us + T * 0.56 us)        (128 x 128 tiles, 1 CTA / SM)
// Small grids want the largest T that still fits one wave of CTAs; large grids settle near T = 12, where the
// halo redundancy ((128 / (128 - 2 HX)) * (128 / (128 - 2 T))) starts to outweigh the saved loads.
static int pick_T(const smk_grid_t* g, int K, int TH, int tmax, int sms)
{
    double best = 1e30; int bestT = 1;
    for (int nl = 1; nl <= K; ++nl) {
        const int T = (K + nl - 1) / nl;
        if (T > tmax) continue;
        double cost = 2.3 * nl;
        int left = K;
        for (int l = 0; l < nl; ++l) {
            const int t = (left + (nl - l) - 1) / (nl - l);       // balanced split
            left -= t;
            const int HX = (t + 3) & ~3;
            const long ctas = (long)ntiles(g->w, 128, HX) * ntiles(g->h, TH, t) * g->batch;
            const long waves = (ctas + sms - 1) / sms;
            cost += waves * (3.0 + 0.56 * t);
        }
        if (cost < best) { best = cost; bestT = T; }
        if (T == 1) break;
    }
    return bestT;
}

// 0: scalar in-place kernel (k_jacobi); else the ping-pong kernel (k_jacobi_packed) with this mask of row pairs
// (bit rr) on f32x2 arithmetic.  Default 6: measured fastest on B200 (tools/micro/jacobi_probe.cu: 886 cycles per
// sweep of a 128 x 128 tile against 928 all-scalar, 920 all-packed; the in-place k_jacobi needs ~1130).
static int use_packed()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("SMK_JACOBI_PACKED"); v = e ? atoi(e) : 6; }
    return v;
}

template <int R, int NW>
static int run_cfg(const smk_grid_t* g, const float* div, float* a, float* b, int K, int T, int* in_scratch, cudaStream_t s)
{
    const int TH = R * NW;
    const bool whole = (g->h <= TH) && (g->w <= 128);
    float* src = a; float* dst = b;
    int left = K;
    int nl = whole ? 1 : (K + T - 1) / T;
    for (int l = 0; l < nl; ++l) {
        const int t = (left + (nl - l) - 1) / (nl - l);           // K sweeps split evenly over the launches
        left -= t;
        const int HX = (t + 3) & ~3;
        const int nx = ntiles(g->w, 128, HX), ny = ntiles(g->h, TH, t);
        dim3 grid(nx, ny, g->batch);
        int rc;
        {
            ProfScope prof_(SMK_PH_JACOBI, s);
#define SMK_PACKED_CASE(M) case M: k_jacobi_packed<8, NW, M><<<grid, NW * 32, 0, s>>>(src, dst, div, g->h, g->w, g->pitch_c, \
                                                           (long long)g->stride_c, t, HX, 128 - 2 * HX, TH - 2 * t); break;
            if (R == 8 && use_packed()) {
                switch (use_packed()) {
                    SMK_PACKED_CASE(6) SMK_PACKED_CASE(9) SMK_PACKED_CASE(2) SMK_PACKED_CASE(7) SMK_PACKED_CASE(16)
                    default: k_jacobi_packed<8, NW, 15><<<grid, NW * 32, 0, s>>>(src, dst, div, g->h, g->w, g->pitch_c,
                                                           (long long)g->stride_c, t, HX, 128 - 2 * HX, TH - 2 * t);
                }
            } else
                k_jacobi<R, NW><<<grid, NW * 32, 0, s>>>(src, dst, div, g->h, g->w, g->pitch_c, (long long)g->stride_c,
                                                         t, HX, 128 - 2 * HX, TH - 2 * t);
            rc = check_launch("k_jacobi");
        }
        if (rc != SMK_OK) return rc;
        float* tmp = src; src = dst; dst = tmp;
    }
    *in_scratch = (src == b) ? 1 : 0;
    return SMK_OK;
}

static int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int launch_jacobi(const smk_grid_t* g, const float* div, float* p, float* scratch, int K, int T, int* in_scratch, cudaStream_t s)
{
    *in_scratch = 0;
    if (K <= 0) return SMK_OK;
    if (g->batch > 65535) return fail(SMK_EUNSUPPORTED, "smk_jacobi: batch %d > 65535", g->batch);
    if (g->w <= 128 && g->h <= 32) return run_cfg<4, 8>(g, div, p, scratch, K, T > 0 ? T : 8, in_scratch, s);
    if (g->w <= 128 && g->h <= 64) return run_cfg<8, 8>(g, div, p, scratch, K, T > 0 ? T : 8, in_scratch, s);
    if (g->w <= 128 && g->h <= 128) return run_cfg<8, 16>(g, div, p, scratch, K, T > 0 ? T : 8, in_scratch, s);
    // large grids: overlapped tiles.  T must leave a positive advance in both directions.
    int tile = 0;                                   // 0 auto, 1: 64 x 128 (8 warps), 2: 128 x 128 (16 warps), 3: 64 x 128 (16 warps)
    if (const char* e = getenv("SMK_JACOBI_TILE")) tile = atoi(e);
    if (tile == 0) {
        // Many waves of CTAs (>= 4 per SM: 4096^2, or a 1/8 slab of 8192^2): 64 x 128 tiles at two CTAs per SM, so that one CTA's tile load / store overlaps
        // the other's sweeps: 7-8 % faster than 128 x 128 tiles at 4096^2 and 8192^2, equal at 2048^2 (tools/tune_jacobi.py).
        const long ctas128 = (long)ntiles(g->w, 128, 12) * ntiles(g->h, 128, 10) * g->batch;
        if (ctas128 >= 4L * sm_count()) tile = 1;
    }
    if (T <= 0) T = (tile == 1 || tile == 3) ? 10 : pick_T(g, K, 128, 24, sm_count());
    if (tile == 1) return run_cfg<8, 8>(g, div, p, scratch, K, min(T, 12), in_scratch, s);
    if (tile == 3) return run_cfg<4, 16>(g, div, p, scratch, K, min(T, 12), in_scratch, s);
    return run_cfg<8, 16>(g, div, p, scratch, K, min(T, 24), in_scratch, s);
}

}  // namespace smk

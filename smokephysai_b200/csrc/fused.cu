// fused.cu -- n whole step()s (navier_stokes.py:151-173) of a simulation that fits on one SM, in ONE kernel.
//
// For grids up to 128 x 128 (the reference's config.yaml default, and BASELINE config 2: hundreds of independent
// 128 x 128 sequences) the entire state of a simulation fits on chip: u, v and density live in shared memory
// (66,048 + 67,584 + 65,536 B), the pressure and the divergence in registers (the 128 x 128 register tile of
// k_jacobi: a warp owns 8 rows, a lane 4 columns).  A CTA of 16 warps owns one simulation and runs every phase
// of every step on it -- buoyancy, the three diffusions, divergence, K Jacobi sweeps, gradient subtract, the three
// sequential advections, decay, the returned (fractal-scaled) frame -- touching HBM only to read the state once
// per launch, write one frame per step and write the state back at the end: 4 B per cell-step (+ 36 B per cell
// per launch) instead of the ~104 B per cell-step of the phase-per-kernel path, and one launch instead of 6 n.
//
// Two thread -> cell mappings are used, both over the warp's 8 rows:
//   strip  : lane l owns columns 4l .. 4l+3 (one LDS.128 / STS.128 per row, neighbours by shuffle) -- buoyancy,
//            diffusion, divergence, Jacobi, gradient subtract, frame output;
//   cyclic : lane l owns columns l, l+32, l+64, l+96 -- advection, whose back-traced gathers then hit 32 distinct
//            banks for sub-cell displacements (the strip mapping would be a 4-way conflict on every gather).
// Out-of-place phases (diffusion, advection) are made in-place-safe through registers: every thread computes
// the new values of its cells, the CTA synchronises, then everyone writes back.  u's staggered row 128 and v's
// staggered column 128 (present when h == 128 / w == 128) are spread over lanes 0..7 of every warp.
// Arithmetic is the same rounded-once sequence as the tiled kernels, so results are bit-identical to them.
//
// Scheduling: one CTA per simulation, or -- when the simulations do not fill whole waves of SMs -- the time-sliced
// schedule (plan_items / pick_seg_len below): the simulation-steps of the call are cut into equal pieces per SM, a
// simulation that straddles a piece boundary is handed from one CTA to another through global memory.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>
#include "common.cuh"
#include "jacobi_core.cuh"

namespace smk {

constexpr int FZ_R = 8, FZ_NW = 16, FZ_THREADS = FZ_NW * 32;
constexpr int FZ_PU = 128, FZ_PV = 132, FZ_PD = 128;                 // shared-memory row pitches
constexpr int FZ_SU = 129 * FZ_PU, FZ_SV = 128 * FZ_PV, FZ_SD = 128 * FZ_PD;     // floats
#ifndef SMK_FZ_PMASK
#define SMK_FZ_PMASK 6
#endif
constexpr int FZ_PMASK = SMK_FZ_PMASK;   // bit rr: row pair rr of a sweep uses f32x2 arithmetic (6: pairs 1, 2 packed, 0, 3 scalar)
#ifndef SMK_FZ_INTERIOR_FIRST
#define SMK_FZ_INTERIOR_FIRST 1
#endif
constexpr int FZ_ORDER = SMK_FZ_INTERIOR_FIRST;   // a sweep computes rows 1..6 before rows 0 and 7 (hides the halo loads): mask 6 + this order is 2.3 % faster on c2 than mask 7 + boundary rows first
#ifndef SMK_FZ_ELEM
#define SMK_FZ_ELEM float2
#endif
// element of the pressure strips (jacobi_core.cuh): float2, or a pinned 64-bit pair.  The 64-bit element removes 6 % of the
// sweep loop's instructions here (no FSEL, fewer MOV) but c2 ran 2 % slower with it (49.1 against 50.1 G cell-steps/s), so float2 stays.
typedef SMK_FZ_ELEM FzElem;
typedef PackedStripT<FzElem> FzStrip;
constexpr size_t FZ_SMEM = (size_t)(FZ_SU + FZ_SV + FZ_SD) * 4 + sizeof(float4) * 2 * 2 * FZ_NW * 32;

// Phase timing for tools/micro/fused_probe.cu only (never defined in the library build): thread 0 of CTA 0
// accumulates clock64() deltas per phase into FusedArgs::ticks.
#ifdef SMK_FUSED_TIMING
#define FZ_TICK(k) do { if (tid == 0 && blockIdx.x == 0) { const long long t1_ = clock64(); a.ticks[k] += t1_ - tick0_; tick0_ = t1_; } } while (0)
#else
#define FZ_TICK(k) do { } while (0)
#endif

struct FusedArgs {
    float* U; float* V; float* D; float* P;              // live state, updated in place
    float* frames; const float* fmul;                     // frames may be NULL; fmul may be NULL
    int h, w, pu, pv, pc;
    long long su_, sv_, sc_, frame_step_stride, frame_batch_stride;
    float dt, c_uv, c_d, decay;
    int K, nsteps;
    const int4* items;                                    // time-sliced schedule (see k_step_fused): (simulation, first step, end step, 0) per CTA, else NULL
    unsigned* progress;                                   // [nsims] steps completed, zeroed before the launch (time-sliced schedule only)
    unsigned* ticket;                                     // next item of `items` to hand out (zeroed with progress)
    unsigned spin_budget;                                 // polls a consumer CTA spends on a hand-over before it gives up
#ifdef SMK_FUSED_TIMING
    long long* ticks;
#endif
};

__device__ __forceinline__ float4 zlds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void zsts4(float* p, const float4 v) { *reinterpret_cast<float4*>(p) = v; }

// diffusion_step at one cell: f + c*((((up+down)+left)+right) - 4f)                    navier_stokes.py:69-72
__device__ __forceinline__ float zdiff1(float f, float up, float dn, float l, float r, float c)
{
    float s = up + dn;
    s = s + l;
    s = s + r;
    return f + c * (s - 4.0f * f);
}
// ... with replicate padding by index clamp (:57-66), any cell of a rows x cols field in shared memory
__device__ __forceinline__ float zdiff_cell(const float* F, int pitch, int rows, int cols, int i, int j, float c)
{
    const int iu = max(i - 1, 0), id = min(i + 1, rows - 1), jl = max(j - 1, 0), jr = min(j + 1, cols - 1);
    return zdiff1(F[i * pitch + j], F[iu * pitch + j], F[id * pitch + j], F[i * pitch + jl], F[i * pitch + jr], c);
}

// Strip-mapped diffusion of the thread's 8 x 4 cells of a rows x cols field (rows <= 129, cols <= 129).  Rows slide
// through registers (LDS.128); the left / right neighbours of the strip come from the adjacent lanes; replicate
// padding (:57-66) is an index clamp on the rows and a select on the columns.  Cells outside the field get
// finite garbage that the caller does not store.  In the FULL instantiation rows / cols are literals.
template <int PITCH>
__device__ __forceinline__ void zdiffuse_strip(const float* F, const int rows, const int cols, const float c, float4 (&out)[FZ_R],
                                               const int r0, const int c0, const int lane)
{
    const int rlast = rows - 1, clast = cols - 1;
    float4 up = zlds4(F + min(max(r0 - 1, 0), rlast) * PITCH + c0);
    float4 cur = zlds4(F + min(r0, rlast) * PITCH + c0);
#pragma unroll
    for (int r = 0; r < FZ_R; ++r) {
        const int i = r0 + r;
        const float4 dn = zlds4(F + min(i + 1, rlast) * PITCH + c0);
        float left = __shfl_up_sync(0xffffffffu, cur.w, 1);
        float right = __shfl_down_sync(0xffffffffu, cur.x, 1);
        if (lane == 0) left = cur.x;
        if (lane == 31 && clast >= 128) right = F[min(i, rlast) * PITCH + 128];       // the staggered column of v
        out[r].x = zdiff1(cur.x, up.x, dn.x, left, (c0 + 1 <= clast) ? cur.y : cur.x, c);
        out[r].y = zdiff1(cur.y, up.y, dn.y, cur.x, (c0 + 2 <= clast) ? cur.z : cur.y, c);
        out[r].z = zdiff1(cur.z, up.z, dn.z, cur.y, (c0 + 3 <= clast) ? cur.w : cur.z, c);
        out[r].w = zdiff1(cur.w, up.w, dn.w, cur.z, (c0 + 4 <= clast) ? right : cur.w, c);
        up = cur; cur = dn;
    }
}

// zero the components of a strip float4 that lie at or beyond column `cols` (the padding columns stay zero)
__device__ __forceinline__ float4 zmask4(float4 v, const int c0, const int cols)
{
    if (c0 + 0 >= cols) v.x = 0.f;
    if (c0 + 1 >= cols) v.y = 0.f;
    if (c0 + 2 >= cols) v.z = 0.f;
    if (c0 + 3 >= cols) v.w = 0.f;
    return v;
}

// Four IEEE quotients n / d (navier_stokes.py:136 divides by dt).  nvcc's div.rn.f32 has a fast path (three FFMAs
// around a reciprocal) and, for operands or quotients near the denormal range, a ~100-instruction subroutine; the
// thin ring where the diffusing velocities underflow sends whole warps there every step.  Those warps take an
// fp64 division instead: (float)((double)n / (double)d) is the correctly rounded fp32 quotient for every finite
// n, d (53 >= 2*24 + 2 bits makes the double rounding innocuous, denormal results included: a quotient of two
// 24-bit significands is either exactly a rounding midpoint or at least 2^-48 away from it in relative terms).
__device__ __forceinline__ bool zdiv_odd(const float x)
{
    const float ax = fabsf(x);
    return ax != 0.0f && !(ax >= 1e-30f && ax <= 1e30f);
}
__device__ __forceinline__ float4 zdiv4(float4 n, const float d)
{
#ifdef SMK_PROBE_NODIV               // tools/micro only: what the step costs without the division (wrong results)
    n.x *= d; n.y *= d; n.z *= d; n.w *= d;
    return n;
#endif
    if (__any_sync(0xffffffffu, zdiv_odd(n.x) || zdiv_odd(n.y) || zdiv_odd(n.z) || zdiv_odd(n.w))) {
        const double dd = (double)d;
        n.x = (float)((double)n.x / dd); n.y = (float)((double)n.y / dd);
        n.z = (float)((double)n.z / dd); n.w = (float)((double)n.w / dd);
    } else {
        n.x = n.x / d; n.y = n.y / d; n.z = n.z / d; n.w = n.w / d;
    }
    return n;
}
// (Round 2 tried to hoist the reciprocal refinement of div.rn.f32 out of the quotients -- MUFU.RCP + two FFMA once per kernel, then
//  the three-FFMA tail per quotient, bit-identical to the compiler's own fast path and verified on the CPU over all 2^31 numerators
//  of dt = 0.01 -- with the validity range enforced by an explicit per-warp test instead of FCHK.  The divergence phase shrank from
//  612 to ~470 executed instructions, but c2 ran 1.8 % SLOWER (50.6 against 51.5 G cell-steps/s): the range test needs three
//  compares per value where FCHK needs one instruction, and the compiler had merged the two branches below into one stream.)

// advection_step at one cell (navier_stokes.py:74-131): same restatement as k_advect (stencil.cu) with the
// field, u and v in shared memory.  rows x cols is the advected field, h x w the cell grid.
template <int PITCH>
__device__ __forceinline__ float zadvect_cell(const float* F, const int rows, const int cols, const float* su, const float* sv,
                                              const int h, const int w, const int i, const int j, const float dt)
{
    // a9: u_i = 0.5*U[i][j] + 0.5*U[i][j+1] for j <= w-2 and i <= h-1, else 0;  v_i likewise along rows
    float ui = 0.0f, vi = 0.0f;
    if (j <= w - 2 && i <= h - 1) ui = 0.5f * su[i * FZ_PU + j] + 0.5f * su[i * FZ_PU + j + 1];
    if (i <= h - 2 && j <= w - 1) vi = 0.5f * sv[i * FZ_PV + j] + 0.5f * sv[(i + 1) * FZ_PV + j];
    const float xmax = (float)(cols - 1), ymax = (float)(rows - 1);
    float px = (float)j - dt * ui;
    float py = (float)i - dt * vi;
    px = fminf(fmaxf(px, 0.0f), xmax);
    py = fminf(fmaxf(py, 0.0f), ymax);
    const float fx0 = floorf(px), fy0 = floorf(py);
    const float fx1 = fminf(fx0 + 1.0f, xmax), fy1 = fminf(fy0 + 1.0f, ymax);
    const int dx = (fx1 != fx0) ? 1 : 0, dy = (fy1 != fy0) ? PITCH : 0;
    const float ax = fx1 - px, bx = px - fx0, ay = fy1 - py, by = py - fy0;
    const float* q = F + (int)__fmaf_rn(fy0, (float)PITCH, fx0);           // exact in fp32 (< 2^24), one conversion
    float s = (ax * ay) * q[0] + (bx * ay) * q[dx];
    s = s + (ax * by) * q[dy];
    s = s + (bx * by) * q[dy + dx];
    return s;
}

// The same for the two cells (i, j) and (i, j + 32) of a thread's cyclic-mapped row at once, with the fp32 adds and
// multiplies issued as f32x2 instructions (__fadd2_rn / __fmul2_rn: two IEEE-rounded results per instruction, so
// the per-cell operation sequence and rounding are those of zadvect_cell; x - y is x + (-y)).  Clamps, floors,
// conversions and the gathers stay scalar -- and so does every ADD THAT CONSUMES A PRODUCT: ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (the scalar forms are left alone), which would
// break the rounded-once contract.  Packed: the multiplies, and the adds whose operands are not products.  Velocity samples that the reference's interpolation zeroes are loaded as
// 0, which makes u_i / v_i exactly 0.  `scale` (the 0.995 decay, :171) is applied when SCALE is set.
__device__ __forceinline__ float2 zneg2(const float2 a) { return make_float2(-a.x, -a.y); }

template <int PITCH, bool SCALE>
__device__ __forceinline__ float2 zadvect_pair(const float* F, const int rows, const int cols, const float* su, const float* sv,
                                               const int h, const int w, const int i, const int j, const float dt, const float scale)
{
    const int j1 = j + 32;
    const bool cu0 = (j <= w - 2 && i <= h - 1), cu1 = (j1 <= w - 2 && i <= h - 1);
    const bool cv0 = (i <= h - 2 && j <= w - 1), cv1 = (i <= h - 2 && j1 <= w - 1);
    const float2 ua = make_float2(cu0 ? su[i * FZ_PU + j] : 0.f, cu1 ? su[i * FZ_PU + j1] : 0.f);
    const float2 ub = make_float2(cu0 ? su[i * FZ_PU + j + 1] : 0.f, cu1 ? su[i * FZ_PU + j1 + 1] : 0.f);
    const float2 va = make_float2(cv0 ? sv[i * FZ_PV + j] : 0.f, cv1 ? sv[i * FZ_PV + j1] : 0.f);
    const float2 vb = make_float2(cv0 ? sv[(i + 1) * FZ_PV + j] : 0.f, cv1 ? sv[(i + 1) * FZ_PV + j1] : 0.f);
    const float2 half2 = make_float2(0.5f, 0.5f), dt2 = make_float2(dt, dt), one2 = make_float2(1.0f, 1.0f);
    const float2 hua = __fmul2_rn(half2, ua), hub = __fmul2_rn(half2, ub);
    const float2 hva = __fmul2_rn(half2, va), hvb = __fmul2_rn(half2, vb);
    const float2 ui = make_float2(hua.x + hub.x, hua.y + hub.y);
    const float2 vi = make_float2(hva.x + hvb.x, hva.y + hvb.y);
    const float xmax = (float)(cols - 1), ymax = (float)(rows - 1);
    const float2 du = __fmul2_rn(dt2, ui), dv = __fmul2_rn(dt2, vi);
    float2 px = make_float2((float)j - du.x, (float)j1 - du.y);
    float2 py = make_float2((float)i - dv.x, (float)i - dv.y);
    px.x = fminf(fmaxf(px.x, 0.0f), xmax); px.y = fminf(fmaxf(px.y, 0.0f), xmax);
    py.x = fminf(fmaxf(py.x, 0.0f), ymax); py.y = fminf(fmaxf(py.y, 0.0f), ymax);
    const float2 fx0 = make_float2(floorf(px.x), floorf(px.y)), fy0 = make_float2(floorf(py.x), floorf(py.y));
    float2 fx1 = __fadd2_rn(fx0, one2), fy1 = __fadd2_rn(fy0, one2);
    fx1.x = fminf(fx1.x, xmax); fx1.y = fminf(fx1.y, xmax);
    fy1.x = fminf(fy1.x, ymax); fy1.y = fminf(fy1.y, ymax);
    const float2 ax = __fadd2_rn(fx1, zneg2(px)), bx = __fadd2_rn(px, zneg2(fx0));
    const float2 ay = __fadd2_rn(fy1, zneg2(py)), by = __fadd2_rn(py, zneg2(fy0));
    // flat index y0 * PITCH + x0 formed in fp32 (integers below 2^24: exact) and converted once: the conversion pipe
    // delivers 16 results per clock per SM, an eighth of the FP32 rate (tools/micro/fp32_pipes.cu)
    const float* q0 = F + (int)__fmaf_rn(fy0.x, (float)PITCH, fx0.x);
    const float* q1 = F + (int)__fmaf_rn(fy0.y, (float)PITCH, fx0.y);
    const int dx0 = (fx1.x != fx0.x) ? 1 : 0, dy0 = (fy1.x != fy0.x) ? PITCH : 0;
    const int dx1 = (fx1.y != fx0.y) ? 1 : 0, dy1 = (fy1.y != fy0.y) ? PITCH : 0;
    const float2 f00 = make_float2(q0[0], q1[0]), f01 = make_float2(q0[dx0], q1[dx1]);
    const float2 f10 = make_float2(q0[dy0], q1[dy1]), f11 = make_float2(q0[dy0 + dx0], q1[dy1 + dx1]);
    const float2 t00 = __fmul2_rn(__fmul2_rn(ax, ay), f00), t01 = __fmul2_rn(__fmul2_rn(bx, ay), f01);
    const float2 t10 = __fmul2_rn(__fmul2_rn(ax, by), f10), t11 = __fmul2_rn(__fmul2_rn(bx, by), f11);
    float2 s = make_float2(t00.x + t01.x, t00.y + t01.y);
    s.x = s.x + t10.x; s.y = s.y + t10.y;
    s.x = s.x + t11.x; s.y = s.y + t11.y;
    if (SCALE) s = __fmul2_rn(s, make_float2(scale, scale));
    return s;
}

// FULL: h == w == 128 (every strip cell is a grid cell and the staggered extras exist); otherwise any h, w <= 128.
template <bool FULL>
__global__ void __launch_bounds__(FZ_THREADS, 1)
k_step_fused(const FusedArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float* su = smem;
    float* sv = su + FZ_SU;
    float* sd = sv + FZ_SV;
    float4 (*halo)[2][FZ_NW][32] = reinterpret_cast<float4 (*)[2][FZ_NW][32]>(sd + FZ_SD);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = FULL ? 128 : a.h, w = FULL ? 128 : a.w;
    const int pu = FULL ? 128 : a.pu, pv = FULL ? 132 : a.pv, pc = FULL ? 128 : a.pc;
    // Work item of this CTA.  Classic schedule (items == NULL): CTA b runs all nsteps of simulation b.  Time-sliced schedule:
    // the nsims x nsteps simulation-steps are laid on a line (simulation-major) and cut into equal pieces, one piece per SM
    // (McNaughton's wrap-around rule); a piece is walked from its END, so a simulation cut by a piece boundary is started
    // (steps 0..) as the FIRST item of the lower piece and finished as the LAST item of the next piece, after the first has
    // published the state (progress[b], release / acquire at GPU scope).  Every item is a CTA of its own; the host orders
    // them by planned start time: an item that waits always waits for an item of lower table index (no deadlock, see below).
    // 256 simulations x 20 steps on 148 SMs take 35 step-times instead of 40 that way.
    // The item is claimed with a ticket (atomicAdd at CTA entry), not read at blockIdx: tickets are handed out in the order
    // CTAs actually START, so "the item I wait for has a lower ticket" means "its CTA is already running or done" whatever
    // order the hardware dispatches the grid in (MPS, a co-resident kernel, a future scheduler).
    int4 item = make_int4((int)blockIdx.x, 0, a.nsteps, 0);
    if (a.items) {
        if (threadIdx.x == 0) reinterpret_cast<int*>(sd + FZ_SD)[0] = (int)atomicAdd(a.ticket, 1u);
        __syncthreads();
        item = __ldg(a.items + reinterpret_cast<const int*>(sd + FZ_SD)[0]);
        __syncthreads();                // the halo buffer the ticket travelled through is written again further down
    }
    const size_t b = (size_t)item.x;
    const int t_begin = item.y, t_end = item.z;
    float* __restrict__ gU = a.U + b * a.su_;
    float* __restrict__ gV = a.V + b * a.sv_;
    float* __restrict__ gD = a.D + b * a.sc_;
    float* __restrict__ gP = a.P + b * a.sc_;
    const int r0 = warp * FZ_R, c0 = lane * 4;
    // the staggered extras: u[128][xe] and v[xe][128], xe = 8*warp + lane for lanes 0..7
    const int xe = 8 * warp + lane;
    const bool xu = (h == 128) && lane < 8 && xe < w;
    const bool xv = (w == 128) && lane < 8 && xe < h;
    const float dt = a.dt;
#ifdef SMK_FUSED_TIMING
    long long tick0_ = clock64();
#endif

    if (t_begin > 0) {                  // the first steps of this simulation ran in another CTA: wait for its state
        if (tid == 0) {
            unsigned done, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(a.progress + b) : "memory");
                if (done < (unsigned)t_begin) {
                    // Backstop only: the producer holds a lower ticket, so it is resident and this wait ends after at most
                    // t_begin steps of its work.  The budget grows with that work (about 2^18 polls ~ 0.1 s per awaited
                    // step-sweep block, floor 2^25 ~ 6 s) so a long K or a slow clock cannot trip it; if it does expire
                    // the launch fails (sticky error the host sees at its next call) instead of hanging the GPU.
                    __nanosleep(200);
                    if (++spins > a.spin_budget) __trap();
                }
            } while (done < (unsigned)t_begin);
        }
        __syncthreads();
    }

    // ---- load the state: u, v, density to shared memory, pressure to registers ---------------------------
    if (!FULL) {
        for (int k = tid; k < (int)(FZ_SU + FZ_SV + FZ_SD) / 4; k += FZ_THREADS) zsts4(smem + 4 * k, make_float4(0.f, 0.f, 0.f, 0.f));
        __syncthreads();
    }
    {
        const int gu4 = pu >> 2, gv4 = pv >> 2, gc4 = pc >> 2;
        for (int k = tid; k < (h + 1) * gu4; k += FZ_THREADS) {
            const int i = k / gu4, g = k - i * gu4;
            zsts4(su + i * FZ_PU + 4 * g, __ldcg(reinterpret_cast<const float4*>(gU + (size_t)i * pu + 4 * g)));
        }
        for (int k = tid; k < h * gv4; k += FZ_THREADS) {
            const int i = k / gv4, g = k - i * gv4;
            zsts4(sv + i * FZ_PV + 4 * g, __ldcg(reinterpret_cast<const float4*>(gV + (size_t)i * pv + 4 * g)));
        }
        for (int k = tid; k < h * gc4; k += FZ_THREADS) {
            const int i = k / gc4, g = k - i * gc4;
            zsts4(sd + i * FZ_PD + 4 * g, __ldcg(reinterpret_cast<const float4*>(gD + (size_t)i * pc + 4 * g)));
        }
    }
    // pressure: the packed register strip of jacobi_core.cuh (a float2 pairs row r with row r + 4)
    FzStrip P = {};
    const bool colin = c0 < pc;
    unsigned ringmask = 0;              // rows of the strip on the Dirichlet ring or outside the grid
#pragma unroll
    for (int r = 0; r < FZ_R; ++r) {
        const int i = r0 + r;
        packed_set_row(P, r, (colin && i < h) ? __ldcg(reinterpret_cast<const float4*>(gP + (size_t)i * pc + c0)) : make_float4(0.f, 0.f, 0.f, 0.f));
        if (i < 1 || i > h - 2) ringmask |= 1u << r;
    }
    FzElem M[4];                        // 0.25 inside, 0 on the ring columns / outside
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float m = (c0 + c >= 1 && c0 + c <= w - 2) ? 0.25f : 0.f;
        M[c] = pe_make<FzElem>(make_float2(m, m));
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // frame output of the step just finished (:173, fractal_generator.py:62) and buoyancy of the next one
    // (:154-155: v[:, :-1] += dt * (density * 0.1)); both read the thread's own strip of density.
    // (mrow: the thread's eight float4 of the fractal multiplier, loaded by the caller ahead of time so that the L2
    // latency hides behind the density write-back and its barriers)
    auto frame_and_buoyancy = [&](float* frame, const bool buoy, const float4 (&mrow)[FZ_R]) {
#pragma unroll
        for (int r = 0; r < FZ_R; ++r) {
            const int i = r0 + r;
            if (FULL || (i < h && colin)) {
                const float4 d4 = zlds4(sd + i * FZ_PD + c0);
                if (frame) {
                    float4 fr = d4;
                    if (a.fmul) {
                        const float4 m = mrow[r];
                        fr.x = fr.x + m.x * fr.x; fr.y = fr.y + m.y * fr.y; fr.z = fr.z + m.z * fr.z; fr.w = fr.w + m.w * fr.w;
                    }
                    *reinterpret_cast<float4*>(frame + (size_t)i * pc + c0) = fr;
                }
                if (buoy) {
                    float4 v4 = zlds4(sv + i * FZ_PV + c0);
                    if (FULL || c0 + 0 < w) v4.x = v4.x + dt * (d4.x * 0.1f);
                    if (FULL || c0 + 1 < w) v4.y = v4.y + dt * (d4.y * 0.1f);
                    if (FULL || c0 + 2 < w) v4.z = v4.z + dt * (d4.z * 0.1f);
                    if (FULL || c0 + 3 < w) v4.w = v4.w + dt * (d4.w * 0.1f);
                    zsts4(sv + i * FZ_PV + c0, v4);
                }
            }
        }
    };

    {
        const float4 none[FZ_R] = {};
        frame_and_buoyancy(nullptr, true, none);
    }
    __syncthreads();
    FZ_TICK(0);

    for (int t = t_begin; t < t_end; ++t) {
        // ---- a4 diffusion of u (rows 0..h), v (cols 0..w), density                    navier_stokes.py:158-160
        {
            float4 R[FZ_R];
            float X = 0.f;
            // u
            zdiffuse_strip<FZ_PU>(su, h + 1, w, a.c_uv, R, r0, c0, lane);
            if (xu) X = zdiff_cell(su, FZ_PU, h + 1, w, 128, xe, a.c_uv);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < FZ_R; ++r) { const int i = r0 + r; if (FULL || (i <= h && colin)) zsts4(su + i * FZ_PU + c0, FULL ? R[r] : zmask4(R[r], c0, w)); }
            if (xu) su[128 * FZ_PU + xe] = X;
            // v (reads sv only: no barrier needed after the u write-back)
            zdiffuse_strip<FZ_PV>(sv, h, w + 1, a.c_uv, R, r0, c0, lane);
            if (xv) X = zdiff_cell(sv, FZ_PV, h, w + 1, xe, 128, a.c_uv);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < FZ_R; ++r) { const int i = r0 + r; if (FULL || (i < h && c0 <= w)) zsts4(sv + i * FZ_PV + c0, FULL ? R[r] : zmask4(R[r], c0, w + 1)); }
            if (xv) sv[xe * FZ_PV + 128] = X;
            // density
            zdiffuse_strip<FZ_PD>(sd, h, w, a.c_d, R, r0, c0, lane);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < FZ_R; ++r) { const int i = r0 + r; if (FULL || (i < h && colin)) zsts4(sd + i * FZ_PD + c0, FULL ? R[r] : zmask4(R[r], c0, w)); }
            __syncthreads();
        }

        FZ_TICK(1);
        // ---- a5 divergence into registers: (((u[i+1][j] - u[i][j]) + v[i][j+1]) - v[i][j]) / dt           :136
        FzStrip ND = {};                // minus the divergence (x - div == x + (-div))
        {
            float4 ua = zlds4(su + r0 * FZ_PU + c0);
#pragma unroll
            for (int r = 0; r < FZ_R; ++r) {
                const int i = r0 + r;
                const float4 ub = zlds4(su + (i + 1) * FZ_PU + c0);
                const float4 va = zlds4(sv + i * FZ_PV + c0);
                const float vr = sv[i * FZ_PV + c0 + 4];
                float4 o;
                o.x = ((ub.x - ua.x) + va.y) - va.x;
                o.y = ((ub.y - ua.y) + va.z) - va.y;
                o.z = ((ub.z - ua.z) + va.w) - va.z;
                o.w = ((ub.w - ua.w) + vr) - va.w;
                o = zdiv4(o, dt);
                if (!FULL) {
                    if (i >= h) o = zero4;
                    if (c0 + 0 >= w) o.x = 0.f;
                    if (c0 + 1 >= w) o.y = 0.f;
                    if (c0 + 2 >= w) o.z = 0.f;
                    if (c0 + 3 >= w) o.w = 0.f;
                }
                packed_set_row(ND, r, make_float4(-o.x, -o.y, -o.z, -o.w));
                ua = ub;
            }
        }

        FZ_TICK(2);
        // ---- a6 K Jacobi sweeps on the register tile (see jacobi.cu / jacobi_core.cuh)                :139-145
        // two sweeps per trip, ping-pong between the register sets P and Q; sweep s reads halo[s & 1]
        halo[0][0][warp][lane] = packed_row(P, 0);
        halo[0][1][warp][lane] = packed_row(P, 7);
        __syncthreads();
        {
            FzStrip Q;
            int s = 0;
            for (; s + 1 < a.K; s += 2) {
                {
                    const float4 up = warp > 0 ? halo[0][1][warp - 1][lane] : zero4;
                    const float4 dn = warp < FZ_NW - 1 ? halo[0][0][warp + 1][lane] : zero4;
                    sweep_packed<FZ_PMASK, 0, false, FZ_ORDER>(P, Q, ND, up, dn, M, ringmask, &halo[1][0][warp][lane], &halo[1][1][warp][lane]);
                    __syncthreads();
                }
                {
                    const float4 up = warp > 0 ? halo[1][1][warp - 1][lane] : zero4;
                    const float4 dn = warp < FZ_NW - 1 ? halo[1][0][warp + 1][lane] : zero4;
                    sweep_packed<FZ_PMASK, 0, false, FZ_ORDER>(Q, P, ND, up, dn, M, ringmask, &halo[0][0][warp][lane], &halo[0][1][warp][lane]);
                    __syncthreads();
                }
            }
            if (s < a.K) {                  // odd K: one more sweep, the result moves back into P
                const float4 up = warp > 0 ? halo[0][1][warp - 1][lane] : zero4;
                const float4 dn = warp < FZ_NW - 1 ? halo[0][0][warp + 1][lane] : zero4;
                sweep_packed<FZ_PMASK, 0, false, FZ_ORDER>(P, Q, ND, up, dn, M, ringmask, &halo[1][0][warp][lane], &halo[1][1][warp][lane]);
                P = Q;
                __syncthreads();
            }
        }

        FZ_TICK(3);
        // ---- a7 gradient subtract, in place on the thread's strip                                     :148-149
        {
            // halo[K & 1][1][warp - 1] holds the final last row of the warp above: the "row above" of row r0
            float4 pabove = warp > 0 ? halo[a.K & 1][1][warp - 1][lane] : zero4;
#pragma unroll
            for (int r = 0; r < FZ_R; ++r) {
                const int i = r0 + r;
                const float4 pc4 = packed_row(P, r);
                const float pleft = __shfl_up_sync(0xffffffffu, pc4.w, 1);
                if (FULL || (i < h && colin)) {
                    float4 u4 = zlds4(su + i * FZ_PU + c0);
                    float4 v4 = zlds4(sv + i * FZ_PV + c0);
                    if (i >= 1) {
                        if (FULL || c0 + 0 < w) u4.x = u4.x - dt * (pc4.x - pabove.x);
                        if (FULL || c0 + 1 < w) u4.y = u4.y - dt * (pc4.y - pabove.y);
                        if (FULL || c0 + 2 < w) u4.z = u4.z - dt * (pc4.z - pabove.z);
                        if (FULL || c0 + 3 < w) u4.w = u4.w - dt * (pc4.w - pabove.w);
                        zsts4(su + i * FZ_PU + c0, u4);
                    }
                    if (c0 >= 1) v4.x = v4.x - dt * (pc4.x - pleft);
                    if (FULL || c0 + 1 < w) v4.y = v4.y - dt * (pc4.y - pc4.x);
                    if (FULL || c0 + 2 < w) v4.z = v4.z - dt * (pc4.z - pc4.y);
                    if (FULL || c0 + 3 < w) v4.w = v4.w - dt * (pc4.w - pc4.z);
                    zsts4(sv + i * FZ_PV + c0, v4);
                }
                pabove = pc4;
            }
        }
        __syncthreads();

        FZ_TICK(4);
        float4 mrow[FZ_R] = {};
        // ---- a10/a11 advection (cyclic mapping): u by (u, v); v by (u', v); density by (u', v'), decay  :166-171
        {
            // a thread's row of four cyclic cells is two pairs: columns (lane, lane + 32) and (lane + 64, lane + 96)
            float2 R[FZ_R][2];
            float X = 0.f;
#pragma unroll
            for (int r = 0; r < FZ_R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    R[r][kp] = zadvect_pair<FZ_PU, false>(su, h + 1, w, su, sv, h, w, i, j, dt, 1.0f);
                }
            if (xu) X = zadvect_cell<FZ_PU>(su, h + 1, w, su, sv, h, w, 128, xe, dt);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < FZ_R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    if (FULL || (i <= h && j < w)) su[i * FZ_PU + j] = R[r][kp].x;
                    if (FULL || (i <= h && j + 32 < w)) su[i * FZ_PU + j + 32] = R[r][kp].y;
                }
            if (xu) su[128 * FZ_PU + xe] = X;
            __syncthreads();

#pragma unroll
            for (int r = 0; r < FZ_R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    R[r][kp] = zadvect_pair<FZ_PV, false>(sv, h, w + 1, su, sv, h, w, i, j, dt, 1.0f);
                }
            if (xv) X = zadvect_cell<FZ_PV>(sv, h, w + 1, su, sv, h, w, xe, 128, dt);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < FZ_R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    if (FULL || (i < h && j <= w)) sv[i * FZ_PV + j] = R[r][kp].x;
                    if (FULL || (i < h && j + 32 <= w)) sv[i * FZ_PV + j + 32] = R[r][kp].y;
                }
            if (xv) sv[xe * FZ_PV + 128] = X;
            __syncthreads();

#pragma unroll
            for (int r = 0; r < FZ_R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    R[r][kp] = zadvect_pair<FZ_PD, true>(sd, h, w, su, sv, h, w, i, j, dt, a.decay);
                }
            __syncthreads();
            if (a.frames && a.fmul) {
#pragma unroll
                for (int r = 0; r < FZ_R; ++r) {
                    const int i = r0 + r;
                    if (FULL || (i < h && colin)) mrow[r] = __ldg(reinterpret_cast<const float4*>(a.fmul + (size_t)i * pc + c0));
                }
            }
#pragma unroll
            for (int r = 0; r < FZ_R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    if (FULL || (i < h && j < w)) sd[i * FZ_PD + j] = R[r][kp].x;
                    if (FULL || (i < h && j + 32 < w)) sd[i * FZ_PD + j + 32] = R[r][kp].y;
                }
            __syncthreads();
        }

        FZ_TICK(5);
        // ---- a11 returned copy of this step + buoyancy of the next
        float* frame = a.frames ? a.frames + b * a.frame_batch_stride + (size_t)t * a.frame_step_stride : nullptr;
        frame_and_buoyancy(frame, t + 1 < t_end, mrow);
        __syncthreads();
        FZ_TICK(6);
    }

    // ---- write back: pressure from registers, u, v, density from shared memory ------------------------------
#pragma unroll
    for (int r = 0; r < FZ_R; ++r) {
        const int i = r0 + r;
        if (i < h && colin) *reinterpret_cast<float4*>(gP + (size_t)i * pc + c0) = packed_row(P, r);
    }
    {
        const int gu4 = pu >> 2, gv4 = pv >> 2, gc4 = pc >> 2;
        for (int k = tid; k < (h + 1) * gu4; k += FZ_THREADS) {
            const int i = k / gu4, g = k - i * gu4;
            *reinterpret_cast<float4*>(gU + (size_t)i * pu + 4 * g) = zlds4(su + i * FZ_PU + 4 * g);
        }
        for (int k = tid; k < h * gv4; k += FZ_THREADS) {
            const int i = k / gv4, g = k - i * gv4;
            *reinterpret_cast<float4*>(gV + (size_t)i * pv + 4 * g) = zlds4(sv + i * FZ_PV + 4 * g);
        }
        for (int k = tid; k < h * gc4; k += FZ_THREADS) {
            const int i = k / gc4, g = k - i * gc4;
            *reinterpret_cast<float4*>(gD + (size_t)i * pc + 4 * g) = zlds4(sd + i * FZ_PD + 4 * g);
        }
    }
    if (t_end < a.nsteps) {             // the rest of this simulation runs in another CTA: publish the state
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.progress + b), "r"((unsigned)t_end) : "memory");
        }
    }
    FZ_TICK(7);
}

// =====================================================================================================================
// k_step_cluster<NC> -- ONE 128 x 128 simulation on a cluster of NC (2 or 4) CTAs, i.e. on NC SMs.
//
// k_step_fused gives a simulation one SM: with fewer simulations than SMs (BASELINE config 2 strong-scaled over 8 GPUs is 32
// per GPU; a single interactive simulation is 1) most of the chip idles and a step takes the 41 us one SM needs.  Here the rows
// are split over the CTAs of a thread-block cluster: CTA c owns rows [c RB, (c+1) RB), RB = 128 / NC, keeps them -- plus CH
// halo rows on either side -- in its shared memory and its part of the pressure in registers, exactly as k_step_fused does for
// the whole grid (same strip / cyclic mappings, same per-cell arithmetic: results are bit-identical).  What crosses CTAs goes
// through distributed shared memory:
//   * after a phase has rewritten a field, a CTA stores its first / last CH rows into the neighbours' halo rows (generic
//     stores through cluster.map_shared_rank) and the cluster barrier publishes them;
//   * a Jacobi sweep computes its boundary rows first, stores them into the neighbouring warps' (shared memory) and CTAs'
//     (DSMEM) halo lines, ARRIVES on the cluster barrier, computes the interior rows, then WAITS: the ~380-cycle barrier
//     latency hides under the interior arithmetic, and the barrier doubles as the CTA barrier of the sweep;
//   * an advection back-trace that leaves the stored rows (more than about CH cells) gathers from the owner CTA's shared
//     memory directly (ld through map_shared_rank); a cluster barrier separates everybody's gathers from the write-back.
// Fields are addressed by GLOBAL row through "virtual" base pointers (storage base minus the first stored row), so the cell
// code is k_step_fused's with r0 = c RB + 8 warp.
// =====================================================================================================================
constexpr int CH = 2;                                   // halo rows kept on either side of a CTA's band

template <int NC> struct ClusterCfg {
    // 16 warps per CTA whatever the split: a warp owns R = 8 / NC rows (lane = 4 columns), a thread 4 R cells.  (A first version
    // kept k_step_fused's 8 rows per warp, i.e. 16 / NC warps per CTA: one warp per scheduler cannot hide the shared-memory and
    // FP latencies of this code, and four SMs ran a simulation SLOWER than one -- 49 against 40 us per step.)
    static constexpr int RB = 128 / NC, NW = 16, NT = NW * 32, R = RB / NW;
    static constexpr int SU_ROWS = RB + 2 * CH + 1, SV_ROWS = RB + 2 * CH, SD_ROWS = RB + 2 * CH;
    static constexpr int SU = SU_ROWS * FZ_PU, SV = SV_ROWS * FZ_PV, SD = SD_ROWS * FZ_PD;       // floats
    static constexpr size_t SMEM = (size_t)(SU + SV + SD) * 4 + sizeof(float4) * (2 * 2 * NW * 32 + 2 * 2 * 2 * 32) + 4 * sizeof(unsigned long long);
};

// zdiffuse_strip for a strip of R rows
template <int PITCH, int R>
__device__ __forceinline__ void cdiffuse_strip(const float* F, const int rows, const int cols, const float c, float4 (&out)[R],
                                               const int r0, const int c0, const int lane)
{
    const int rlast = rows - 1, clast = cols - 1;
    float4 up = zlds4(F + min(max(r0 - 1, 0), rlast) * PITCH + c0);
    float4 cur = zlds4(F + min(r0, rlast) * PITCH + c0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r0 + r;
        const float4 dn = zlds4(F + min(i + 1, rlast) * PITCH + c0);
        float left = __shfl_up_sync(0xffffffffu, cur.w, 1);
        float right = __shfl_down_sync(0xffffffffu, cur.x, 1);
        if (lane == 0) left = cur.x;
        if (lane == 31 && clast >= 128) right = F[min(i, rlast) * PITCH + 128];       // the staggered column of v
        out[r].x = zdiff1(cur.x, up.x, dn.x, left, (c0 + 1 <= clast) ? cur.y : cur.x, c);
        out[r].y = zdiff1(cur.y, up.y, dn.y, cur.x, (c0 + 2 <= clast) ? cur.z : cur.y, c);
        out[r].z = zdiff1(cur.z, up.z, dn.z, cur.y, (c0 + 3 <= clast) ? cur.w : cur.z, c);
        out[r].w = zdiff1(cur.w, up.w, dn.w, cur.z, (c0 + 4 <= clast) ? right : cur.w, c);
        up = cur; cur = dn;
    }
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
// Execution barrier only (no release fence: ptxas implements the cluster-scope release as MEMBAR.ALL.GPU, ~0.7 us here): for
// the "everybody has finished READING the old values" points, where no data has to become visible.
__device__ __forceinline__ void cluster_sync_exec()
{
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned csmem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared::cluster address of the same location in CTA `rank`
__device__ __forceinline__ unsigned cluster_map32(const unsigned local, const unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
// 16 bytes into another CTA's shared memory; the store itself signals the destination mbarrier (complete_tx 16 bytes): no fence,
// no separate arrive -- data and signal travel together (STAS.128)
__device__ __forceinline__ void st_async_f4(const unsigned remote_data, const float4 v, const unsigned remote_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 :: "r"(remote_data), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
                    "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void cmbar_wait(const unsigned bar, const unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cmbar_arm(const unsigned bar, const unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// the same shared-memory location in CTA `rank` of this cluster, as a generic pointer
template <class T>
__device__ __forceinline__ T* cluster_map(T* local, const unsigned rank)
{
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<unsigned long long>(local)), "r"(rank));
    return reinterpret_cast<T*>(out);
}

// a field of the cluster kernel: virtual bases (index by GLOBAL row) of this CTA's copy and of the two neighbours' copies
struct CField {
    float* mine; float* up; float* dn;       // up / dn are NULL at the ends of the cluster
    int first, last;                          // global rows [first, last] this CTA owns
};
// store one float4 / float of an owned row, and into the neighbours' halo rows when the row is one they keep
template <int PITCH>
__device__ __forceinline__ void cput4(const CField& f, const int g, const int x, const float4 v)
{
    zsts4(f.mine + g * PITCH + x, v);
    if (f.up && g < f.first + CH) zsts4(f.up + g * PITCH + x, v);
    if (f.dn && g > f.last - CH) zsts4(f.dn + g * PITCH + x, v);
}
template <int PITCH>
__device__ __forceinline__ void cput1(const CField& f, const int g, const int x, const float v)
{
    f.mine[g * PITCH + x] = v;
    if (f.up && g < f.first + CH) f.up[g * PITCH + x] = v;
    if (f.dn && g > f.last - CH) f.dn[g * PITCH + x] = v;
}

// zadvect_pair (above) for a CTA that stores rows [slo, shi] of the advected field: same operations in the same order;
// a pair whose four corner rows are not all stored gathers from the owner CTAs through `far`.
template <int PITCH, bool SCALE, int NC, class Far>
__device__ __forceinline__ float2 cadvect_pair(const float* F, const int rows, const int cols, const float* su, const float* sv,
                                               const int i, const int j, const float dt, const float scale,
                                               const float slo, const float shi, const Far& far)
{
    constexpr int h = 128, w = 128;
    const int j1 = j + 32;
    const bool cu0 = (j <= w - 2 && i <= h - 1), cu1 = (j1 <= w - 2 && i <= h - 1);
    const bool cv0 = (i <= h - 2 && j <= w - 1), cv1 = (i <= h - 2 && j1 <= w - 1);
    const float2 ua = make_float2(cu0 ? su[i * FZ_PU + j] : 0.f, cu1 ? su[i * FZ_PU + j1] : 0.f);
    const float2 ub = make_float2(cu0 ? su[i * FZ_PU + j + 1] : 0.f, cu1 ? su[i * FZ_PU + j1 + 1] : 0.f);
    const float2 va = make_float2(cv0 ? sv[i * FZ_PV + j] : 0.f, cv1 ? sv[i * FZ_PV + j1] : 0.f);
    const float2 vb = make_float2(cv0 ? sv[(i + 1) * FZ_PV + j] : 0.f, cv1 ? sv[(i + 1) * FZ_PV + j1] : 0.f);
    const float2 half2 = make_float2(0.5f, 0.5f), dt2 = make_float2(dt, dt), one2 = make_float2(1.0f, 1.0f);
    const float2 hua = __fmul2_rn(half2, ua), hub = __fmul2_rn(half2, ub);
    const float2 hva = __fmul2_rn(half2, va), hvb = __fmul2_rn(half2, vb);
    const float2 ui = make_float2(hua.x + hub.x, hua.y + hub.y);
    const float2 vi = make_float2(hva.x + hvb.x, hva.y + hvb.y);
    const float xmax = (float)(cols - 1), ymax = (float)(rows - 1);
    const float2 du = __fmul2_rn(dt2, ui), dv = __fmul2_rn(dt2, vi);
    float2 px = make_float2((float)j - du.x, (float)j1 - du.y);
    float2 py = make_float2((float)i - dv.x, (float)i - dv.y);
    px.x = fminf(fmaxf(px.x, 0.0f), xmax); px.y = fminf(fmaxf(px.y, 0.0f), xmax);
    py.x = fminf(fmaxf(py.x, 0.0f), ymax); py.y = fminf(fmaxf(py.y, 0.0f), ymax);
    const float2 fx0 = make_float2(floorf(px.x), floorf(px.y)), fy0 = make_float2(floorf(py.x), floorf(py.y));
    float2 fx1 = __fadd2_rn(fx0, one2), fy1 = __fadd2_rn(fy0, one2);
    fx1.x = fminf(fx1.x, xmax); fx1.y = fminf(fx1.y, xmax);
    fy1.x = fminf(fy1.x, ymax); fy1.y = fminf(fy1.y, ymax);
    const float2 ax = __fadd2_rn(fx1, zneg2(px)), bx = __fadd2_rn(px, zneg2(fx0));
    const float2 ay = __fadd2_rn(fy1, zneg2(py)), by = __fadd2_rn(py, zneg2(fy0));
    const int dx0 = (fx1.x != fx0.x) ? 1 : 0, dy0 = (fy1.x != fy0.x) ? 1 : 0;
    const int dx1 = (fx1.y != fx0.y) ? 1 : 0, dy1 = (fy1.y != fy0.y) ? 1 : 0;
    float2 f00, f01, f10, f11;
    if (fy0.x >= slo && fy1.x <= shi && fy0.y >= slo && fy1.y <= shi) {
        const float* q0 = F + (int)__fmaf_rn(fy0.x, (float)PITCH, fx0.x);
        const float* q1 = F + (int)__fmaf_rn(fy0.y, (float)PITCH, fx0.y);
        f00 = make_float2(q0[0], q1[0]); f01 = make_float2(q0[dx0], q1[dx1]);
        f10 = make_float2(q0[dy0 * PITCH], q1[dy1 * PITCH]); f11 = make_float2(q0[dy0 * PITCH + dx0], q1[dy1 * PITCH + dx1]);
    } else {
        const int ya = (int)fy0.x, xa = (int)fx0.x, yb = (int)fy0.y, xb = (int)fx0.y;
        f00 = make_float2(far(ya, xa), far(yb, xb)); f01 = make_float2(far(ya, xa + dx0), far(yb, xb + dx1));
        f10 = make_float2(far(ya + dy0, xa), far(yb + dy1, xb)); f11 = make_float2(far(ya + dy0, xa + dx0), far(yb + dy1, xb + dx1));
    }
    const float2 t00 = __fmul2_rn(__fmul2_rn(ax, ay), f00), t01 = __fmul2_rn(__fmul2_rn(bx, ay), f01);
    const float2 t10 = __fmul2_rn(__fmul2_rn(ax, by), f10), t11 = __fmul2_rn(__fmul2_rn(bx, by), f11);
    float2 s = make_float2(t00.x + t01.x, t00.y + t01.y);
    s.x = s.x + t10.x; s.y = s.y + t10.y;
    s.x = s.x + t11.x; s.y = s.y + t11.y;
    if (SCALE) s = __fmul2_rn(s, make_float2(scale, scale));
    return s;
}

template <int NC>
__global__ void __launch_bounds__(ClusterCfg<NC>::NT, 1)
k_step_cluster(const FusedArgs a)
{
    typedef ClusterCfg<NC> C;
    constexpr int RB = C::RB, NW = C::NW, NT = C::NT, R = C::R;
    constexpr int h = 128, w = 128, pu = 128, pv = 132, pc = 128;
    extern __shared__ __align__(16) float smem[];
    float* su_s = smem;                       // storage: row 0 of each array is global row g0 = c RB - CH
    float* sv_s = su_s + C::SU;
    float* sd_s = sv_s + C::SV;
    float4 (*halo)[2][NW][32] = reinterpret_cast<float4 (*)[2][NW][32]>(sd_s + C::SD);
    // rows from the neighbouring CTAs: xp[slot][0: from above, 1: from below][0: the row nearer to this band, 1: the one beyond][lane]
    float4 (*xp)[2][2][32] = reinterpret_cast<float4 (*)[2][2][32]>(reinterpret_cast<float4*>(sd_s + C::SD) + 2 * 2 * NW * 32);
    unsigned long long* mb = reinterpret_cast<unsigned long long*>(reinterpret_cast<float4*>(sd_s + C::SD) + 2 * 2 * NW * 32 + 2 * 2 * 2 * 32);   // mb[slot * 2 + side]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = (int)cluster_rank();
    const size_t b = blockIdx.x / NC;
    const int g0 = c * RB - CH;               // first stored row (may be negative: rows outside the grid are never touched)
    const int r0 = c * RB + warp * R, c0 = lane * 4;
    const bool last_cta = c == NC - 1;
    float* __restrict__ gU = a.U + b * a.su_;
    float* __restrict__ gV = a.V + b * a.sv_;
    float* __restrict__ gD = a.D + b * a.sc_;
    float* __restrict__ gP = a.P + b * a.sc_;
    const float dt = a.dt;
#ifdef SMK_FUSED_TIMING
    long long tick0_ = clock64();
#define CL_TICK(k) do { if (tid == 0 && blockIdx.x == 0) { const long long t1_ = clock64(); a.ticks[k] += t1_ - tick0_; tick0_ = t1_; } } while (0)
#else
#define CL_TICK(k) do { } while (0)
#endif

    CField fu, fv, fd;
    fu.mine = su_s - g0 * FZ_PU; fv.mine = sv_s - g0 * FZ_PV; fd.mine = sd_s - g0 * FZ_PD;
    fu.up = fv.up = fd.up = fu.dn = fv.dn = fd.dn = nullptr;
    if (c > 0) {
        const int gn = (c - 1) * RB - CH;
        fu.up = cluster_map(su_s, c - 1) - gn * FZ_PU; fv.up = cluster_map(sv_s, c - 1) - gn * FZ_PV; fd.up = cluster_map(sd_s, c - 1) - gn * FZ_PD;
    }
    if (!last_cta) {
        const int gn = (c + 1) * RB - CH;
        fu.dn = cluster_map(su_s, c + 1) - gn * FZ_PU; fv.dn = cluster_map(sv_s, c + 1) - gn * FZ_PV; fd.dn = cluster_map(sd_s, c + 1) - gn * FZ_PD;
    }
    fu.first = fv.first = fd.first = c * RB;
    fv.last = fd.last = c * RB + RB - 1;
    fu.last = last_cta ? 128 : c * RB + RB - 1;
    float* const su = fu.mine; float* const sv = fv.mine; float* const sd = fd.mine;
    // Jacobi rows between CTAs, TWO rows per side every two sweeps (see the sweep loop): this CTA's first two rows go to the upper
    // neighbour's xp[slot][1][.][lane] and signal its mb[slot][1]; the last two rows go to the lower neighbour's xp[slot][0][.][lane]
    // / mb[slot][0] (st.async: data + signal).  slot = exchange number & 1.
#ifdef SMK_CL_NOEXCH                          // probe builds only (tools/micro/cluster_probe.cu): what a sweep costs without the rows from the neighbouring CTAs
    const bool has_up = false, has_dn = false;
#else
    const bool has_up = c > 0, has_dn = !last_cta;
#endif
    const unsigned rxp_up = has_up ? cluster_map32(csmem(&xp[0][1][0][lane]), c - 1) : 0u, rmb_up = has_up ? cluster_map32(csmem(&mb[1]), c - 1) : 0u;
    const unsigned rxp_dn = has_dn ? cluster_map32(csmem(&xp[0][0][0][lane]), c + 1) : 0u, rmb_dn = has_dn ? cluster_map32(csmem(&mb[0]), c + 1) : 0u;
    constexpr unsigned XP_SLOT = 2 * 2 * 32 * sizeof(float4), XP_ROW = 32 * sizeof(float4), MB_SLOT = 2 * sizeof(unsigned long long);
    const unsigned mb_local = csmem(mb);
    unsigned nex = 0;                         // exchanges posted so far (the same number in every thread of the cluster)
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb_local + 8u * k) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // two rows of 32 lanes x 16 bytes; every barrier is armed for its first use here and re-armed by its consumer after each use
        for (int k = 0; k < 4; ++k)
            if ((k & 1) ? has_dn : has_up) cmbar_arm(mb_local + 8u * k, 1024u);
    }
    // rows of the stored windows that exist in the grid
    const int slo = max(g0, 0);
    // (u keeps one more storage row than the CH halo rows below the band -- it is row 128 for the last CTA -- but only the halo
    //  rows are refreshed by the neighbour: the row after them is stale and must not be gathered from)
    const int shi_u = last_cta ? 128 : c * RB + RB + CH - 1, shi_c = min(g0 + C::SV_ROWS - 1, 127);
    // u's staggered row 128 is owned by the last CTA: its 128 columns are spread over lanes 0..7 of that CTA's 16 warps
    const int xe = 8 * warp + lane;
    const bool xu = last_cta && lane < 8;
    // v's staggered column 128: one cell per owned row, lanes 0..R-1 of the warp that owns the row
    const bool xv = lane < R;
    const int xr = r0 + lane;

    // ---- load: stored rows of u, v, density (own rows and the halo rows, all straight from global memory), own pressure rows
    for (int k = tid; k < (shi_u - slo + 1) * (pu / 4); k += NT) {
        const int i = slo + k / (pu / 4), g = k % (pu / 4);
        zsts4(su + i * FZ_PU + 4 * g, __ldcg(reinterpret_cast<const float4*>(gU + (size_t)i * pu + 4 * g)));
    }
    for (int k = tid; k < (shi_c - slo + 1) * (pv / 4); k += NT) {
        const int i = slo + k / (pv / 4), g = k % (pv / 4);
        zsts4(sv + i * FZ_PV + 4 * g, __ldcg(reinterpret_cast<const float4*>(gV + (size_t)i * pv + 4 * g)));
    }
    for (int k = tid; k < (shi_c - slo + 1) * (pc / 4); k += NT) {
        const int i = slo + k / (pc / 4), g = k % (pc / 4);
        zsts4(sd + i * FZ_PD + 4 * g, __ldcg(reinterpret_cast<const float4*>(gD + (size_t)i * pc + 4 * g)));
    }
    float4 P[R];
#pragma unroll
    for (int r = 0; r < R; ++r) P[r] = __ldcg(reinterpret_cast<const float4*>(gP + (size_t)(r0 + r) * pc + c0));
    float cm[4];                              // 0.25 inside, 0 on the Dirichlet ring columns
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) cm[cc] = (c0 + cc >= 1 && c0 + cc <= w - 2) ? 0.25f : 0.f;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // buoyancy (:154-155) on the owned rows of v and, redundantly, on the stored halo rows (they stay current without an exchange);
    // the frame of the step just finished comes from the owned rows of density
    auto frame_and_buoyancy = [&](float* frame, const bool buoy, const float4 (&mrow)[R]) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = r0 + r;
            const float4 d4 = zlds4(sd + i * FZ_PD + c0);
            if (frame) {
                float4 fr = d4;
                if (a.fmul) {
                    const float4 m = mrow[r];
                    fr.x = fr.x + m.x * fr.x; fr.y = fr.y + m.y * fr.y; fr.z = fr.z + m.z * fr.z; fr.w = fr.w + m.w * fr.w;
                }
                *reinterpret_cast<float4*>(frame + (size_t)i * pc + c0) = fr;
            }
            if (buoy) {
                float4 v4 = zlds4(sv + i * FZ_PV + c0);
                v4.x = v4.x + dt * (d4.x * 0.1f); v4.y = v4.y + dt * (d4.y * 0.1f);
                v4.z = v4.z + dt * (d4.z * 0.1f); v4.w = v4.w + dt * (d4.w * 0.1f);
                zsts4(sv + i * FZ_PV + c0, v4);
            }
        }
        if (buoy && tid < 2 * CH * 32) {                                 // the 2 CH halo rows: one float4 per thread
            const int hr = tid >> 5;
            const int i = hr < CH ? c * RB - CH + hr : c * RB + RB + (hr - CH);
            if (i >= 0 && i <= h - 1) {
                const float4 d4 = zlds4(sd + i * FZ_PD + c0);
                float4 v4 = zlds4(sv + i * FZ_PV + c0);
                v4.x = v4.x + dt * (d4.x * 0.1f); v4.y = v4.y + dt * (d4.y * 0.1f);
                v4.z = v4.z + dt * (d4.z * 0.1f); v4.w = v4.w + dt * (d4.w * 0.1f);
                zsts4(sv + i * FZ_PV + c0, v4);
            }
        }
    };
    {
        const float4 none[R] = {};
        frame_and_buoyancy(nullptr, true, none);
    }
    cluster_sync_all();          // every CTA of the cluster is past its loads before anybody stores into a neighbour
    CL_TICK(0);

    // gathers of back-traces that leave the stored rows: from the owner CTA's shared memory (its OWN rows are current)
    auto far_u = [&](const int y, const int x) { const int rk = min(y / RB, NC - 1); return cluster_map(su_s, rk)[(y - (rk * RB - CH)) * FZ_PU + x]; };
    auto far_v = [&](const int y, const int x) { const int rk = y / RB; return cluster_map(sv_s, rk)[(y - (rk * RB - CH)) * FZ_PV + x]; };
    auto far_d = [&](const int y, const int x) { const int rk = y / RB; return cluster_map(sd_s, rk)[(y - (rk * RB - CH)) * FZ_PD + x]; };

    for (int t = 0; t < a.nsteps; ++t) {
        // ---- a4 diffusion of u, v, density: strip mapping, halo rows give the rows above / below the band          :158-160
        {
            float4 Rs[R];
            float X = 0.f;
            cdiffuse_strip<FZ_PU, R>(su, h + 1, w, a.c_uv, Rs, r0, c0, lane);
            if (xu) X = zdiff_cell(su, FZ_PU, h + 1, w, 128, xe, a.c_uv);
            cluster_sync_exec();                       // everybody has read the old u (own rows and halo copies)
#pragma unroll
            for (int r = 0; r < R; ++r) cput4<FZ_PU>(fu, r0 + r, c0, Rs[r]);
            if (xu) su[128 * FZ_PU + xe] = X;
            cdiffuse_strip<FZ_PV, R>(sv, h, w + 1, a.c_uv, Rs, r0, c0, lane);
            if (xv) X = zdiff_cell(sv, FZ_PV, h, w + 1, xr, 128, a.c_uv);
            cluster_sync_exec();
#pragma unroll
            for (int r = 0; r < R; ++r) cput4<FZ_PV>(fv, r0 + r, c0, Rs[r]);
            if (xv) cput1<FZ_PV>(fv, xr, 128, X);
            cdiffuse_strip<FZ_PD, R>(sd, h, w, a.c_d, Rs, r0, c0, lane);
            cluster_sync_exec();
#pragma unroll
            for (int r = 0; r < R; ++r) cput4<FZ_PD>(fd, r0 + r, c0, Rs[r]);
            cluster_sync_all();                       // the new rows and halo rows are visible
        }
        CL_TICK(1);

        // ---- a5 divergence into registers                                                                          :136
        float4 Dv[R];
        {
            float4 ua = zlds4(su + r0 * FZ_PU + c0);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = r0 + r;
                const float4 ub = zlds4(su + (i + 1) * FZ_PU + c0);
                const float4 va = zlds4(sv + i * FZ_PV + c0);
                const float vr = sv[i * FZ_PV + c0 + 4];
                float4 o;
                o.x = ((ub.x - ua.x) + va.y) - va.x;
                o.y = ((ub.y - ua.y) + va.z) - va.y;
                o.z = ((ub.z - ua.z) + va.w) - va.z;
                o.w = ((ub.w - ua.w) + vr) - va.w;
                Dv[r] = zdiv4(o, dt);
                ua = ub;
            }
        }
        // ... and of the row just above / below the band (halo rows of u and v), which the first and the last warp sweep redundantly
        float4 Dx = zero4;
        if ((warp == 0 && has_up) || (warp == NW - 1 && has_dn)) {
            const int i = warp == 0 ? c * RB - 1 : c * RB + RB;
            const float4 ua = zlds4(su + i * FZ_PU + c0), ub = zlds4(su + (i + 1) * FZ_PU + c0);
            const float4 va = zlds4(sv + i * FZ_PV + c0);
            const float vr = sv[i * FZ_PV + c0 + 4];
            float4 o;
            o.x = ((ub.x - ua.x) + va.y) - va.x;
            o.y = ((ub.y - ua.y) + va.z) - va.y;
            o.z = ((ub.z - ua.z) + va.w) - va.z;
            o.w = ((ub.w - ua.w) + vr) - va.w;
            Dx = zdiv4(o, dt);
        }

        CL_TICK(2);
        // ---- a6 K Jacobi sweeps                                                                                   :139-145
        // Within a CTA: a warp computes its boundary rows first and posts them to the neighbouring warps through shared memory, the CTA
        // barrier at the end of the sweep publishes them (two copies of every line, by sweep parity).
        // Between CTAs: a sweep of a band needs the neighbouring band's boundary row of the sweep before, and one SM -> SM hop costs
        // ~700 cycles from st.async to the waiter's wake-up -- with an exchange per sweep a sweep took 845 cycles, 60 % of the step
        // (tools/micro/cluster_probe.cu; a cluster barrier per sweep was worse still: barrier.cluster.arrive.release compiles to
        // MEMBAR.ALL.GPU + UCGABAR, ~0.7 us; {value, tag} 8-byte DSMEM stores polled by the reader took 1420 cycles per sweep).
        // So the hop is paid every SECOND sweep: an exchange carries the two rows next to the band (state s, s even); the first
        // and the last warp sweep the row just outside the band redundantly in sweep s (they hold its divergence, Dx), which
        // gives them the state s + 1 of that row for sweep s + 1; after sweep s + 1 the next two rows travel.  Two slots
        // (exchange number & 1): a slot is rewritten two exchanges after it was read, and its writer cannot be two exchanges
        // ahead of its reader (it needs the reader's rows of the exchange in between).
        auto post_cta = [&]() {                             // this band's first / last two rows (current state) to the neighbours
            const unsigned slot = nex & 1u;
            if (warp == 0 && has_up) {                      // they are the upper neighbour's rows below its band: nearer row first
                st_async_f4(rxp_up + slot * XP_SLOT, P[0], rmb_up + slot * MB_SLOT);
                st_async_f4(rxp_up + slot * XP_SLOT + XP_ROW, P[1], rmb_up + slot * MB_SLOT);
            }
            if (warp == NW - 1 && has_dn) {                 // the lower neighbour's rows above its band: nearer row (my last) first
                st_async_f4(rxp_dn + slot * XP_SLOT, P[R - 1], rmb_dn + slot * MB_SLOT);
                st_async_f4(rxp_dn + slot * XP_SLOT + XP_ROW, P[R - 2], rmb_dn + slot * MB_SLOT);
            }
            ++nex;
        };
        // the latest exchange from the neighbour on `side` (0 above, 1 below): wait for it, read both rows, re-arm the barrier
        auto from_cta = [&](const int side, float4& near_row, float4& far_row) {
            const unsigned x = nex - 1u, slot = x & 1u;     // slot k is used by exchanges k, k + 2, ...: the m-th use has phase parity m & 1
            cmbar_wait(mb_local + 8u * (slot * 2 + side), (x >> 1) & 1u);
            near_row = xp[slot][side][0][lane];
            far_row = xp[slot][side][1][lane];
            __syncwarp();
            if (lane == 0) cmbar_arm(mb_local + 8u * (slot * 2 + side), 1024u);
        };
        static_assert(R >= 2, "an exchange carries two rows of one warp");
        halo[0][0][warp][lane] = P[0];
        halo[0][1][warp][lane] = P[R - 1];
        post_cta();
        __syncthreads();
        float4 X1 = zero4;                                  // first warp: state of the row above the band; last warp: of the row below it
        // A sweep is issue-bound here (4 warps per scheduler, ~40 flops per thread), so everything that does not change from sweep to
        // sweep is hoisted: ring masks of the two boundary rows, the addresses of the lines, and the parity (two sweeps per trip).
        const bool ok0 = r0 >= 1 && r0 <= h - 2, okL = r0 + R - 1 >= 1 && r0 + R - 1 <= h - 2;
        const float m0[4] = {ok0 ? cm[0] : 0.f, ok0 ? cm[1] : 0.f, ok0 ? cm[2] : 0.f, ok0 ? cm[3] : 0.f};
        const float mL[4] = {okL ? cm[0] : 0.f, okL ? cm[1] : 0.f, okL ? cm[2] : 0.f, okL ? cm[3] : 0.f};
        const float4* const rd_up = &halo[0][1][warp > 0 ? warp - 1 : 0][lane];         // + HPAR for the lines of parity 1
        const float4* const rd_dn = &halo[0][0][warp < NW - 1 ? warp + 1 : 0][lane];
        float4* const wr_first = &halo[0][0][warp][lane];
        float4* const wr_last = &halo[0][1][warp][lane];
        constexpr int HPAR = 2 * NW * 32;                   // float4 between the two parities of a line
        const bool edge_up = warp == 0 && has_up, edge_dn = warp == NW - 1 && has_dn;
        auto sweep = [&](auto parity) {
            constexpr int q = decltype(parity)::value;
            float4 up = warp > 0 ? rd_up[q * HPAR] : zero4;
            float4 dn = warp < NW - 1 ? rd_dn[q * HPAR] : zero4;
            if (q == 0) {
                // even sweep: the neighbour's two rows arrive; sweep the nearer one here too (it is an interior row of the grid)
                if (edge_up) {
                    float4 x2;
                    from_cta(0, up, x2);
                    X1 = stencil_row(x2, up, P[0], Dx, cm[0], cm[1], cm[2], cm[3]);
                }
                if (edge_dn) {
                    float4 x2;
                    from_cta(1, dn, x2);
                    X1 = stencil_row(P[R - 1], dn, x2, Dx, cm[0], cm[1], cm[2], cm[3]);
                }
            } else {
                if (edge_up) up = X1;
                if (edge_dn) dn = X1;
            }
            // boundary rows first (for R = 2 every row is one), posted at once; then the rows in between.  A row on the ring (global
            // row 0 or 127) stays zero: its mask multiplies by 0 instead of 0.25, like in every other Jacobi kernel here.
            const float4 o0 = P[0], oL = P[R - 1];
            const float4 n0 = stencil_row(up, o0, P[1], Dv[0], m0[0], m0[1], m0[2], m0[3]);
            const float4 nL = stencil_row(P[R - 2], oL, dn, Dv[R - 1], mL[0], mL[1], mL[2], mL[3]);
            wr_first[(q ^ 1) * HPAR] = n0;
            wr_last[(q ^ 1) * HPAR] = nL;
            float4 prev = o0;
#pragma unroll
            for (int r = 1; r < R - 1; ++r) {                                      // interior rows are never ring rows (R divides the band)
                const float4 cur = P[r];
                const float4 below = (r < R - 2) ? P[r + 1] : oL;
                P[r] = stencil_row(prev, cur, below, Dv[r], cm[0], cm[1], cm[2], cm[3]);
                prev = cur;
            }
            P[0] = n0;
            P[R - 1] = nL;
            if (q == 1) post_cta();                         // the state after an even number of sweeps travels
#ifndef SMK_CL_NOBAR                          // probe builds only
            __syncthreads();
#endif
        };
        {
            int s = 0;
            for (; s + 1 < a.K; s += 2) {
                sweep(std::integral_constant<int, 0>());
                sweep(std::integral_constant<int, 1>());
            }
            if (s < a.K) sweep(std::integral_constant<int, 0>());
        }

        CL_TICK(3);
        // ---- a7 gradient subtract on the owned rows, then the neighbours' halo rows of u and v                   :148-149
        {
            // the row above this warp's strip after K sweeps: from the warp above (lines of parity K & 1); for the first warp of a band
            // from the CTA above -- after an odd K it is the row swept redundantly here (X1), after an even K it comes with the exchange
            // posted after the last sweep, which the last warp has to consume as well to keep its barrier in step
            const int q = a.K & 1;
            float4 pabove = warp > 0 ? halo[q][1][warp - 1][lane] : zero4;
            if (warp == 0 && has_up) {
                if (a.K & 1) pabove = X1;
                else { float4 x2; from_cta(0, pabove, x2); }
            }
            if (warp == NW - 1 && has_dn && !(a.K & 1)) { float4 x1, x2; from_cta(1, x1, x2); }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = r0 + r;
                const float4 pc4 = P[r];
                const float pleft = __shfl_up_sync(0xffffffffu, pc4.w, 1);
                float4 u4 = zlds4(su + i * FZ_PU + c0);
                float4 v4 = zlds4(sv + i * FZ_PV + c0);
                if (i >= 1) {
                    u4.x = u4.x - dt * (pc4.x - pabove.x); u4.y = u4.y - dt * (pc4.y - pabove.y);
                    u4.z = u4.z - dt * (pc4.z - pabove.z); u4.w = u4.w - dt * (pc4.w - pabove.w);
                }
                if (c0 >= 1) v4.x = v4.x - dt * (pc4.x - pleft);
                v4.y = v4.y - dt * (pc4.y - pc4.x); v4.z = v4.z - dt * (pc4.z - pc4.y); v4.w = v4.w - dt * (pc4.w - pc4.z);
                cput4<FZ_PU>(fu, i, c0, u4);
                cput4<FZ_PV>(fv, i, c0, v4);
                pabove = pc4;
            }
        }
        cluster_sync_all();

        CL_TICK(4);
        float4 mrow[R] = {};
        // ---- a10/a11 advection (cyclic mapping): u by (u, v); v by (u', v); density by (u', v'), decay              :166-171
        {
            float2 Ra[R][2];
            float X = 0.f;
            const float ulo = (float)slo, uhi = (float)shi_u, chi = (float)shi_c;
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp)
                    Ra[r][kp] = cadvect_pair<FZ_PU, false, NC>(su, h + 1, w, su, sv, r0 + r, lane + 64 * kp, dt, 1.0f, ulo, uhi, far_u);
            if (xu) X = zadvect_cell<FZ_PU>(su, h + 1, w, su, sv, h, w, 128, xe, dt);       // row 128 back-traces to itself (zero velocity there)
            cluster_sync_exec();                       // every gather of the cluster (near and far) is done
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    cput1<FZ_PU>(fu, i, j, Ra[r][kp].x);
                    cput1<FZ_PU>(fu, i, j + 32, Ra[r][kp].y);
                }
            if (xu) su[128 * FZ_PU + xe] = X;
            cluster_sync_all();

#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp)
                    Ra[r][kp] = cadvect_pair<FZ_PV, false, NC>(sv, h, w + 1, su, sv, r0 + r, lane + 64 * kp, dt, 1.0f, ulo, chi, far_v);
            // the staggered column 128 of v: both interpolated velocities are zero there (navier_stokes.py:97-109), so the cell
            // back-traces to itself and only touches its own row
            if (xv) X = zadvect_cell<FZ_PV>(sv, h, w + 1, su, sv, h, w, xr, 128, dt);
            cluster_sync_exec();
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    cput1<FZ_PV>(fv, i, j, Ra[r][kp].x);
                    cput1<FZ_PV>(fv, i, j + 32, Ra[r][kp].y);
                }
            if (xv) cput1<FZ_PV>(fv, xr, 128, X);
            cluster_sync_all();

#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp)
                    Ra[r][kp] = cadvect_pair<FZ_PD, true, NC>(sd, h, w, su, sv, r0 + r, lane + 64 * kp, dt, a.decay, ulo, chi, far_d);
            cluster_sync_exec();
            if (a.frames && a.fmul) {
#pragma unroll
                for (int r = 0; r < R; ++r) mrow[r] = __ldg(reinterpret_cast<const float4*>(a.fmul + (size_t)(r0 + r) * pc + c0));
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    const int i = r0 + r, j = lane + 64 * kp;
                    cput1<FZ_PD>(fd, i, j, Ra[r][kp].x);
                    cput1<FZ_PD>(fd, i, j + 32, Ra[r][kp].y);
                }
            cluster_sync_all();
        }

        CL_TICK(5);
        // ---- a11 returned copy of this step + buoyancy of the next
        float* frame = a.frames ? a.frames + b * a.frame_batch_stride + (size_t)t * a.frame_step_stride : nullptr;
        frame_and_buoyancy(frame, t + 1 < a.nsteps, mrow);
        __syncthreads();
        CL_TICK(6);
    }

    // ---- write back the owned rows: pressure from registers, u, v, density from shared memory ------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) *reinterpret_cast<float4*>(gP + (size_t)(r0 + r) * pc + c0) = P[r];
    {
        const int first = c * RB;
        for (int k = tid; k < (fu.last - first + 1) * (pu / 4); k += NT) {
            const int i = first + k / (pu / 4), g = k % (pu / 4);
            *reinterpret_cast<float4*>(gU + (size_t)i * pu + 4 * g) = zlds4(su + i * FZ_PU + 4 * g);
        }
        for (int k = tid; k < RB * (pv / 4); k += NT) {
            const int i = first + k / (pv / 4), g = k % (pv / 4);
            *reinterpret_cast<float4*>(gV + (size_t)i * pv + 4 * g) = zlds4(sv + i * FZ_PV + 4 * g);
        }
        for (int k = tid; k < RB * (pc / 4); k += NT) {
            const int i = first + k / (pc / 4), g = k % (pc / 4);
            *reinterpret_cast<float4*>(gD + (size_t)i * pc + 4 * g) = zlds4(sd + i * FZ_PD + 4 * g);
        }
    }
    cluster_sync_all();          // no CTA leaves (and frees its shared memory) while a neighbour may still read or write it
    CL_TICK(7);
#undef CL_TICK
}

// does the current device give one CTA the 226.5 KB the fused kernel needs (B200: 227 KB)?
static bool device_has_room()
{
    static int ok[64] = {};                 // 0 unknown, 1 yes, -1 no
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (ok[dev] == 0) {
        int optin = 0;
        ok[dev] = (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
                   (size_t)optin >= FZ_SMEM) ? 1 : -1;
    }
    return ok[dev] == 1;
}

bool fused_supported(const smk_grid_t* g)
{
    return g->h >= 2 && g->w >= 2 && g->h <= 128 && g->w <= 128 && g->pitch_u <= FZ_PU && g->pitch_v <= FZ_PV && g->pitch_c <= FZ_PD &&
           (g->gh == 0 || (g->gh == g->h && g->row0 == 0)) && device_has_room();
}

// Time-sliced schedule (see k_step_fused): steps per piece, or 0 for one CTA per simulation.  It pays when the simulations do
// not fill a whole number of CTA waves: ceil(total / SMs) step-times (+ about one for the extra state hand-overs) against
// ceil(nsims / SMs) x nsteps.  SMK_FUSED_SLICE = 0 disables it, a positive value forces that many steps per piece (tests).
static int pick_seg_len(const smk_grid_t* g, int nsteps, const void* scratch, cudaStream_t s)
{
    if (!scratch || !aligned16(scratch)) return 0;
    // the schedule is uploaded from pageable host memory, which a stream capture does not allow
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return 0;
    const int64_t total = (int64_t)g->batch * nsteps;
    if (total > 0x3fffffff) return 0;
    if (env().fused_slice != SMK_ENV_UNSET) return env().fused_slice > 0 ? env().fused_slice : 0;
    int dev = 0, nsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) return 0;
    if (g->batch <= nsm) return 0;
    const int64_t L = (total + nsm - 1) / nsm;
    const int64_t classic = (int64_t)((g->batch + nsm - 1) / nsm) * nsteps;
    return (L + 1 < classic) ? (int)L : 0;
}

// The items of the time-sliced schedule in planned start order: pieces of L steps of the simulation-major line, each walked
// from its end.  An item that does not begin at step 0 is preceded (lower index) by the item that holds the steps before it:
// that one starts earlier in a piece of lower index, or -- a simulation spanning three or more pieces -- is a whole piece.
struct FzPlanned { int4 item; int64_t start; int piece; };
static void plan_items(int nsims, int nsteps, int L, std::vector<int4>& out)
{
    const int64_t total = (int64_t)nsims * nsteps;
    std::vector<FzPlanned> v;
    for (int64_t lo = 0, c = 0; lo < total; lo += L, ++c) {
        int64_t x = std::min<int64_t>(lo + L, total), start = 0;
        while (x > lo) {
            const int64_t b = (x - 1) / nsteps, ilo = std::max<int64_t>(lo, b * nsteps);
            v.push_back({make_int4((int)b, (int)(ilo - b * nsteps), (int)(x - b * nsteps), 0), start, (int)c});
            start += x - ilo;
            x = ilo;
        }
    }
    // planned start time first; among equal start times the lower piece first (the item a later piece waits for)
    std::stable_sort(v.begin(), v.end(), [](const FzPlanned& p, const FzPlanned& q) { return p.start != q.start ? p.start < q.start : p.piece < q.piece; });
    out.resize(v.size());
    for (size_t k = 0; k < v.size(); ++k) out[k] = v[k].item;
}

int fused_plan(int nsims, int nsteps, int piece_len, int32_t* items, int capacity, int* count)
{
    std::vector<int4> plan;
    plan_items(nsims, nsteps, piece_len, plan);
    *count = (int)plan.size();
    if ((int)plan.size() > capacity) return fail(SMK_EINVAL, "smk_fused_plan: %zu items, capacity %d", plan.size(), capacity);
    for (size_t k = 0; k < plan.size(); ++k) {
        items[4 * k] = plan[k].x; items[4 * k + 1] = plan[k].y; items[4 * k + 2] = plan[k].z; items[4 * k + 3] = plan[k].w;
    }
    return SMK_OK;
}

// CTAs per simulation for a call of `batch` full-size (128 x 128) simulations: 0 = one CTA per simulation (k_step_fused), 2 or 4 =
// k_step_cluster.  SMK_FUSED_CLUSTER = 0 / 2 / 4 forces.
int pick_cluster(const int batch)
{
    int nsm = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) return 0;
    const int forced = env().fused_cluster;
    if (forced != SMK_ENV_UNSET) return (forced == 2 || forced == 4) ? forced : 0;
    // Measured (tools/cluster_latency.py, K = 40, us per step of a 20-step call): one CTA per simulation 39.7 at any batch up to 148;
    // clusters of four 27.3 - 28.3 up to 32 simulations, 55 from 48 on (a B200 co-schedules 33 clusters of four: a second wave);
    // clusters of two 41.2: no gain, because a sweep's critical path is the round trip of a boundary row between two SMs, not
    // the arithmetic.  So: four CTAs per simulation while all clusters are co-resident, else one.
    return (batch <= nsm / 4 - 4) ? 4 : 0;
}

template <int NC>
static int launch_cluster_nc(const int batch, const FusedArgs& a, cudaStream_t s)
{
    typedef ClusterCfg<NC> C;
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 63;
    if (!attr_set[dev] || dev == 63) {
        const cudaError_t e = cudaFuncSetAttribute(k_step_cluster<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return fail((int)e, "k_step_cluster: cannot opt in to %zu B of shared memory: %s", (size_t)C::SMEM, cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)batch * NC); cfg.blockDim = dim3(C::NT); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = s;
    cudaLaunchAttribute attr = {};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = NC; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    ProfScope prof_(SMK_PH_STEP_FUSED, s);
    (void)cudaLaunchKernelEx(&cfg, k_step_cluster<NC>, a);
    return check_launch("k_step_cluster");
}

static int launch_cluster(const int nc, const int batch, const FusedArgs& a, cudaStream_t s)
{
    return nc == 4 ? launch_cluster_nc<4>(batch, a, s) : launch_cluster_nc<2>(batch, a, s);
}

int launch_steps_fused(const smk_grid_t* g, float* u, float* v, float* d, float* p, int nsteps, float* frames,
                       int64_t frame_step_stride, int64_t frame_batch_stride, const float* fmul,
                       float dt, float c_uv, float c_d, float decay, int K, float* scratch, cudaStream_t s)
{
    if (!fused_supported(g)) return fail(SMK_EUNSUPPORTED, "k_step_fused: needs a non-slab grid of at most 128 x 128 cells, got %d x %d", g->h, g->w);
    if (nsteps <= 0) return SMK_OK;
    const bool full = g->h == 128 && g->w == 128 && g->pitch_u == 128 && g->pitch_v == 132 && g->pitch_c == 128;
    // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 63;
    if (!attr_set[dev] || dev == 63) {
        cudaError_t e = cudaFuncSetAttribute(k_step_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM);
        if (e != cudaSuccess) return fail((int)e, "k_step_fused: cannot opt in to %zu B of shared memory: %s", FZ_SMEM, cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    FusedArgs a;
    a.U = u; a.V = v; a.D = d; a.P = p; a.frames = frames; a.fmul = fmul;
    a.h = g->h; a.w = g->w; a.pu = g->pitch_u; a.pv = g->pitch_v; a.pc = g->pitch_c;
    a.su_ = g->stride_u; a.sv_ = g->stride_v; a.sc_ = g->stride_c;
    a.frame_step_stride = frame_step_stride; a.frame_batch_stride = frame_batch_stride;
    a.dt = dt; a.c_uv = c_uv; a.c_d = c_d; a.decay = decay; a.K = K; a.nsteps = nsteps;
    a.items = nullptr; a.progress = nullptr; a.ticket = nullptr; a.spin_budget = 1u << 25;
    if (full) {
        const int nc = pick_cluster(g->batch);
        if (nc) return launch_cluster(nc, g->batch, a, s);
    }
    int ctas = g->batch;
    const int seg_len = pick_seg_len(g, nsteps, scratch, s);
    if (seg_len > 0) {
        // `scratch` (the divergence array, which this kernel does not use) holds the per-simulation progress counters
        // and, behind them, the item table; the plan depends on (simulations, steps, piece length) only and is cached
        static thread_local std::vector<int4> plan;
        static thread_local int plan_key[3] = {0, 0, 0};
        if (plan_key[0] != g->batch || plan_key[1] != nsteps || plan_key[2] != seg_len) {
            plan_items(g->batch, nsteps, seg_len, plan);
            plan_key[0] = g->batch; plan_key[1] = nsteps; plan_key[2] = seg_len;
        }
        const size_t counters = (sizeof(unsigned) * ((size_t)g->batch + 1) + 15) & ~(size_t)15;      // progress[batch], ticket
        const size_t need = counters + sizeof(int4) * plan.size();
        const size_t have = sizeof(float) * (size_t)g->stride_c * (size_t)(g->batch - 1) + sizeof(float) * (size_t)g->h * (size_t)g->pitch_c;
        if (need <= have) {
            a.progress = reinterpret_cast<unsigned*>(scratch);
            a.ticket = a.progress + g->batch;
            const unsigned long long work = (unsigned long long)nsteps * (unsigned long long)(K > 0 ? K : 1);
            a.spin_budget = (unsigned)std::min<unsigned long long>(0xfffffff0ull, std::max<unsigned long long>(1ull << 25, work << 14));
            a.items = reinterpret_cast<const int4*>(reinterpret_cast<char*>(scratch) + counters);
            cudaError_t e = cudaMemsetAsync(a.progress, 0, counters, s);
            // pageable source: the runtime stages the table before it returns, the cached vector may change afterwards
            if (e == cudaSuccess) e = cudaMemcpyAsync(const_cast<int4*>(a.items), plan.data(), sizeof(int4) * plan.size(), cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return fail((int)e, "k_step_fused: upload of the time-sliced schedule: %s", cudaGetErrorString(e));
            ctas = (int)plan.size();
        }
    }
    ProfScope prof_(SMK_PH_STEP_FUSED, s);
    if (full) k_step_fused<true><<<ctas, FZ_THREADS, FZ_SMEM, s>>>(a);
    else      k_step_fused<false><<<ctas, FZ_THREADS, FZ_SMEM, s>>>(a);
    return check_launch("k_step_fused");
}

}  // namespace smk

// jacobi_probe.cu -- where does a Jacobi sweep of the register-strip kernel spend its cycles?  Times the sweep loop
// of one 128 x 128 tile per SM with pieces knocked out (results are then wrong: timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I smokephysai_b200/csrc -o build/jacobi_probe tools/micro/jacobi_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "jacobi_core.cuh"

namespace smk { int fail(int c, const char*, ...) { return c; } int check_launch(const char*) { return 0; } void prof_mark(int, cudaStream_t, bool) {} }
using namespace smk;

// DBG bit 0: no shuffles, bit 1: no CTA barrier, bit 2: no halo LDS/STS, bit 3: no FP (only exchange)
// STAG: 0 every warp runs the sweep in the same order (IFIRST); else two groups of warps run different orders, so that the shuffle /
// halo part of one group overlaps the arithmetic of the other: 1: (warp >> 2) & 1 picks ORDER 1 / 0 (two warps of each group on every
// scheduler), 2: warp & 1 (whole schedulers), 3: (warp >> 2) & 1 picks ORDER 2 / 1, 4: (warp >> 2) & 1 picks ORDER 2 / 0
template <int PMASK, int DBG, int NW, int IFIRST = 0, int STAG = 0>
__global__ void __launch_bounds__(NW * 32, 1)
k_probe(const float* __restrict__ pin, float* __restrict__ pout, const float* __restrict__ div, const int T)
{
    __shared__ float4 halo[2][2][NW][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gj = lane * 4, gi0 = warp * 8;
    const size_t boff = (size_t)blockIdx.x * 128 * 128;
    pin += boff; pout += boff; div += boff;
    PackedStrip A, B, ND;
    unsigned ringmask = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int gi = gi0 + r;
        const float4 p4 = *reinterpret_cast<const float4*>(pin + (size_t)gi * 128 + gj);
        const float4 d4 = *reinterpret_cast<const float4*>(div + (size_t)gi * 128 + gj);
        if (gi < 1 || gi > 126) ringmask |= 1u << r;
        packed_set_row(A, r, p4);
        packed_set_row(ND, r, make_float4(-d4.x, -d4.y, -d4.z, -d4.w));
    }
    float2 M[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { const float m = (gj + c >= 1 && gj + c <= 126) ? 0.25f : 0.f; M[c] = make_float2(m, m); }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    halo[0][0][warp][lane] = packed_row(A, 0);
    halo[0][1][warp][lane] = packed_row(A, 7);
    __syncthreads();
    for (int s = 0; s + 1 < T; s += 2) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 up = zero4, dn = zero4;
            if (!(DBG & 4)) {
                up = warp > 0 ? halo[half][1][warp - 1][lane] : zero4;
                dn = warp < NW - 1 ? halo[half][0][warp + 1][lane] : zero4;
            }
            float4* pf = (DBG & 4) ? nullptr : &halo[half ^ 1][0][warp][lane];
            float4* pl = (DBG & 4) ? nullptr : &halo[half ^ 1][1][warp][lane];
            if (STAG == 0) {
                if (half == 0) sweep_packed<PMASK, DBG, false, IFIRST>(A, B, ND, up, dn, M, ringmask, pf, pl);
                else           sweep_packed<PMASK, DBG, false, IFIRST>(B, A, ND, up, dn, M, ringmask, pf, pl);
            } else {
                constexpr int OA = (STAG == 3 || STAG == 4) ? 2 : 1, OB = STAG == 3 ? 1 : 0;
                const bool grp = STAG == 2 ? (warp & 1) : ((warp >> 2) & 1);
                if (grp) {
                    if (half == 0) sweep_packed<PMASK, DBG, false, OA>(A, B, ND, up, dn, M, ringmask, pf, pl);
                    else           sweep_packed<PMASK, DBG, false, OA>(B, A, ND, up, dn, M, ringmask, pf, pl);
                } else {
                    if (half == 0) sweep_packed<PMASK, DBG, false, OB>(A, B, ND, up, dn, M, ringmask, pf, pl);
                    else           sweep_packed<PMASK, DBG, false, OB>(B, A, ND, up, dn, M, ringmask, pf, pl);
                }
            }
            if (!(DBG & 2)) __syncthreads();
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) *reinterpret_cast<float4*>(pout + (size_t)(gi0 + r) * 128 + gj) = packed_row(A, r);
}

// neighbour-to-neighbour synchronisation instead of the CTA barrier: warp w waits only for the rows of warps w-1 / w+1
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
    asm volatile("{ .reg .pred p; WAIT_LOOP: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra WAIT_LOOP; }"
                 :: "r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int PMASK, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_probe_desync(const float* __restrict__ pin, float* __restrict__ pout, const float* __restrict__ div, const int T)
{
    __shared__ float4 halo[2][2][NW][32];
    __shared__ unsigned long long bar[2][NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gj = lane * 4, gi0 = warp * 8;
    const size_t boff = (size_t)blockIdx.x * 128 * 128;
    pin += boff; pout += boff; div += boff;
    PackedStrip A, B, ND;
    unsigned ringmask = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int gi = gi0 + r;
        const float4 p4 = *reinterpret_cast<const float4*>(pin + (size_t)gi * 128 + gj);
        const float4 d4 = *reinterpret_cast<const float4*>(div + (size_t)gi * 128 + gj);
        if (gi < 1 || gi > 126) ringmask |= 1u << r;
        packed_set_row(A, r, p4);
        packed_set_row(ND, r, make_float4(-d4.x, -d4.y, -d4.z, -d4.w));
    }
    float2 M[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { const float m = (gj + c >= 1 && gj + c <= 126) ? 0.25f : 0.f; M[c] = make_float2(m, m); }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < 2 * NW) mbar_init(&bar[0][0] + threadIdx.x, 32);
    __syncthreads();
    halo[0][0][warp][lane] = packed_row(A, 0);
    halo[0][1][warp][lane] = packed_row(A, 7);
    mbar_arrive(&bar[0][warp]);
    for (int g = 0; g + 1 < T; g += 2) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const unsigned parity = ((g + half) >> 1) & 1;
            float4 up = zero4, dn = zero4;
            if (warp > 0) { mbar_wait(&bar[half][warp - 1], parity); up = halo[half][1][warp - 1][lane]; }
            if (warp < NW - 1) { mbar_wait(&bar[half][warp + 1], parity); dn = halo[half][0][warp + 1][lane]; }
            float4* pf = &halo[half ^ 1][0][warp][lane];
            float4* pl = &halo[half ^ 1][1][warp][lane];
            if (half == 0) sweep_packed<PMASK, 0, true>(A, B, ND, up, dn, M, ringmask, pf, pl, &bar[half ^ 1][warp]);
            else           sweep_packed<PMASK, 0, true>(B, A, ND, up, dn, M, ringmask, pf, pl, &bar[half ^ 1][warp]);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) *reinterpret_cast<float4*>(pout + (size_t)(gi0 + r) * 128 + gj) = packed_row(A, r);
}

template <int PMASK>
void run_desync(const char* name, float* p, float* q, float* d, int nb)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[2];
    const int Ts[2] = {200, 1000};
    for (int k = 0; k < 2; ++k) {
        k_probe_desync<PMASK, 16><<<nb, 512>>>(p, q, d, Ts[k]);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_probe_desync<PMASK, 16><<<nb, 512>>>(p, q, d, Ts[k]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[k], e0, e1);
    }
    const double us_per_sweep = (ms[1] - ms[0]) * 1e3 / (Ts[1] - Ts[0]);
    printf("%-64s %7.4f us/sweep  (%5.0f cycles at 1965 MHz)  %s\n", name, us_per_sweep, us_per_sweep * 1965, cudaGetErrorString(cudaGetLastError()));
}

template <int PMASK, int DBG, int IFIRST = 0, int STAG = 0>
void run(const char* name, float* p, float* q, float* d, int nb)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[2];
    const int Ts[2] = {200, 1000};
    for (int k = 0; k < 2; ++k) {
        k_probe<PMASK, DBG, 16, IFIRST, STAG><<<nb, 512>>>(p, q, d, Ts[k]);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_probe<PMASK, DBG, 16, IFIRST, STAG><<<nb, 512>>>(p, q, d, Ts[k]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[k], e0, e1);
    }
    const double us_per_sweep = (ms[1] - ms[0]) * 1e3 / (Ts[1] - Ts[0]);
    printf("%-64s %7.4f us/sweep  (%5.0f cycles at 1965 MHz)  %s\n", name, us_per_sweep, us_per_sweep * 1965, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t n = (size_t)sms * 128 * 128;
    float *p, *q, *d;
    cudaMalloc(&p, n * 4); cudaMalloc(&q, n * 4); cudaMalloc(&d, n * 4);
    cudaMemset(p, 0, n * 4); cudaMemset(d, 0, n * 4);
    run<6, 0, 0, 1>("half packed, staggered: warp groups of 4 alternate interior-first / boundary-first", p, q, d, sms);
    run<7, 0, 0, 1>("mask 7, staggered: warp groups of 4 alternate interior-first / boundary-first", p, q, d, sms);
    run<15, 0, 0, 1>("packed, staggered: warp groups of 4 alternate interior-first / boundary-first", p, q, d, sms);
    run<6, 0, 0, 2>("half packed, staggered by warp parity (whole schedulers)", p, q, d, sms);
    run<6, 0, 0, 3>("half packed, staggered: groups of 4 alternate order 2 / order 1", p, q, d, sms);
    run<6, 0, 0, 4>("half packed, staggered: groups of 4 alternate order 2 / order 0", p, q, d, sms);
    run<7, 0, 0, 3>("mask 7, staggered: groups of 4 alternate order 2 / order 1", p, q, d, sms);
    run_desync<16>("scalar, neighbour mbarriers instead of the CTA barrier", p, q, d, sms);
    run_desync<6>("half packed, neighbour mbarriers", p, q, d, sms);
    run_desync<15>("packed, neighbour mbarriers", p, q, d, sms);
    run<16, 0, 1>("scalar, interior rows first", p, q, d, sms);
    run<6, 0, 1>("half packed, interior rows first", p, q, d, sms);
    run<7, 0, 1>("mask 7, interior rows first", p, q, d, sms);
    run<15, 0, 1>("packed, interior rows first", p, q, d, sms);
    run<6, 0, 2>("half packed, order pair1 / boundary+post / pair2", p, q, d, sms);
    run<15, 0, 2>("packed, order pair1 / boundary+post / pair2", p, q, d, sms);
    run<16, 0, 2>("scalar, order pair1 / boundary+post / pair2", p, q, d, sms);
    run<16, 0>("scalar ping-pong, complete", p, q, d, sms);
    run<15, 0>("packed, complete", p, q, d, sms);
    run<6, 0>("half packed, complete", p, q, d, sms);
    run<16, 1>("scalar, no shuffles", p, q, d, sms);
    run<16, 2>("scalar, no barrier", p, q, d, sms);
    run<16, 4>("scalar, no halo LDS/STS", p, q, d, sms);
    run<16, 7>("scalar, FP only (no shuffles, barrier, halo)", p, q, d, sms);
    run<15, 7>("packed, FP only", p, q, d, sms);
    run<6, 7>("half packed, FP only", p, q, d, sms);
    run<16, 8>("no FP (shuffles + halo + barrier only)", p, q, d, sms);
    run<16, 9>("no FP, no shuffles (halo + barrier only)", p, q, d, sms);
    run<16, 6>("scalar, shuffles but no barrier, no halo", p, q, d, sms);
    run<15, 6>("packed, shuffles but no barrier, no halo", p, q, d, sms);
    run<15, 2>("packed, no barrier", p, q, d, sms);
    run<15, 1>("packed, no shuffles", p, q, d, sms);
    return 0;
}

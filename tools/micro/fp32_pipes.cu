// fp32_pipes.cu -- what the sm_100a FP32 pipes sustain per SM per clock for the instruction forms the Jacobi
// sweep can be written in (all bit-exact alternatives of add / mul).  nvcc -arch=sm_100a -o build/fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
#define NACC 16

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, float seed, float b)
{
    b = b + (float)threadIdx.x * 1e-9f; seed = seed + (float)threadIdx.x * 1e-9f;
    float x[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) x[k] = seed + threadIdx.x + k;
    float2 y[NACC / 2];
#pragma unroll
    for (int k = 0; k < NACC / 2; ++k) y[k] = make_float2(x[2 * k], x[2 * k + 1]);
    const float2 b2 = make_float2(b, b);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(b));
            if (MODE == 1) asm volatile("add.rn.f32 %0, %0, 0f3F800000;" : "+f"(x[k]));
            if (MODE == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(b));
            if (MODE == 3) asm volatile("mul.rn.f32 %0, %0, 0f3E800000;" : "+f"(x[k]));
            if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(b), "f"(seed));
            if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(x[k]) : "f"(b));
            if (MODE == 6 && (k & 1) == 0)
                asm volatile("{ .reg .b64 a, c; mov.b64 a, {%0, %1}; mov.b64 c, {%2, %3}; add.rn.f32x2 a, a, c; mov.b64 {%0, %1}, a; }"
                             : "+f"(y[k / 2].x), "+f"(y[k / 2].y) : "f"(b2.x), "f"(b2.y));
            if (MODE == 7) {       // alternate scalar add and packed add: 1 scalar + 1 packed per 3 results
                if ((k & 3) == 0)
                    asm volatile("{ .reg .b64 a, c; mov.b64 a, {%0, %1}; mov.b64 c, {%2, %3}; add.rn.f32x2 a, a, c; mov.b64 {%0, %1}, a; }"
                                 : "+f"(y[k / 2].x), "+f"(y[k / 2].y) : "f"(b2.x), "f"(b2.y));
                else if ((k & 3) >= 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(b));
            }
            if (MODE == 8) {       // alternate add r,r and fma r,imm,r
                if (k & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(b));
                else asm volatile("fma.rn.f32 %0, %0, 0f3F800000, %1;" : "+f"(x[k]) : "f"(b));
            }
            if (MODE == 10) asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+f"(x[k]));                       // FRND.FLOOR
            if (MODE == 11) { int t; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(t) : "f"(x[k])); asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(x[k]) : "r"(t)); }   // F2I + I2F
            if (MODE == 12) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(b));                  // FMNMX
            if (MODE == 13) { asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+f"(x[k])); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[(k + 3) % NACC]) : "f"(b)); }  // FRND + FADD
            if (MODE == 9) {       // add with two varying register operands (neighbour accumulators)
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(x[(k + 5) % NACC]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NACC; ++k) s += x[k];
#pragma unroll
    for (int k = 0; k < NACC / 2; ++k) s += y[k].x + y[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double results_per_thread_iter, int warps_per_sm)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 1024 * 2);
    const int threads = warps_per_sm >= 16 ? 512 : warps_per_sm * 32;
    const int blocks = sms * (warps_per_sm >= 16 ? warps_per_sm / 16 : 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 1.0f, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0f, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double results = results_per_thread_iter * ITER * (double)blocks * threads;
    const double per_clk_sm = results / (ms * 1e-3) / (khz * 1e3) / sms;
    printf("%-44s warps/SM %2d: %7.3f ms  %6.1f fp32 results / clk / SM (at %d MHz nominal)\n", name, warps_per_sm, ms, per_clk_sm, khz / 1000);
    cudaFree(out);
}

int main()
{
    for (int w : {16, 32}) {
        run<0>("FADD r,r", NACC, w);
        run<1>("FADD r,imm", NACC, w);
        run<2>("FMUL r,r", NACC, w);
        run<3>("FMUL r,imm", NACC, w);
        run<4>("FFMA r,r,r", NACC, w);
        run<5>("FFMA r,imm(1.0),r  (== FADD)", NACC, w);
        run<6>("FADD2 (add.f32x2)", NACC, w);
        run<7>("mix: 1 FADD2 + 2 FADD per 4 results", NACC, w);
        run<8>("mix: FADD r,r + FFMA r,1.0,r", NACC, w);
        run<9>("FADD r,r' (two varying regs)", NACC, w);
        run<10>("FRND.FLOOR", NACC, w);
        run<11>("F2I + I2F (counted as 2 results)", 2 * NACC, w);
        run<12>("FMNMX", NACC, w);
        run<13>("FRND + FADD interleaved (2 results)", 2 * NACC, w);
    }
    return 0;
}

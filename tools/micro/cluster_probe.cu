// cluster_probe.cu -- per-phase clock64() breakdown of k_step_cluster<NC> (one 128 x 128 simulation on NC SMs), CTA 0 of cluster 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -DSMK_FUSED_TIMING -DPROBE_NC=4 -I smokephysai_b200/csrc -o build/cluster_probe4 tools/micro/cluster_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "common.cuh"
namespace smk {
int fail(int c, const char*, ...) { return c; }
int check_launch(const char*) { return 0; }
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static EnvCfg g_env_probe = {SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET, SMK_ENV_UNSET};
const EnvCfg& env() { return g_env_probe; }
void prof_mark(int, cudaStream_t, bool) {}
}
#include "fused.cu"
using namespace smk;
#ifndef PROBE_NC
#define PROBE_NC 4
#endif

int main(int argc, char** argv)
{
    const int K = argc > 1 ? atoi(argv[1]) : 40, nsteps = argc > 2 ? atoi(argv[2]) : 8, B = argc > 3 ? atoi(argv[3]) : 32;
    const size_t nu = 129 * 128, nv = 128 * 132, nc = 128 * 128;
    std::vector<float> hd(B * nc, 0.f);
    for (int b = 0; b < B; ++b)
        for (int e = 0; e < 2; ++e) {
            const int cx = 30 + (b * 7 + e * 41) % 70, cy = 30 + (b * 13 + e * 29) % 70;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 128; ++j) {
                    const float d2 = (float)((i - cy) * (i - cy) + (j - cx) * (j - cx));
                    if (d2 <= 64.f) hd[b * nc + i * 128 + j] += 1.5f * expf(-d2 / (2.f * (8.f / 3.f) * (8.f / 3.f)));
                }
        }
    float *u, *v, *d, *p, *fr; long long* ticks;
    cudaMalloc(&u, B * nu * 4); cudaMalloc(&v, B * nv * 4); cudaMalloc(&d, B * nc * 4); cudaMalloc(&p, B * nc * 4);
    cudaMalloc(&fr, (size_t)B * nsteps * nc * 4); cudaMalloc(&ticks, 8 * 8);
    cudaMemset(u, 0, B * nu * 4); cudaMemset(v, 0, B * nv * 4); cudaMemset(p, 0, B * nc * 4);
    cudaMemcpy(d, hd.data(), B * nc * 4, cudaMemcpyHostToDevice);
    FusedArgs a;
    a.U = u; a.V = v; a.D = d; a.P = p; a.frames = fr; a.fmul = nullptr;
    a.h = 128; a.w = 128; a.pu = 128; a.pv = 132; a.pc = 128; a.su_ = nu; a.sv_ = nv; a.sc_ = nc;
    a.frame_step_stride = nc; a.frame_batch_stride = (long long)nsteps * nc;
    a.dt = 0.01f; a.c_uv = (float)(0.01 * 0.001); a.c_d = (float)(0.01 * (0.001 * 0.1)); a.decay = 0.995f; a.K = K; a.nsteps = nsteps;
    a.items = nullptr; a.progress = nullptr; a.ticket = nullptr; a.spin_budget = 0;
    a.ticks = ticks;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(ticks, 0, 64);
        launch_cluster_nc<PROBE_NC>(B, a, 0);
        cudaDeviceSynchronize();
    }
    long long t[8];
    cudaMemcpy(t, ticks, 64, cudaMemcpyDeviceToHost);
    const char* names[8] = {"load + first buoyancy", "diffusion x3", "divergence", "jacobi", "gradient subtract", "advection x3", "frame + buoyancy", "store"};
    long long per_step = 0;
    for (int k = 1; k <= 6; ++k) per_step += t[k];
    printf("clusters of %d, %d simulations, K=%d, %d steps per launch, %s\n", PROBE_NC, B, K, nsteps, cudaGetErrorString(cudaGetLastError()));
    for (int k = 0; k < 8; ++k) {
        const bool once = (k == 0 || k == 7);
        printf("  %-24s %9.0f cycles %s  %5.1f %% of a step\n", names[k], once ? (double)t[k] : (double)t[k] / nsteps, once ? "per launch" : "per step  ",
               100.0 * (once ? (double)t[k] : (double)t[k] / nsteps) / ((double)per_step / nsteps));
    }
    printf("  step total %.0f cycles\n", (double)per_step / nsteps);
    return 0;
}

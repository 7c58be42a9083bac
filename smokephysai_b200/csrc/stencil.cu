// stencil.cu -- every phase of step() except the Jacobi sweeps (navier_stokes.py:151-173), plus the
// emitter splat, the divergence norms and the fractal multiplier field.  All kernels are HBM/L2-bound
// stencils or gathers: coalesced row-major access, shared-memory staging with halos where a phase reuses
// neighbours, no tensor cores (nothing here is a contraction).
#include "common.cuh"

namespace smk {

// =====================================================================================================
// a3 + a4 + a5 fused: buoyancy, three diffusions, divergence -- one read of (u, v, d), one write of
// (u', v', d', div).  Tile FTH x FTW cells; halo 1 for the diffusion + 1 row of u' / col of v' for div.
// =====================================================================================================
constexpr int FTH = 16, FTW = 64, FTHREADS = 256;

__global__ void __launch_bounds__(FTHREADS)
k_forces_diffuse_div(const float* __restrict__ U, const float* __restrict__ V, const float* __restrict__ D,
                     float* __restrict__ Uo, float* __restrict__ Vo, float* __restrict__ Do, float* __restrict__ DIV,
                     const int h, const int w, const int pu, const int pv, const int pc,
                     const long long su_, const long long sv_, const long long sc_,
                     const float dt, const float c_uv, const float c_d)
{
    __shared__ float su[FTH + 3][FTW + 2];     // u rows i0-1 .. i0+FTH+1, cols j0-1 .. j0+FTW   (replicate-clamped)
    __shared__ float sv[FTH + 2][FTW + 3];     // v+buoyancy rows i0-1 .. i0+FTH, cols j0-1 .. j0+FTW+1
    __shared__ float sd[FTH + 2][FTW + 3];     // d, same window as sv
    __shared__ float su1[FTH + 1][FTW];        // diffused u, rows i0 .. i0+FTH
    __shared__ float sv1[FTH][FTW + 1];        // diffused v, cols j0 .. j0+FTW

    const int tid = threadIdx.x;
    const int i0 = blockIdx.y * FTH, j0 = blockIdx.x * FTW;
    const size_t b = blockIdx.z;
    U += b * su_; Uo += b * su_; V += b * sv_; Vo += b * sv_; D += b * sc_; Do += b * sc_;
    if (DIV) DIV += b * sc_;

    for (int k = tid; k < (FTH + 3) * (FTW + 2); k += FTHREADS) {
        const int r = k / (FTW + 2), c = k % (FTW + 2);
        su[r][c] = __ldg(U + (size_t)clampi(i0 - 1 + r, 0, h) * pu + clampi(j0 - 1 + c, 0, w - 1));
    }
    for (int k = tid; k < (FTH + 2) * (FTW + 3); k += FTHREADS) {
        const int r = k / (FTW + 3), c = k % (FTW + 3);
        const int ci = clampi(i0 - 1 + r, 0, h - 1);
        const int cjv = clampi(j0 - 1 + c, 0, w), cjd = min(cjv, w - 1);
        const float dv = __ldg(D + (size_t)ci * pc + cjd);
        float vv = __ldg(V + (size_t)ci * pv + cjv);
        if (cjv < w) {                       // v[:, :-1] += dt * (density * 0.1)      navier_stokes.py:154-155
            const float bu = dv * 0.1f;
            vv = vv + dt * bu;
        }
        sv[r][c] = vv;
        sd[r][c] = dv;
    }
    __syncthreads();

    // diffusion_step: f + c*((((up+down)+left)+right) - 4f)                          navier_stokes.py:69-72
    for (int k = tid; k < (FTH + 1) * FTW; k += FTHREADS) {
        const int r = k / FTW, c = k % FTW;
        const int i = i0 + r, j = j0 + c;
        const float f = su[r + 1][c + 1];
        float s = su[r][c + 1] + su[r + 2][c + 1];
        s = s + su[r + 1][c];
        s = s + su[r + 1][c + 2];
        const float o = f + c_uv * (s - 4.0f * f);
        su1[r][c] = o;
        if (i <= h && j < w && (r < FTH || i == h)) Uo[(size_t)i * pu + j] = o;
    }
    for (int k = tid; k < FTH * (FTW + 1); k += FTHREADS) {
        const int r = k / (FTW + 1), c = k % (FTW + 1);
        const int i = i0 + r, j = j0 + c;
        const float f = sv[r + 1][c + 1];
        float s = sv[r][c + 1] + sv[r + 2][c + 1];
        s = s + sv[r + 1][c];
        s = s + sv[r + 1][c + 2];
        const float o = f + c_uv * (s - 4.0f * f);
        sv1[r][c] = o;
        if (i < h && j <= w && (c < FTW || j == w)) Vo[(size_t)i * pv + j] = o;
    }
    for (int k = tid; k < FTH * FTW; k += FTHREADS) {
        const int r = k / FTW, c = k % FTW;
        const int i = i0 + r, j = j0 + c;
        const float f = sd[r + 1][c + 1];
        float s = sd[r][c + 1] + sd[r + 2][c + 1];
        s = s + sd[r + 1][c];
        s = s + sd[r + 1][c + 2];
        if (i < h && j < w) Do[(size_t)i * pc + j] = f + c_d * (s - 4.0f * f);
    }
    if (DIV == nullptr) return;
    __syncthreads();
    // div = (((u[i+1][j] - u[i][j]) + v[i][j+1]) - v[i][j]) / dt                      navier_stokes.py:136
    for (int k = tid; k < FTH * FTW; k += FTHREADS) {
        const int r = k / FTW, c = k % FTW;
        const int i = i0 + r, j = j0 + c;
        if (i < h && j < w) {
            float s = su1[r + 1][c] - su1[r][c];
            s = s + sv1[r][c + 1];
            s = s - sv1[r][c];
            DIV[(size_t)i * pc + j] = s / dt;
        }
    }
}

int launch_forces_diffuse_div(const smk_grid_t* g, const float* u, const float* v, const float* d,
                              float* uo, float* vo, float* dout, float* div, float dt, float c_uv, float c_d, cudaStream_t s)
{
    dim3 grid((g->w + FTW - 1) / FTW, (g->h + FTH - 1) / FTH, g->batch);
    ProfScope prof_(SMK_PH_FORCES_DIFFUSE_DIV, s);
    k_forces_diffuse_div<<<grid, FTHREADS, 0, s>>>(u, v, d, uo, vo, dout, div, g->h, g->w, g->pitch_u, g->pitch_v, g->pitch_c,
                                                   g->stride_u, g->stride_v, g->stride_c, dt, c_uv, c_d);
    return check_launch("k_forces_diffuse_div");
}

// ---- a4 alone: diffusion_step on one field (unit hook for NavierStokesSimulator.diffusion_step) -------
__global__ void k_diffuse(const float* __restrict__ F, float* __restrict__ O, const int rows, const int cols,
                          const int pitch, const long long stride, const float c)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= rows || j >= cols) return;
    F += (size_t)blockIdx.z * stride; O += (size_t)blockIdx.z * stride;
    const int iu = max(i - 1, 0), id = min(i + 1, rows - 1), jl = max(j - 1, 0), jr = min(j + 1, cols - 1);
    const float f = F[(size_t)i * pitch + j];
    float s = F[(size_t)iu * pitch + j] + F[(size_t)id * pitch + j];
    s = s + F[(size_t)i * pitch + jl];
    s = s + F[(size_t)i * pitch + jr];
    O[(size_t)i * pitch + j] = f + c * (s - 4.0f * f);
}

int launch_diffuse(const float* in, float* out, int rows, int cols, int pitch, int batch, int64_t stride, float c, cudaStream_t s)
{
    dim3 grid((cols + 31) / 32, (rows + 7) / 8, batch), blk(32, 8);
    ProfScope prof_(SMK_PH_OTHER, s);
    k_diffuse<<<grid, blk, 0, s>>>(in, out, rows, cols, pitch, stride, c);
    return check_launch("k_diffuse");
}

// ---- a5 alone ----------------------------------------------------------------------------------------
__global__ void k_divergence(const float* __restrict__ U, const float* __restrict__ V, float* __restrict__ DIV,
                             const int h, const int w, const int pu, const int pv, const int pc,
                             const long long su_, const long long sv_, const long long sc_, const float dt)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= h || j >= w) return;
    const size_t b = blockIdx.z;
    U += b * su_; V += b * sv_; DIV += b * sc_;
    float s = U[(size_t)(i + 1) * pu + j] - U[(size_t)i * pu + j];
    s = s + V[(size_t)i * pv + j + 1];
    s = s - V[(size_t)i * pv + j];
    DIV[(size_t)i * pc + j] = s / dt;
}

int launch_divergence(const smk_grid_t* g, const float* u, const float* v, float* div, float dt, cudaStream_t s)
{
    dim3 grid((g->w + 31) / 32, (g->h + 7) / 8, g->batch), blk(32, 8);
    ProfScope prof_(SMK_PH_OTHER, s);
    k_divergence<<<grid, blk, 0, s>>>(u, v, div, g->h, g->w, g->pitch_u, g->pitch_v, g->pitch_c,
                                      g->stride_u, g->stride_v, g->stride_c, dt);
    return check_launch("k_divergence");
}

// ---- a7: gradient subtract, in place (navier_stokes.py:148-149) -----------------------------------------
__global__ void k_project(const float* __restrict__ P, float* __restrict__ U, float* __restrict__ V,
                          const int h, const int w, const int pu, const int pv, const int pc,
                          const long long su_, const long long sv_, const long long sc_, const float dt)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= h || j >= w) return;
    const size_t b = blockIdx.z;
    P += b * sc_; U += b * su_; V += b * sv_;
    const float pc0 = P[(size_t)i * pc + j];
    if (i >= 1) {        // u[1:-1, :] -= dt * (p[1:, :] - p[:-1, :])
        const float gr = pc0 - P[(size_t)(i - 1) * pc + j];
        U[(size_t)i * pu + j] = U[(size_t)i * pu + j] - dt * gr;
    }
    if (j >= 1) {        // v[:, 1:-1] -= dt * (p[:, 1:] - p[:, :-1])
        const float gr = pc0 - P[(size_t)i * pc + j - 1];
        V[(size_t)i * pv + j] = V[(size_t)i * pv + j] - dt * gr;
    }
}

int launch_project(const smk_grid_t* g, const float* p, float* u, float* v, float dt, cudaStream_t s)
{
    dim3 grid((g->w + 31) / 32, (g->h + 7) / 8, g->batch), blk(32, 8);
    ProfScope prof_(SMK_PH_PROJECT, s);
    k_project<<<grid, blk, 0, s>>>(p, u, v, g->h, g->w, g->pitch_u, g->pitch_v, g->pitch_c,
                                   g->stride_u, g->stride_v, g->stride_c, dt);
    return check_launch("k_project");
}

// ---- a8: bilinear_interpolate at arbitrary coordinates (unit hook) ---------------------------------------
__global__ void k_bilerp(const float* __restrict__ F, const int rows, const int cols, const int pitch,
                         const float* __restrict__ Y, const float* __restrict__ X, float* __restrict__ O,
                         const long long n, const int mode)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float y = Y[k], x = X[k];
    if (mode == 1) x = clampf(x + 0.5f, 0.0f, (float)(cols - 1));        // navier_stokes.py:100-101
    if (mode == 2) y = clampf(y + 0.5f, 0.0f, (float)(rows - 1));        // navier_stokes.py:107-108
    const GlobalField f{F, pitch};
    O[k] = bilerp(f, rows, cols, y, x);
}

int launch_bilerp(const float* f, int rows, int cols, int pitch, const float* y, const float* x, float* out, int64_t n, int mode, cudaStream_t s)
{
    if (n <= 0) return SMK_OK;
    ProfScope prof_(SMK_PH_OTHER, s);
    k_bilerp<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(f, rows, cols, pitch, y, x, out, n, mode);
    return check_launch("k_bilerp");
}

// ---- a10 (+a9, a11 epilogue): advection_step of one field by (u, v) ---------------------------------------
__global__ void __launch_bounds__(256)
k_advect(const float* __restrict__ F, float* __restrict__ O, const int rows, const int cols, const int pitch,
         const long long stride, const float* __restrict__ U, const float* __restrict__ V,
         const int h, const int w, const int pu, const int pv, const long long su_, const long long sv_,
         const float dt, const int has_scale, const float scale,
         float* __restrict__ frame, const long long frame_stride, const int frame_pitch, const float* __restrict__ fmul)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= rows || j >= cols) return;
    const size_t b = blockIdx.z;
    const GlobalField f{F + b * stride, pitch}, fu{U + b * su_, pu}, fv{V + b * sv_, pv};
    const float ui = interp_u_at(fu, h, w, i, j);                        // navier_stokes.py:84
    const float vi = interp_v_at(fv, h, w, i, j);                        // navier_stokes.py:85
    float px = (float)j - dt * ui;                                       // :87
    float py = (float)i - dt * vi;                                       // :88
    px = clampf(px, 0.0f, (float)(cols - 1));                            // :91
    py = clampf(py, 0.0f, (float)(rows - 1));                            // :92
    float val = bilerp(f, rows, cols, py, px);                           // :95
    if (has_scale) val = val * scale;                                    // :171
    O[b * stride + (size_t)i * pitch + j] = val;
    if (frame) {                                                         // :173 (+ fractal_generator.py:62)
        float fr = val;
        if (fmul) fr = val + fmul[(size_t)i * frame_pitch + j] * val;
        frame[b * frame_stride + (size_t)i * frame_pitch + j] = fr;
    }
}

int launch_advect(const smk_grid_t* g, const float* field, float* out, int rows, int cols, int pitch, int64_t stride,
                  const float* u, const float* v, float dt, float scale, float* frame, int64_t frame_stride,
                  const float* fmul, cudaStream_t s)
{
    dim3 grid((cols + 31) / 32, (rows + 7) / 8, g->batch), blk(32, 8);
    ProfScope prof_(rows == g->h + 1 ? SMK_PH_ADVECT_U : (cols == g->w + 1 ? SMK_PH_ADVECT_V : SMK_PH_ADVECT_D), s);
    k_advect<<<grid, blk, 0, s>>>(field, out, rows, cols, pitch, stride, u, v, g->h, g->w, g->pitch_u, g->pitch_v,
                                  g->stride_u, g->stride_v, dt, scale != 1.0f ? 1 : 0, scale,
                                  frame, frame_stride, g->pitch_c, fmul);
    return check_launch("k_advect");
}

// ---- a2: add_smoke_source, batched ---------------------------------------------------------------------
__global__ void k_splat(float* __restrict__ Dn, const int h, const int w, const int pc, const long long sc_,
                        const smk_source_t* __restrict__ src, const int32_t* __restrict__ off)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= h || j >= w) return;
    const int b = blockIdx.z;
    const int s0 = off[b], s1 = off[b + 1];
    if (s0 == s1) return;
    float* cell = Dn + (size_t)b * sc_ + (size_t)i * pc + j;
    float acc = *cell;
    for (int k = s0; k < s1; ++k) {
        const smk_source_t e = src[k];
        const long long dx = (long long)j - e.x, dy = (long long)i - e.y;
        const float dist = sqrtf((float)(dx * dx + dy * dy));            // navier_stokes.py:45
        if (dist <= (float)e.radius) {                                   // :46
            const double r3 = (double)e.radius / 3.0;
            const float denom = (float)(2.0 * (r3 * r3));
            const float d2 = dist * dist;
            acc = acc + e.intensity * expf((-d2) / denom);               // :48
        }
    }
    *cell = acc;
}

int launch_splat(const smk_grid_t* g, float* density, const smk_source_t* src, const int32_t* off, cudaStream_t s)
{
    dim3 grid((g->w + 31) / 32, (g->h + 7) / 8, g->batch), blk(32, 8);
    ProfScope prof_(SMK_PH_SPLAT, s);
    k_splat<<<grid, blk, 0, s>>>(density, g->h, g->w, g->pitch_c, g->stride_c, src, off);
    return check_launch("k_splat");
}

// ---- divergence norms: per simulation max|div| and sum div^2 (un-normalised u,v differences) -------------
__global__ void __launch_bounds__(256)
k_div_norms(const float* __restrict__ U, const float* __restrict__ V, float* __restrict__ out,
            const int h, const int w, const int pu, const int pv, const long long su_, const long long sv_)
{
    const size_t b = blockIdx.z;
    U += b * su_; V += b * sv_;
    float mx = 0.f, ss = 0.f;
    for (int i = blockIdx.y; i < h; i += gridDim.y)
        for (int j = blockIdx.x * 256 + threadIdx.x; j < w; j += gridDim.x * 256) {
            float s = U[(size_t)(i + 1) * pu + j] - U[(size_t)i * pu + j];
            s = s + V[(size_t)i * pv + j + 1];
            s = s - V[(size_t)i * pv + j];
            mx = fmaxf(mx, fabsf(s));
            ss += s * s;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    __shared__ float smx[8], sss[8];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) { smx[wp] = mx; sss[wp] = ss; }
    __syncthreads();
    if (wp == 0) {
        mx = lane < 8 ? smx[lane] : 0.f; ss = lane < 8 ? sss[lane] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) {
            atomicMax(reinterpret_cast<int*>(out + 2 * b), __float_as_int(mx));   // mx >= 0: int order == float order
            atomicAdd(out + 2 * b + 1, ss);
        }
    }
}

int launch_div_norms(const smk_grid_t* g, const float* u, const float* v, float* out, cudaStream_t s)
{
    dim3 grid((g->w + 255) / 256, min(g->h, 64), g->batch);
    ProfScope prof_(SMK_PH_OTHER, s);
    k_div_norms<<<grid, 256, 0, s>>>(u, v, out, g->h, g->w, g->pitch_u, g->pitch_v, g->stride_u, g->stride_v);
    return check_launch("k_div_norms");
}

// ---- a13: FractalGenerator fields (constant per grid shape; computed once and cached by the host) -------
// Output index [a][b] pairs px[a]/mx[a] with py[b]/my[b] (torch.meshgrid(x_w, y_h, indexing='ij'),
// fractal_generator.py:17-19,:38-42): shape (na=w, nb=h).  Any of perlin / mandel / mul may be NULL.
__global__ void k_fractal_fields(float* __restrict__ perlin, float* __restrict__ mandel, float* __restrict__ mul,
                                 const int na, const int nb, const int pitch, const float intensity, const int iterations,
                                 const float* __restrict__ px, const float* __restrict__ py,
                                 const float* __restrict__ mx, const float* __restrict__ my)
{
    const int bcol = blockIdx.x * 32 + threadIdx.x, a = blockIdx.y * 8 + threadIdx.y;
    if (a >= na || bcol >= nb) return;
    float P = 0.f, M = 0.f;
    if (perlin || mul) {
        // octaves: noise += (amp*sin(f*X))*cos(f*Y); (noise+1)/2                 fractal_generator.py:21-31
        const float X = px[a], Y = py[bcol];
        float noise = 0.f, amp = 1.f, freq = 1.f;
        for (int o = 0; o < 6; ++o) {
            float t = amp * sinf(freq * X);
            t = t * cosf(freq * Y);
            noise = noise + t;
            amp *= 0.5f; freq *= 2.0f;
        }
        P = (noise + 1.0f) / 2.0f;
        if (perlin) perlin[(size_t)a * pitch + bcol] = P;
    }
    if (mandel || mul) {
        // escape count of z <- z^2 + c while |z| <= 2, c = mx[a] + i*my[b]       fractal_generator.py:44-51
        // z^2 as ATen's vectorised complex square: re = a*a - b*b, im = a*b + b*a, each op rounded.
        const float cr = mx[a], ci = my[bcol];
        float zr = 0.f, zi = 0.f, cnt = 0.f;
        for (int it = 0; it < iterations; ++it) {
            // |z| <= 2 with |z| the fp32-rounded hypot  <=>  zr^2+zi^2 <= (2 + 2^-23)^2, evaluated in double
            const double s2 = (double)zr * (double)zr + (double)zi * (double)zi;
            if (!(s2 <= 4.0 + 0x1p-21 + 0x1p-46)) break;      // once outside, z is frozen: it stays outside
            const float aa = zr * zr, bb = zi * zi;
            const float re = aa - bb;
            const float im = zr * zi + zi * zr;
            zr = re + cr; zi = im + ci;
            cnt = (float)it;
        }
        M = cnt / (float)iterations;
        if (mandel) mandel[(size_t)a * pitch + bcol] = M;
    }
    if (mul) {
        const float F = 0.7f * P + 0.3f * M;                                      // :59
        mul[(size_t)a * pitch + bcol] = intensity * F;                            // :62
    }
}

int launch_fractal_fields(float* perlin, float* mandel, float* mul, int na, int nb, int pitch, float intensity, int iterations,
                          const float* px, const float* py, const float* mx, const float* my, cudaStream_t s)
{
    dim3 grid((nb + 31) / 32, (na + 7) / 8), blk(32, 8);
    ProfScope prof_(SMK_PH_OTHER, s);
    k_fractal_fields<<<grid, blk, 0, s>>>(perlin, mandel, mul, na, nb, pitch, intensity, iterations, px, py, mx, my);
    return check_launch("k_fractal_fields");
}

__global__ void k_apply_mul(const float* __restrict__ F, const float* __restrict__ M, float* __restrict__ O,
                            const int rows, const int cols, const int pitch, const long long stride)
{
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= rows || j >= cols) return;
    const size_t o = (size_t)blockIdx.z * stride + (size_t)i * pitch + j;
    const float f = F[o];
    O[o] = f + M[(size_t)i * pitch + j] * f;
}

int launch_apply_mul(const float* f, const float* mul, float* out, int rows, int cols, int pitch, int batch, int64_t stride, cudaStream_t s)
{
    dim3 grid((cols + 31) / 32, (rows + 7) / 8, batch), blk(32, 8);
    ProfScope prof_(SMK_PH_OTHER, s);
    k_apply_mul<<<grid, blk, 0, s>>>(f, mul, out, rows, cols, pitch, stride);
    return check_launch("k_apply_mul");
}

}  // namespace smk

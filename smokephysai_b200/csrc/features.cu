// features.cu -- SURVEY.md s8(f) rank 2: the per-frame ingredients of SmokeSimulator.get_chaos_features
// (smoke_simulator.py:47-140) on the device, for many frames per launch:
//   * box counts of the above-mean mask at scales 2, 4, 8, 16, 32           (compute_fractal_dimension :89-124)
//   * the 256-bin histogram over [0, 1] with torch.histogram's bin rule      (compute_entropy :126-140)
//   * squared L2 distances between consecutive frames                        (compute_lyapunov_exponent :67-87)
// The reference evaluates these with a Python double loop (one .sum() per box), a CPU histogram and 19
// .item() syncs per call; they dominate dataset generation once the solver is fast.
// Counts are integers and therefore exact; the mean and the distances are accumulated in double.
#include "common.cuh"

namespace smk {

constexpr int FEAT_THREADS = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wp] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];          // same order in every thread
    return t;
}

// One CTA per frame.  box_counts[frame][5], hist[frame][nbins] (both zeroed by the caller), mean_out[frame].
__global__ void __launch_bounds__(FEAT_THREADS)
k_frame_features(const float* __restrict__ frames, const long long frame_stride, const int h, const int w, const int pitch,
                 const float* __restrict__ edges, const int nbins, const float lo, const float hi,
                 int* __restrict__ box_counts, int* __restrict__ hist, float* __restrict__ mean_out)
{
    extern __shared__ int shist[];                       // nbins ints
    __shared__ double sred[FEAT_THREADS / 32];
    __shared__ int sbox[5];
    const float* F = frames + (size_t)blockIdx.x * frame_stride;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, nw = blockDim.x >> 5;

    for (int k = tid; k < nbins; k += blockDim.x) shist[k] = 0;
    if (tid < 5) sbox[tid] = 0;
    __syncthreads();

    // ---- pass 1: mean (current.mean(), smoke_simulator.py:96) and the histogram (:134) ----------------------
    double acc = 0.0;
    for (int i = wp; i < h; i += nw) {
        const float* row = F + (size_t)i * pitch;
        for (int j = lane; j < w; j += 32) {
            const float x = row[j];
            acc += (double)x;
            if (x >= lo && x <= hi) {
                // torch.histogram (ATen HistogramKernel.cpp, LINEAR_INTERPOLATION_WITH_LOCAL_SEARCH): a linear guess,
                // then upper_bound over the neighbouring edges; the right-most bin includes its right edge.
                long long pos = (long long)((x - lo) * (float)nbins / (hi - lo));
                long long pmin = pos - 1 > 0 ? pos - 1 : 0;
                long long pmax = pos + 2 < nbins + 1 ? pos + 2 : nbins + 1;
                long long q = pmin;
                while (q < pmax && !(x < edges[q])) ++q;                  // upper_bound in [pmin, pmax)
                pos = q - 1;
                if (pos == nbins) pos -= 1;
                if (pos >= 0 && pos < nbins) atomicAdd(&shist[(int)pos], 1);
            }
        }
    }
    const double total = block_sum(acc, sred);
    const float mean = (float)(total / ((double)h * (double)w));
    if (tid == 0 && mean_out) mean_out[blockIdx.x] = mean;

    // ---- pass 2: box counting (:97-115).  One warp per 32 x 32 aligned super block; lane l holds the bit mask of
    // row l (bit c: pixel > mean).  For scale s the reference crops to (h//s)*s x (w//s)*s and counts boxes with any
    // set pixel: OR over s rows by xor-shuffles, OR over s columns by shifts, popcount of every s-th bit.
    const int nby = (h + 31) / 32, nbx = (w + 31) / 32;
    int cnt[5] = {0, 0, 0, 0, 0};
    for (int blk = wp; blk < nby * nbx; blk += nw) {
        const int Y = (blk / nbx) * 32, X = (blk % nbx) * 32;
        const int i = Y + lane;
        unsigned m = 0;
        if (i < h) {
            const float* row = F + (size_t)i * pitch + X;
#pragma unroll 8
            for (int c = 0; c < 32; ++c)
                if (X + c < w && row[c] > mean) m |= 1u << c;
        }
#pragma unroll
        for (int lvl = 0; lvl < 5; ++lvl) {
            const int s = 2 << lvl;
            const int vh = (h / s) * s, vw = (w / s) * s;                 // cropped extent at this scale
            unsigned mm = (i < vh) ? m : 0u;
            const int cols = vw - X;                                      // valid columns in this super block
            if (cols < 32) mm &= cols > 0 ? ((1u << cols) - 1u) : 0u;
            for (int o = 1; o < s; o <<= 1) mm |= __shfl_xor_sync(0xffffffffu, mm, o);      // OR over the s rows
            for (int o = 1; o < s; o <<= 1) mm |= mm >> o;                                    // OR over the s columns
            unsigned pick = 0;                                            // every s-th bit
            for (int c = 0; c < 32; c += s) pick |= 1u << c;
            if ((lane & (s - 1)) == 0) cnt[lvl] += __popc(mm & pick);
        }
    }
#pragma unroll
    for (int lvl = 0; lvl < 5; ++lvl) {
        int v = cnt[lvl];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&sbox[lvl], v);
    }
    __syncthreads();
    if (tid < 5) box_counts[(size_t)blockIdx.x * 5 + tid] = sbox[tid];
    for (int k = tid; k < nbins; k += blockDim.x) hist[(size_t)blockIdx.x * nbins + k] = shist[k];
}

int launch_frame_features(const float* frames, int64_t frame_stride, int nframes, int h, int w, int pitch,
                          const float* edges, int nbins, float lo, float hi, int* box_counts, int* hist, float* mean_out,
                          cudaStream_t s)
{
    if (nframes <= 0) return SMK_OK;
    ProfScope prof_(SMK_PH_OTHER, s);
    k_frame_features<<<nframes, FEAT_THREADS, nbins * sizeof(int), s>>>(frames, frame_stride, h, w, pitch, edges, nbins, lo, hi,
                                                                        box_counts, hist, mean_out);
    return check_launch("k_frame_features");
}

// sumsq[n] = sum over the frame of (F[n+1] - F[n])^2 in double, n < nframes - 1 (frames frame_stride apart).
__global__ void __launch_bounds__(FEAT_THREADS)
k_frame_distances(const float* __restrict__ frames, const long long frame_stride, const int h, const int w, const int pitch,
                  double* __restrict__ sumsq)
{
    __shared__ double sred[FEAT_THREADS / 32];
    const float* A = frames + (size_t)blockIdx.x * frame_stride;
    const float* B = A + frame_stride;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double acc = 0.0;
    for (int i = wp; i < h; i += nw)
        for (int j = lane; j < w; j += 32) {
            const float d = B[(size_t)i * pitch + j] - A[(size_t)i * pitch + j];      // states[1:] - states[:-1] in fp32 (:74)
            acc += (double)d * (double)d;
        }
    const double t = block_sum(acc, sred);
    if (threadIdx.x == 0) sumsq[blockIdx.x] = t;
}

int launch_frame_distances(const float* frames, int64_t frame_stride, int nframes, int h, int w, int pitch, double* sumsq, cudaStream_t s)
{
    if (nframes <= 1) return SMK_OK;
    ProfScope prof_(SMK_PH_OTHER, s);
    k_frame_distances<<<nframes - 1, FEAT_THREADS, 0, s>>>(frames, frame_stride, h, w, pitch, sumsq);
    return check_launch("k_frame_distances");
}

}  // namespace smk

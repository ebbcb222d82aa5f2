// Shared by the 3-D (CostRegNet, conv3d.cu) and 2-D (FeatureNet, featurenet.cu) convolution stacks:
// train-mode InPlaceABN folded into the consumer's load ("normalise on load").
#pragma once
#include "common.cuh"

namespace mvsn {

constexpr float kBnEps = 1e-5f;
constexpr float kSlope = 0.01f;
constexpr int kMaxCin = 64;

struct ActSrc {                 // one input tensor of a layer, stored raw + its batch statistics
    const float* x;             // [C][D][H][W]
    const double* stats;        // [C][2] sum, sum of squares over `count` voxels; null = plain tensor
    const float* gamma;         // [C]
    const float* beta;          // [C]
    double count;
    const float* rmean;         // eval mode: running mean / variance [C] used instead of the batch statistics
    const float* rvar;          // (null in train mode)
};


__device__ __forceinline__ float act(float x, float sc, float sh) {
    const float y = fmaf(x, sc, sh);
    return y > 0.f ? y : y * kSlope;
}

// per-channel (scale, shift) of a source: y = leaky(x * scale + shift)
__device__ inline void load_norm(const ActSrc& s, int C, float* sc, float* sh, int tid, int nthreads) {
    for (int c = tid; c < C; c += nthreads) {
        if (s.stats) {
            double mean, var;
            if (s.rmean) {                                               // F.batch_norm(training=False)
                mean = (double)s.rmean[c]; var = (double)s.rvar[c];
            } else {
                mean = s.stats[2 * c] / s.count;
                var = s.stats[2 * c + 1] / s.count - mean * mean;        // biased, as F.batch_norm(training=True)
            }
            var = var > 0.0 ? var : 0.0;
            const double inv = 1.0 / sqrt(var + (double)kBnEps);
            const double g = fabs((double)s.gamma[c]) + (double)kBnEps;
            sc[c] = (float)(g * inv);
            sh[c] = (float)((double)s.beta[c] - mean * g * inv);
        } else {
            sc[c] = 1.f; sh[c] = 0.f;
        }
    }
}

// train mode side effect of F.batch_norm: running = (1 - momentum) running + momentum batch (variance UNBIASED, n / (n - 1))
__global__ void bn_update_running_kernel(const double* __restrict__ stats, double count, int C, float momentum,
                                         float* __restrict__ rmean, float* __restrict__ rvar);

}  // namespace mvsn

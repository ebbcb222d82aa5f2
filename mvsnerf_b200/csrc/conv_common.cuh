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
};


__device__ __forceinline__ float act(float x, float sc, float sh) {
    const float y = fmaf(x, sc, sh);
    return y > 0.f ? y : y * kSlope;
}

// per-channel (scale, shift) of a source: y = leaky(x * scale + shift)
__device__ inline void load_norm(const ActSrc& s, int C, float* sc, float* sh, int tid, int nthreads) {
    for (int c = tid; c < C; c += nthreads) {
        if (s.stats) {
            const double mean = s.stats[2 * c] / s.count;
            double var = s.stats[2 * c + 1] / s.count - mean * mean;     // biased, as F.batch_norm(training=True)
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

}  // namespace mvsn

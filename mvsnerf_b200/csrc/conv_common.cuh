// Shared by the 3-D (CostRegNet, conv3d.cu) and 2-D (FeatureNet, featurenet.cu) convolution stacks:
// train-mode InPlaceABN folded into the consumer's load ("normalise on load").
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace mvsn {

constexpr float kBnEps = 1e-5f;
constexpr float kSlope = 0.01f;
constexpr int kMaxCin = 64;

// Batch statistics are accumulated as 64-bit FIXED-POINT integers (2^-16 resolution) in the `stats` slots: per CTA the
// warps' fp32 partials are converted and summed as integers, then one global integer atomic per CTA and channel: integer addition is associative, so the statistics --
// and therefore the whole encoding volume -- are bit-identical from run to run and from rank to rank (round 1 used
// fp32 / fp64 floating-point atomics whose summation order made replicated builds differ in the last bits).
// Range: |sum x^2| < 2^47 = 1.4e14 per channel; resolution 1.5e-5 per warp partial, far below the fp32 rounding of the
// partial itself.
constexpr double kStatScale = 65536.0;
__device__ __forceinline__ unsigned long long stat_fx(float partial) {
    return (unsigned long long)__double2ll_rn((double)partial * kStatScale);
}
__device__ __forceinline__ double stat_value(const double* slot) {
    return (double)(*reinterpret_cast<const long long*>(slot)) * (1.0 / kStatScale);
}

struct ActSrc {                 // one input tensor of a layer, stored raw + its batch statistics
    const float* x;             // [C][D][H][W]
    const double* stats;        // [C][2] 64-bit slots: fixed-point sum, sum of squares over `count` voxels (stat_value); null = plain tensor
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
                mean = stat_value(s.stats + 2 * c) / s.count;
                var = stat_value(s.stats + 2 * c + 1) / s.count - mean * mean;        // biased, as F.batch_norm(training=True)
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

struct ConvArgs {               // one 3-D (transposed) convolution layer of CostRegNet (conv3d.cu, conv0_tc.cu)
    ActSrc in0, in1;            // in1.x == null unless the layer input is a skip sum
    int Cin, Din, Hin, Win;
    const float* w;             // Conv3d [Cout][Cin][27] or ConvTranspose3d [Cin][Cout][27]
    int Cout, Dout, Hout, Wout;
    float* out;                 // [Cout][Dout][Hout][Wout] raw
    double* stats_out;          // [Cout][2] fixed-point slots
};
size_t conv0_tc_workspace_bytes();
int launch_conv0_tc(const ConvArgs& a, void* wimg, cudaStream_t st);

// train mode side effect of F.batch_norm: running = (1 - momentum) running + momentum batch (variance UNBIASED, n / (n - 1));
// one launch for all BatchNorm layers of a network (blockIdx.x = layer)
struct BnUpdateArgs {
    const double* stats[10];
    double count[10];
    int C[10];
    float* rmean[10];
    float* rvar[10];
    float momentum;
};
__global__ void bn_update_running_kernel(BnUpdateArgs a);

}  // namespace mvsn

// fp32 FFMA building blocks of the per-sample MLP shared by the forward render kernel (render_fp32.cu) and the
// fine-tuning backward kernel (render_bwd.cu): a 128-row tile GEMM with 8x8 register blocking whose B operand is
// streamed global -> shared with cp.async double buffering, and the fused pass epilogues.
#pragma once
#include "common.cuh"

namespace mvsn {

constexpr int TILE_M = 128;
constexpr int PE_LD  = 68;    // 63 PE channels + 1 zero + 4 pad (row stride = 16 banks mod 32)
constexpr int H_LD   = 132;
constexpr int FEAT_LD = 36;   // 20 features + 12 zeros + 4 pad
constexpr int HV_LD  = 68;
constexpr int KCHUNK = 32;


__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[MR][NC/16] += A[16 MR][K] (smem, row stride lda) * Wt[K][NC] (global, streamed through sW).
// Thread (ty, tx) = (tid/16, tid%16) owns rows {4ty..4ty+3, 64+4ty..} (MR = 8; MR = 4: the first four only, i.e. a
// 64-row A) and columns {4tx..4tx+3, NC/2+4tx..} (NC=128) or {4tx..4tx+3} (NC=64).
template <int NC, int MR>
__device__ __forceinline__ void gemm_pass(float (&acc)[MR][NC / 16], const float* sA, int lda, int K,
                                          const float* __restrict__ gW, float* sW, int tid) {
    static_assert(MR == 8 || MR == 4, "8 (128-row A) or 4 (64-row A) rows per thread");
    constexpr int NT = NC / 16;
    constexpr int V4_PER_CHUNK = KCHUNK * NC / 4;
    const int ty = tid >> 4, tx = tid & 15;
    const int nchunks = K / KCHUNK;
    auto issue = [&](int c) {
        const float4* src = reinterpret_cast<const float4*>(gW + (size_t)c * KCHUNK * NC);
        float4* dst = reinterpret_cast<float4*>(sW + (c & 1) * KCHUNK * NC);
        for (int i = tid; i < V4_PER_CHUNK; i += 256) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    issue(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) { issue(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* w = sW + (c & 1) * KCHUNK * NC;
        const float* a0 = sA + (ty * 4) * lda + c * KCHUNK;
        const float* a1 = sA + (64 + ty * 4) * lda + c * KCHUNK;
#pragma unroll 2
        for (int kk = 0; kk < KCHUNK; kk += 4) {
            float4 av[MR];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                av[r] = *reinterpret_cast<const float4*>(a0 + r * lda + kk);
                if constexpr (MR == 8) av[4 + r] = *reinterpret_cast<const float4*>(a1 + r * lda + kk);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float b[NT];
                float4 b0 = *reinterpret_cast<const float4*>(w + (kk + j) * NC + tx * 4);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                if constexpr (NT == 8) {
                    float4 b1 = *reinterpret_cast<const float4*>(w + (kk + j) * NC + NC / 2 + tx * 4);
                    b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
                }
#pragma unroll
                for (int r = 0; r < MR; ++r) {
                    float a = j == 0 ? av[r].x : j == 1 ? av[r].y : j == 2 ? av[r].z : av[r].w;
#pragma unroll
                    for (int n = 0; n < NT; ++n) acc[r][n] = fmaf(a, b[n], acc[r][n]);
                }
            }
        }
        __syncthreads();   // everyone is done with this weight buffer (and, on the last chunk, with sA)
    }
}

template <int MR, int NT> __device__ __forceinline__ void zero_acc(float (&acc)[MR][NT]) {
#pragma unroll
    for (int r = 0; r < MR; ++r)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[r][n] = 0.f;
}

// epilogue for N=128 passes.  MODE 0: out = acc + bias ; MODE 1: out = relu((acc + bias) * mod)
template <int MODE>
__device__ __forceinline__ void store_pass128(const float (&acc)[8][8], const float* __restrict__ bias,
                                              const float* s_mod, float* s_out, int tid) {
    const int ty = tid >> 4, tx = tid & 15;
    float4 bl = __ldg(reinterpret_cast<const float4*>(bias + tx * 4));
    float4 bh = __ldg(reinterpret_cast<const float4*>(bias + 64 + tx * 4));
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int row = (r < 4 ? 0 : 64) + ty * 4 + (r & 3);
        float4 lo = make_float4(acc[r][0] + bl.x, acc[r][1] + bl.y, acc[r][2] + bl.z, acc[r][3] + bl.w);
        float4 hi = make_float4(acc[r][4] + bh.x, acc[r][5] + bh.y, acc[r][6] + bh.z, acc[r][7] + bh.w);
        if (MODE == 1) {
            float4 ml = *reinterpret_cast<const float4*>(s_mod + row * H_LD + tx * 4);
            float4 mh = *reinterpret_cast<const float4*>(s_mod + row * H_LD + 64 + tx * 4);
            lo.x = fmaxf(lo.x * ml.x, 0.f); lo.y = fmaxf(lo.y * ml.y, 0.f);
            lo.z = fmaxf(lo.z * ml.z, 0.f); lo.w = fmaxf(lo.w * ml.w, 0.f);
            hi.x = fmaxf(hi.x * mh.x, 0.f); hi.y = fmaxf(hi.y * mh.y, 0.f);
            hi.z = fmaxf(hi.z * mh.z, 0.f); hi.w = fmaxf(hi.w * mh.w, 0.f);
        }
        *reinterpret_cast<float4*>(s_out + row * H_LD + tx * 4) = lo;
        *reinterpret_cast<float4*>(s_out + row * H_LD + 64 + tx * 4) = hi;
    }
}


}  // namespace mvsn

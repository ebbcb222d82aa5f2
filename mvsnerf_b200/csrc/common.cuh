// Shared host/device helpers for the mvsnerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>

#include <nvtx3/nvToolsExt.h>                 // header-only NVTX v3: a no-op function-pointer check unless a tool is attached

#include "../../include/mvsnerf_b200.h"

namespace mvsn {

// ---- error channel (thread-local text, never throws) -------------------------------------
void set_error(const char* fmt, ...);

#define MVSN_CUDA_CHECK(expr)                                                         \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            ::mvsn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                    \
            return MVSN_ECUDA;                                                        \
        }                                                                             \
    } while (0)

#define MVSN_REQUIRE(cond, code, ...)          \
    do {                                       \
        if (!(cond)) {                         \
            ::mvsn::set_error(__VA_ARGS__);    \
            return (code);                     \
        }                                      \
    } while (0)

// NVTX range over a C-ABI entry point (the reference has no tracing; nsys / ncu --nvtx see these)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define MVSN_RANGE(name) ::mvsn::NvtxRange mvsn_nvtx_range_(name)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();   // cached multiprocessor count of the current device

// ---- fp32 MLP weight image (MVSN_MLP_FP32), offsets in floats ----------------------------
// All GEMM weights are stored transposed, Wt[k][n] = W[n][k], K padded with zero rows so the
// kernel never branches on K; biases follow each matrix.
namespace w32 {
constexpr int WB   = 0;                      // pts_bias^T       [32][128]  (rows 20..31 zero)
constexpr int BB   = WB + 32 * 128;          // pts_bias.bias    [128]
constexpr int W0   = BB + 128;               // pts_linears.0^T  [64][128]  (row 63 zero)
constexpr int B0   = W0 + 64 * 128;
constexpr int W1   = B0 + 128;               // pts_linears.1..4 [128][128] + bias, 4 blocks
constexpr int LSTR = 128 * 128 + 128;
constexpr int W5   = W1 + 4 * LSTR;          // pts_linears.5^T  [192][128]: rows 0..62 pe, 63 zero, 64..191 h
constexpr int B5   = W5 + 192 * 128;
constexpr int WA   = B5 + 128;               // alpha_linear     [128]
constexpr int BA   = WA + 128;               // alpha bias       [4] (1 used)
constexpr int WF   = BA + 4;                 // feature_linear^T [128][128]
constexpr int BF   = WF + 128 * 128;
constexpr int WV   = BF + 128;               // views_linears.0^T, feature part [128][64]
constexpr int WVD  = WV + 128 * 64;          // views_linears.0, dir part       [3][64] (+1 zero row)
constexpr int BV   = WVD + 4 * 64;
constexpr int WR   = BV + 64;                // rgb_linear       [3][64] (+1 zero row)
constexpr int BR   = WR + 4 * 64;            // rgb bias         [4] (3 used)
constexpr int TOTAL = BR + 4;
}  // namespace w32

}  // namespace mvsn

// Diagnostic entry: one 128 x N x K fp16 GEMM on the tcgen05 path, built from exactly the pieces the
// fused render kernel uses (SWIZZLE_128B K-major operand tiles written from registers, descriptor
// K-stepping, an optional no-swizzle 16-wide "bias" K-step, tcgen05.commit -> mbarrier,
// tcgen05.ld epilogue).  tests/test_gpu_umma.py checks it against a plain fp32 matmul.
#include "common.cuh"
#include "umma.cuh"

namespace mvsn {

using namespace umma;

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ B, const __half* __restrict__ Bc,
                     int N, int K, float* __restrict__ D, int reps, long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B tiles: 1024-B aligned
    const int nkb = K / 64;
    uint8_t* sA = smem;                                   // nkb blocks of 128 rows x 128 B
    uint8_t* sB = sA + nkb * 128 * 128;                   // nkb blocks of N rows x 128 B  (N % 8 == 0)
    uint8_t* sBc = sB + nkb * N * 128;                    // [N x 16] no-swizzle: 8-row groups of 2 x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_holder;
    const int tid = threadIdx.x, warp = tid >> 5;

    // operands: global (row-major [rows][K]) -> swizzled K-blocks
    for (int i = tid; i < 128 * K / 8; i += 128) {        // 16-byte chunks
        const int row = i / (K / 8), kc = i % (K / 8), kb = kc / 8, c = kc % 8;
        const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)row * K + kc * 8);
        *reinterpret_cast<uint4*>(sA + kb * 128 * 128 + sw128_offset(row, c * 8)) = v;
    }
    for (int i = tid; i < N * K / 8; i += 128) {
        const int row = i / (K / 8), kc = i % (K / 8), kb = kc / 8, c = kc % 8;
        const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)row * K + kc * 8);
        *reinterpret_cast<uint4*>(sB + kb * N * 128 + sw128_offset(row, c * 8)) = v;
    }
    if (Bc) {
        for (int i = tid; i < N * 2; i += 128) {          // row n, K half hf: core matrix (n/8, hf), row n%8
            const int n = i >> 1, hf = i & 1;
            const uint4 v = *reinterpret_cast<const uint4*>(Bc + (size_t)n * 16 + hf * 8);
            *reinterpret_cast<uint4*>(sBc + (n >> 3) * 256 + hf * 128 + (n & 7) * 16) = v;
        }
    }
    fence_proxy_async();
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_holder;

    if (tid == 0) {
        const uint32_t idesc = idesc_f16(128, N);
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t a0 = smem_u32(sA + kb * 128 * 128), b0 = smem_u32(sB + kb * N * 128);
            for (int ks = 0; ks < 4; ++ks) {
                mma_f16(tmem, desc_sw128(a0 + ks * 32), desc_sw128(b0 + ks * 32), idesc, acc);
                acc = 1;
            }
        }
        if (Bc)   // D += A[:, 16:32] * Bc^T : A chunk from the swizzled block, B chunk from the no-swizzle tile
            mma_f16(tmem, desc_sw128(smem_u32(sA) + 32), desc_nosw(smem_u32(sBc), 128, 256), idesc, 1);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {                     // throughput probe: re-issue the K-steps into other columns
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t a0 = smem_u32(sA + kb * 128 * 128), b0 = smem_u32(sB + kb * N * 128);
                for (int ks = 0; ks < 4; ++ks)
                    mma_f16(tmem + 256, desc_sw128(a0 + ks * 32), desc_sw128(b0 + ks * 32), idesc, 1);
            }
        }
        mma_commit(&bar);
        if (reps > 0) { mbar_wait(&bar, 0); cycles[0] = clock64() - t0; }
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace mvsn

using namespace mvsn;

#ifdef MVSN_BUILD_PROBES   // issue/commit timing probe: libmvsnerf_b200_probes.so only (tools/umma_probe.py)
extern "C" int mvsn_selftest_umma_probe(const void* A, const void* B, int N, int K, float* D, int reps, long long* cycles, void* stream) {
    const size_t smem = (size_t)(K / 64) * (128 + N) * 128 + (size_t)N * 32 + 1024;
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
        static_cast<const __half*>(A), static_cast<const __half*>(B), nullptr, N, K, D, reps, cycles);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}
#endif

extern "C" int mvsn_selftest_umma(const void* A, const void* B, const void* Bc, int N, int K, float* D, void* stream) {
    MVSN_REQUIRE(A && B && D, MVSN_ENULL, "mvsn_selftest_umma: NULL argument");
    MVSN_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 64 == 0 && K >= 64 && K <= 384, MVSN_EBADSHAPE,
                 "mvsn_selftest_umma: N=%d (16..256, %%16) K=%d (64..256, %%64)", N, K);
    const size_t smem = (size_t)(K / 64) * (128 + N) * 128 + (size_t)N * 32 + 1024;
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
        static_cast<const __half*>(A), static_cast<const __half*>(B), static_cast<const __half*>(Bc), N, K, D, 0, nullptr);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

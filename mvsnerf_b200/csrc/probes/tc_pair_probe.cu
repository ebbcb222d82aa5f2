// ROUND-2 PROBE -- NOT on any product path.  Run once on a B200 at the end of round 1 (tools/umma_pair_probe.py):
// max |D - A B^T| = 4.8e-6 / 6.7e-6 / 1.5e-5 for K = 64 / 128 / 256 on all 256 rows, i.e. the pair MMA with a
// split B operand works as written here.  It is the smallest kernel that exercises what the
// planned CTA-pair render kernel needs from `cta_group::2`:
//   * a 2-CTA cluster, TMEM allocated with cta_group::2 in both CTAs;
//   * ONE tcgen05.mma.cta_group::2 per K-step, issued by the leader CTA: M = 256 (128 rows of A from each
//     CTA's own shared memory), N = 128 with the B operand SPLIT across the pair (each CTA holds 64 of the 128
//     rows of B -- the point of the exercise: half the weight staging per SM);
//   * completion multicast to both CTAs' mbarriers, each CTA reading its own 128 accumulator rows.
// D[256, 128] = A[256, K] * B[128, K]^T, fp16 operands, fp32 accumulate.  tools/umma_pair_probe.py runs it
// under a timeout and compares with torch; mbar_wait traps instead of hanging.
#include "../common.cuh"
#include "../umma.cuh"

namespace mvsn {
using namespace umma;

namespace {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_pair_probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, int K, float* __restrict__ D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_holder;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int nkb = K / 64;
    uint8_t* sA = smem;                                   // nkb blocks of [128 x 64] SW128 (16 KB each): rows rank*128 ..
    uint8_t* sB = smem + (size_t)nkb * 16384;             // nkb blocks of [ 64 x 64] SW128 ( 8 KB each): rows rank*64 ..

    // operand tiles (generic stores, then made visible to the async proxy)
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half*>(sA + (size_t)(k >> 6) * 16384 + sw128_offset(r, k & 63)) = A[(size_t)(rank * 128 + r) * K + k];
    }
    for (int i = tid; i < 64 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half*>(sB + (size_t)(k >> 6) * 8192 + sw128_offset(r, k & 63)) = B[(size_t)(rank * 64 + r) * K + k];
    }
    if (tid == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) { tmem_alloc_pair(&tmem_holder, 128); tmem_relinquish_pair(); }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // both CTAs: tiles written, barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem = tmem_holder;

    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = idesc_f16(256, 128);
            for (int kb = 0; kb < nkb; ++kb) {
                const uint64_t da = desc_sw128(smem_u32(sA + (size_t)kb * 16384));
                const uint64_t db = desc_sw128(smem_u32(sB + (size_t)kb * 8192));
                for (int ks = 0; ks < 4; ++ks)            // a K-step of 16 elements advances the start address by 32 B = 2 units
                    mma_f16_pair(tmem, da + 2 * ks, db + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
            }
            mma_commit_pair(&done_bar);
        }
        __syncwarp();
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    // this CTA's 128 accumulator rows: warp w reads lanes 32 w .. 32 w + 31, four 32-column groups
    for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        float* o = D + (size_t)(rank * 128 + tid) * 128 + c * 32;
        for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // nobody frees TMEM / exits while the peer may still be reading
    if (warp == 0) tmem_dealloc_pair(tmem, 128);
}

}  // namespace
}  // namespace mvsn

// Not declared in include/mvsnerf_b200.h on purpose: a bring-up probe, not part of the ABI.
extern "C" int mvsn_probe_umma_pair(const void* A, const void* B, int K, float* D, void* stream) {
    using namespace mvsn;
    MVSN_REQUIRE(A && B && D, MVSN_ENULL, "mvsn_probe_umma_pair: NULL argument");
    MVSN_REQUIRE(K % 64 == 0 && K >= 64 && K <= 256, MVSN_EBADSHAPE, "mvsn_probe_umma_pair: K=%d (64..256, %%64)", K);
    const size_t smem = (size_t)(K / 64) * (16384 + 8192) + 1024;
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(umma_pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_pair_probe_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(static_cast<const __half*>(A), static_cast<const __half*>(B), K, D);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

// BRING-UP PROBE -- NOT on any product path (built into libmvsnerf_b200_probes.so only).
// A miniature of the data flow the round-2 render kernel (csrc/render_tc2.cu) relies on, so every
// hardware mechanism that is new relative to round 1 is checked in isolation before the big kernel:
//   * the MMA's A operand read from TENSOR MEMORY ("TS" form): fp16 activations packed two per
//     32-bit column, lane = row, written by the row's own thread with tcgen05.st;
//   * the same with cta_group::2 (M = 256 over a CTA pair, B rows split across the pair);
//   * the per-round hand-off chain  writer threads -> (remote) mbarrier arrive -> leader's issuer ->
//     MMAs (one smem-A block + one TMEM-A block into the same accumulator) -> multicast commit ->
//     writers read the accumulator and overwrite both operands for the next round.
// Round i computes D_i = X_i[M,64] * B1[128,64]^T + H_i[M,128] * B2[128,128]^T  (fp16 in, fp32 out).
#define MVSN_MBAR_TIMEOUT_NS 2000000000ull
#include "../common.cuh"
#include "../umma.cuh"

namespace mvsn {
using namespace umma;
namespace {

struct TsShared {
    uint64_t in_ready;      // writer warps (4 per CTA) -> issuer of the leader CTA
    uint64_t acc_ready;     // commit -> writers of each CTA
    uint32_t tmem_holder;
};

template <bool PAIR>
__global__ void __launch_bounds__(160, 1)
ts_probe_kernel(const __half* __restrict__ X, const __half* __restrict__ Hh, const __half* __restrict__ B1,
                const __half* __restrict__ B2, int rounds, float* __restrict__ D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ TsShared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    constexpr int NCTA = PAIR ? 2 : 1;
    constexpr int M = 128 * NCTA;
    constexpr int NB = 128 / NCTA;                       // B rows held by this CTA
    uint8_t* sX = smem;                                  // [128 x 64] SW128                    16 KB
    uint8_t* sB1 = smem + 16384;                         // [NB x 64] SW128                     <= 16 KB
    uint8_t* sB2 = smem + 32768;                         // 2 K-blocks of [NB x 64] SW128       <= 32 KB

    for (int i = tid; i < NB * 64; i += 160) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<__half*>(sB1 + sw128_offset(r, k)) = B1[(size_t)(rank * NB + r) * 64 + k];
    }
    for (int i = tid; i < NB * 128; i += 160) {
        const int r = i / 128, k = i % 128;
        *reinterpret_cast<__half*>(sB2 + (k >> 6) * (NB * 128) + sw128_offset(r, k & 63)) = B2[(size_t)(rank * NB + r) * 128 + k];
    }
    if (tid == 0) {
        mbar_init(&sh.in_ready, 4 * NCTA);
        mbar_init(&sh.acc_ready, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (warp == 4) {
        if (PAIR) { tmem_alloc_pair(&sh.tmem_holder, 256); tmem_relinquish_pair(); }
        else      { tmem_alloc(&sh.tmem_holder, 256); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_holder;
    constexpr uint32_t COL_ACC = 0, COL_H = 128;         // H: 64 columns = 128 fp16 per row

    if (warp < 4) {
        // ---- writer / reader threads: one accumulator row each -------------------------------------
        const int row = warp * 32 + lane;
        const size_t grow = (size_t)rank * 128 + row;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t leader_in_ready = PAIR ? mapa_u32(smem_u32(&sh.in_ready), 0) : 0u;
        for (int it = 0; it < rounds; ++it) {
            // X_i row -> swizzled smem block (generic proxy), H_i row -> TMEM (packed fp16 pairs)
            const __half* xr = X + ((size_t)it * M + grow) * 64;
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(sX + sw128_offset(row, c * 8)) = *reinterpret_cast<const uint4*>(xr + c * 8);
            const uint32_t* hr = reinterpret_cast<const uint32_t*>(Hh + ((size_t)it * M + grow) * 128);
            for (int c = 0; c < 4; ++c) {
                uint32_t v[16];
                for (int j = 0; j < 16; ++j) v[j] = hr[c * 16 + j];
                tmem_st16(lane_base + COL_H + c * 16, v);
            }
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(leader_in_ready);
                else mbar_arrive(&sh.in_ready);
            }
            mbar_wait(&sh.acc_ready, it & 1);
            tc_fence_after();
            for (int c = 0; c < 8; ++c) {
                uint32_t r[16];
                tmem_ld16(lane_base + COL_ACC + c * 16, r);
                tmem_ld_wait();
                float* o = D + ((size_t)it * M + grow) * 128 + c * 16;
                for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(r[j]);
            }
            tc_fence_before();
        }
    } else if (rank == 0) {
        // ---- issuer (leader CTA) ------------------------------------------------------------------------
        const bool leader = elect_one();
        const uint32_t idesc = idesc_f16(M, 128);
        for (int it = 0; it < rounds; ++it) {
            if (PAIR) mbar_wait_cluster(&sh.in_ready, it & 1); else mbar_wait(&sh.in_ready, it & 1);
            tc_fence_after();
            if (leader) {
                const uint64_t dx = desc_sw128(smem_u32(sX)), db1 = desc_sw128(smem_u32(sB1));
                for (int ks = 0; ks < 4; ++ks) {
                    if (PAIR) mma_f16_pair(tmem + COL_ACC, dx + 2 * ks, db1 + 2 * ks, idesc, ks ? 1u : 0u);
                    else      mma_f16(tmem + COL_ACC, dx + 2 * ks, db1 + 2 * ks, idesc, ks ? 1u : 0u);
                }
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t db2 = desc_sw128(smem_u32(sB2 + kb * (NB * 128)));
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t a = tmem + COL_H + (uint32_t)((kb * 4 + ks) * 8);
                        if (PAIR) mma_f16_ts_pair(tmem + COL_ACC, a, db2 + 2 * ks, idesc, 1u);
                        else      mma_f16_ts(tmem + COL_ACC, a, db2 + 2 * ks, idesc, 1u);
                    }
                }
                if (PAIR) mma_commit_pair(&sh.acc_ready); else mma_commit(&sh.acc_ready);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 4) {
        if (PAIR) tmem_dealloc_pair(tmem, 256); else tmem_dealloc(tmem, 256);
    }
}

}  // namespace
}  // namespace mvsn

// X [rounds, M, 64], H [rounds, M, 128], B1 [128, 64], B2 [128, 128] fp16; D [rounds, M, 128] fp32; M = pair ? 256 : 128
extern "C" int mvsn_probe_umma_ts(const void* X, const void* H, const void* B1, const void* B2, int rounds, int pair,
                                  float* D, void* stream) {
    using namespace mvsn;
    MVSN_REQUIRE(X && H && B1 && B2 && D && rounds > 0, MVSN_ENULL, "mvsn_probe_umma_ts: bad argument");
    const int smem = 65536 + 1024;
    cudaStream_t st = (cudaStream_t)stream;
    if (pair) {
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(ts_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        MVSN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ts_probe_kernel<true>, static_cast<const __half*>(X), static_cast<const __half*>(H),
                                           static_cast<const __half*>(B1), static_cast<const __half*>(B2), rounds, D));
    } else {
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(ts_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ts_probe_kernel<false><<<1, 160, smem, st>>>(static_cast<const __half*>(X), static_cast<const __half*>(H),
                                                     static_cast<const __half*>(B1), static_cast<const __half*>(B2), rounds, D);
    }
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

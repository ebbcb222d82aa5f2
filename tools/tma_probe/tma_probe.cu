// Stand-alone probe of tensor-map TMA loads on sm_100a (tools/tma_probe/run.sh): which way of issuing
// cp.async.bulk.tensor works here?  Each variant runs in its own process (a faulting variant kills only itself).
//   argv[1]: variant  0 = 2-D map, __grid_constant__ param        1 = 2-D map in global memory
//                     2 = 4-D map (conv0 shape), param, in-bounds  3 = 4-D, param, negative / out-of-range coordinates
//                     4 = 4-D, param, .shared::cta destination     5 = 4-D, global-memory map, OOB coordinates
//                     6 = 4-D, param, x = -4 (16-byte aligned), y = z = -1     7 = 4-D, param, x = 28, y = 22 (box runs past W and H)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, bool CTA_DST>
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    if (RANK == 2) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                     :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
    } else if (!CTA_DST) {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                     :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                     :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
    }
}

template <int RANK, bool CTA_DST>
__device__ void body(const CUtensorMap* map, int c0, int c1, int c2, int bytes, float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        tma_load<RANK, CTA_DST>(smem, map, c0, c1, c2, 0, &bar);
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\nselp.b32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const float* s = reinterpret_cast<const float*>(smem);
    float acc = 0.f;
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) acc += s[i];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) { out[0] = acc; out[1] = s[0]; out[2] = s[33]; }
}

__global__ void k_param_2d(const __grid_constant__ CUtensorMap map, int c0, int c1, int bytes, float* out) { body<2, false>(&map, c0, c1, 0, bytes, out); }
__global__ void k_global_2d(const CUtensorMap* map, int c0, int c1, int bytes, float* out) { body<2, false>(map, c0, c1, 0, bytes, out); }
__global__ void k_param_4d(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, float* out) { body<4, false>(&map, c0, c1, c2, bytes, out); }
__global__ void k_param_4d_cta(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, float* out) { body<4, true>(&map, c0, c1, c2, bytes, out); }
__global__ void k_global_4d(const CUtensorMap* map, int c0, int c1, int c2, int bytes, float* out) { body<4, false>(map, c0, c1, c2, bytes, out); }

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int C = 41, D = 16, H = 24, W = 48;
    std::vector<float> h((size_t)C * D * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000) * 0.001f;
    float *d_in, *d_out;
    CK(cudaMalloc(&d_in, h.size() * 4)); CK(cudaMalloc(&d_out, 16));
    CK(cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    CUtensorMap m2, m4;
    {
        const cuuint64_t gd[2] = {(cuuint64_t)W, (cuuint64_t)H * D * C}, gs[1] = {(cuuint64_t)W * 4};
        const cuuint32_t box[2] = {32, 4}, es[2] = {1, 1};
        CUresult r = encode(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_in, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 2d: %d\n", (int)r);
    }
    {
        const cuuint64_t gd[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C};
        const cuuint64_t gs[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
        const cuuint32_t box[4] = {32, 4, 1, 41}, es[4] = {1, 1, 1, 1};
        CUresult r = encode(&m4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_in, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 4d: %d\n", (int)r);
    }
    CUtensorMap* d_map;
    CK(cudaMalloc(&d_map, sizeof(CUtensorMap)));
    const int smem = 32768;
    CK(cudaFuncSetAttribute(k_param_4d, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (variant == 0) k_param_2d<<<1, 32, smem>>>(m2, 8, 3, 32 * 4 * 4, d_out);
    if (variant == 1) { CK(cudaMemcpy(d_map, &m2, sizeof(m2), cudaMemcpyHostToDevice)); k_global_2d<<<1, 32, smem>>>(d_map, 8, 3, 32 * 4 * 4, d_out); }
    if (variant == 2) k_param_4d<<<1, 32, smem>>>(m4, 8, 4, 2, 41 * 128 * 4, d_out);
    if (variant == 3) k_param_4d<<<1, 32, smem>>>(m4, -1, -1, -1, 41 * 128 * 4, d_out);
    if (variant == 4) k_param_4d_cta<<<1, 32, smem>>>(m4, 8, 4, 2, 41 * 128 * 4, d_out);
    if (variant == 5) { CK(cudaMemcpy(d_map, &m4, sizeof(m4), cudaMemcpyHostToDevice)); k_global_4d<<<1, 32, smem>>>(d_map, 30, 22, 15, 41 * 128 * 4, d_out); }
    if (variant == 6) k_param_4d<<<1, 32, smem>>>(m4, -4, -1, -1, 41 * 128 * 4, d_out);
    if (variant == 7) k_param_4d<<<1, 32, smem>>>(m4, 28, 22, 15, 41 * 128 * 4, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    float o[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) cudaMemcpy(o, d_out, 12, cudaMemcpyDeviceToHost);
    // expected s[0] for the in-bounds 4-D variants: element (c 0, z 2, y 4, x 8)
    const size_t idx = ((size_t)2 * H + 4) * W + 8;
    printf("variant %d: %s  sum %.4f  s[0] %.4f (in-bounds 4-D expects %.4f)  s[33] %.4f\n", variant, cudaGetErrorString(e), o[0], o[1],
           (float)(idx % 1000) * 0.001f, o[2]);
    return e == cudaSuccess ? 0 : 1;
}

// K-C, tensor-core mode (MVSN_MLP_TC_HALF): the fused per-ray render kernel with the per-sample MLP
// on the 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM).
//
// One persistent CTA per SM, 320 threads in four roles:
//   warps 0-3   slot 0 group : front end (ray march, NDC, trilinear + colour gather, positional
//   warps 4-7   slot 1 group   encoding -> fp16 operand tiles in shared memory), per-layer
//                              epilogues (TMEM -> registers -> modulation/ReLU -> fp16 -> smem) and
//                              alpha compositing; one thread per sample row
//   warp 8      MMA issuer   : one thread issues every tcgen05.mma of both slots and the commits
//   warp 9      weight loader: one thread streams the pre-swizzled weight image L2 -> smem ring with
//                              1-D bulk async copies (cp.async.bulk + mbarrier complete_tx)
// Two 128-sample tiles ("slots") are in flight per CTA and run the nine GEMM phases in lock step,
// so each streamed weight chunk serves both tiles and one slot's epilogue overlaps the other
// slot's MMAs.  Biases ride inside the GEMMs (a constant-one operand column), the multiplicative
// feature modulation (models.py:199-203) is itself a GEMM whose result stays in TMEM for the six
// trunk layers.  Samples never touch HBM between the gather and the composited pixel.
//
// Replaces renderer.rendering (renderer.py:138-165) and callees; see include/mvsnerf_b200.h.
#include "render_frontend.cuh"
#include "umma.cuh"

namespace mvsn {

using namespace umma;

namespace tcw {     // weight image of MVSN_MLP_TC_HALF: chunks in consumption order (bytes)
constexpr int NCHUNK = 18;
__host__ __device__ constexpr int chunk_bytes(int c) {
    return c == 0 || c == 1 || c == 2 || c == 4 || c == 6 || c == 8 || c == 10 || c == 11 || c == 12 ? 16384
         : c == 3 || c == 5 || c == 7 || c == 9 ? 20480
         : c == 13 ? 18432 : c == 14 ? 23040 : c == 15 ? 8192 : c == 16 ? 10240 : 2048;
}
__host__ __device__ constexpr int chunk_offset(int c) {
    int o = 0;
    for (int i = 0; i < c; ++i) o += chunk_bytes(i);
    return o;
}
constexpr int STREAM_BYTES = chunk_offset(NCHUNK);          // 291 328
constexpr int TAIL_OFFSET = STREAM_BYTES;                   // fp32 tail: rgb_linear.bias[3], 0
constexpr int TOTAL_BYTES = STREAM_BYTES + 16;
constexpr int STAGE_BYTES = 24576;
constexpr int NSTAGE = 4;
}  // namespace tcw

constexpr int TC_THREADS = 320;
constexpr int SLOT_BYTES = 65536;                           // PE 16K | H0 16K | H1 16K | MISC 16K
constexpr int OFF_PE = 0, OFF_H0 = 16384, OFF_H1 = 32768, OFF_MISC = 49152;
constexpr int RING_OFFSET = 2 * SLOT_BYTES;
constexpr int TC_SMEM_BYTES = RING_OFFSET + tcw::NSTAGE * tcw::STAGE_BYTES + 1024;

struct TcShared {
    uint64_t in_ready[2];       // slot group (128 arrivals) -> MMA issuer: operand tile written
    uint64_t acc_ready[2];      // tcgen05.commit -> slot group: accumulator complete
    uint64_t w_full[tcw::NSTAGE];
    uint64_t w_empty[tcw::NSTAGE];
    uint32_t tmem_base;
    float scan[2][4][8];        // per slot, per warp: cross-warp compositing scratch
    float carry[2][8];          // S > 128: transmittance / partial sums carried between chunks
    Cams cams;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- epilogue: relu(acc * mod) -> fp16 -> H tile (K columns c0 .. c0+31 of row `row`) -----------
template <bool MODULATE, bool RELU>
__device__ __forceinline__ void epilogue_cols32(uint32_t t_acc, uint32_t t_mod, int c0, uint8_t* slot, int row) {
    uint32_t a[32], m[32];
    tmem_ld32(t_acc + c0, a);
    if (MODULATE) tmem_ld32(t_mod + c0, m);
    tmem_ld_wait();
    uint8_t* blk = slot + (c0 < 64 ? OFF_H0 : OFF_H1);
    const int kc0 = (c0 & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) {             // 4 chunks of 8 columns
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x0 = __uint_as_float(a[q * 8 + 2 * j]), x1 = __uint_as_float(a[q * 8 + 2 * j + 1]);
            if (MODULATE) { x0 *= __uint_as_float(m[q * 8 + 2 * j]); x1 *= __uint_as_float(m[q * 8 + 2 * j + 1]); }
            if (RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
            p[j] = pack_h2(x0, x1);
        }
        *reinterpret_cast<uint4*>(blk + sw128_offset(row, (kc0 + q) * 8)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
}

struct TcIO { RenderIO io; };

template <bool FAST>
__global__ void __launch_bounds__(TC_THREADS, 1)
render_tc_kernel(const SceneDev sc, const RenderIO io, const uint8_t* __restrict__ wimg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ TcShared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    load_cams(sc, &sh.cams, tid);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&sh.in_ready[s], 128); mbar_init(&sh.acc_ready[s], 1); }
        for (int i = 0; i < tcw::NSTAGE; ++i) { mbar_init(&sh.w_full[i], 1); mbar_init(&sh.w_empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    // ---- work decomposition (identical in every role) ---------------------------------------------
    const int N = io.N, S = io.S;
    const int R = S <= 128 ? 128 / S : 1;                    // rays per tile (S | 128 or 128 | S)
    const int nchunks = S <= 128 ? 1 : S / 128;
    const int G = (N + R - 1) / R;                           // ray groups
    const int pairs_total = (G + 1) / 2;
    const int my_pairs = blockIdx.x < pairs_total ? (pairs_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int npass = my_pairs * nchunks;                    // one tile per active slot per pass
    auto group_of = [&](int pass, int s) { return ((pass / nchunks) * (int)gridDim.x + (int)blockIdx.x) * 2 + s; };

    if (warp < 8) {
        // =========================== slot group: front end + epilogues + compositing ===============
        const int s = warp >> 2, row = tid & 127, wq = warp & 3;
        uint8_t* slot = smem + s * SLOT_BYTES;
        const uint32_t t_acc = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(s * 256);
        const uint32_t t_mod = t_acc + 128;
        uint32_t par_acc = 0;
        const float br0 = __ldg(reinterpret_cast<const float*>(wimg + tcw::TAIL_OFFSET));
        const float br1 = __ldg(reinterpret_cast<const float*>(wimg + tcw::TAIL_OFFSET) + 1);
        const float br2 = __ldg(reinterpret_cast<const float*>(wimg + tcw::TAIL_OFFSET) + 2);

        for (int pass = 0; pass < npass; ++pass) {
            const int g = group_of(pass, s), chunk = pass % nchunks;
            if (g >= G) continue;                            // this slot idles in the last pass
            // -------------------------- front end -------------------------------------------------
            int r_in, s_idx;
            if (S <= 128) { r_in = row / S; s_idx = row - r_in * S; } else { r_in = 0; s_idx = chunk * 128 + row; }
            const int ray = g * R + r_in;
            const bool valid = ray < N;
            float pe[3] = {0.f, 0.f, 0.f}, feat[20], dir[3] = {0.f, 0.f, 0.f}, zv = 0.f;
#pragma unroll
            for (int i = 0; i < 20; ++i) feat[i] = 0.f;
            const size_t si = (size_t)ray * S + s_idx;
            if (valid) {
                float px, py, pz, dx, dy, dz;
                if (FAST) {
                    const float4* rp = reinterpret_cast<const float4*>(io.rays + (size_t)ray * 8);
                    float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                    dx = r0.w; dy = r1.x; dz = r1.y;
                    const float near = r1.z, far = r1.w, t = __ldg(io.t_steps + s_idx);
                    if (!io.rg.lindisp) zv = __fadd_rn(__fmul_rn(near, 1.f - t), __fmul_rn(far, t));
                    else zv = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, near), 1.f - t),
                                                       __fmul_rn(__fdiv_rn(1.f, far), t)));
                    px = __fadd_rn(r0.x, __fmul_rn(dx, zv));
                    py = __fadd_rn(r0.y, __fmul_rn(dy, zv));
                    pz = __fadd_rn(r0.z, __fmul_rn(dz, zv));
                    ndc_of_point(sc, sh.cams, io.rg, px, py, pz, pe[0], pe[1], pe[2]);
                } else {
                    px = __ldg(io.pts + si * 3); py = __ldg(io.pts + si * 3 + 1); pz = __ldg(io.pts + si * 3 + 2);
                    pe[0] = __ldg(io.ndc + si * 3); pe[1] = __ldg(io.ndc + si * 3 + 1); pe[2] = __ldg(io.ndc + si * 3 + 2);
                    zv = __ldg(io.z + si);
                    dx = __ldg(io.dirs + (size_t)ray * 3); dy = __ldg(io.dirs + (size_t)ray * 3 + 1);
                    dz = __ldg(io.dirs + (size_t)ray * 3 + 2);
                }
                view_dir(sh.cams, dx, dy, dz, dir);
                sample_volume(sc, pe[0], pe[1], pe[2], feat);
#pragma unroll
                for (int v = 0; v < 3; ++v) sample_color(sc, sh.cams, v, px, py, pz, feat + 8 + 4 * v);
                if (io.input_feat) {
                    float4* o = reinterpret_cast<float4*>(io.input_feat + si * 20);
#pragma unroll
                    for (int i = 0; i < 5; ++i)
                        o[i] = make_float4(feat[4 * i], feat[4 * i + 1], feat[4 * i + 2], feat[4 * i + 3]);
                }
            }
            {   // positional encoding -> PE tile: [x(3), sin(2^k x) k-major, cos(2^k x) k-major, 1]
                float v[64];
                v[0] = pe[0]; v[1] = pe[1]; v[2] = pe[2];
                float f = 1.f;
#pragma unroll
                for (int k = 0; k < 10; ++k) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) sincosf(pe[j] * f, &v[3 + 3 * k + j], &v[33 + 3 * k + j]);
                    f *= 2.f;
                }
                v[63] = 1.f;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(slot + OFF_PE + sw128_offset(row, c * 8)) =
                        make_uint4(pack_h2(v[c * 8], v[c * 8 + 1]), pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                                   pack_h2(v[c * 8 + 4], v[c * 8 + 5]), pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
            }
            {   // MISC tile: [feat 0..19, 1, 0 x11 | dir 0..2, 1, 0 x12 | unused 16]
                uint8_t* m = slot + OFF_MISC;
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 0)) =
                    make_uint4(pack_h2(feat[0], feat[1]), pack_h2(feat[2], feat[3]), pack_h2(feat[4], feat[5]), pack_h2(feat[6], feat[7]));
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 8)) =
                    make_uint4(pack_h2(feat[8], feat[9]), pack_h2(feat[10], feat[11]), pack_h2(feat[12], feat[13]), pack_h2(feat[14], feat[15]));
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 16)) =
                    make_uint4(pack_h2(feat[16], feat[17]), pack_h2(feat[18], feat[19]), pack_h2(1.f, 0.f), 0u);
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 24)) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 32)) =
                    make_uint4(pack_h2(dir[0], dir[1]), pack_h2(dir[2], 1.f), 0u, 0u);
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 40)) = make_uint4(0u, 0u, 0u, 0u);
            }
            fence_proxy_async();
            mbar_arrive(&sh.in_ready[s]);

            // -------------------------- trunk: ops 0..5 -> h ------------------------------------------
#pragma unroll 1
            for (int op = 0; op < 6; ++op) {
                mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < 128; c0 += 32) epilogue_cols32<true, true>(t_acc, t_mod, c0, slot, row);
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&sh.in_ready[s]);
            }
            // -------------------------- op 6: feature (128) + sigma (col 128) ---------------------------
            float sigma;
            {
                mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < 128; c0 += 32) epilogue_cols32<false, false>(t_acc, t_mod, c0, slot, row);
                uint32_t r16[16];
                tmem_ld16(t_acc + 128, r16);
                tmem_ld_wait();
                sigma = fmaxf(__uint_as_float(r16[0]), 0.f);
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&sh.in_ready[s]);
            }
            // -------------------------- op 7: views layer (64) ---------------------------------------------
            {
                mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                tc_fence_after();
                epilogue_cols32<false, true>(t_acc, t_mod, 0, slot, row);
                epilogue_cols32<false, true>(t_acc, t_mod, 32, slot, row);
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&sh.in_ready[s]);
            }
            // -------------------------- op 8: rgb ----------------------------------------------------------
            float cr, cg, cb;
            {
                mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                tc_fence_after();
                uint32_t r16[16];
                tmem_ld16(t_acc, r16);
                tmem_ld_wait();
                tc_fence_before();
                cr = __fdiv_rn(1.f, 1.f + expf(-(__uint_as_float(r16[0]) + br0)));
                cg = __fdiv_rn(1.f, 1.f + expf(-(__uint_as_float(r16[1]) + br1)));
                cb = __fdiv_rn(1.f, 1.f + expf(-(__uint_as_float(r16[2]) + br2)));
            }
            // -------------------------- compositing (renderer.py:18-26,65-92) ----------------------------------
            {
                const float alpha = 1.f - expf(-sigma);
                const float fac = (1.f - alpha) + 1e-10f;
                const int seg = S < 32 ? S : 32;                  // scan segment inside a warp
                const int ls = lane & (seg - 1);
                float incl = fac;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float u = __shfl_up_sync(0xffffffffu, incl, off);
                    if (off < seg && ls >= off) incl *= u;
                }
                float T = __shfl_up_sync(0xffffffffu, incl, 1);
                if (ls == 0) T = 1.f;
                if (S > 32) {                                     // rays span S/32 (<= 4) warps of this group
                    if (lane == 31) sh.scan[s][wq][0] = incl;
                    named_bar_sync(1 + s, 128);
                    const int wpr = S >= 128 ? 4 : S / 32;        // warps per ray inside a tile
                    const int w0 = wq - (wq % wpr);
                    for (int w = w0; w < wq; ++w) T *= sh.scan[s][w][0];
                    if (nchunks > 1 && chunk > 0) T *= sh.carry[s][0];
                    named_bar_sync(1 + s, 128);
                }
                const float wgt = alpha * T;
                if (valid) {
                    if (io.alpha) io.alpha[si] = alpha;
                    if (io.weights) io.weights[si] = wgt;
                }
                float v0 = wgt * cr, v1 = wgt * cg, v2 = wgt * cb, v3 = wgt * zv, v4 = wgt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    if (off < seg) {
                        v0 += __shfl_xor_sync(0xffffffffu, v0, off); v1 += __shfl_xor_sync(0xffffffffu, v1, off);
                        v2 += __shfl_xor_sync(0xffffffffu, v2, off); v3 += __shfl_xor_sync(0xffffffffu, v3, off);
                        v4 += __shfl_xor_sync(0xffffffffu, v4, off);
                    }
                }
                if (S > 32) {
                    if (lane == 0) {
                        sh.scan[s][wq][1] = v0; sh.scan[s][wq][2] = v1; sh.scan[s][wq][3] = v2;
                        sh.scan[s][wq][4] = v3; sh.scan[s][wq][5] = v4;
                    }
                    named_bar_sync(1 + s, 128);
                    const int wpr = S >= 128 ? 4 : S / 32;
                    if (lane == 0 && (wq % wpr) == 0) {
                        for (int w = wq + 1; w < wq + wpr; ++w) {
                            v0 += sh.scan[s][w][1]; v1 += sh.scan[s][w][2]; v2 += sh.scan[s][w][3];
                            v3 += sh.scan[s][w][4]; v4 += sh.scan[s][w][5];
                        }
                        if (nchunks > 1) {
                            float tall = sh.scan[s][0][0] * sh.scan[s][1][0] * sh.scan[s][2][0] * sh.scan[s][3][0];
                            if (chunk > 0) {
                                tall *= sh.carry[s][0];
                                v0 += sh.carry[s][1]; v1 += sh.carry[s][2]; v2 += sh.carry[s][3];
                                v3 += sh.carry[s][4]; v4 += sh.carry[s][5];
                            }
                            sh.carry[s][0] = tall; sh.carry[s][1] = v0; sh.carry[s][2] = v1; sh.carry[s][3] = v2;
                            sh.carry[s][4] = v3; sh.carry[s][5] = v4;
                        }
                    }
                    named_bar_sync(1 + s, 128);
                }
                const bool writer = S > 32 ? (lane == 0 && (wq % (S >= 128 ? 4 : S / 32)) == 0) : (ls == 0);
                if (writer && valid && chunk == nchunks - 1) {
                    if (sc.white_bkgd) { const float bg = 1.f - v4; v0 += bg; v1 += bg; v2 += bg; }
                    io.rgb[(size_t)ray * 3 + 0] = v0; io.rgb[(size_t)ray * 3 + 1] = v1; io.rgb[(size_t)ray * 3 + 2] = v2;
                    io.depth[ray] = v3;
                }
            }
        }
    } else if (warp == 8) {
        // =========================== MMA issuer ========================================================
        if (lane == 0) {
            uint32_t par_in[2] = {0, 0};
            const uint32_t sbase = smem_u32(smem);
            const uint32_t ring = sbase + RING_OFFSET;
            constexpr uint32_t ID128 = idesc_f16(128, 128), ID144 = idesc_f16(128, 144),
                               ID64 = idesc_f16(128, 64), ID16 = idesc_f16(128, 16);
            uint32_t nchunk_base = 0;                         // global chunk counter at the start of the pass
            for (int pass = 0; pass < npass; ++pass) {
                const bool act[2] = {group_of(pass, 0) < G, group_of(pass, 1) < G};
                auto stage_addr = [&](int c) { return ring + ((nchunk_base + c) % tcw::NSTAGE) * tcw::STAGE_BYTES; };
                auto wait_full = [&](int c) {
                    const uint32_t n = nchunk_base + c;
                    mbar_wait(&sh.w_full[n % tcw::NSTAGE], (n / tcw::NSTAGE) & 1);
                };
                auto release = [&](int c) { mma_commit(&sh.w_empty[(nchunk_base + c) % tcw::NSTAGE]); };
                // K-steps over one 64-wide swizzled K-block (nsteps x 16 columns)
                auto block = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, int nsteps, uint32_t& accum) {
                    for (int ks = 0; ks < nsteps; ++ks) {
                        mma_f16(d, desc_sw128(a + ks * 32), desc_sw128(b + ks * 32), idesc, accum);
                        accum = 1;
                    }
                };
                for (int op = 0; op < 9; ++op) {
                    for (int s = 0; s < 2; ++s) {
                        if (!act[s]) continue;
                        const bool first = (s == 0) || !act[0], last = (s == 1) || !act[1];
                        mbar_wait(&sh.in_ready[s], par_in[s]); par_in[s] ^= 1;
                        const uint32_t sl = sbase + s * SLOT_BYTES;
                        const uint32_t d_acc = tmem + s * 256, d_mod = d_acc + 128;
                        uint32_t accum = 0;
                        if (op == 0) {
                            if (first) { wait_full(0); wait_full(1); }
                            tc_fence_after();
                            block(d_mod, sl + OFF_MISC, stage_addr(0), ID128, 2, accum);        // modulation (K = 20 + 1)
                            accum = 0;
                            block(d_acc, sl + OFF_PE, stage_addr(1), ID128, 4, accum);          // layer 0 (K = 63 + 1)
                            mma_commit(&sh.acc_ready[s]);
                            if (last) { release(0); release(1); }
                        } else if (op <= 4) {                                                    // layers 1..4
                            const int c = 2 * op;
                            if (first) { wait_full(c); wait_full(c + 1); }
                            tc_fence_after();
                            block(d_acc, sl + OFF_H0, stage_addr(c), ID128, 4, accum);
                            block(d_acc, sl + OFF_H1, stage_addr(c + 1), ID128, 4, accum);
                            mma_f16(d_acc, desc_sw128(sl + OFF_MISC + 32), desc_nosw(stage_addr(c + 1) + 16384, 128, 256), ID128, 1);
                            mma_commit(&sh.acc_ready[s]);
                            if (last) { release(c); release(c + 1); }
                        } else if (op == 5) {                                                    // layer 5: [pe | h]
                            if (first) { wait_full(10); wait_full(11); wait_full(12); }
                            tc_fence_after();
                            block(d_acc, sl + OFF_PE, stage_addr(10), ID128, 4, accum);
                            block(d_acc, sl + OFF_H0, stage_addr(11), ID128, 4, accum);
                            block(d_acc, sl + OFF_H1, stage_addr(12), ID128, 4, accum);
                            mma_commit(&sh.acc_ready[s]);
                            if (last) { release(10); release(11); release(12); }
                        } else if (op == 6) {                                                    // feature (128) + sigma
                            if (first) { wait_full(13); wait_full(14); }
                            tc_fence_after();
                            block(d_acc, sl + OFF_H0, stage_addr(13), ID144, 4, accum);
                            block(d_acc, sl + OFF_H1, stage_addr(14), ID144, 4, accum);
                            mma_f16(d_acc, desc_sw128(sl + OFF_MISC + 32), desc_nosw(stage_addr(14) + 18432, 128, 256), ID144, 1);
                            mma_commit(&sh.acc_ready[s]);
                            if (last) { release(13); release(14); }
                        } else if (op == 7) {                                                    // views: [feature | dir]
                            if (first) { wait_full(15); wait_full(16); }
                            tc_fence_after();
                            block(d_acc, sl + OFF_H0, stage_addr(15), ID64, 4, accum);
                            block(d_acc, sl + OFF_H1, stage_addr(16), ID64, 4, accum);
                            mma_f16(d_acc, desc_sw128(sl + OFF_MISC + 64), desc_nosw(stage_addr(16) + 8192, 128, 256), ID64, 1);
                            mma_commit(&sh.acc_ready[s]);
                            if (last) { release(15); release(16); }
                        } else {                                                                 // rgb (N = 16, 3 used)
                            if (first) wait_full(17);
                            tc_fence_after();
                            block(d_acc, sl + OFF_H0, stage_addr(17), ID16, 4, accum);
                            mma_commit(&sh.acc_ready[s]);
                            if (last) release(17);
                        }
                    }
                }
                nchunk_base += tcw::NCHUNK;
            }
        }
    } else {
        // =========================== weight loader ========================================================
        if (lane == 0) {
            uint8_t* ring = smem + RING_OFFSET;
            uint32_t n = 0;
            for (int pass = 0; pass < npass; ++pass) {
                for (int c = 0; c < tcw::NCHUNK; ++c, ++n) {
                    const uint32_t st = n % tcw::NSTAGE;
                    mbar_wait(&sh.w_empty[st], ((n / tcw::NSTAGE) & 1) ^ 1);
                    const uint32_t bytes = (uint32_t)tcw::chunk_bytes(c);
                    mbar_arrive_expect_tx(&sh.w_full[st], bytes);
                    bulk_load(ring + st * tcw::STAGE_BYTES, wimg + tcw::chunk_offset(c), bytes, &sh.w_full[st]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

int launch_render_tc(const SceneDev& sc, const RenderIO& io, bool fast, const void* wimg, cudaStream_t stream) {
    const int S = io.S;
    MVSN_REQUIRE((S <= 128 && 128 % S == 0) || (S > 128 && S % 128 == 0), MVSN_EUNSUPPORTED,
                 "tensor-core render mode needs N_samples dividing 128 or a multiple of 128 (got %d); use MVSN_MLP_FP32", S);
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    const int R = S <= 128 ? 128 / S : 1;
    const int G = (io.N + R - 1) / R;
    const int pairs = (G + 1) / 2;
    const int grid = pairs < sm_count() ? pairs : sm_count();
    if (grid <= 0) return MVSN_OK;
    const uint8_t* w = static_cast<const uint8_t*>(wimg);
    if (fast) render_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(sc, io, w);
    else      render_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(sc, io, w);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

// ------------------------------------------------------------------------------------------------
// weight image packer (fp32 nn.Linear tensors -> fp16 pre-swizzled chunks)
// ------------------------------------------------------------------------------------------------
struct MlpPtrsTc { const float* p[MVSN_N_MLP_TENSORS]; };

__device__ __forceinline__ uint32_t nosw_offset(int r, int k) {      // [R x 16] no-swizzle K-major tile
    return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

__global__ void pack_mlp_tc_kernel(MlpPtrsTc w, uint8_t* __restrict__ out) {
    // tensor indices: 0..11 pts_linears (w,b) x6; 12,13 pts_bias; 14,15 views; 16,17 feature; 18,19 alpha; 20,21 rgb
    const int c = blockIdx.x;
    uint8_t* dst = out + tcw::chunk_offset(c);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid * 16; i < tcw::chunk_bytes(c); i += nt * 16) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    auto put = [&](uint32_t off, float v) { *reinterpret_cast<__half*>(dst + off) = __float2half_rn(v); };
    if (c == 0) {
        for (int i = tid; i < 128 * 21; i += nt) { const int r = i / 21, k = i % 21;
            put(sw128_offset(r, k), k < 20 ? w.p[12][r * 20 + k] : w.p[13][r]); }
    } else if (c == 1) {
        for (int i = tid; i < 128 * 64; i += nt) { const int r = i / 64, k = i % 64;
            put(sw128_offset(r, k), k < 63 ? w.p[0][r * 63 + k] : w.p[1][r]); }
    } else if (c >= 2 && c <= 9) {
        const int l = c / 2, kb = c & 1;                      // layer 1..4
        for (int i = tid; i < 128 * 64; i += nt) { const int r = i / 64, k = i % 64;
            put(sw128_offset(r, k), w.p[2 * l][r * 128 + kb * 64 + k]); }
        if (kb == 1) for (int r = tid; r < 128; r += nt) put(16384 + nosw_offset(r, 4), w.p[2 * l + 1][r]);
    } else if (c == 10) {
        for (int i = tid; i < 128 * 64; i += nt) { const int r = i / 64, k = i % 64;
            put(sw128_offset(r, k), k < 63 ? w.p[10][r * 191 + k] : w.p[11][r]); }
    } else if (c == 11 || c == 12) {
        const int kb = c - 11;
        for (int i = tid; i < 128 * 64; i += nt) { const int r = i / 64, k = i % 64;
            put(sw128_offset(r, k), w.p[10][r * 191 + 63 + kb * 64 + k]); }
    } else if (c == 13 || c == 14) {
        const int kb = c - 13;
        for (int i = tid; i < 129 * 64; i += nt) { const int r = i / 64, k = i % 64;
            put(sw128_offset(r, k), r < 128 ? w.p[16][r * 128 + kb * 64 + k] : w.p[18][kb * 64 + k]); }
        if (kb == 1) for (int r = tid; r < 129; r += nt) put(18432 + nosw_offset(r, 4), r < 128 ? w.p[17][r] : w.p[19][0]);
    } else if (c == 15 || c == 16) {
        const int kb = c - 15;
        for (int i = tid; i < 64 * 64; i += nt) { const int r = i / 64, k = i % 64;
            put(sw128_offset(r, k), w.p[14][r * 131 + kb * 64 + k]); }
        if (kb == 1) for (int i = tid; i < 64 * 4; i += nt) { const int r = i / 4, k = i % 4;
            put(8192 + nosw_offset(r, k), k < 3 ? w.p[14][r * 131 + 128 + k] : w.p[15][r]); }
    } else if (c == 17) {
        for (int i = tid; i < 3 * 64; i += nt) { const int r = i / 64, k = i % 64; put(sw128_offset(r, k), w.p[20][r * 64 + k]); }
        if (tid < 4) reinterpret_cast<float*>(out + tcw::TAIL_OFFSET)[tid] = tid < 3 ? w.p[21][tid] : 0.f;
    }
}

size_t mlp_tc_packed_bytes() { return tcw::TOTAL_BYTES; }

int pack_mlp_tc(const float* const* w, void* packed, cudaStream_t stream) {
    MlpPtrsTc p;
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i) p.p[i] = w[i];
    pack_mlp_tc_kernel<<<tcw::NCHUNK, 256, 0, stream>>>(p, static_cast<uint8_t*>(packed));
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // namespace mvsn

// K-C, tensor-core mode (MVSN_MLP_TC_HALF): the fused per-ray render kernel with the per-sample MLP
// on the 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM).
//
// One persistent CTA per SM, 576 threads in four roles:
//   warps 0-7   slot 0 group : front end (ray march, NDC, trilinear + colour gather, positional
//   warps 8-15  slot 1 group   encoding -> fp16 operand tiles in shared memory), per-layer
//                              epilogues (TMEM -> registers -> modulation/ReLU -> fp16 -> smem) and
//                              alpha compositing; two threads per sample row (column halves)
//   warp 16     MMA issuer   : one thread issues every tcgen05.mma of both slots and the commits
//   warp 17     weight loader: one thread streams the pre-swizzled weight image L2 -> smem ring with
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
constexpr int STAGE_BYTES = 23552;                          // >= largest chunk (23 040), 1024-aligned
constexpr int NSTAGE = 4;
}  // namespace tcw

constexpr int TC_THREADS = 576;                           // 16 slot warps + MMA issuer + weight loader
constexpr int SLOT_BYTES = 65536;                           // PE 16K | H0 16K | H1 16K | MISC 16K
constexpr int OFF_PE = 0, OFF_H0 = 16384, OFF_H1 = 32768, OFF_MISC = 49152;
constexpr int RING_OFFSET = 2 * SLOT_BYTES;
constexpr int XCH_OFFSET = RING_OFFSET + tcw::NSTAGE * tcw::STAGE_BYTES;   // per slot: float4 (alpha, r, g, b) x 128 rows
constexpr int TC_SMEM_BYTES = XCH_OFFSET + 2 * 2048 + 1024;

struct TcShared {
    uint64_t in_ready[2];       // slot group (256 arrivals) -> MMA issuer: operand tile written
    uint64_t acc_ready[2];      // tcgen05.commit -> slot group: accumulator complete
    uint64_t w_full[tcw::NSTAGE];
    uint64_t w_free[4];         // GEMM op k (global index) signals w_free[k & 3] when its weight chunks may be overwritten;
                                // four barriers in rotation keep a late waiter at most one phase behind (no parity aliasing)
    uint32_t tmem_base;
    Cams cams;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// tcgen05.wait::ld that also names the destination registers, so no use can be scheduled above it
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// two fp32 -> packed fp16x2 (low half = first argument), saturating, optionally clamped at zero by the converter
template <bool RELU>
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) {
    uint32_t r;
    // .satfinite: a value beyond the fp16 range becomes +-65504 instead of inf (inf x 0-weight = NaN in the next layer)
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// 16 accumulator columns -> (x mod) -> relu -> fp16 -> two 16-byte chunks of an operand K-block.
// `rowaddr` = shared address of the row inside the (1024-aligned) K-block with the row's swizzle
// phase already folded in, so chunk kc lives at rowaddr ^ (kc * 16).
template <bool MODULATE, bool RELU>
__device__ __forceinline__ void emit16(const uint32_t (&a)[16], const uint32_t (&m)[16], uint32_t rowaddr, int kc) {
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float2 x = make_float2(__uint_as_float(a[2 * j]), __uint_as_float(a[2 * j + 1]));
        if (MODULATE) x = __fmul2_rn(x, make_float2(__uint_as_float(m[2 * j]), __uint_as_float(m[2 * j + 1])));
        p[j] = cvt_h2<RELU>(x.x, x.y);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(rowaddr ^ (uint32_t)(kc * 16)), "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(rowaddr ^ (uint32_t)((kc + 1) * 16)), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]) : "memory");
}

// epilogue over NCOL (multiple of 16) accumulator columns starting at t_acc / t_mod, written to
// K-block `blk` starting at 16-byte chunk kc0; TMEM loads of chunk i+1 fly while chunk i is processed
template <int NCOL, bool MODULATE, bool RELU>
__device__ __forceinline__ void epilogue(uint32_t t_acc, uint32_t t_mod, uint8_t* blk, int row, int kc0) {
    const uint32_t rowaddr = smem_u32(blk) + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (row & 7) * 16);
    uint32_t a[2][16], m[2][16];
    tmem_ld16(t_acc, a[0]);
    if (MODULATE) tmem_ld16(t_mod, m[0]);
#pragma unroll
    for (int i = 0; i < NCOL / 16; ++i) {
        tmem_wait16(a[i & 1]);
        if (MODULATE) tmem_wait16(m[i & 1]);
        if (i + 1 < NCOL / 16) {
            tmem_ld16(t_acc + (i + 1) * 16, a[(i + 1) & 1]);
            if (MODULATE) tmem_ld16(t_mod + (i + 1) * 16, m[(i + 1) & 1]);
        }
        emit16<MODULATE, RELU>(a[i & 1], m[i & 1], rowaddr, kc0 + 2 * i);
    }
}

// ---- GEMM schedule of one tile (op = one accumulator phase) --------------------------------------
// op 0: modulation (MISC x chunk 0 -> TMEM cols 128..) + layer 0 (PE x chunk 1); ops 1-4: layers 1-4;
// op 5: layer 5 [PE | H]; op 6: feature + sigma (N = 144); op 7: views layer (N = 64); op 8: rgb (N = 16)
__constant__ int c_op_nblk[9] = {2, 2, 2, 2, 2, 3, 2, 2, 1};
__constant__ uint32_t c_op_idesc[9] = {idesc_f16(128, 128), idesc_f16(128, 128), idesc_f16(128, 128), idesc_f16(128, 128),
                                       idesc_f16(128, 128), idesc_f16(128, 128), idesc_f16(128, 144), idesc_f16(128, 64),
                                       idesc_f16(128, 16)};
__constant__ int c_blk_chunk[9][3] = {{0, 1, 0}, {2, 3, 0}, {4, 5, 0}, {6, 7, 0}, {8, 9, 0}, {10, 11, 12}, {13, 14, 0}, {15, 16, 0}, {17, 0, 0}};
__constant__ uint32_t c_blk_aoff[9][3] = {{49152, 0, 0}, {16384, 32768, 0}, {16384, 32768, 0}, {16384, 32768, 0}, {16384, 32768, 0},
                                          {0, 16384, 32768}, {16384, 32768, 0}, {16384, 32768, 0}, {16384, 0, 0}};
__constant__ int c_chunk_op[18] = {0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8};
__constant__ uint32_t c_op_bias_aoff[9] = {0, 49152 + 32, 49152 + 32, 49152 + 32, 49152 + 32, 0, 49152 + 32, 49152 + 64, 0};
__constant__ uint32_t c_op_bias_boff[9] = {0, 16384, 16384, 16384, 16384, 0, 18432, 8192, 0};

// debug timeline: role r writes (clock << 8 | event) into trace[r * 1024 + i]
#ifdef MVSN_TC_TRACE
#define TC_TRACE(ev) do { if (tr && tr_n < 1023) tr[tr_n++] = (clock64() << 8) | (long long)(ev); } while (0)
#else
#define TC_TRACE(ev) do { } while (0)
#endif

template <bool FAST>
__global__ void __launch_bounds__(TC_THREADS, 1)
render_tc_kernel(const SceneDev sc, const RenderIO io, const uint8_t* __restrict__ wimg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ TcShared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    load_cams(sc, &sh.cams, tid);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sh.in_ready[s], 256);
            mbar_init(&sh.acc_ready[s], 1);
        }
        for (int i = 0; i < tcw::NSTAGE; ++i) mbar_init(&sh.w_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&sh.w_free[i], 1);
        fence_barrier_init();
    }
    if (warp == 16) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;
#ifdef MVSN_TC_TRACE
    // trace roles: 0..3 = (slot, part) warp 0 lane 0 ; 4 = MMA issuer ; 5 = loader
    long long* tr = nullptr; int tr_n = 0;
    if (io.trace && blockIdx.x == 0 && lane == 0) {
        if (warp < 16 && (warp & 3) == 0) tr = io.trace + (warp >> 2) * 1024;
        else if (warp >= 16) tr = io.trace + (4 + warp - 16) * 1024;
    }
#endif

    // ---- work decomposition (identical in every role) ---------------------------------------------
    const int N = io.N, S = io.S;
    // A tile is RT rays x SP = 128/RT consecutive samples: row = sub * RT + ray_in, so the 32 lanes of a
    // warp are (for RT = 32) adjacent rays at the SAME sample index -- their volume / image taps fall in
    // the same few cache lines.  A slot walks the NT tiles of its RT-ray group front to back and the
    // compositing state (transmittance, sums) is carried in registers.  Any S works.
    const int RT = io.rays_per_tile, SP = 128 / RT;
    const int rt_shift = 31 - __clz(RT);
    const int NT = (S + SP - 1) / SP;                        // tiles (= passes) per ray group
    const int G = (N + RT - 1) / RT;                         // ray groups
    const int pairs_total = (G + 1) / 2;
    const int my_pairs = blockIdx.x < pairs_total ? (pairs_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int npass = my_pairs * NT;                         // one tile per active slot per pass
    auto group_of = [&](int pass, int s) { return ((pass / NT) * (int)gridDim.x + (int)blockIdx.x) * 2 + s; };

    if (warp < 16) {
        // =========================== slot group: front end + epilogues + compositing ===============
        // part 0 (warps 0-3 of the slot): accumulator columns 0..63, volume fetch, sin half of the
        // encoding, compositing.  part 1 (warps 4-7): columns 64..127, colour fetch, cos half.
        const int s = warp >> 3, part = (warp >> 2) & 1, wq = warp & 3, row = wq * 32 + lane;
        uint8_t* slot = smem + s * SLOT_BYTES;
        const uint32_t t_acc = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(s * 256);
        const uint32_t t_mod = t_acc + 128;
        uint32_t par_acc = 0;
        float cT = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f;   // compositing state of ray `row` (row < RT)

        const float br0 = __ldg(reinterpret_cast<const float*>(wimg + tcw::TAIL_OFFSET));
        const float br1 = __ldg(reinterpret_cast<const float*>(wimg + tcw::TAIL_OFFSET) + 1);
        const float br2 = __ldg(reinterpret_cast<const float*>(wimg + tcw::TAIL_OFFSET) + 2);
        uint8_t* hblk = slot + (part ? OFF_H1 : OFF_H0);

        // The loop is rotated by one tile: between the views-layer epilogue (op 7) and the rgb result (op 8) of
        // the tile in flight, every thread already builds its share of the NEXT tile's operand tiles (PE / MISC
        // are last read by ops 5 / 7), so the front end hides under the rgb GEMM and the compositing.
        float sigma = 0.f;
        int g_cur = 0, tile_cur = 0;
        bool act_cur = false;
        for (int pass = -1; pass < npass; ++pass) {
            if (act_cur) {
                // -------------------------- trunk: ops 0..5 -> h ------------------------------------------
    #pragma unroll 1
                for (int op = 0; op < 6; ++op) {
                    mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                    TC_TRACE(10 + op);
                    tc_fence_after();
                    epilogue<64, true, true>(t_acc + part * 64, t_mod + part * 64, hblk, row, 0);
                    tc_fence_before();
                    fence_proxy_async();
                    TC_TRACE(30 + op);
                    mbar_arrive(&sh.in_ready[s]);
                }
                // -------------------------- op 6: feature (128) + sigma (col 128) ---------------------------
                sigma = 0.f;
                {
                    mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                    tc_fence_after();
                    epilogue<64, false, false>(t_acc + part * 64, t_mod, hblk, row, 0);
                    if (part == 0) {
                        uint32_t r16[16];
                        tmem_ld16(t_acc + 128, r16);
                        tmem_wait16(r16);
                        sigma = fmaxf(__uint_as_float(r16[0]), 0.f);
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    mbar_arrive(&sh.in_ready[s]);
                }
                // -------------------------- op 7: views layer (64 = 2 x 32) ------------------------------------
                {
                    mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                    tc_fence_after();
                    epilogue<32, false, true>(t_acc + part * 32, t_mod, slot + OFF_H0, row, part * 4);
                    tc_fence_before();
                    fence_proxy_async();
                    mbar_arrive(&sh.in_ready[s]);
                }
            }
            const int np = pass + 1;
            const int g = np < npass ? group_of(np, s) : G, tile = np < npass ? np % NT : 0;
            const bool act_next = g < G;                     // a slot only idles in the final passes
            if (act_next) {
                TC_TRACE(1);
                // -------------------------- front end -------------------------------------------------
                const int r_in = row & (RT - 1), s_idx = tile * SP + (row >> rt_shift);
                const int ray = g * RT + r_in;
                const bool valid = ray < N && s_idx < S;
                const size_t si = (size_t)ray * S + s_idx;
                float nx = 0.f, ny = 0.f, nz = 0.f, zv = 0.f;
                float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
                if (valid) {
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
                        ndc_of_point<false>(sc, sh.cams, io.rg, px, py, pz, nx, ny, nz);
                    } else {
                        px = __ldg(io.pts + si * 3); py = __ldg(io.pts + si * 3 + 1); pz = __ldg(io.pts + si * 3 + 2);
                        nx = __ldg(io.ndc + si * 3); ny = __ldg(io.ndc + si * 3 + 1); nz = __ldg(io.ndc + si * 3 + 2);
                        zv = __ldg(io.z + si);
                        dx = __ldg(io.dirs + (size_t)ray * 3); dy = __ldg(io.dirs + (size_t)ray * 3 + 1);
                        dz = __ldg(io.dirs + (size_t)ray * 3 + 2);
                    }
                }
                const float nd[3] = {nx, ny, nz};
                TC_TRACE(3);
                if (part == 0) {
                    // PE cols 0..31 = [x y z | sin(2^k x) for the first 29 of 30].  (part 0 also owns the rgb
                    // epilogue + compositing of the previous tile, so it gets the light half of the front end)
                    float v[32];
                    v[0] = nx; v[1] = ny; v[2] = nz;
                    float f = 1.f;
    #pragma unroll
                    for (int k = 0; k < 10; ++k) {
    #pragma unroll
                        for (int j = 0; j < 3; ++j)
                            if (3 + 3 * k + j < 32) v[3 + 3 * k + j] = __sinf(nd[j] * f);
                        f *= 2.f;
                    }
    #pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(slot + OFF_PE + sw128_offset(row, c * 8)) =
                            make_uint4(pack_h2(v[c * 8], v[c * 8 + 1]), pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                                       pack_h2(v[c * 8 + 4], v[c * 8 + 5]), pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
                } else {
                    // all gathers: volume (8) + colour (12) features + view direction -> MISC cols 0..47 ;
                    // PE cols 32..63 = [sin(512 z) | cos | 1].  This half starts while part 0 is still compositing.
                    float feat[20], dir[3] = {0.f, 0.f, 0.f};
    #pragma unroll
                    for (int i = 0; i < 20; ++i) feat[i] = 0.f;
                    if (valid) {
                        view_dir<false>(sh.cams, dx, dy, dz, dir);
                        sample_volume(sc, nx, ny, nz, feat);
    #pragma unroll
                        for (int v = 0; v < 3; ++v) sample_color<false>(sc, sh.cams, v, px, py, pz, feat + 8 + 4 * v);
                        if (io.input_feat) {
                            float4* o = reinterpret_cast<float4*>(io.input_feat + si * 20);
    #pragma unroll
                            for (int i = 0; i < 5; ++i)
                                o[i] = make_float4(feat[4 * i], feat[4 * i + 1], feat[4 * i + 2], feat[4 * i + 3]);
                        }
                    }
                    TC_TRACE(4);
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
                    TC_TRACE(5);
                    float v[32];
                    v[0] = __sinf(nz * 512.f);
                    float f = 1.f;
    #pragma unroll
                    for (int k = 0; k < 10; ++k) {
    #pragma unroll
                        for (int j = 0; j < 3; ++j) v[1 + 3 * k + j] = __cosf(nd[j] * f);
                        f *= 2.f;
                    }
                    v[31] = 1.f;
    #pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(slot + OFF_PE + sw128_offset(row, 32 + c * 8)) =
                            make_uint4(pack_h2(v[c * 8], v[c * 8 + 1]), pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                                       pack_h2(v[c * 8 + 4], v[c * 8 + 5]), pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
                }
                fence_proxy_async();
                TC_TRACE(2);
            }
            if (act_cur) {
                if (part == 1) {
                    mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;   // op 8 retired (keeps the barrier phases aligned)
                } else {
                    // -------------------------- op 8: rgb ----------------------------------------------------------
                    float cr, cg, cb;
                    {
                        TC_TRACE(48);
                        mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                        TC_TRACE(49);
                        tc_fence_after();
                        uint32_t r16[16];
                        tmem_ld16(t_acc, r16);
                        tmem_wait16(r16);
                        tc_fence_before();
                        cr = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[0]) + br0)));
                        cg = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[1]) + br1)));
                        cb = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[2]) + br2)));
                    }
                    // -------------------------- compositing (renderer.py:18-26,65-92) ----------------------------------
                    // every row publishes (alpha, r, g, b); the first RT threads of the slot then walk their ray's
                    // SP samples of this tile front to back -- the reference's sequential cumprod order.
                    TC_TRACE(50);
                    {
                        float4* xch = reinterpret_cast<float4*>(smem + XCH_OFFSET + s * 2048);
                        xch[row] = make_float4(1.f - __expf(-sigma), cr, cg, cb);
                        named_bar_sync(1 + s, 128);
                        TC_TRACE(51);
                        if (row < RT) {
                            if (tile_cur == 0) { cT = 1.f; c0 = c1 = c2 = c3 = c4 = 0.f; }
                            const int cray = g_cur * RT + row;
                            if (cray < N) {
                                float znear = 0.f, zfar = 0.f;
                                if (FAST) { const float4 r1 = __ldg(reinterpret_cast<const float4*>(io.rays + (size_t)cray * 8) + 1); znear = r1.z; zfar = r1.w; }
                                for (int sub = 0; sub < SP; ++sub) {
                                    const int sj = tile_cur * SP + sub;
                                    if (sj >= S) break;
                                    const float4 v = xch[sub * RT + row];
                                    float z;
                                    if (FAST) {
                                        const float t = __ldg(io.t_steps + sj);
                                        if (!io.rg.lindisp) z = __fadd_rn(__fmul_rn(znear, 1.f - t), __fmul_rn(zfar, t));
                                        else z = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, znear), 1.f - t), __fmul_rn(__fdiv_rn(1.f, zfar), t)));
                                    } else {
                                        z = __ldg(io.z + (size_t)cray * S + sj);
                                    }
                                    const float wgt = v.x * cT;
                                    if (io.alpha) io.alpha[(size_t)cray * S + sj] = v.x;
                                    if (io.weights) io.weights[(size_t)cray * S + sj] = wgt;
                                    c0 = fmaf(wgt, v.y, c0); c1 = fmaf(wgt, v.z, c1); c2 = fmaf(wgt, v.w, c2);
                                    c3 = fmaf(wgt, z, c3); c4 += wgt;
                                    cT *= (1.f - v.x) + 1e-10f;
                                }
                                if (tile_cur == NT - 1) {
                                    float o0 = c0, o1 = c1, o2 = c2;
                                    if (sc.white_bkgd) { const float bg = 1.f - c4; o0 += bg; o1 += bg; o2 += bg; }
                                    store_pixel(io, cray, o0, o1, o2, c3);
                                }
                            }
                        }
                        TC_TRACE(53);
                    }
                }
            }
            if (act_next) mbar_arrive(&sh.in_ready[s]);
            g_cur = g; tile_cur = tile; act_cur = act_next;
        }
    } else if (warp == 16) {
        // =========================== MMA issuer ========================================================
        // the whole warp walks the schedule (waits are warp-uniform); one elected lane issues
        const bool leader = elect_one();
        {
            uint32_t par_in[2] = {0, 0};
            const uint32_t sbase = smem_u32(smem);
            const uint32_t ring = sbase + RING_OFFSET;
            // descriptors: hi word constant per layout, lo word = (addr >> 4) | LBO field; a K-step adds 32 B = 2 units
            constexpr uint32_t HI_SW = (uint32_t)(desc_sw128(0) >> 32), HI_NS = (uint32_t)(desc_nosw(0, 128, 256) >> 32);
            constexpr uint32_t LO_SW = (uint32_t)desc_sw128(0), LO_NS = (uint32_t)desc_nosw(0, 128, 256);
            auto dsw = [&](uint32_t addr) { return ((uint64_t)HI_SW << 32) | (uint64_t)(LO_SW | (addr >> 4)); };
            auto dns = [&](uint32_t addr) { return ((uint64_t)HI_NS << 32) | (uint64_t)(LO_NS | (addr >> 4)); };
            uint32_t nchunk_base = 0;                         // global chunk counter at the start of the pass
            uint32_t op_base = 0;                             // global GEMM-op counter at the start of the pass
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass, op_base += 9) {
                const bool act[2] = {group_of(pass, 0) < G, group_of(pass, 1) < G};
                // one generic body per (op, slot), driven by the op tables (kept small on purpose: this code
                // runs once per pass and must not evict the epilogue loop from the instruction cache)
#pragma unroll 1
                for (int op = 0; op < 9; ++op) {
                    const int nblk = c_op_nblk[op];
                    const uint32_t idesc = c_op_idesc[op];
#pragma unroll 1
                    for (int s = 0; s < 2; ++s) {
                        if (!act[s]) continue;
                        const bool first = (s == 0) || !act[0], last = (s == 1) || !act[1];
                        if (first) {                                      // weights first: they normally landed long ago
#pragma unroll 1
                            for (int b = 0; b < nblk; ++b) {
                                const uint32_t n = nchunk_base + c_blk_chunk[op][b];
                                mbar_wait(&sh.w_full[n % tcw::NSTAGE], (n / tcw::NSTAGE) & 1);
                            }
                        }
                        TC_TRACE(100 + op * 2 + s);
                        mbar_wait(&sh.in_ready[s], par_in[s]); par_in[s] ^= 1;
                        TC_TRACE(140 + op * 2 + s);
                        tc_fence_after();
                        const uint32_t sl = sbase + s * SLOT_BYTES;
                        const uint32_t d_acc = tmem + s * 256;
                        TC_TRACE(60 + op * 2 + s);
                        if (leader) {
                            uint32_t last_stage = 0;
#pragma unroll 1
                            for (int b = 0; b < nblk; ++b) {
                                const uint32_t st = ring + ((nchunk_base + c_blk_chunk[op][b]) % tcw::NSTAGE) * tcw::STAGE_BYTES;
                                last_stage = st;
                                const bool to_mod = (op == 0 && b == 0);          // modulation GEMM: K = 32, own TMEM columns
                                const uint32_t d = to_mod ? d_acc + 128 : d_acc;
                                const uint64_t da = dsw(sl + c_blk_aoff[op][b]), db = dsw(st);
                                const uint32_t acc0 = (b > 0 && op != 0) ? 1u : 0u;
                                mma_f16(d, da, db, idesc, acc0);
                                mma_f16(d, da + 2, db + 2, idesc, 1);
                                if (!to_mod) {
                                    mma_f16(d, da + 4, db + 4, idesc, 1);
                                    mma_f16(d, da + 6, db + 6, idesc, 1);
                                }
                            }
                            if (c_op_bias_aoff[op])                               // constant-one column x bias row ("bias step")
                                mma_f16(d_acc, dsw(sl + c_op_bias_aoff[op]), dns(last_stage + c_op_bias_boff[op]), idesc, 1);
                            TC_TRACE(80 + op * 2 + s);
                            mma_commit(&sh.acc_ready[s]);
                            if (last) mma_commit(&sh.w_free[(op_base + op) & 3]);   // this op's chunks are free once these MMAs retire
                        }
                        __syncwarp();
                    }
                }
                TC_TRACE(200);
                nchunk_base += tcw::NCHUNK;
            }
        }
    } else {
        // =========================== weight loader ========================================================
        if (elect_one()) {
            uint8_t* ring = smem + RING_OFFSET;
            uint32_t n = 0;                                 // chunks issued
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
#pragma unroll 1
                for (int c = 0; c < tcw::NCHUNK; ++c, ++n) {
                    const uint32_t st = n % tcw::NSTAGE;
                    if (n >= (uint32_t)tcw::NSTAGE) {
                        // the stage's previous tenant is chunk n - NSTAGE; wait until its op has been released
                        const uint32_t pn = n - tcw::NSTAGE;
                        const uint32_t need = (pn / tcw::NCHUNK) * 9 + (uint32_t)c_chunk_op[pn % tcw::NCHUNK];   // global op index
                        mbar_wait(&sh.w_free[need & 3], (need >> 2) & 1);
                    }
                    const uint32_t bytes = (uint32_t)tcw::chunk_bytes(c);
                    mbar_arrive_expect_tx(&sh.w_full[st], bytes);
                    bulk_load(ring + st * tcw::STAGE_BYTES, wimg + tcw::chunk_offset(c), bytes, &sh.w_full[st]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 512);
}

int launch_render_tc(const SceneDev& sc, const RenderIO& io_in, bool fast, const void* wimg, cudaStream_t stream) {
    RenderIO io = io_in;
    static bool attr_set[64] = {false};                   // once per device, not per launch
    int dev = 0;
    MVSN_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        if (dev < 64) attr_set[dev] = true;
    }
    // rays per tile: 32 (best gather locality) unless the batch is too small to give every SM a pair of groups
    int rt = 32;
    while (rt > 4 && (io.N + rt - 1) / rt < 2 * sm_count()) rt >>= 1;
    io.rays_per_tile = rt;
    const int G = (io.N + rt - 1) / rt;
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

// K-C, fp32-grade tensor-core mode (MVSN_MLP_TC_SPLIT): the fused per-ray render kernel with the MLP on
// tcgen05 tensor cores at fp32-class accuracy (north-star fp32 gate, RGB Linf <= 1e-4).
//
// Every GEMM operand is carried as a two-term fp16 split  x = hi + lo  (hi = fp16(x), lo = fp16(x - hi),
// 22 significant bits) and every K-step issues three MMAs into the same fp32 TMEM accumulator:
//     D += A_hi * B_hi + A_hi * B_lo + A_lo * B_hi          (the dropped lo*lo term is ~2^-22 relative)
// Simulated on a 128x160 scene this reproduces the fp32 reference to 4.8e-7 RGB Linf, against
// 6e-4 for single fp16 operands.  Everything outside the GEMMs (gather, encoding, modulation multiply,
// activations, compositing) is fp32 with the same operation order as the FFMA kernel.
//
// fp16 subnormals do not survive the tensor core, so operands are kept in range by exact power-of-two scales:
// activations are stored x16, weights and biases x256 (so the lo terms of weights >= 5e-4 and of activations
// >= 8e-3 are normal numbers); accumulators therefore carry 4096 x the true value and the epilogues undo it
// exactly (2^-20 with the x16 re-scale of the next operand folded in).
//
// Structure = render_tc.cu with ONE tile in flight per CTA (the doubled operand tiles take 128 KB):
//   warps 0-7  slot group  : front end, per-layer epilogues (TMEM -> fp32 math -> hi/lo fp16 tiles), compositing
//   warp 8     MMA issuer  : three tcgen05.mma per K-step, tcgen05.commit hand-offs
//   warp 9     weight loader: streams the (hi | lo) weight chunks through a 2-stage ring with bulk async copies
// Tile = RT adjacent rays x 128/RT samples (lanes = adjacent rays), compositing state carried in registers.
#include "render_frontend.cuh"
#include "umma.cuh"

namespace mvsn {

using namespace umma;

namespace tcs {     // weight image: 16 chunks in consumption order, each stored as [hi image | lo image]
// chunks 0..12 as in the single-fp16 round-1 kernel (modulation, layer 0, layers 1..4 x 2 K-blocks, layer 5 x 3);
// 13, 14: feature_linear and views_linears[0] FOLDED at pack time (no non-linearity between them, models.py:213-218):
//         W' = Wv[:, :128] Wf (64 rows) + alpha_linear as row 64 (sigma), N = 80; chunk 14 also carries the [dir | 1] tile;
// 15: rgb_linear.   8 GEMM phases per tile instead of 9.
constexpr int NCHUNK = 16;
__host__ __device__ constexpr int part_bytes(int c) {           // bytes of ONE part (hi or lo) of chunk c
    return c == 0 || c == 1 || c == 2 || c == 4 || c == 6 || c == 8 || c == 10 || c == 11 || c == 12 ? 16384
         : c == 3 || c == 5 || c == 7 || c == 9 ? 20480
         : c == 13 ? 10240 : c == 14 ? 13312 /* 10 240 + 2 560 padded: the lo image must start 1024-aligned */
         : 2048;
}
__host__ __device__ constexpr int chunk_offset(int c) {
    int o = 0;
    for (int i = 0; i < c; ++i) o += 2 * part_bytes(i);
    return o;
}
constexpr int STREAM_BYTES = chunk_offset(NCHUNK);
constexpr int TAIL_OFFSET = STREAM_BYTES;                   // fp32 tail: rgb_linear.bias[3], 0
constexpr int TOTAL_BYTES = STREAM_BYTES + 16;
constexpr int STAGE_BYTES = 40960;                          // = 2 x 20 480 (largest chunk), 1024-aligned
constexpr int NSTAGE = 2;
constexpr int N_VIEWS_OP = 80;                              // 64 views-layer outputs + sigma + 15 zero rows
}  // namespace tcs

namespace {

constexpr float SA = 16.f, SW = 256.f;                      // operand scales (activations, weights)
constexpr int THREADS = 320;                                // 8 slot warps + MMA issuer + weight loader
// operand tiles (SWIZZLE_128B K-blocks of 128 rows x 64 fp16), hi and lo copies
constexpr int OFF_PE = 0, OFF_H0 = 32768, OFF_H1 = 65536, OFF_MISC = 98304, LO = 16384;
constexpr int RING_OFFSET = 131072;
constexpr int XCH_OFFSET = RING_OFFSET + tcs::NSTAGE * tcs::STAGE_BYTES;
constexpr int SMEM_BYTES = XCH_OFFSET + 2048 + 1024;

struct Shared {
    uint64_t in_ready;          // slot group (256 arrivals) -> MMA issuer: operand tiles written
    uint64_t acc_ready;         // tcgen05.commit -> slot group: accumulator complete
    uint64_t w_full[tcs::NSTAGE];
    uint64_t w_empty[tcs::NSTAGE];
    uint32_t tmem_base;
    Cams cams;
};

__constant__ int c_op_nblk[8] = {2, 2, 2, 2, 2, 3, 2, 1};
__constant__ uint32_t c_op_idesc[8] = {idesc_f16(128, 128), idesc_f16(128, 128), idesc_f16(128, 128), idesc_f16(128, 128),
                                       idesc_f16(128, 128), idesc_f16(128, 128), idesc_f16(128, tcs::N_VIEWS_OP), idesc_f16(128, 16)};
__constant__ int c_blk_chunk[8][3] = {{0, 1, 0}, {2, 3, 0}, {4, 5, 0}, {6, 7, 0}, {8, 9, 0}, {10, 11, 12}, {13, 14, 0}, {15, 0, 0}};
__constant__ uint32_t c_blk_aoff[8][3] = {{OFF_MISC, OFF_PE, 0}, {OFF_H0, OFF_H1, 0}, {OFF_H0, OFF_H1, 0}, {OFF_H0, OFF_H1, 0}, {OFF_H0, OFF_H1, 0},
                                          {OFF_PE, OFF_H0, OFF_H1}, {OFF_H0, OFF_H1, 0}, {OFF_H0, 0, 0}};
// "bias step": A = 16 MISC columns holding a constant one (cols 16..31 for the trunk, 32..47 = [dir, 1] for the folded views
// layer), B = a 16-wide no-swizzle tile stored after the K-block inside the op's last chunk (byte offset inside one part)
__constant__ uint32_t c_op_bias_aoff[8] = {0, OFF_MISC + 32, OFF_MISC + 32, OFF_MISC + 32, OFF_MISC + 32, 0, OFF_MISC + 64, 0};
__constant__ uint32_t c_op_bias_boff[8] = {0, 16384, 16384, 16384, 16384, 0, 10240, 0};
constexpr int OP_VIEWS = 6;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// two fp32 values -> (hi pair, lo pair) packed fp16x2
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// eight fp32 values -> one 16-byte chunk of the hi tile and of the lo tile (chunk kc of row `row`)
__device__ __forceinline__ void store_split8(uint8_t* tile_hi, int row, int kc, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
    const uint32_t off = sw128_offset(row, kc * 8);
    *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(tile_hi + LO + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// epilogue over NCOL accumulator columns: (x modulation) -> relu -> hi/lo fp16 -> K-block `blk` (+LO), chunks kc0..
template <int NCOL, bool MODULATE, bool RELU>
__device__ __forceinline__ void epilogue(uint32_t t_acc, uint32_t t_mod, uint8_t* blk, int row, int kc0, float scale) {
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
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float x = __uint_as_float(a[i & 1][j]) * scale;       // exact: scale is a power of two
            if (MODULATE) x *= __uint_as_float(m[i & 1][j]);
            if (RELU) x = fmaxf(x, 0.f);
            v[j] = x;
        }
        store_split8(blk, row, kc0 + 2 * i, v);
        store_split8(blk, row, kc0 + 2 * i + 1, v + 8);
    }
}

template <bool FAST>
__global__ void __launch_bounds__(THREADS, 1)
render_tcs_kernel(const SceneDev sc, const RenderIO io, const uint8_t* __restrict__ wimg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ Shared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    load_cams(sc, &sh.cams, tid);
    if (tid == 0) {
        mbar_init(&sh.in_ready, 256);
        mbar_init(&sh.acc_ready, 1);
        for (int i = 0; i < tcs::NSTAGE; ++i) { mbar_init(&sh.w_full[i], 1); mbar_init(&sh.w_empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 8) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;                      // accumulator: cols 0..143, modulation: cols 256..383

    // ---- work decomposition: tile = RT rays x SP samples, NT tiles per ray group, groups strided over the grid
    const int N = io.N, S = io.S;
    const int RT = io.rays_per_tile, SP = 128 / RT;
    const int rt_shift = 31 - __clz(RT);
    const int NT = (S + SP - 1) / SP;
    const int G = (N + RT - 1) / RT;
    const int my_groups = (int)blockIdx.x < G ? (G - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int ntiles = my_groups * NT;
    auto group_of = [&](int t) { return (t / NT) * (int)gridDim.x + (int)blockIdx.x; };

    if (warp < 8) {
        // =========================== slot group =========================================================
        const int part = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
        const uint32_t t_acc = tmem + ((uint32_t)(wq * 32) << 16);
        const uint32_t t_mod = t_acc + 256;
        uint32_t par_acc = 0;
        float cT = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f;   // compositing state of ray `row` (row < RT)
        const float br0 = __ldg(reinterpret_cast<const float*>(wimg + tcs::TAIL_OFFSET));
        const float br1 = __ldg(reinterpret_cast<const float*>(wimg + tcs::TAIL_OFFSET) + 1);
        const float br2 = __ldg(reinterpret_cast<const float*>(wimg + tcs::TAIL_OFFSET) + 2);
        uint8_t* hblk = smem + (part ? OFF_H1 : OFF_H0);

        float sigma = 0.f;
        int g_cur = 0, tile_cur = 0;
        bool act_cur = false;
        // rotated by one tile: the next tile's operand tiles are built between op 6 and op 7 (rgb) of the tile in flight
        for (int t = -1; t < ntiles; ++t) {
            if (act_cur) {
                // ---- trunk ops 0..5: h = relu((W h + b) * modulation)
#pragma unroll 1
                for (int op = 0; op < 6; ++op) {
                    mbar_wait(&sh.acc_ready, par_acc); par_acc ^= 1;
                    tc_fence_after();
                    epilogue<64, true, true>(t_acc + part * 64, t_mod + part * 64, hblk, row, 0, SA / ((SA * SW) * (SA * SW)));   // 16 * (acc'/4096) * (mod'/4096)
                    tc_fence_before();
                    fence_proxy_async();
                    mbar_arrive(&sh.in_ready);
                }
                // ---- op 6: folded feature -> views layer (64 cols, relu) + sigma (col 64); hv -> H0 (hi/lo)
                {
                    mbar_wait(&sh.acc_ready, par_acc); par_acc ^= 1;
                    tc_fence_after();
                    epilogue<32, false, true>(t_acc + part * 32, t_mod, smem + OFF_H0, row, part * 4, 1.f / SW);                // 16 hv
                    sigma = 0.f;
                    if (part == 0) {
                        uint32_t r16[16];
                        tmem_ld16(t_acc + 64, r16);
                        tmem_wait16(r16);
                        sigma = fmaxf(__uint_as_float(r16[0]) * (1.f / (SA * SW)), 0.f);
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    mbar_arrive(&sh.in_ready);
                }
            }
            // ---- front end of the next tile (fp32-exact path: same arithmetic as the FFMA kernel)
            const int nt = t + 1;
            const int g = nt < ntiles ? group_of(nt) : G, tile = nt < ntiles ? nt % NT : 0;
            const bool act_next = g < G;
            if (act_next) {
                const int r_in = row & (RT - 1), s_idx = tile * SP + (row >> rt_shift);
                const int ray = g * RT + r_in;
                const bool valid = ray < N && s_idx < S;
                const size_t si = (size_t)ray * S + s_idx;
                float nx = 0.f, ny = 0.f, nz = 0.f;
                float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
                if (valid) {
                    if (FAST) {
                        const float4* rp = reinterpret_cast<const float4*>(io.rays + (size_t)ray * 8);
                        float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                        dx = r0.w; dy = r1.x; dz = r1.y;
                        const float near = r1.z, far = r1.w, tt = __ldg(io.t_steps + s_idx);
                        float zv;
                        if (!io.rg.lindisp) zv = __fadd_rn(__fmul_rn(near, 1.f - tt), __fmul_rn(far, tt));
                        else zv = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, near), 1.f - tt),
                                                           __fmul_rn(__fdiv_rn(1.f, far), tt)));
                        px = __fadd_rn(r0.x, __fmul_rn(dx, zv));
                        py = __fadd_rn(r0.y, __fmul_rn(dy, zv));
                        pz = __fadd_rn(r0.z, __fmul_rn(dz, zv));
                        ndc_of_point<true>(sc, sh.cams, io.rg, px, py, pz, nx, ny, nz);
                    } else {
                        px = __ldg(io.pts + si * 3); py = __ldg(io.pts + si * 3 + 1); pz = __ldg(io.pts + si * 3 + 2);
                        nx = __ldg(io.ndc + si * 3); ny = __ldg(io.ndc + si * 3 + 1); nz = __ldg(io.ndc + si * 3 + 2);
                        dx = __ldg(io.dirs + (size_t)ray * 3); dy = __ldg(io.dirs + (size_t)ray * 3 + 1);
                        dz = __ldg(io.dirs + (size_t)ray * 3 + 2);
                    }
                }
                const float nd[3] = {nx, ny, nz};
                if (part == 0) {
                    // PE cols 0..31 = [x y z | sin(2^k x), first 29 of 30]
                    float v[32];
                    v[0] = nx; v[1] = ny; v[2] = nz;
                    float f = 1.f;
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            if (3 + 3 * k + j < 32) v[3 + 3 * k + j] = sinf(nd[j] * f);
                        f *= 2.f;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= SA;
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_split8(smem + OFF_PE, row, c, v + 8 * c);
                } else {
                    // all gathers -> MISC cols 0..47 ; PE cols 32..63 = [sin(512 z) | cos | 1]
                    float feat[24], dir[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 24; ++i) feat[i] = 0.f;
                    if (valid) {
                        view_dir<true>(sh.cams, dx, dy, dz, dir);
                        sample_volume(sc, nx, ny, nz, feat);
#pragma unroll
                        for (int v = 0; v < 3; ++v) sample_color<true>(sc, sh.cams, v, px, py, pz, feat + 8 + 4 * v);
                        if (io.input_feat) {
                            float4* o = reinterpret_cast<float4*>(io.input_feat + si * 20);
#pragma unroll
                            for (int i = 0; i < 5; ++i)
                                o[i] = make_float4(feat[4 * i], feat[4 * i + 1], feat[4 * i + 2], feat[4 * i + 3]);
                        }
                    }
                    feat[20] = 1.f;                                       // the constant-one column of the bias steps
#pragma unroll
                    for (int i = 0; i < 24; ++i) feat[i] *= SA;
                    store_split8(smem + OFF_MISC, row, 0, feat);
                    store_split8(smem + OFF_MISC, row, 1, feat + 8);
                    store_split8(smem + OFF_MISC, row, 2, feat + 16);
                    const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    store_split8(smem + OFF_MISC, row, 3, z8);
                    const float d8[8] = {dir[0] * SA, dir[1] * SA, dir[2] * SA, SA, 0.f, 0.f, 0.f, 0.f};
                    store_split8(smem + OFF_MISC, row, 4, d8);
                    store_split8(smem + OFF_MISC, row, 5, z8);
                    float v[32];
                    v[0] = sinf(nz * 512.f);
                    float f = 1.f;
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) v[1 + 3 * k + j] = cosf(nd[j] * f);
                        f *= 2.f;
                    }
                    v[31] = 1.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= SA;
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_split8(smem + OFF_PE, row, 4 + c, v + 8 * c);
                }
                fence_proxy_async();
            }
            if (act_cur) {
                if (part == 1) {
                    mbar_wait(&sh.acc_ready, par_acc); par_acc ^= 1;     // op 7 retired (keeps the barrier phases aligned)
                } else {
                    // ---- op 7: rgb
                    float cr, cg, cb;
                    {
                        mbar_wait(&sh.acc_ready, par_acc); par_acc ^= 1;
                        tc_fence_after();
                        uint32_t r16[16];
                        tmem_ld16(t_acc, r16);
                        tmem_wait16(r16);
                        tc_fence_before();
                        const float us = 1.f / (SA * SW);
                        cr = __fdiv_rn(1.f, 1.f + expf(-(__uint_as_float(r16[0]) * us + br0)));
                        cg = __fdiv_rn(1.f, 1.f + expf(-(__uint_as_float(r16[1]) * us + br1)));
                        cb = __fdiv_rn(1.f, 1.f + expf(-(__uint_as_float(r16[2]) * us + br2)));
                    }
                    // ---- compositing (renderer.py:18-26,65-92): sequential per ray, the reference's cumprod order
                    float4* xch = reinterpret_cast<float4*>(smem + XCH_OFFSET);
                    xch[row] = make_float4(1.f - expf(-sigma), cr, cg, cb);
                    named_bar_sync(1, 128);
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
                                    const float tt = __ldg(io.t_steps + sj);
                                    if (!io.rg.lindisp) z = __fadd_rn(__fmul_rn(znear, 1.f - tt), __fmul_rn(zfar, tt));
                                    else z = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, znear), 1.f - tt), __fmul_rn(__fdiv_rn(1.f, zfar), tt)));
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
                }
            }
            if (act_next) mbar_arrive(&sh.in_ready);
            g_cur = g; tile_cur = tile; act_cur = act_next;
        }
    } else if (warp == 8) {
        // =========================== MMA issuer ==========================================================
        const bool leader = elect_one();
        uint32_t par_in = 0;
        const uint32_t sbase = smem_u32(smem);
        const uint32_t ring = sbase + RING_OFFSET;
        constexpr uint32_t HI_SW = (uint32_t)(desc_sw128(0) >> 32), HI_NS = (uint32_t)(desc_nosw(0, 128, 256) >> 32);
        constexpr uint32_t LO_SW = (uint32_t)desc_sw128(0), LO_NS = (uint32_t)desc_nosw(0, 128, 256);
        auto dsw = [&](uint32_t addr) { return ((uint64_t)HI_SW << 32) | (uint64_t)(LO_SW | (addr >> 4)); };
        auto dns = [&](uint32_t addr) { return ((uint64_t)HI_NS << 32) | (uint64_t)(LO_NS | (addr >> 4)); };
        uint32_t n = 0;                                        // global chunk counter
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll 1
            for (int op = 0; op < 8; ++op) {
                const int nblk = c_op_nblk[op];
                const uint32_t idesc = c_op_idesc[op];
                mbar_wait(&sh.in_ready, par_in); par_in ^= 1;
#pragma unroll 1
                for (int b = 0; b < nblk; ++b, ++n) {
                    const int c = c_blk_chunk[op][b];
                    const uint32_t st = n % tcs::NSTAGE;
                    mbar_wait(&sh.w_full[st], (n / tcs::NSTAGE) & 1);
                    tc_fence_after();
                    const uint32_t w_hi = ring + st * tcs::STAGE_BYTES, w_lo = w_hi + (uint32_t)tcs::part_bytes(c);
                    const bool to_mod = (op == 0 && b == 0);          // modulation GEMM: K = 32, own TMEM columns
                    const uint32_t d = to_mod ? tmem + 256 : tmem;
                    const uint32_t a_hi = sbase + c_blk_aoff[op][b];
                    const uint64_t dah = dsw(a_hi), dal = dsw(a_hi + LO), dbh = dsw(w_hi), dbl = dsw(w_lo);
                    const int nsteps = to_mod ? 2 : 4;
                    const bool last_blk = (b == nblk - 1);
                    if (leader) {
                        uint32_t accum = (b > 0 && op != 0) ? 1u : 0u;
#pragma unroll 1
                        for (int ks = 0; ks < nsteps; ++ks) {
                            mma_f16(d, dah + 2 * ks, dbh + 2 * ks, idesc, accum);      // hi * hi
                            mma_f16(d, dah + 2 * ks, dbl + 2 * ks, idesc, 1);          // hi * lo
                            mma_f16(d, dal + 2 * ks, dbh + 2 * ks, idesc, 1);          // lo * hi
                            accum = 1;
                        }
                        if (last_blk && c_op_bias_aoff[op]) {                          // bias step: [.., 1] x bias rows
                            const uint32_t ab = sbase + c_op_bias_aoff[op];
                            const uint32_t boff = c_op_bias_boff[op];
                            mma_f16(d, dsw(ab), dns(w_hi + boff), idesc, 1);
                            mma_f16(d, dsw(ab), dns(w_lo + boff), idesc, 1);
                            if (op == OP_VIEWS) mma_f16(d, dsw(ab + LO), dns(w_hi + boff), idesc, 1);   // dir_lo * W_hi
                        }
                        if (last_blk) mma_commit(&sh.acc_ready);
                        mma_commit(&sh.w_empty[st]);                                   // strict full/empty alternation per stage
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 9) {
        // =========================== weight loader =========================================================
        if (elect_one()) {
            uint8_t* ring = smem + RING_OFFSET;
            uint32_t n = 0;
#pragma unroll 1
            for (int t = 0; t < ntiles; ++t) {
#pragma unroll 1
                for (int c = 0; c < tcs::NCHUNK; ++c, ++n) {
                    const uint32_t st = n % tcs::NSTAGE;
                    if (n >= (uint32_t)tcs::NSTAGE) mbar_wait(&sh.w_empty[st], ((n / tcs::NSTAGE) - 1u) & 1u);
                    const uint32_t bytes = 2u * (uint32_t)tcs::part_bytes(c);
                    mbar_arrive_expect_tx(&sh.w_full[st], bytes);
                    bulk_load(ring + st * tcs::STAGE_BYTES, wimg + tcs::chunk_offset(c), bytes, &sh.w_full[st]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// weight image packer: fp32 nn.Linear tensors -> (hi | lo) fp16 pre-swizzled chunks
// ------------------------------------------------------------------------------------------------
struct MlpPtrs { const float* p[MVSN_N_MLP_TENSORS]; };

__device__ __forceinline__ uint32_t nosw_offset(int r, int k) {
    return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

__global__ void pack_mlp_tcs_kernel(MlpPtrs w, uint8_t* __restrict__ out) {
    const int c = blockIdx.x;
    uint8_t* dst = out + tcs::chunk_offset(c);
    const int pb = tcs::part_bytes(c);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid * 16; i < 2 * pb; i += nt * 16) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    auto put = [&](uint32_t off, float v) {
        v *= SW;
        const __half h = __float2half_rn(v);
        *reinterpret_cast<__half*>(dst + off) = h;
        *reinterpret_cast<__half*>(dst + pb + off) = __float2half_rn(v - __half2float(h));
    };
    if (c == 0) {
        for (int i = tid; i < 128 * 21; i += nt) { const int r = i / 21, k = i % 21;
            put(sw128_offset(r, k), k < 20 ? w.p[12][r * 20 + k] : w.p[13][r]); }
    } else if (c == 1) {
        for (int i = tid; i < 128 * 64; i += nt) { const int r = i / 64, k = i % 64;
            put(sw128_offset(r, k), k < 63 ? w.p[0][r * 63 + k] : w.p[1][r]); }
    } else if (c >= 2 && c <= 9) {
        const int l = c / 2, kb = c & 1;
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
        // folded views layer: rows n < 64: W'[n][k] = sum_j Wv[n][j] Wf[j][k] (fp64) ; row 64: alpha_linear ; rows 65..79 zero
        const int kb = c - 13;
        for (int i = tid; i < 65 * 64; i += nt) {
            const int r = i / 64, k = i % 64, kk = kb * 64 + k;
            float v;
            if (r < 64) {
                double acc = 0.0;
                for (int j = 0; j < 128; ++j) acc += (double)w.p[14][r * 131 + j] * (double)w.p[16][j * 128 + kk];
                v = (float)acc;
            } else {
                v = w.p[18][kk];
            }
            put(sw128_offset(r, k), v);
        }
        if (kb == 1)            // [dir | 1] tile: k < 3: Wv[:, 128 + k] ; k = 3: b' = Wv[:, :128] bf + bv (row 64: alpha bias)
            for (int i = tid; i < 65 * 4; i += nt) {
                const int r = i / 4, k = i % 4;
                float v = 0.f;
                if (r < 64) {
                    if (k < 3) v = w.p[14][r * 131 + 128 + k];
                    else {
                        double acc = (double)w.p[15][r];
                        for (int j = 0; j < 128; ++j) acc += (double)w.p[14][r * 131 + j] * (double)w.p[17][j];
                        v = (float)acc;
                    }
                } else if (k == 3) {
                    v = w.p[19][0];
                }
                put(10240 + nosw_offset(r, k), v);
            }
    } else if (c == 15) {
        for (int i = tid; i < 3 * 64; i += nt) { const int r = i / 64, k = i % 64; put(sw128_offset(r, k), w.p[20][r * 64 + k]); }
        if (tid < 4) reinterpret_cast<float*>(out + tcs::TAIL_OFFSET)[tid] = tid < 3 ? w.p[21][tid] : 0.f;
    }
}

}  // namespace

size_t mlp_tcs_packed_bytes() { return tcs::TOTAL_BYTES; }

int pack_mlp_tcs(const float* const* w, void* packed, cudaStream_t stream) {
    MlpPtrs p;
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i) p.p[i] = w[i];
    pack_mlp_tcs_kernel<<<tcs::NCHUNK, 256, 0, stream>>>(p, static_cast<uint8_t*>(packed));
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

int launch_render_tcs(const SceneDev& sc, const RenderIO& io_in, bool fast, const void* wimg, cudaStream_t stream) {
    RenderIO io = io_in;
    static bool attr_set[64] = {false};                   // once per device, not per launch
    int dev = 0;
    MVSN_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tcs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tcs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        if (dev < 64) attr_set[dev] = true;
    }
    int rt = 32;                                             // rays per tile, as in the single-fp16 kernel
    while (rt > 4 && (io.N + rt - 1) / rt < sm_count()) rt >>= 1;
    io.rays_per_tile = rt;
    const int G = (io.N + rt - 1) / rt;
    const int grid = G < sm_count() ? G : sm_count();
    if (grid <= 0) return MVSN_OK;
    const uint8_t* w = static_cast<const uint8_t*>(wimg);
    if (fast) render_tcs_kernel<true><<<grid, THREADS, SMEM_BYTES, stream>>>(sc, io, w);
    else      render_tcs_kernel<false><<<grid, THREADS, SMEM_BYTES, stream>>>(sc, io, w);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // namespace mvsn

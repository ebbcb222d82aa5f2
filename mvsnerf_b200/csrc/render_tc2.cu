// K-C, tensor-core mode (MVSN_MLP_TC_HALF), round-2 kernel: the fused per-ray render kernel on CTA PAIRS
// (tcgen05 cta_group::2) with the whole weight set RESIDENT in shared memory and the hidden activations
// living in TENSOR MEMORY.
//
// What changed against the round-1 kernel (render_tc.cu, kept as MVSN_MLP_TC_HALF_V1) and why:
//   * a cluster of two CTAs issues every MMA as ONE M = 256 instruction (128 sample rows per CTA) whose B operand
//     (the weights) is split by output row across the pair, so each SM holds only HALF of every weight matrix:
//     125 KB of fp16 -- the full MLP stays resident.  No weight streaming at all (round 1 re-streamed 291 KB from
//     L2 per pair of tiles: 2.8 TB/s of L2 traffic and a loader whose latency sat on the critical path);
//   * the 128-wide hidden activation H never touches shared memory: the epilogue thread of a row converts its
//     accumulator row to fp16 and writes it back into TMEM (tcgen05.st), and the next layer's MMA takes its A
//     operand straight from TMEM.  No swizzled st.shared, no generic->async proxy fence on the hot hand-off, and
//     the MMA reads only B from shared memory;
//   * the multiplicative modulation (models.py:199-203) is kept as packed fp16 in TMEM (64 columns instead of
//     128 fp32), halving the epilogue's TMEM reads; it is produced by op 0 into a landing zone that is then
//     re-used for H and for the packed modulation itself;
//   * front end and epilogue are different warps: four producer warps per slot build the NEXT tile's positional
//     encoding / gathered features while the four epilogue warps of the slot walk the current tile's layers, so
//     the ~2.7 k-cycle front end is off the tensor pipe's critical path;
//   * feature_linear and views_linears[0] have no non-linearity between them (models.py:213-218), so they are
//     folded at pack time into ONE GEMM from h:  W' = Wv[:, :128] Wf,  b' = Wv[:, :128] bf + bv  (fp64 on the
//     packer).  sigma rides as an extra output row of that GEMM.  8 GEMM phases per tile instead of 9.
//
// Roles per CTA (544 threads):  warps 0-3 / 4-7  epilogue + compositing of slot 0 / 1 (thread = sample row,
// warp % 4 = its TMEM lane quarter);  warps 8-11 / 12-15  producers of slot 0 / 1 (thread = sample row);
// warp 16  MMA issuer (leader CTA only).  Two 128-row tiles ("slots") are in flight per CTA, i.e. four per pair.
//
// TMEM (512 columns, per slot 256):  [0,128) fp32 accumulator | [128,192) H as packed fp16 (A operand) |
// [192,256) modulation as packed fp16.   Columns [128,256) double as the fp32 landing zone of the modulation
// GEMM during op 0.
// SMEM: [0, 125 KB) this CTA's half of the weights | per slot PE 16 KB + MISC 16 KB (SW128 K-blocks, A operands
// of the layers that read the encoding / gathered features) | 4 KB compositing exchange.
//
// Replaces renderer.rendering (renderer.py:138-165) and callees; see include/mvsnerf_b200.h.
#include "render_frontend.cuh"
#include "umma.cuh"

namespace mvsn {

using namespace umma;

namespace tc2 {     // per-CTA half of the weight image (bytes); the packed buffer is [rank 0 | rank 1 | fp32 tail]
constexpr int W_MOD = 0;                        // [64 x 64] SW128: K 0..19 pts_bias.weight, K 20 its bias
constexpr int W_L0 = 8192;                      // [64 x 64] SW128: K 0..62 pts_linears.0, K 63 its bias
constexpr int W_L1 = 16384;                     // layers 1..4: h K-block 0 | h K-block 1 | [64 x 16] bias tile
constexpr int L_STRIDE = 18432;
constexpr int W_L5 = W_L1 + 4 * L_STRIDE;       // pe K-block (K 63 = bias) | h K-block 0 | h K-block 1
constexpr int W_V = W_L5 + 3 * 8192;            // folded views layer + sigma, N = 80 -> 40 rows: kb0 | kb1 | [40 x 16] dir/bias tile
constexpr int W_V_TILE = 10240;
constexpr int W_RGB = W_V + 12288;              // [8 x 64] SW128 (N = 16 -> 8 rows per CTA)
constexpr int HALF_BYTES = W_RGB + 1024;        // 128 000
constexpr int TAIL_OFFSET = 2 * HALF_BYTES;     // fp32: rgb_linear.bias[3], 0
constexpr int TOTAL_BYTES = TAIL_OFFSET + 16;
static_assert(W_L5 % 1024 == 0 && W_V % 1024 == 0 && W_RGB % 1024 == 0 && HALF_BYTES % 1024 == 0, "SW128 blocks need 1024-byte alignment");
constexpr int N_VIEWS_OP = 80;                  // 64 views-layer outputs + sigma (+ 15 zero rows)
}  // namespace tc2

namespace {

constexpr int THREADS = 544;                    // 8 epilogue warps + 8 producer warps + MMA issuer
constexpr int OFF_SLOT = tc2::HALF_BYTES;       // slot s: PE at +s*32768, MISC at +16384
constexpr int SLOT_BYTES = 32768, OFF_PE = 0, OFF_MISC = 16384;
constexpr int OFF_XCH = OFF_SLOT + 2 * SLOT_BYTES;
constexpr int SMEM_BYTES = OFF_XCH + 2 * 2048 + 1024;
constexpr uint32_t COL_H = 128, COL_MOD = 192, COL_LZ = 128;   // TMEM columns inside a slot's 256

struct Shared {
    uint64_t in_ready[2];       // epilogue warps of BOTH CTAs (8 arrivals) -> issuer: H written / accumulator free
    uint64_t pe_ready[2];       // producer warps of BOTH CTAs (8 arrivals) -> issuer: PE + MISC tiles written
    uint64_t acc_ready[2];      // tcgen05.commit (multicast) -> epilogue warps: accumulator complete
    uint64_t pe_free[2];        // commit after op 5: the PE tile may be overwritten
    uint64_t misc_free[2];      // commit after op 6: the MISC tile may be overwritten
    uint64_t w_full;            // this CTA's weight half has landed
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
__device__ __forceinline__ void tmem_wait8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// two fp32 -> packed fp16x2 (low half = first argument), saturating to the largest finite fp16
template <bool RELU>
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) {
    uint32_t r;
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// relu(a * b) on packed fp16 pairs, clamped to the largest finite fp16 (an fp16 overflow must not turn into NaN in
// the next layer: inf * 0 weight)
__device__ __forceinline__ uint32_t mul_relu_h2(uint32_t a, uint32_t b) {
    uint32_t r;
    const uint32_t zero = 0u, maxh = 0x7BFF7BFFu;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;\n" : "=r"(r) : "r"(a), "r"(b), "r"(zero));
    asm("min.f16x2 %0, %1, %2;\n" : "=r"(r) : "r"(r), "r"(maxh));
    return r;
}

// Generic hidden-layer epilogue of one row: acc[0,128) (fp32) x mod (packed fp16) -> relu -> packed fp16 -> H.
// 8 chunks of 16 accumulator columns; the TMEM loads of chunk i+1 fly while chunk i is converted and stored.
__device__ __forceinline__ void epilogue_hidden(uint32_t t_acc, uint32_t t_mod, uint32_t t_h) {
    uint32_t a[2][16], m[2][8];
    tmem_ld16(t_acc, a[0]);
    tmem_ld8(t_mod, m[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        tmem_wait16(a[i & 1]);
        tmem_wait8(m[i & 1]);
        if (i + 1 < 8) {
            tmem_ld16(t_acc + (i + 1) * 16, a[(i + 1) & 1]);
            tmem_ld8(t_mod + (i + 1) * 8, m[(i + 1) & 1]);
        }
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = mul_relu_h2(cvt_h2<false>(__uint_as_float(a[i & 1][2 * j]), __uint_as_float(a[i & 1][2 * j + 1])), m[i & 1][j]);
        tmem_st8(t_h + i * 8, o);
    }
}

// debug timeline: role r writes (clock << 8 | event) into trace[r * 1024 + i]
#ifdef MVSN_TC_TRACE
#define TC_TRACE(ev) do { if (tr && tr_n < 1023) tr[tr_n++] = (clock64() << 8) | (long long)(ev); } while (0)
#else
#define TC_TRACE(ev) do { } while (0)
#endif

template <bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
render_tc2_kernel(const SceneDev sc, const RenderIO io, const uint8_t* __restrict__ wimg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ Shared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();

    // ---- prologue: cameras, barriers, TMEM, resident weights --------------------------------------
    load_cams(sc, &sh.cams, tid);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sh.in_ready[s], 8);
            mbar_init(&sh.pe_ready[s], 8);
            mbar_init(&sh.acc_ready[s], 1);
            mbar_init(&sh.pe_free[s], 1);
            mbar_init(&sh.misc_free[s], 1);
        }
        mbar_init(&sh.w_full, 1);
        fence_barrier_init();
    }
    if (warp == 16) { tmem_alloc_pair(&sh.tmem_base, 512); tmem_relinquish_pair(); }
    __syncthreads();
    if (tid == 0) {
        // this CTA's half of the weights: 4 bulk copies of 32 000 B, one transaction barrier
        mbar_arrive_expect_tx(&sh.w_full, (uint32_t)tc2::HALF_BYTES);
        const uint8_t* src = wimg + (size_t)rank * tc2::HALF_BYTES;
        for (int i = 0; i < 4; ++i) bulk_load(smem + i * 32000, src + i * 32000, 32000u, &sh.w_full);
    }
    mbar_wait(&sh.w_full, 0);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();             // both CTAs: barriers initialised, TMEM allocated, weights resident
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;
#ifdef MVSN_TC_TRACE
    // trace roles: 0,1 = epilogue warp 0 of slot 0/1 ; 2,3 = producer warp 0 of slot 0/1 ; 4 = MMA issuer
    long long* tr = nullptr; int tr_n = 0;
    if (io.trace && blockIdx.x == 0 && lane == 0) {
        if (warp < 16 && (warp & 3) == 0) tr = io.trace + (warp >> 2) * 1024;
        else if (warp == 16) tr = io.trace + 4 * 1024;
    }
#endif

    // ---- work decomposition (identical in every role and in both CTAs of a pair) ---------------------
    // A tile is RT rays x SP = 128/RT consecutive samples: row = sub * RT + ray_in, so the 32 lanes of a warp are
    // (RT >= 32) adjacent rays at the SAME sample index -- their volume / image taps share cache lines.  A
    // (CTA, slot) context walks the NT tiles of its RT-ray group front to back; contexts are dealt
    //   group = ((gpass * npairs + pair) * 2 + slot) * 2 + rank.
    const int N = io.N, S = io.S;
    const int RT = io.rays_per_tile, SP = 128 / RT;
    const int rt_shift = 31 - __clz(RT);
    const int NT = (S + SP - 1) / SP;
    const int G = (N + RT - 1) / RT;
    const int npairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
    const int n_gpass = (G + npairs * 4 - 1) / (npairs * 4);
    auto group_of = [&](int gp, int s, int r) { return ((gp * npairs + pair) * 2 + s) * 2 + r; };
    auto slot_active = [&](int gp, int s) { return group_of(gp, s, 0) < G; };     // rank 0 holds the lower group

    if (warp < 8) {
        // =========================== epilogue + compositing warps =====================================
        const int s = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
        const uint32_t t_acc = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(s * 256);
        const uint32_t t_h = t_acc + COL_H, t_mod = t_acc + COL_MOD, t_lz = t_acc + COL_LZ;
        const uint32_t leader_in_ready = mapa_u32(smem_u32(&sh.in_ready[s]), 0);
        uint32_t par_acc = 0;
        float cT = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f;   // compositing state of ray `row` (row < RT)
        const float br0 = __ldg(reinterpret_cast<const float*>(wimg + tc2::TAIL_OFFSET));
        const float br1 = __ldg(reinterpret_cast<const float*>(wimg + tc2::TAIL_OFFSET) + 1);
        const float br2 = __ldg(reinterpret_cast<const float*>(wimg + tc2::TAIL_OFFSET) + 2);
        float4* xch = reinterpret_cast<float4*>(smem + OFF_XCH + s * 2048);

        // hand the slot back to the issuer: everything this warp wrote to TMEM is complete and ordered
        auto signal = [&]() {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_in_ready);
        };
        signal();                                           // "accumulator free" for the very first op 0

#pragma unroll 1
        for (int gp = 0; gp < n_gpass; ++gp) {
            if (!slot_active(gp, s)) continue;
            const int g = group_of(gp, s, (int)rank);
#pragma unroll 1
            for (int tile = 0; tile < NT; ++tile) {
                // ---- op 0: modulation GEMM landed in [128,256) as fp32, layer 0 in the accumulator -------
                mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                TC_TRACE(10);
                tc_fence_after();
                {
                    // landing zone [128,256) holds mod[n] as fp32 in column 128 + n; the packed form goes to [192,256).
                    // Upper half first (n = 64..127 lives exactly where the packed tile will be written), then the lower
                    // half chunk by chunk straight into the freed columns.
                    uint32_t mu[32], t[2][16];
                    tmem_ld16(t_lz + 64, t[0]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        tmem_wait16(t[i & 1]);
                        tmem_ld16(i + 1 < 4 ? t_lz + 64 + (i + 1) * 16 : t_lz, t[(i + 1) & 1]);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            mu[i * 8 + j] = cvt_h2<false>(__uint_as_float(t[i & 1][2 * j]), __uint_as_float(t[i & 1][2 * j + 1]));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        tmem_wait16(t[i & 1]);
                        if (i + 1 < 4) tmem_ld16(t_lz + (i + 1) * 16, t[(i + 1) & 1]);
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            o[j] = cvt_h2<false>(__uint_as_float(t[i & 1][2 * j]), __uint_as_float(t[i & 1][2 * j + 1]));
                        tmem_st8(t_mod + i * 8, o);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = mu[i * 8 + j];
                        tmem_st8(t_mod + 32 + i * 8, o);
                    }
                    tmem_st_wait();
                }
                epilogue_hidden(t_acc, t_mod, t_h);
                tmem_st_wait();
                TC_TRACE(30);
                signal();
                // ---- ops 1..5: trunk -------------------------------------------------------------------------
#pragma unroll 1
                for (int op = 1; op < 6; ++op) {
                    mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                    TC_TRACE(10 + op);
                    tc_fence_after();
                    epilogue_hidden(t_acc, t_mod, t_h);
                    tmem_st_wait();
                    TC_TRACE(30 + op);
                    signal();
                }
                // ---- op 6: folded views layer (64 columns, relu) + sigma (column 64) -------------------------
                float sigma;
                {
                    mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                    TC_TRACE(16);
                    tc_fence_after();
                    uint32_t a[2][16];
                    tmem_ld16(t_acc, a[0]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        tmem_wait16(a[i & 1]);
                        tmem_ld16(t_acc + (i + 1) * 16, a[(i + 1) & 1]);     // i == 3 fetches columns 64..79 (sigma)
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            o[j] = cvt_h2<true>(__uint_as_float(a[i & 1][2 * j]), __uint_as_float(a[i & 1][2 * j + 1]));
                        tmem_st8(t_h + i * 8, o);
                    }
                    tmem_wait16(a[0]);
                    sigma = fmaxf(__uint_as_float(a[0][0]), 0.f);
                    tmem_st_wait();
                    TC_TRACE(36);
                    signal();
                }
                // ---- op 7: rgb ---------------------------------------------------------------------------
                float cr, cg, cb;
                {
                    mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
                    TC_TRACE(17);
                    tc_fence_after();
                    uint32_t r16[16];
                    tmem_ld16(t_acc, r16);
                    tmem_wait16(r16);
                    signal();                                  // accumulator free: the next tile's op 0 may be issued
                    cr = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[0]) + br0)));
                    cg = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[1]) + br1)));
                    cb = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[2]) + br2)));
                }
                // ---- compositing (renderer.py:18-26,65-92) --------------------------------------------------
                // every row publishes (alpha, r, g, b); the first RT threads of the slot then walk their ray's SP
                // samples of this tile front to back -- the reference's sequential cumprod order.
                xch[row] = make_float4(1.f - __expf(-sigma), cr, cg, cb);
                named_bar_sync(1 + s, 128);
                if (row < RT) {
                    if (tile == 0) { cT = 1.f; c0 = c1 = c2 = c3 = c4 = 0.f; }
                    const int cray = g * RT + row;
                    if (cray < N) {
                        float znear = 0.f, zfar = 0.f;
                        if (FAST) { const float4 r1 = __ldg(reinterpret_cast<const float4*>(io.rays + (size_t)cray * 8) + 1); znear = r1.z; zfar = r1.w; }
                        for (int sub = 0; sub < SP; ++sub) {
                            const int sj = tile * SP + sub;
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
                        if (tile == NT - 1) {
                            float o0 = c0, o1 = c1, o2 = c2;
                            if (sc.white_bkgd) { const float bg = 1.f - c4; o0 += bg; o1 += bg; o2 += bg; }
                            store_pixel(io, cray, o0, o1, o2, c3);
                        }
                    }
                }
                TC_TRACE(53);
            }
        }
    } else if (warp < 16) {
        // =========================== producer warps: front end of the NEXT tile ===========================
        const int s = (warp - 8) >> 2, row = ((warp - 8) & 3) * 32 + lane;
        uint8_t* slot = smem + OFF_SLOT + s * SLOT_BYTES;
        const uint32_t leader_pe_ready = mapa_u32(smem_u32(&sh.pe_ready[s]), 0);
        uint32_t tcount = 0;                                 // tiles produced so far by this slot
#pragma unroll 1
        for (int gp = 0; gp < n_gpass; ++gp) {
            if (!slot_active(gp, s)) continue;
            const int g = group_of(gp, s, (int)rank);
#pragma unroll 1
            for (int tile = 0; tile < NT; ++tile, ++tcount) {
                TC_TRACE(1);
                const int r_in = row & (RT - 1), s_idx = tile * SP + (row >> rt_shift);
                const int ray = g * RT + r_in;
                const bool valid = g < G && ray < N && s_idx < S;
                const size_t si = (size_t)ray * S + s_idx;
                float nx = 0.f, ny = 0.f, nz = 0.f;
                float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
                if (valid) {
                    if (FAST) {
                        const float4* rp = reinterpret_cast<const float4*>(io.rays + (size_t)ray * 8);
                        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                        dx = r0.w; dy = r1.x; dz = r1.y;
                        const float near = r1.z, far = r1.w, t = __ldg(io.t_steps + s_idx);
                        float zv;
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
                        dx = __ldg(io.dirs + (size_t)ray * 3); dy = __ldg(io.dirs + (size_t)ray * 3 + 1);
                        dz = __ldg(io.dirs + (size_t)ray * 3 + 2);
                    }
                }
                // ---- gathers: volume (8) + colour (12) features + view direction -> 6 packed chunks of MISC ----
                uint4 mc[3];
                uint32_t md0, md1;
                {
                    float feat[20], dir[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 20; ++i) feat[i] = 0.f;
                    if (valid) {
                        view_dir<false>(sh.cams, dx, dy, dz, dir);
                        sample_volume_2pass(sc, nx, ny, nz, feat);
#pragma unroll
                        for (int v = 0; v < 3; ++v) sample_color<false>(sc, sh.cams, v, px, py, pz, feat + 8 + 4 * v);
                        if (io.input_feat) {
                            float4* o = reinterpret_cast<float4*>(io.input_feat + si * 20);
#pragma unroll
                            for (int i = 0; i < 5; ++i)
                                o[i] = make_float4(feat[4 * i], feat[4 * i + 1], feat[4 * i + 2], feat[4 * i + 3]);
                        }
                    }
                    mc[0] = make_uint4(pack_h2(feat[0], feat[1]), pack_h2(feat[2], feat[3]), pack_h2(feat[4], feat[5]), pack_h2(feat[6], feat[7]));
                    mc[1] = make_uint4(pack_h2(feat[8], feat[9]), pack_h2(feat[10], feat[11]), pack_h2(feat[12], feat[13]), pack_h2(feat[14], feat[15]));
                    mc[2] = make_uint4(pack_h2(feat[16], feat[17]), pack_h2(feat[18], feat[19]), pack_h2(1.f, 0.f), 0u);
                    md0 = pack_h2(dir[0], dir[1]); md1 = pack_h2(dir[2], 1.f);
                }
                TC_TRACE(4);
                // ---- positional encoding: [x y z | sin(2^k x) k-major | cos(2^k x) k-major | 1] -> 8 packed chunks ----
                uint4 pc[8];
                {
                    const float nd[3] = {nx, ny, nz};
                    // element i of the 64-wide encoding row (compile-time index under full unrolling); computed chunk by
                    // chunk so that only 8 fp32 values are live next to the packed result
                    auto pe = [&](int i) -> float {
                        if (i < 3) return nd[i];
                        if (i < 33) return __sinf(nd[(i - 3) % 3] * (float)(1 << ((i - 3) / 3)));
                        if (i < 63) return __cosf(nd[(i - 33) % 3] * (float)(1 << ((i - 33) / 3)));
                        return 1.f;
                    };
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        pc[c] = make_uint4(pack_h2(pe(c * 8), pe(c * 8 + 1)), pack_h2(pe(c * 8 + 2), pe(c * 8 + 3)),
                                           pack_h2(pe(c * 8 + 4), pe(c * 8 + 5)), pack_h2(pe(c * 8 + 6), pe(c * 8 + 7)));
                }
                TC_TRACE(5);
                // ---- hand-over: the previous tile's last readers of PE (op 5) and MISC (op 6) must have retired ----
                if (tcount > 0) mbar_wait(&sh.pe_free[s], (tcount - 1) & 1);
#pragma unroll
                for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(slot + OFF_PE + sw128_offset(row, c * 8)) = pc[c];
                if (tcount > 0) mbar_wait(&sh.misc_free[s], (tcount - 1) & 1);
                uint8_t* m = slot + OFF_MISC;
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 0)) = mc[0];
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 8)) = mc[1];
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 16)) = mc[2];
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 24)) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 32)) = make_uint4(md0, md1, 0u, 0u);
                *reinterpret_cast<uint4*>(m + sw128_offset(row, 40)) = make_uint4(0u, 0u, 0u, 0u);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_pe_ready);
                TC_TRACE(2);
            }
        }
    } else if (rank == 0) {
        // =========================== MMA issuer (leader CTA) ================================================
        // the whole warp walks the schedule (waits are warp-uniform); one elected lane issues
        const bool leader = elect_one();
        const uint32_t sbase = smem_u32(smem);
        constexpr uint32_t HI_SW = (uint32_t)(desc_sw128(0) >> 32), HI_NS = (uint32_t)(desc_nosw(0, 128, 256) >> 32);
        constexpr uint32_t LO_SW = (uint32_t)desc_sw128(0), LO_NS = (uint32_t)desc_nosw(0, 128, 256);
        auto dsw = [&](uint32_t addr) { return ((uint64_t)HI_SW << 32) | (uint64_t)(LO_SW | (addr >> 4)); };
        auto dns = [&](uint32_t addr) { return ((uint64_t)HI_NS << 32) | (uint64_t)(LO_NS | (addr >> 4)); };
        constexpr uint32_t ID128 = idesc_f16(256, 128), IDV = idesc_f16(256, tc2::N_VIEWS_OP), IDR = idesc_f16(256, 16);
        uint32_t par_in[2] = {0, 0}, par_pe[2] = {0, 0};
#pragma unroll 1
        for (int gp = 0; gp < n_gpass; ++gp) {
            const bool act[2] = {slot_active(gp, 0), slot_active(gp, 1)};
#pragma unroll 1
            for (int tile = 0; tile < NT; ++tile) {
#pragma unroll 1
                for (int op = 0; op < 8; ++op) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        if (!act[s]) continue;
                        if (op == 0) { mbar_wait_cluster(&sh.pe_ready[s], par_pe[s]); par_pe[s] ^= 1; }
                        TC_TRACE(100 + op * 2 + s);
                        mbar_wait_cluster(&sh.in_ready[s], par_in[s]); par_in[s] ^= 1;
                        TC_TRACE(140 + op * 2 + s);
                        tc_fence_after();
                        if (leader) {
                            const uint32_t d_acc = tmem + (uint32_t)(s * 256);
                            const uint32_t a_h = d_acc + COL_H;
                            const uint32_t pe = sbase + OFF_SLOT + s * SLOT_BYTES + OFF_PE, misc = pe + OFF_MISC;
                            if (op == 0) {
                                // modulation: [feat | 1] (K = 32) x pts_bias -> landing zone ; layer 0: PE (K = 64) -> accumulator
                                const uint64_t dm = dsw(misc), dwm = dsw(sbase + tc2::W_MOD);
                                mma_f16_pair(d_acc + COL_LZ, dm, dwm, ID128, 0u);
                                mma_f16_pair(d_acc + COL_LZ, dm + 2, dwm + 2, ID128, 1u);
                                const uint64_t dp = dsw(pe), dw0 = dsw(sbase + tc2::W_L0);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) mma_f16_pair(d_acc, dp + 2 * ks, dw0 + 2 * ks, ID128, ks ? 1u : 0u);
                            } else if (op <= 4) {
                                const uint32_t wl = sbase + tc2::W_L1 + (op - 1) * tc2::L_STRIDE;
#pragma unroll
                                for (int kb = 0; kb < 2; ++kb) {
                                    const uint64_t dw = dsw(wl + kb * 8192);
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks)
                                        mma_f16_ts_pair(d_acc, a_h + (uint32_t)((kb * 4 + ks) * 8), dw + 2 * ks, ID128, (kb | ks) ? 1u : 0u);
                                }
                                mma_f16_pair(d_acc, dsw(misc) + 2, dns(wl + 16384), ID128, 1u);      // bias: constant-one column of MISC
                            } else if (op == 5) {
                                const uint64_t dp = dsw(pe), dw5 = dsw(sbase + tc2::W_L5);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) mma_f16_pair(d_acc, dp + 2 * ks, dw5 + 2 * ks, ID128, ks ? 1u : 0u);
#pragma unroll
                                for (int kb = 0; kb < 2; ++kb) {
                                    const uint64_t dw = dsw(sbase + tc2::W_L5 + 8192 + kb * 8192);
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks)
                                        mma_f16_ts_pair(d_acc, a_h + (uint32_t)((kb * 4 + ks) * 8), dw + 2 * ks, ID128, 1u);
                                }
                            } else if (op == 6) {
#pragma unroll
                                for (int kb = 0; kb < 2; ++kb) {
                                    const uint64_t dw = dsw(sbase + tc2::W_V + kb * 5120);
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks)
                                        mma_f16_ts_pair(d_acc, a_h + (uint32_t)((kb * 4 + ks) * 8), dw + 2 * ks, IDV, (kb | ks) ? 1u : 0u);
                                }
                                mma_f16_pair(d_acc, dsw(misc) + 4, dns(sbase + tc2::W_V + tc2::W_V_TILE), IDV, 1u);   // [dir | 1] x [Wv dir part | b']
                            } else {
                                const uint64_t dw = dsw(sbase + tc2::W_RGB);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) mma_f16_ts_pair(d_acc, a_h + (uint32_t)(ks * 8), dw + 2 * ks, IDR, ks ? 1u : 0u);
                            }
                            TC_TRACE(80 + op * 2 + s);
                            mma_commit_pair(&sh.acc_ready[s]);
                            if (op == 5) mma_commit_pair(&sh.pe_free[s]);
                            if (op == 6) mma_commit_pair(&sh.misc_free[s]);
                        }
                        __syncwarp();
                    }
                }
                TC_TRACE(200);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();             // nobody frees TMEM / exits while the peer may still signal or be read
    if (warp == 16) tmem_dealloc_pair(tmem, 512);
}

}  // namespace

int launch_render_tc2(const SceneDev& sc, const RenderIO& io_in, bool fast, const void* wimg, cudaStream_t stream) {
    RenderIO io = io_in;
    static bool attr_set[64] = {false};
    int dev = 0;
    MVSN_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        if (dev < 64) attr_set[dev] = true;
    }
    const int max_pairs = sm_count() / 2;
    // rays per tile: 32 (best gather locality) unless the batch is too small to give every (CTA, slot) context a group
    int rt = 32;
    while (rt > 4 && (io.N + rt - 1) / rt < 4 * max_pairs) rt >>= 1;
    io.rays_per_tile = rt;
    const int G = (io.N + rt - 1) / rt;
    int pairs = (G + 3) / 4;
    if (pairs > max_pairs) pairs = max_pairs;
    if (pairs <= 0) return MVSN_OK;
    const uint8_t* w = static_cast<const uint8_t*>(wimg);
    if (fast) render_tc2_kernel<true><<<2 * pairs, THREADS, SMEM_BYTES, stream>>>(sc, io, w);
    else      render_tc2_kernel<false><<<2 * pairs, THREADS, SMEM_BYTES, stream>>>(sc, io, w);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

// ------------------------------------------------------------------------------------------------
// weight image packer (fp32 nn.Linear tensors -> per-CTA halves of fp16 pre-swizzled operand tiles)
// ------------------------------------------------------------------------------------------------
struct MlpPtrsTc2 { const float* p[MVSN_N_MLP_TENSORS]; };

__device__ __forceinline__ uint32_t nosw_offset2(int r, int k) {      // [R x 16] no-swizzle K-major tile
    return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// blockIdx.x = rank * 9 + region ; regions: 0 mod, 1 L0, 2..5 L1..4, 6 L5, 7 folded views + sigma, 8 rgb
__global__ void pack_mlp_tc2_kernel(MlpPtrsTc2 w, uint8_t* __restrict__ out) {
    // tensor indices: 0..11 pts_linears (w,b) x6; 12,13 pts_bias; 14,15 views; 16,17 feature; 18,19 alpha; 20,21 rgb
    const int rank = blockIdx.x / 9, reg = blockIdx.x % 9;
    const int tid = threadIdx.x, nt = blockDim.x;
    int off, bytes;
    if (reg == 0) { off = tc2::W_MOD; bytes = 8192; }
    else if (reg == 1) { off = tc2::W_L0; bytes = 8192; }
    else if (reg <= 5) { off = tc2::W_L1 + (reg - 2) * tc2::L_STRIDE; bytes = tc2::L_STRIDE; }
    else if (reg == 6) { off = tc2::W_L5; bytes = 3 * 8192; }
    else if (reg == 7) { off = tc2::W_V; bytes = 12288; }
    else { off = tc2::W_RGB; bytes = 1024; }
    uint8_t* dst = out + (size_t)rank * tc2::HALF_BYTES + off;
    for (int i = tid * 16; i < bytes; i += nt * 16) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    auto put = [&](uint32_t o, float v) { *reinterpret_cast<__half*>(dst + o) = __float2half_rn(v); };
    if (reg == 0) {
        for (int i = tid; i < 64 * 21; i += nt) { const int lr = i / 21, k = i % 21, n = rank * 64 + lr;
            put(sw128_offset(lr, k), k < 20 ? w.p[12][n * 20 + k] : w.p[13][n]); }
    } else if (reg == 1) {
        for (int i = tid; i < 64 * 64; i += nt) { const int lr = i / 64, k = i % 64, n = rank * 64 + lr;
            put(sw128_offset(lr, k), k < 63 ? w.p[0][n * 63 + k] : w.p[1][n]); }
    } else if (reg <= 5) {
        const int l = reg - 1;                                  // layer 1..4
        for (int i = tid; i < 64 * 128; i += nt) { const int lr = i / 128, k = i % 128, n = rank * 64 + lr;
            put((k >> 6) * 8192 + sw128_offset(lr, k & 63), w.p[2 * l][n * 128 + k]); }
        for (int lr = tid; lr < 64; lr += nt) put(16384 + nosw_offset2(lr, 4), w.p[2 * l + 1][rank * 64 + lr]);
    } else if (reg == 6) {
        for (int i = tid; i < 64 * 64; i += nt) { const int lr = i / 64, k = i % 64, n = rank * 64 + lr;
            put(sw128_offset(lr, k), k < 63 ? w.p[10][n * 191 + k] : w.p[11][n]); }
        for (int i = tid; i < 64 * 128; i += nt) { const int lr = i / 128, k = i % 128, n = rank * 64 + lr;
            put(8192 + (k >> 6) * 8192 + sw128_offset(lr, k & 63), w.p[10][n * 191 + 63 + k]); }
    } else if (reg == 7) {
        // rows n < 64: W'[n][k] = sum_j Wv[n][j] Wf[j][k], b'[n] = sum_j Wv[n][j] bf[j] + bv[n] (fp64) ; row 64: alpha_linear
        for (int i = tid; i < 40 * 128; i += nt) {
            const int lr = i / 128, k = i % 128, n = rank * 40 + lr;
            float v = 0.f;
            if (n < 64) {
                double acc = 0.0;
                for (int j = 0; j < 128; ++j) acc += (double)w.p[14][n * 131 + j] * (double)w.p[16][j * 128 + k];
                v = (float)acc;
            } else if (n == 64) {
                v = w.p[18][k];
            }
            put((k >> 6) * 5120 + sw128_offset(lr, k & 63), v);
        }
        for (int i = tid; i < 40 * 4; i += nt) {
            const int lr = i / 4, k = i % 4, n = rank * 40 + lr;
            float v = 0.f;
            if (n < 64) {
                if (k < 3) v = w.p[14][n * 131 + 128 + k];
                else {
                    double acc = (double)w.p[15][n];
                    for (int j = 0; j < 128; ++j) acc += (double)w.p[14][n * 131 + j] * (double)w.p[17][j];
                    v = (float)acc;
                }
            } else if (n == 64 && k == 3) {
                v = w.p[19][0];
            }
            put(tc2::W_V_TILE + nosw_offset2(lr, k), v);
        }
    } else {
        for (int i = tid; i < 8 * 64; i += nt) { const int lr = i / 64, k = i % 64, n = rank * 8 + lr;
            put(sw128_offset(lr, k), n < 3 ? w.p[20][n * 64 + k] : 0.f); }
        if (rank == 0 && tid < 4) reinterpret_cast<float*>(out + tc2::TAIL_OFFSET)[tid] = tid < 3 ? w.p[21][tid] : 0.f;
    }
}

size_t mlp_tc2_packed_bytes() { return tc2::TOTAL_BYTES; }

int pack_mlp_tc2(const float* const* w, void* packed, cudaStream_t stream) {
    MlpPtrsTc2 p;
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i) p.p[i] = w[i];
    pack_mlp_tc2_kernel<<<18, 256, 0, stream>>>(p, static_cast<uint8_t*>(packed));
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // namespace mvsn

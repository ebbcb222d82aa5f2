// K-B0 on tensor cores: CostRegNet's first layer (conv0: 41 -> 8 channels, 3x3x3, 83 of the net's 111 GFLOP at the
// BASELINE config; models.py:757 with ConvBnReLU3D models.py:674-685) as a tcgen05 GEMM at fp32-grade accuracy.
//
// A 3x3x3 convolution with 8 output channels is a bad implicit GEMM in the usual orientation (M = voxels, N = 8,
// K = 41 * 27): with N = 8 the tensor core idles behind the A-operand traffic.  It is a good one turned inside out:
//
//     Y[p][t * 8 + co] = sum_ci  in[p][ci] * W[co][ci][t]          M = input positions, N = 27 taps x 8 = 216, K = 41
//     out[o][co]       = sum_t   Y[o + off_t][t * 8 + co]            (27 shifted adds per output, done by the epilogue)
//
// so every input position is staged ONCE (K = 41 -> 48 columns of one 128-row operand tile) and multiplied against the
// whole resident weight matrix (224 x 48), N = 224 keeps the MMA at full rate, and the "im2col" never exists.
//
// fp32-grade accuracy (the volume gate is 1e-4 |v|max): 2-term fp16 operand split, three MMAs per K-step
// (hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator), operands carried with exact power-of-two scales (x16 inputs,
// x256 weights) so that the lo terms stay normal fp16 numbers -- the scheme of render_tcs.cu.
//
// One persistent CTA per SM walks bricks of 6 x 10 x 24 output voxels (halo brick 8 x 12 x 32 positions = 24 operand
// tiles of 128 positions; a tile = 4 x-rows of 32 positions, so a warp of the epilogue = one (z, y) row, lane = x):
//   warps 16-19 producers: one TMA tensor-map load per tile (box 32 x 4 y 1 z 41 channels = 21 KB fp32, out-of-volume
//                          positions zero-filled by the hardware = the convolution's padding) into a 2-stage ring; then
//                          each thread turns its position's 41 values into the hi / lo fp16 rows of the operand tile
//   warp  20   issuer    : 9 tcgen05.mma (M 128, N 224, K 16) per tile into one of two TMEM accumulators, commits
//   warps 0-15 epilogue  : four groups of four warps, one per PAIR of output channels (disjoint output cells, own
//                          named barrier); per (dy, dz): the dx sum of its columns by two warp shuffles -> += into the brick's
//                          output tile in shared memory (row owned by the warp within a step; a named barrier between
//                          steps keeps the accumulation order fixed => bit-reproducible); at the end of a brick the
//                          tile is written out (raw, pre-BatchNorm) and the batch statistics accumulated
// Replaces conv0_k3_kernel (FFMA, 1.80 ms at 8x128x176x208); same ConvArgs contract (raw output + fixed-point sums).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "conv_common.cuh"
#include "umma.cuh"

namespace mvsn {

using namespace umma;

namespace c0 {
constexpr int CIN = 41, COUT = 8;
constexpr int BZ = 6, BY = 10, BX = 24;                 // output brick
constexpr int XOFF = 4;                                 // lane of the first output column: the tile's x window starts at x0 - 4,
                                                        // because a TMA box must start 16-byte aligned in the innermost dimension
constexpr int HZ = BZ + 2, HY = BY + 2, HX = 32;        // halo brick (x: one warp = 24 outputs + the two halo columns + 3 + 3 spare)
static_assert(BX % 4 == 0 && XOFF % 4 == 0 && XOFF >= 1 && XOFF + BX + 1 <= HX, "x window");
constexpr int TILES = HZ * HY * HX / 128;               // 24 operand tiles per brick
constexpr int NCOL = 224;                               // 27 taps x 8 channels = 216, padded to a multiple of 16
constexpr float SA = 16.f, SW = 256.f, INV_SCALE = 1.f / (16.f * 256.f);
constexpr int W_PART = NCOL * 128;                      // bytes of the hi (or lo) weight image: [224][64] fp16, SW128
constexpr int A_PART = 128 * 128;                       // bytes of the hi (or lo) operand tile
constexpr int OFF_W = 0;                                // hi | lo
constexpr int OFF_A = 2 * W_PART;                       // 2 buffers x (hi | lo)
constexpr int OFF_OUT = OFF_A + 4 * A_PART;             // out_s [8][BZ][BY][32] fp32
constexpr int OUT_FLOATS = COUT * BZ * BY * 32;
constexpr int STAGE_BYTES = CIN * 128 * 4;              // one TMA box: [41 c][4 y][32 x] fp32 = 20 992 B
constexpr int STAGE_STRIDE = 21504;                     // 128-byte aligned slot
constexpr int OFF_STAGE = OFF_OUT + OUT_FLOATS * 4;     // 2 slots
constexpr int SMEM_BYTES = OFF_STAGE + 2 * STAGE_STRIDE + 1024;
static_assert(OFF_STAGE % 128 == 0 && STAGE_STRIDE % 128 == 0 && STAGE_STRIDE >= STAGE_BYTES, "TMA destinations are 128-byte aligned");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int THREADS = 672;                            // 16 epilogue warps (four channel pairs) + 4 producers + issuer
constexpr int PRODUCER_WARP0 = 16, ISSUER_WARP = 20;
static_assert(OFF_A % 1024 == 0 && W_PART % 1024 == 0, "SW128 tiles need 1024-byte alignment");
static_assert(HY % 4 == 0, "a tile is four y-rows of the halo brick");
}  // namespace c0

long long* debug_trace_buffer();

namespace {

// trace build: per-role cycle totals of CTA 0 -> trace[role * 8 + slot]  (tools/conv0_profile.py)
#ifdef MVSN_TC_TRACE
#define C0_T0() const long long _t0 = clock64()
#define C0_ACC(role, slot) do { if (prof && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 128 || threadIdx.x == 512 || threadIdx.x == 640)) prof[(role) * 8 + (slot)] += clock64() - _t0; } while (0)
#else
#define C0_T0() do { } while (0)
#define C0_ACC(role, slot) do { } while (0)
#endif

struct SharedC0 {
    uint64_t a_full[2];         // producers (4 warps) -> issuer
    uint64_t a_free[2];         // tcgen05.commit -> producers: the operand tile has been consumed
    uint64_t acc_ready[2];      // tcgen05.commit -> epilogue
    uint64_t acc_free[2];       // epilogue (4 warps) -> issuer
    uint64_t w_full;
    uint64_t stage_full[2];     // TMA complete_tx -> producers
    uint64_t stage_free[2];     // producers (4 warps) -> TMA thread
    uint32_t tmem_base;
};

// one tile of the cost volume: box {32 x, 4 y, 1 z, 41 c} at (x, y, z, 0); coordinates may lie outside the volume (zero fill)
__device__ __forceinline__ void tma_load_tile(void* smem_dst, const CUtensorMap* tmap, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(z), "r"(0), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_ld2_c0(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(taddr));
}
// wait for the outstanding tcgen05.ld of this warp; tying the registers to empty asm statements keeps their uses below it
__device__ __forceinline__ void tmem_wait54(uint32_t (&r)[54]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 54; i += 6)
        asm volatile("" : "+r"(r[i]), "+r"(r[i + 1]), "+r"(r[i + 2]), "+r"(r[i + 3]), "+r"(r[i + 4]), "+r"(r[i + 5]) :: "memory");
}
__device__ __forceinline__ void split2_c0(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Epilogue of one PAIR of output channels (G = 0..3: channels 2G, 2G + 1), four warps = the four (z, y) rows of a tile,
// lane = x.  216 accumulator columns = 3 dy x 3 dz x 3 dx x 8 co; this group needs 2 columns of each of the 27 taps: all
// 27 tcgen05.ld.x2 are issued at once (54 registers) and waited for ONCE -- a TMEM round trip is ~450 cycles, paying it
// per dy-group was what bounded the kernel.  Inside a dy-group the three dz steps hit three different z planes and the
// four warps four different y rows: nothing collides; consecutive dy-groups overlap by rows, hence ONE named barrier (of
// this channel pair) per dy-group keeps the accumulation order fixed, which makes the sums bit-reproducible.
template <int G>
__device__ __forceinline__ void epilogue_role(const ConvArgs& a, SharedC0& sh, float* out_s, uint32_t tmem, int w, int lane,
                                              int nbricks, int nbx, int nby, long long* prof) {
    using namespace c0;
    const int D = a.Din, H = a.Hin, W = a.Win;
    const uint32_t t_lane = tmem + ((uint32_t)(w * 32) << 16);
    float st_s = 0.f, st_q = 0.f;                               // batch statistics of channel 2G + (w >> 1), half the rows
    uint32_t it = 0;
#pragma unroll 1
    for (int b = blockIdx.x; b < nbricks; b += gridDim.x) {
        const int x0 = (b % nbx) * BX, y0 = ((b / nbx) % nby) * BY, z0 = (b / (nbx * nby)) * BZ;
#pragma unroll 1
        for (int m = 0; m < TILES; ++m, ++it) {
            const uint32_t buf = it & 1u;
            { C0_T0(); mbar_wait(&sh.acc_ready[buf], (it >> 1) & 1u); C0_ACC(G & 1, 0); }
            C0_T0();
            tc_fence_after();
            const int z = m / (HY / 4), y = (m % (HY / 4)) * 4 + w;           // this warp's row of the halo brick
            const uint32_t t_acc = t_lane + buf * 256u + G * 2;
            uint32_t yv[54];                                               // [(dy * 9 + dz * 3 + dx) * 2 + c]
#pragma unroll
            for (int q = 0; q < 27; ++q) tmem_ld2_c0(t_acc + q * 8, yv[q * 2], yv[q * 2 + 1]);
            tmem_wait54(yv);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.acc_free[buf]);                  // the accumulator is in registers: release it early
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int yi = y - dy;
                float acc[3][2];
                float* cell[3];
                bool ok[3];
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {                            // all loads of the three steps first
                    const int zi = z - dz;
                    ok[dz] = (unsigned)zi < (unsigned)BZ && (unsigned)yi < (unsigned)BY;
                    cell[dz] = out_s + ((G * 2 * BZ + zi) * BY + yi) * 32 + lane;      // + c * BZ * BY * 32
#pragma unroll
                    for (int c = 0; c < 2; ++c) acc[dz][c] = ok[dz] ? cell[dz][c * (BZ * BY * 32)] : 0.f;
                }
#pragma unroll
                for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int q = (dy * 9 + dz * 3) * 2 + c;
                        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(yv[q]), 1);
                        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(yv[q + 4]), 1);
                        acc[dz][c] += __uint_as_float(yv[q + 2]) + left + right;          // still x4096: undone at write-out
                    }
#pragma unroll
                for (int dz = 0; dz < 3; ++dz)
                    if (ok[dz]) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) cell[dz][c * (BZ * BY * 32)] = acc[dz][c];
                    }
                bar_sync_named(1 + G, 128);
            }
            C0_ACC(G & 1, 1);
        }
        C0_T0();
        // ---- brick complete: this warp writes half the rows of channel 2G + (w >> 1) (raw outputs), accumulates their
        // statistics, clears them
        const size_t plane = (size_t)H * W, vol = plane * D;
        const int co = G * 2 + (w >> 1), gx = x0 + lane - XOFF;
        const bool xok = lane >= XOFF && lane < XOFF + BX && gx < W;
        constexpr int HALF_ROWS = BZ * BY / 2;
        float* rows = out_s + ((size_t)co * BZ * BY + (w & 1) * HALF_ROWS) * 32 + lane;
        float* gout = a.out + (size_t)co * vol + gx;
#pragma unroll 1
        for (int j0 = 0; j0 < HALF_ROWS; j0 += 5) {
            float v[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) v[u] = rows[(j0 + u) * 32] * INV_SCALE;    // exact: a power of two
#pragma unroll
            for (int u = 0; u < 5; ++u) rows[(j0 + u) * 32] = 0.f;
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int j = (w & 1) * HALF_ROWS + j0 + u, zi = j / BY, yi = j - zi * BY, gz = z0 + zi, gy = y0 + yi;
                if (xok && gz < D && gy < H) {
                    gout[(size_t)gz * plane + (size_t)gy * W] = v[u];
                    st_s += v[u]; st_q = fmaf(v[u], v[u], st_q);
                }
            }
        }
        bar_sync_named(1 + G, 128);              // every row cleared before the next brick accumulates
        C0_ACC(G & 1, 2);
    }
    // batch statistics: fixed order per thread, fixed-order warp reduction, integer (fixed-point) atomics
    float s = st_s, q = st_q;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if (lane == 0) {
        unsigned long long* st = reinterpret_cast<unsigned long long*>(a.stats_out) + 2 * (G * 2 + (w >> 1));
        atomicAdd(st, stat_fx(s));
        atomicAdd(st + 1, stat_fx(q));
    }
}

__global__ void __launch_bounds__(c0::THREADS, 1)
conv0_tc_kernel(const ConvArgs a, const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ wimg, long long* prof) {
    using namespace c0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ SharedC0 sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.Din, H = a.Hin, W = a.Win;
    const int nbx = (W + BX - 1) / BX, nby = (H + BY - 1) / BY, nbz = (D + BZ - 1) / BZ;
    const int nbricks = nbx * nby * nbz;
    float* out_s = reinterpret_cast<float*>(smem + OFF_OUT);

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sh.a_full[b], 4); mbar_init(&sh.a_free[b], 1);
            mbar_init(&sh.acc_ready[b], 1); mbar_init(&sh.acc_free[b], 16);
        }
        mbar_init(&sh.w_full, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&sh.stage_full[b], 1); mbar_init(&sh.stage_free[b], 4); }
        fence_barrier_init();
    }
    if (warp == ISSUER_WARP) { tmem_alloc(&sh.tmem_base, 512); tmem_relinquish(); }
    for (int i = tid; i < OUT_FLOATS; i += THREADS) out_s[i] = 0.f;
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&sh.w_full, 2u * W_PART);
        bulk_load(smem + OFF_W, wimg, W_PART, &sh.w_full);
        bulk_load(smem + OFF_W + W_PART, wimg + W_PART, W_PART, &sh.w_full);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    auto brick_origin = [&](int b, int& z0, int& y0, int& x0) {
        x0 = (b % nbx) * BX; y0 = ((b / nbx) % nby) * BY; z0 = (b / (nbx * nby)) * BZ;
    };

    if (warp < 16) {
        // =========================== epilogue: shifted tap sums -> output tile =========================
        if (warp < 4)       epilogue_role<0>(a, sh, out_s, tmem, warp, lane, nbricks, nbx, nby, prof);
        else if (warp < 8)  epilogue_role<1>(a, sh, out_s, tmem, warp - 4, lane, nbricks, nbx, nby, prof);
        else if (warp < 12) epilogue_role<2>(a, sh, out_s, tmem, warp - 8, lane, nbricks, nbx, nby, prof);
        else                epilogue_role<3>(a, sh, out_s, tmem, warp - 12, lane, nbricks, nbx, nby, prof);
    } else if (warp < ISSUER_WARP) {
        // =========================== producers: TMA-staged tile -> hi / lo operand rows ====================
        const int row = (warp - PRODUCER_WARP0) * 32 + lane;
        // the first producer warp issues the TMA loads: the whole warp waits for the ring slot (warp-uniform), one elected lane issues
        const bool tma_warp = warp == PRODUCER_WARP0;
        const bool tma_leader = tma_warp ? elect_one() : false;
        auto issue_tma = [&](int b, int m, uint32_t it) {                 // tile (b, m) = the it-th tile of this CTA
            const uint32_t sb = it & 1u;
            mbar_wait(&sh.stage_free[sb], ((it >> 1) & 1u) ^ 1u);
            if (tma_leader) {
                int z0, y0, x0;
                brick_origin(b, z0, y0, x0);
                mbar_arrive_expect_tx(&sh.stage_full[sb], (uint32_t)STAGE_BYTES);
                tma_load_tile(smem + OFF_STAGE + sb * STAGE_STRIDE, &tmap, x0 - XOFF, y0 - 1 + (m % (HY / 4)) * 4, z0 - 1 + m / (HY / 4),
                              &sh.stage_full[sb]);
            }
            __syncwarp();
        };
        uint32_t it = 0;
        int b = blockIdx.x, m = 0;
        if (tma_warp && b < nbricks) issue_tma(b, m, 0);
#pragma unroll 1
        while (b < nbricks) {
            int b2 = b, m2 = m + 1;
            if (m2 == TILES) { m2 = 0; b2 += gridDim.x; }
            if (tma_warp && b2 < nbricks) issue_tma(b2, m2, it + 1);        // the next tile streams in while this one is converted
            const uint32_t sb = it & 1u, buf = it & 1u;
            { C0_T0(); mbar_wait(&sh.stage_full[sb], (it >> 1) & 1u); C0_ACC(2, 2); }
            const float* st = reinterpret_cast<const float*>(smem + OFF_STAGE + sb * STAGE_STRIDE) + row;
            { C0_T0(); mbar_wait(&sh.a_free[buf], ((it >> 1) & 1u) ^ 1u); C0_ACC(2, 0); }
            C0_T0();
            uint8_t* hi = smem + OFF_A + buf * (2 * A_PART);
#pragma unroll
            for (int k2 = 0; k2 < 3; ++k2) {                               // 16 channels per round: 16 independent loads in flight
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = k2 * 16 + j < CIN ? st[(k2 * 16 + j) * 128] : 0.f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split2_c0(v[half * 8 + 2 * j] * SA, v[half * 8 + 2 * j + 1] * SA, h[j], l[j]);
                    const uint32_t off = sw128_offset(row, (k2 * 2 + half) * 8);
                    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(hi + A_PART + off) = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.stage_free[sb]);                 // the staged tile has been consumed
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.a_full[buf]);
            C0_ACC(2, 1);
            ++it; b = b2; m = m2;
        }
    } else {
        // =========================== MMA issuer ===========================================================
        mbar_wait(&sh.w_full, 0);
        const bool leader = elect_one();
        const uint32_t sbase = smem_u32(smem);
        const uint64_t dw_hi = desc_sw128(sbase + OFF_W), dw_lo = desc_sw128(sbase + OFF_W + W_PART);
        constexpr uint32_t IDESC = idesc_f16(128, NCOL);
        uint32_t it = 0;
#pragma unroll 1
        for (int b = blockIdx.x; b < nbricks; b += gridDim.x) {
#pragma unroll 1
            for (int m = 0; m < TILES; ++m, ++it) {
                const uint32_t buf = it & 1u;
                { C0_T0(); mbar_wait(&sh.a_full[buf], (it >> 1) & 1u); C0_ACC(3, 0); }
                { C0_T0(); mbar_wait(&sh.acc_free[buf], ((it >> 1) & 1u) ^ 1u); C0_ACC(3, 1); }
                C0_T0();
                tc_fence_after();
                if (leader) {
                    const uint64_t da_hi = desc_sw128(sbase + OFF_A + buf * (2 * A_PART));
                    const uint64_t da_lo = desc_sw128(sbase + OFF_A + buf * (2 * A_PART) + A_PART);
                    const uint32_t d = tmem + buf * 256u;
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) mma_f16(d, da_hi + 2 * ks, dw_hi + 2 * ks, IDESC, ks ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) mma_f16(d, da_hi + 2 * ks, dw_lo + 2 * ks, IDESC, 1u);
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) mma_f16(d, da_lo + 2 * ks, dw_hi + 2 * ks, IDESC, 1u);
                    mma_commit(&sh.acc_ready[buf]);
                    mma_commit(&sh.a_free[buf]);
                }
                __syncwarp();
                C0_ACC(3, 2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ISSUER_WARP) tmem_dealloc(tmem, 512);
}

// weights [8][41][3][3][3] -> B operand [224 n][64 k = ci] fp16, SWIZZLE_128B, x256, as (hi | lo)
__global__ void pack_conv0_tc_kernel(const float* __restrict__ w, uint8_t* __restrict__ out) {
    using namespace c0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NCOL * 64; i += gridDim.x * blockDim.x) {
        // column n = ((dy * 3 + dz) * 3 + dx) * 8 + co: the three dz taps of one dy are 72 contiguous accumulator columns
        const int n = i >> 6, k = i & 63, g = n >> 3, co = n & 7;
        const int dy = g / 9, dz = (g / 3) % 3, dx = g % 3, t = dz * 9 + dy * 3 + dx;
        const float v = (n < 27 * COUT && k < CIN) ? w[(size_t)co * CIN * 27 + (size_t)k * 27 + t] * SW : 0.f;
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const uint32_t off = sw128_offset(n, k);
        *reinterpret_cast<__half*>(out + off) = h;
        *reinterpret_cast<__half*>(out + W_PART + off) = l;
    }
}

}  // namespace

size_t conv0_tc_workspace_bytes() { return 2 * (size_t)c0::W_PART; }

int launch_conv0_tc(const ConvArgs& a, void* wimg, cudaStream_t st) {
    using namespace c0;
    MVSN_REQUIRE(a.Cin == CIN && a.Cout == COUT, MVSN_EBADSHAPE, "conv0_tc: built for 41 -> 8 channels");
    MVSN_REQUIRE(aligned16(wimg), MVSN_EALIGN, "conv0_tc: weight image must be 16-byte aligned");
    static bool attr_set[64] = {false};
    int dev = 0;
    MVSN_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        if (dev < 64) attr_set[dev] = true;
    }
    pack_conv0_tc_kernel<<<14, 256, 0, st>>>(a.w, static_cast<uint8_t*>(wimg));
    MVSN_CUDA_CHECK(cudaGetLastError());
    // tensor map of the input [41][D][H][W] fp32 (innermost first); box = one operand tile
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MVSN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        MVSN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, MVSN_ECUDA, "conv0_tc: cuTensorMapEncodeTiled is not available");
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    MVSN_REQUIRE(a.Win % 4 == 0 && aligned16(a.in0.x), MVSN_EALIGN, "conv0_tc: the input needs 16-byte aligned rows (W %% 4 == 0)");
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {(cuuint64_t)a.Win, (cuuint64_t)a.Hin, (cuuint64_t)a.Din, (cuuint64_t)CIN};
    const cuuint64_t gstride[3] = {(cuuint64_t)a.Win * 4, (cuuint64_t)a.Win * a.Hin * 4, (cuuint64_t)a.Win * a.Hin * a.Din * 4};
    const cuuint32_t box[4] = {32, 4, 1, (cuuint32_t)CIN}, estr[4] = {1, 1, 1, 1};
    const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.in0.x), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVSN_REQUIRE(cr == CUDA_SUCCESS, MVSN_ECUDA, "conv0_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    const int nbricks = ((a.Win + BX - 1) / BX) * ((a.Hin + BY - 1) / BY) * ((a.Din + BZ - 1) / BZ);
    const int grid = nbricks < sm_count() ? nbricks : sm_count();
    conv0_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(a, tmap, static_cast<const uint8_t*>(wimg), debug_trace_buffer());
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // namespace mvsn

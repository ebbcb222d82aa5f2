// WORK IN PROGRESS (round 2) -- NOT part of libmvsnerf_b200.so, NOT validated on hardware.
// Built only by `python -m mvsnerf_b200.build --wip` into libmvsnerf_b200_wip.so, where it answers to
// mlp_mode = 3 (lib.MLP_TC_PAIR_WIP).  Written at the end of round 1, after the GPU budget was spent, from the
// arithmetic in DESIGN.md section 7; only its building block (the CTA-pair MMA with a split B operand,
// tc_pair_probe.cu) has run on a B200.  Expect bring-up bugs; validate against the oracle (5e-3 gate) and
// against MVSN_MLP_TC_HALF before anything else.
//
// K-C, tensor-core mode with THREE tiles in flight per SM on CTA pairs (`cta_group::2`):
//   * cluster of 2 CTAs (one per SM of a TPC), 448 threads each:
//       warps 0-11  three slot groups x 4 warps; thread = one sample row of the slot's 128-sample tile: front end
//                   (ray march, NDC, trilinear + colour gathers, positional encoding -> fp16 operand tiles), all
//                   epilogues (128 accumulator columns), compositing
//       warp 12     leader CTA: MMA issuer (ONE M=256 tcgen05.mma.cta_group::2 per K-step covers slot s of both
//                   CTAs); peer CTA: relays "my half of the weight chunk has landed" to the leader
//       warp 13     weight loader: each CTA streams ITS HALF of every weight chunk (B operand split across the pair)
//   * TMEM per CTA: 3 x 128 accumulator columns + 48 columns for the three sigma results.  The multiplicative
//     modulation does not live in TMEM any more: it is the first GEMM phase of a tile and is kept by each thread
//     as 64 packed fp16x2 registers, applied with cvt.rn.f16x2 + fma.rn.relu.f16x2.
//   * shared memory per CTA: 3 x 56 KB operand tiles (PE 16 K | H 32 K | MISC 8 K no-swizzle) + 4 x 12 KB ring of
//     half chunks + 6 KB compositing exchange = 222 KB.
// Ten GEMM phases per tile: modulation, layer 0, layers 1-4, layer 5 [pe|h], feature (+ sigma as an N=16 MMA),
// views layer, rgb.  Slots run the phases in lock step so a weight chunk serves all six tiles of the pair.
#include "../render_frontend.cuh"
#include "../umma.cuh"

namespace mvsn {
using namespace umma;

namespace tpw {     // weight image: for chunk c, rank r: HB(c) contiguous bytes at OFF(c) + r * HB(c)
constexpr int NCHUNK = 18, NOP = 10;
__host__ __device__ constexpr int half_bytes(int c) {
    return c <= 1 ? 8192 : c <= 9 ? ((c & 1) ? 10240 : 8192) : c <= 12 ? 8192 : c == 13 ? 9216 : c == 14 ? 11520
         : c == 15 ? 4096 : c == 16 ? 5120 : 1024;
}
__host__ __device__ constexpr int chunk_offset(int c) {
    int o = 0;
    for (int i = 0; i < c; ++i) o += 2 * half_bytes(i);
    return o;
}
constexpr int STREAM_BYTES = chunk_offset(NCHUNK);          // 291 328 (the same data as the single-CTA image)
constexpr int TAIL_OFFSET = STREAM_BYTES;                   // fp32 tail: rgb_linear.bias[3], 0
constexpr int TOTAL_BYTES = STREAM_BYTES + 16;
constexpr int STAGE_BYTES = 12288;                          // >= largest half chunk (11 520), 1024-aligned
constexpr int NSTAGE = 4;
}  // namespace tpw

namespace {

constexpr int TP_THREADS = 448;                             // 12 slot warps + issuer/relay + loader
constexpr int NSLOT = 3;
constexpr int SLOT_BYTES = 57344;                           // PE 16K | H0 16K | H1 16K | MISC 8K
constexpr int OFF_PE = 0, OFF_H0 = 16384, OFF_H1 = 32768, OFF_MISC = 49152;
constexpr int RING_OFFSET = NSLOT * SLOT_BYTES;
constexpr int XCH_OFFSET = RING_OFFSET + tpw::NSTAGE * tpw::STAGE_BYTES;
constexpr int TP_SMEM_BYTES = XCH_OFFSET + NSLOT * 2048 + 1024;
constexpr int TMEM_COLS = 512, T_SIGMA = 384;

struct TpShared {
    uint64_t in_ready[NSLOT];   // LEADER's copy is the live one: 256 arrivals (128 rows of slot s in each CTA)
    uint64_t acc_ready[NSLOT];  // multicast tcgen05.commit -> both CTAs
    uint64_t w_full[tpw::NSTAGE];   // this CTA's half chunk landed
    uint64_t w_peer[tpw::NSTAGE];   // LEADER only: the peer's half landed (relayed by the peer's warp 12)
    uint64_t w_free[4];             // multicast commit: GEMM op k (global index) -> w_free[k & 3]
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
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
template <bool RELU>
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) {
    uint32_t r;
    if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    else      asm("cvt.rn.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t hmul2_relu(uint32_t a, uint32_t b) {      // relu(a * b) per fp16 lane
    uint32_t r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;\n" : "=r"(r) : "r"(a), "r"(b), "r"(0u));
    return r;
}

// ---- cluster plumbing -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of THIS CTA's layout) inside CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// wait on a LOCAL mbarrier whose arrivals come from the peer CTA as well: acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t tries = 0;; ++tries) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2, 0xF4240;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t"
            "}\n" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (tries > 4000u) __trap();
    }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit_pair(uint64_t* bar) {      // arrives on `bar` (same offset) in BOTH CTAs
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// no-swizzle K-major tiles: 8-row x 16-byte core matrices, LBO (K direction) 128 B
__device__ __forceinline__ uint32_t misc_offset(int r, int k) {      // [128 x 32] A tile, SBO 512
    return (uint32_t)((r >> 3) * 512 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
__host__ __device__ constexpr uint32_t nosw16_offset(int r, int k) {  // [R x 16] B tile, SBO 256
    return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// ---- epilogues: this thread's row, all 128 accumulator columns ----------------------------------------
// op 0: modulation GEMM result -> 64 packed fp16x2 registers
__device__ __forceinline__ void epilogue_mod(uint32_t t_acc, uint32_t (&mh)[64]) {
    uint32_t a[2][16];
    tmem_ld16(t_acc, a[0]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        tmem_wait16(a[i & 1]);
        if (i + 1 < 8) tmem_ld16(t_acc + (i + 1) * 16, a[(i + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            mh[i * 8 + j] = cvt_h2<false>(__uint_as_float(a[i & 1][2 * j]), __uint_as_float(a[i & 1][2 * j + 1]));
    }
}
// NCOL accumulator columns -> [x mod] -> [relu] -> fp16 -> K-blocks H0 (cols 0-63) / H1 (cols 64-127) of the slot
template <int NCOL, bool MODULATE, bool RELU>
__device__ __forceinline__ void epilogue_h(uint32_t t_acc, const uint32_t (&mh)[64], uint8_t* h0, int row) {
    const uint32_t ra = smem_u32(h0) + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (row & 7) * 16);
    uint32_t a[2][16];
    tmem_ld16(t_acc, a[0]);
#pragma unroll
    for (int i = 0; i < NCOL / 16; ++i) {
        tmem_wait16(a[i & 1]);
        if (i + 1 < NCOL / 16) tmem_ld16(t_acc + (i + 1) * 16, a[(i + 1) & 1]);
        uint32_t p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x = __uint_as_float(a[i & 1][2 * j]), y = __uint_as_float(a[i & 1][2 * j + 1]);
            if (MODULATE) p[j] = hmul2_relu(cvt_h2<false>(x, y), mh[i * 8 + j]);
            else p[j] = cvt_h2<RELU>(x, y);
        }
        const uint32_t base = ra + (uint32_t)((i >> 2) * 16384);
        const int kc = (i & 3) * 2;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(base ^ (uint32_t)(kc * 16)), "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(base ^ (uint32_t)((kc + 1) * 16)), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]) : "memory");
    }
}

// ---- GEMM schedule of one tile ---------------------------------------------------------------------------
// op 0 modulation (MISC K-steps 0,1 x chunk 0) ; op 1 layer 0 (PE x chunk 1) ; ops 2-5 layers 1-4 ; op 6 layer 5
// [PE | H] ; op 7 feature (N=128) + sigma (N=16 -> TMEM cols 384 + 16 s) ; op 8 views (N=64) ; op 9 rgb (N=16)
__constant__ int c_nblk[10] = {1, 1, 2, 2, 2, 2, 3, 2, 2, 1};
__constant__ int c_chunk0[10] = {0, 1, 2, 4, 6, 8, 10, 13, 15, 17};          // first chunk of the op
__constant__ int c_opN[10] = {128, 128, 128, 128, 128, 128, 128, 128, 64, 16};
__constant__ uint32_t c_aoff[10][3] = {{OFF_MISC, 0, 0}, {OFF_PE, 0, 0}, {OFF_H0, OFF_H1, 0}, {OFF_H0, OFF_H1, 0}, {OFF_H0, OFF_H1, 0},
                                       {OFF_H0, OFF_H1, 0}, {OFF_PE, OFF_H0, OFF_H1}, {OFF_H0, OFF_H1, 0}, {OFF_H0, OFF_H1, 0}, {OFF_H0, 0, 0}};
// bias step: A = MISC K-step 1 (cols 16..31: feat16-19, 1, dir0-2), B = no-swizzle [N/2 x 16] tile inside the LAST chunk of the op
__constant__ int c_bias_boff[10] = {-1, -1, 8192, 8192, 8192, 8192, -1, 9216, 4096, -1};
__constant__ int c_chunk_op[18] = {0, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 6, 7, 7, 8, 8, 9};

}  // namespace

template <bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TP_THREADS, 1)
render_tc_pair_kernel(const SceneDev sc, const RenderIO io, const uint8_t* __restrict__ wimg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ TpShared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();

    load_cams(sc, &sh.cams, tid);
    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&sh.in_ready[s], 256); mbar_init(&sh.acc_ready[s], 1); }
        for (int i = 0; i < tpw::NSTAGE; ++i) { mbar_init(&sh.w_full[i], 1); mbar_init(&sh.w_peer[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&sh.w_free[i], 1);
        fence_barrier_init();
    }
    if (warp == 12) { tmem_alloc_pair(&sh.tmem_base, TMEM_COLS); tmem_relinquish_pair(); }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // both CTAs: barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    // ---- work decomposition (identical in every role and in both CTAs of the pair) ---------------------
    const int N = io.N, S = io.S;
    constexpr int RT = 32, SP = 4;                           // tile = 32 adjacent rays x 4 consecutive samples
    const int NT = (S + SP - 1) / SP;                        // tiles (= passes) per ray group
    const int G = (N + RT - 1) / RT;                         // ray groups
    const int P = (int)gridDim.x >> 1, p = (int)blockIdx.x >> 1;
    const int G6 = (G + 5) / 6;                              // a pair takes 6 groups per round (2 CTAs x 3 slots)
    const int rounds = p < G6 ? (G6 - p + P - 1) / P : 0;
    const int npass = rounds * NT;
    auto group_of = [&](int pass, int r, int s) { return (((pass / NT) * P + p) * 2 + r) * NSLOT + s; };

    if (warp < 12) {
        // =========================== slot group ==============================================================
        const int s = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
        uint8_t* slot = smem + s * SLOT_BYTES;
        const uint32_t t_acc = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(s * 128);
        const uint32_t t_sig = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(T_SIGMA + s * 16);
        const uint32_t in_ready_leader = map_to_cta(&sh.in_ready[s], 0);
        uint32_t par_acc = 0;
        float cT = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f;   // compositing state of ray `row` (row < RT)
        const float br0 = __ldg(reinterpret_cast<const float*>(wimg + tpw::TAIL_OFFSET));
        const float br1 = __ldg(reinterpret_cast<const float*>(wimg + tpw::TAIL_OFFSET) + 1);
        const float br2 = __ldg(reinterpret_cast<const float*>(wimg + tpw::TAIL_OFFSET) + 2);
        uint32_t mh[64];                                     // this row's modulation, packed fp16x2
#pragma unroll
        for (int i = 0; i < 64; ++i) mh[i] = 0u;

        auto handoff = [&]() {                               // operand tile written / accumulator drained -> issuer
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive_cluster(in_ready_leader);
        };
        auto wait_acc = [&]() {
            mbar_wait(&sh.acc_ready[s], par_acc); par_acc ^= 1;
            tc_fence_after();
        };

        float sigma = 0.f;
        int g_cur = 0, tile_cur = 0;
        bool have_cur = false;
        for (int pass = -1; pass < npass; ++pass) {
            if (have_cur) {
                wait_acc(); epilogue_mod(t_acc, mh); handoff();                                   // op 0
#pragma unroll 1
                for (int op = 1; op <= 6; ++op) {                                                 // layers 0..5
                    wait_acc(); epilogue_h<128, true, false>(t_acc, mh, slot + OFF_H0, row); handoff();
                }
                {                                                                                 // op 7: feature + sigma
                    wait_acc();
                    epilogue_h<128, false, false>(t_acc, mh, slot + OFF_H0, row);
                    uint32_t r16[16];
                    tmem_ld16(t_sig, r16);
                    tmem_wait16(r16);
                    sigma = fmaxf(__uint_as_float(r16[0]), 0.f);
                    handoff();
                }
                { wait_acc(); epilogue_h<64, false, true>(t_acc, mh, slot + OFF_H0, row); handoff(); }   // op 8: views
            }
            const int np = pass + 1;
            const bool have_next = np < npass;
            const int g = have_next ? group_of(np, (int)rank, s) : G, tile = have_next ? np % NT : 0;
            if (have_next) {
                // -------------------------- front end of the next tile (PE / MISC are free after op 8) ----------
                const int r_in = row & (RT - 1), s_idx = tile * SP + (row >> 5);
                const int ray = g * RT + r_in;
                const bool valid = g < G && ray < N && s_idx < S;
                const size_t si = (size_t)ray * S + s_idx;
                float nx = 0.f, ny = 0.f, nz = 0.f;
                float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
                if (valid) {
                    if (FAST) {
                        const float4* rp = reinterpret_cast<const float4*>(io.rays + (size_t)ray * 8);
                        float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
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
                // gathers: volume (8) + colour (12) features, view direction -> MISC cols 0..23 (col 20 = 1)
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
                        for (int i = 0; i < 5; ++i) o[i] = make_float4(feat[4 * i], feat[4 * i + 1], feat[4 * i + 2], feat[4 * i + 3]);
                    }
                }
                uint8_t* m = slot + OFF_MISC;
                *reinterpret_cast<uint4*>(m + misc_offset(row, 0)) =
                    make_uint4(pack_h2(feat[0], feat[1]), pack_h2(feat[2], feat[3]), pack_h2(feat[4], feat[5]), pack_h2(feat[6], feat[7]));
                *reinterpret_cast<uint4*>(m + misc_offset(row, 8)) =
                    make_uint4(pack_h2(feat[8], feat[9]), pack_h2(feat[10], feat[11]), pack_h2(feat[12], feat[13]), pack_h2(feat[14], feat[15]));
                *reinterpret_cast<uint4*>(m + misc_offset(row, 16)) =
                    make_uint4(pack_h2(feat[16], feat[17]), pack_h2(feat[18], feat[19]), pack_h2(1.f, dir[0]), pack_h2(dir[1], dir[2]));
                *reinterpret_cast<uint4*>(m + misc_offset(row, 24)) = make_uint4(0u, 0u, 0u, 0u);
                // positional encoding: [x y z | sin(2^k x) k-major | cos(2^k x) k-major | 1]  (models.py:47-51, F6)
                const float nd[3] = {nx, ny, nz};
                float v[64];
                v[0] = nx; v[1] = ny; v[2] = nz;
                float f = 1.f;
#pragma unroll
                for (int k = 0; k < 10; ++k) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        v[3 + 3 * k + j] = __sinf(nd[j] * f);
                        v[33 + 3 * k + j] = __cosf(nd[j] * f);
                    }
                    f *= 2.f;
                }
                v[63] = 1.f;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(slot + OFF_PE + sw128_offset(row, c * 8)) =
                        make_uint4(pack_h2(v[c * 8], v[c * 8 + 1]), pack_h2(v[c * 8 + 2], v[c * 8 + 3]),
                                   pack_h2(v[c * 8 + 4], v[c * 8 + 5]), pack_h2(v[c * 8 + 6], v[c * 8 + 7]));
            }
            if (have_cur) {
                // -------------------------- op 9: rgb, then compositing (renderer.py:18-26,65-92) ----------------
                wait_acc();
                uint32_t r16[16];
                tmem_ld16(t_acc, r16);
                tmem_wait16(r16);
                tc_fence_before();
                const float cr = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[0]) + br0)));
                const float cg = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[1]) + br1)));
                const float cb = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r16[2]) + br2)));
                float4* xch = reinterpret_cast<float4*>(smem + XCH_OFFSET + s * 2048);
                xch[row] = make_float4(1.f - __expf(-sigma), cr, cg, cb);
                named_bar_sync(1 + s, 128);
                if (row < RT) {
                    if (tile_cur == 0) { cT = 1.f; c0 = c1 = c2 = c3 = c4 = 0.f; }
                    const int cray = g_cur * RT + row;
                    if (g_cur < G && cray < N) {
                        float znear = 0.f, zfar = 0.f;
                        if (FAST) { const float4 r1 = __ldg(reinterpret_cast<const float4*>(io.rays + (size_t)cray * 8) + 1); znear = r1.z; zfar = r1.w; }
                        for (int sub = 0; sub < SP; ++sub) {
                            const int sj = tile_cur * SP + sub;
                            if (sj >= S) break;
                            const float4 q = xch[sub * RT + row];
                            float z;
                            if (FAST) {
                                const float t = __ldg(io.t_steps + sj);
                                if (!io.rg.lindisp) z = __fadd_rn(__fmul_rn(znear, 1.f - t), __fmul_rn(zfar, t));
                                else z = __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, znear), 1.f - t), __fmul_rn(__fdiv_rn(1.f, zfar), t)));
                            } else {
                                z = __ldg(io.z + (size_t)cray * S + sj);
                            }
                            const float wgt = q.x * cT;
                            if (io.alpha) io.alpha[(size_t)cray * S + sj] = q.x;
                            if (io.weights) io.weights[(size_t)cray * S + sj] = wgt;
                            c0 = fmaf(wgt, q.y, c0); c1 = fmaf(wgt, q.z, c1); c2 = fmaf(wgt, q.w, c2);
                            c3 = fmaf(wgt, z, c3); c4 += wgt;
                            cT *= (1.f - q.x) + 1e-10f;
                        }
                        if (tile_cur == NT - 1) {
                            float o0 = c0, o1 = c1, o2 = c2;
                            if (sc.white_bkgd) { const float bg = 1.f - c4; o0 += bg; o1 += bg; o2 += bg; }
                            io.rgb[(size_t)cray * 3 + 0] = o0; io.rgb[(size_t)cray * 3 + 1] = o1; io.rgb[(size_t)cray * 3 + 2] = o2;
                            io.depth[cray] = c3;
                        }
                    }
                }
                named_bar_sync(1 + s, 128);                  // xch is rewritten by the next tile's op 9 only, but keep the
                                                             // compositing threads and the publishers in step
            }
            if (have_next) { fence_proxy_async(); mbar_arrive_cluster(in_ready_leader); }   // next tile's operands are in place
            g_cur = g; tile_cur = tile; have_cur = have_next;
        }
    } else if (warp == 12) {
        if (rank == 0) {
            // =========================== MMA issuer (leader CTA) ================================================
            const bool leader = elect_one();
            uint32_t par_in[NSLOT] = {0, 0, 0};
            const uint32_t sbase = smem_u32(smem), ring = sbase + RING_OFFSET;
            constexpr uint32_t HI_SW = (uint32_t)(desc_sw128(0) >> 32), LO_SW = (uint32_t)desc_sw128(0);
            constexpr uint32_t HI_NA = (uint32_t)(desc_nosw(0, 128, 512) >> 32), LO_NA = (uint32_t)desc_nosw(0, 128, 512);
            constexpr uint32_t HI_NB = (uint32_t)(desc_nosw(0, 128, 256) >> 32), LO_NB = (uint32_t)desc_nosw(0, 128, 256);
            auto dsw = [&](uint32_t addr) { return ((uint64_t)HI_SW << 32) | (uint64_t)(LO_SW | (addr >> 4)); };
            auto dna = [&](uint32_t addr) { return ((uint64_t)HI_NA << 32) | (uint64_t)(LO_NA | (addr >> 4)); };   // MISC (A)
            auto dnb = [&](uint32_t addr) { return ((uint64_t)HI_NB << 32) | (uint64_t)(LO_NB | (addr >> 4)); };   // bias tiles (B)
            uint32_t nchunk_base = 0, op_base = 0;
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass, op_base += tpw::NOP, nchunk_base += tpw::NCHUNK) {
#pragma unroll 1
                for (int op = 0; op < tpw::NOP; ++op) {
                    const int nblk = c_nblk[op];
                    const uint32_t idesc = idesc_f16(256, c_opN[op]);
                    // weights of this op: my half and the peer's half of every chunk
#pragma unroll 1
                    for (int b = 0; b < nblk; ++b) {
                        const uint32_t n = nchunk_base + c_chunk0[op] + b;
                        mbar_wait(&sh.w_full[n % tpw::NSTAGE], (n / tpw::NSTAGE) & 1);
                        mbar_wait_cluster(&sh.w_peer[n % tpw::NSTAGE], (n / tpw::NSTAGE) & 1);
                    }
#pragma unroll 1
                    for (int s = 0; s < NSLOT; ++s) {
                        mbar_wait_cluster(&sh.in_ready[s], par_in[s]); par_in[s] ^= 1;
                        tc_fence_after();
                        const uint32_t sl = sbase + s * SLOT_BYTES;
                        const uint32_t d_acc = tmem + s * 128;
                        if (leader) {
                            uint32_t last_stage = 0;
#pragma unroll 1
                            for (int b = 0; b < nblk; ++b) {
                                const uint32_t st = ring + ((nchunk_base + c_chunk0[op] + b) % tpw::NSTAGE) * tpw::STAGE_BYTES;
                                last_stage = st;
                                const uint64_t db = dsw(st);
                                if (op == 0) {                               // A = MISC (no swizzle): K-steps 0, 1
                                    const uint64_t da = dna(sl + OFF_MISC);
                                    mma_pair(d_acc, da, db, idesc, 0);
                                    mma_pair(d_acc, da + 16, db + 2, idesc, 1);              // +256 B = 16 units
                                } else {
                                    const uint64_t da = dsw(sl + c_aoff[op][b]);
                                    mma_pair(d_acc, da, db, idesc, b > 0 ? 1u : 0u);
                                    mma_pair(d_acc, da + 2, db + 2, idesc, 1);
                                    mma_pair(d_acc, da + 4, db + 4, idesc, 1);
                                    mma_pair(d_acc, da + 6, db + 6, idesc, 1);
                                }
                            }
                            const uint64_t da_bias = dna(sl + OFF_MISC) + 16;                // MISC K-step 1
                            if (c_bias_boff[op] >= 0) mma_pair(d_acc, da_bias, dnb(last_stage + c_bias_boff[op]), idesc, 1);
                            if (op == 7) {                                   // sigma: N = 16, rows 128.. of the feature chunks
                                const uint32_t isig = idesc_f16(256, 16);
                                const uint32_t d_sig = tmem + T_SIGMA + s * 16;
#pragma unroll 1
                                for (int b = 0; b < 2; ++b) {
                                    const uint32_t st = ring + ((nchunk_base + 13 + b) % tpw::NSTAGE) * tpw::STAGE_BYTES;
                                    const uint64_t da = dsw(sl + (b ? OFF_H1 : OFF_H0)), db = dsw(st + 8192);
                                    mma_pair(d_sig, da, db, isig, b > 0 ? 1u : 0u);
                                    mma_pair(d_sig, da + 2, db + 2, isig, 1);
                                    mma_pair(d_sig, da + 4, db + 4, isig, 1);
                                    mma_pair(d_sig, da + 6, db + 6, isig, 1);
                                }
                                mma_pair(d_sig, da_bias, dnb(last_stage + 11264), isig, 1);
                            }
                            commit_pair(&sh.acc_ready[s]);
                            if (s == NSLOT - 1) commit_pair(&sh.w_free[(op_base + op) & 3]);   // this op's chunks may be overwritten
                        }
                        __syncwarp();
                    }
                }
            }
        } else {
            // =========================== peer CTA: relay "my half chunk landed" to the leader ================
            if (elect_one()) {
                const uint32_t total = (uint32_t)npass * tpw::NCHUNK;
                for (uint32_t n = 0; n < total; ++n) {
                    mbar_wait(&sh.w_full[n % tpw::NSTAGE], (n / tpw::NSTAGE) & 1);
                    mbar_arrive_cluster(map_to_cta(&sh.w_peer[n % tpw::NSTAGE], 0));
                }
            }
        }
    } else {
        // =========================== weight loader (both CTAs: own half of every chunk) ======================
        if (elect_one()) {
            uint8_t* ring = smem + RING_OFFSET;
            uint32_t n = 0;
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
#pragma unroll 1
                for (int c = 0; c < tpw::NCHUNK; ++c, ++n) {
                    const uint32_t st = n % tpw::NSTAGE;
                    if (n >= (uint32_t)tpw::NSTAGE) {
                        const uint32_t pn = n - tpw::NSTAGE;          // the stage's previous tenant
                        const uint32_t need = (pn / tpw::NCHUNK) * tpw::NOP + (uint32_t)c_chunk_op[pn % tpw::NCHUNK];
                        mbar_wait(&sh.w_free[need & 3], (need >> 2) & 1);
                    }
                    const uint32_t bytes = (uint32_t)tpw::half_bytes(c);
                    mbar_arrive_expect_tx(&sh.w_full[st], bytes);
                    bulk_load(ring + st * tpw::STAGE_BYTES, wimg + tpw::chunk_offset(c) + rank * bytes, bytes, &sh.w_full[st]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // the peer may still be reading operands / TMEM of this pass
    if (warp == 12) tmem_dealloc_pair(tmem, TMEM_COLS);
}

int launch_render_tc_pair(const SceneDev& sc, const RenderIO& io_in, bool fast, const void* wimg, cudaStream_t stream) {
    RenderIO io = io_in;
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES));
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_tc_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES));
    io.rays_per_tile = 32;
    const int G = (io.N + 31) / 32, G6 = (G + 5) / 6;
    int pairs = sm_count() / 2;
    if (G6 < pairs) pairs = G6;
    if (pairs <= 0) return MVSN_OK;
    const uint8_t* w = static_cast<const uint8_t*>(wimg);
    if (fast) render_tc_pair_kernel<true><<<2 * pairs, TP_THREADS, TP_SMEM_BYTES, stream>>>(sc, io, w);
    else      render_tc_pair_kernel<false><<<2 * pairs, TP_THREADS, TP_SMEM_BYTES, stream>>>(sc, io, w);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

// ------------------------------------------------------------------------------------------------
// weight image packer: the single-CTA image's chunks, each split by output row between the two CTAs
// ------------------------------------------------------------------------------------------------
namespace {
struct MlpPtrsTp { const float* p[MVSN_N_MLP_TENSORS]; };

__global__ void pack_mlp_tc_pair_kernel(MlpPtrsTp w, uint8_t* __restrict__ out) {
    // tensor indices: 0..11 pts_linears (w,b) x6; 12,13 pts_bias; 14,15 views; 16,17 feature; 18,19 alpha; 20,21 rgb
    const int c = blockIdx.x;
    uint8_t* base = out + tpw::chunk_offset(c);
    const int HB = tpw::half_bytes(c);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid * 16; i < 2 * HB; i += nt * 16) *reinterpret_cast<uint4*>(base + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // element (output row n of an N-row operand, k) of a SWIZZLE_128B block that starts `boff` bytes into each half
    auto put_sw = [&](int n, int N, int k, int boff, float v) {
        const int r = n / (N / 2), ln = n - r * (N / 2);
        *reinterpret_cast<__half*>(base + r * HB + boff + sw128_offset(ln, k)) = __float2half_rn(v);
    };
    auto put_ns = [&](int n, int N, int k, int boff, float v) {       // no-swizzle [N/2 x 16] tile
        const int r = n / (N / 2), ln = n - r * (N / 2);
        *reinterpret_cast<__half*>(base + r * HB + boff + nosw16_offset(ln, k)) = __float2half_rn(v);
    };
    if (c == 0) {                                             // modulation: K = 32 used (cols 0..19 feat, 20 = bias)
        for (int i = tid; i < 128 * 21; i += nt) { const int n = i / 21, k = i % 21;
            put_sw(n, 128, k, 0, k < 20 ? w.p[12][n * 20 + k] : w.p[13][n]); }
    } else if (c == 1) {                                      // layer 0: 63 PE columns + bias at column 63
        for (int i = tid; i < 128 * 64; i += nt) { const int n = i / 64, k = i % 64;
            put_sw(n, 128, k, 0, k < 63 ? w.p[0][n * 63 + k] : w.p[1][n]); }
    } else if (c >= 2 && c <= 9) {                            // layers 1..4, two K-blocks each; bias tile with the second
        const int l = c / 2, kb = c & 1;
        for (int i = tid; i < 128 * 64; i += nt) { const int n = i / 64, k = i % 64;
            put_sw(n, 128, k, 0, w.p[2 * l][n * 128 + kb * 64 + k]); }
        if (kb == 1) for (int n = tid; n < 128; n += nt) put_ns(n, 128, 4, 8192, w.p[2 * l + 1][n]);
    } else if (c == 10) {                                     // layer 5, PE part (+ bias at column 63)
        for (int i = tid; i < 128 * 64; i += nt) { const int n = i / 64, k = i % 64;
            put_sw(n, 128, k, 0, k < 63 ? w.p[10][n * 191 + k] : w.p[11][n]); }
    } else if (c == 11 || c == 12) {                          // layer 5, h part
        const int kb = c - 11;
        for (int i = tid; i < 128 * 64; i += nt) { const int n = i / 64, k = i % 64;
            put_sw(n, 128, k, 0, w.p[10][n * 191 + 63 + kb * 64 + k]); }
    } else if (c == 13 || c == 14) {                          // feature (N = 128) + sigma (row 0 of an N = 16 operand)
        const int kb = c - 13;
        for (int i = tid; i < 128 * 64; i += nt) { const int n = i / 64, k = i % 64;
            put_sw(n, 128, k, 0, w.p[16][n * 128 + kb * 64 + k]); }
        for (int k = tid; k < 64; k += nt) put_sw(0, 16, k, 8192, w.p[18][kb * 64 + k]);
        if (kb == 1) {
            for (int n = tid; n < 128; n += nt) put_ns(n, 128, 4, 9216, w.p[17][n]);
            if (tid == 0) put_ns(0, 16, 4, 11264, w.p[19][0]);
        }
    } else if (c == 15 || c == 16) {                          // views layer (N = 64): f part; dirs + bias ride in the bias step
        const int kb = c - 15;
        for (int i = tid; i < 64 * 64; i += nt) { const int n = i / 64, k = i % 64;
            put_sw(n, 64, k, 0, w.p[14][n * 131 + kb * 64 + k]); }
        if (kb == 1) for (int i = tid; i < 64 * 4; i += nt) { const int n = i / 4, j = i % 4;     // MISC cols 20 (=1), 21..23 (dir)
            put_ns(n, 64, 4 + j, 4096, j == 0 ? w.p[15][n] : w.p[14][n * 131 + 128 + (j - 1)]); }
    } else if (c == 17) {                                     // rgb (N = 16, rows 0..2)
        for (int i = tid; i < 3 * 64; i += nt) { const int n = i / 64, k = i % 64; put_sw(n, 16, k, 0, w.p[20][n * 64 + k]); }
        if (tid < 4) reinterpret_cast<float*>(out + tpw::TAIL_OFFSET)[tid] = tid < 3 ? w.p[21][tid] : 0.f;
    }
}
}  // namespace

size_t mlp_tc_pair_packed_bytes() { return tpw::TOTAL_BYTES; }

int pack_mlp_tc_pair(const float* const* w, void* packed, cudaStream_t stream) {
    MlpPtrsTp p;
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i) p.p[i] = w[i];
    pack_mlp_tc_pair_kernel<<<tpw::NCHUNK, 256, 0, stream>>>(p, static_cast<uint8_t*>(packed));
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // namespace mvsn

// K-C backward: the per-scene fine-tuning step of the fused render path as ONE kernel.
//
// Reference: train_mvs_nerf_finetuning_pl.py:140-189 -- `rendering(...)` under autograd, img2mse loss, Adam over the
// 22 MLP tensors (models.py:145-222) and RefVolume.feat_volume (models.py:935-950).  What autograd computes there
// with ~200 library launches per step is done here by
//
//   render_bwd_kernel      per tile of 128 samples: recompute the forward (fp32 FFMA, same passes as render_fp32.cu),
//                          reverse compositing scan, MLP dgrad + wgrad as tiled GEMMs, trilinear scatter-add of the
//                          8 volume-feature gradients into the channels-last volume gradient;
//   mlp_grad_reduce_kernel per-CTA private weight-gradient accumulators -> the 22 tensors in nn.Linear layout;
//   adam_*_kernel          fused Adam (torch.optim.Adam arithmetic) on the MLP tensors and on the volume.
//
// Data flow of a tile (one persistent CTA of 256 threads per SM, 7 tiles per CTA at 1024 rays x 128 samples):
//   * forward recompute keeps what the backward needs in a per-CTA scratch in global memory (704 KB, L2 resident):
//     the layer outputs TRANSPOSED (hT[n][r]) -- that is exactly the A operand of the wgrad GEMM
//     dW^T[k][n] = sum_r x^T[k][r] dpre[r][n], and the element-wise stage reads the same fragment shape from it;
//   * the pre-activation is not stored: h = relu(pre * mod) > 0  =>  pre = h / mod, and the sample does not
//     contribute where h == 0;
//   * every GEMM uses the forward kernel's tuned pattern: A row-major in shared memory, B streamed global -> shared
//     by cp.async.  dgrad:  A = dpre [r][n] (shared), B = W[n][k] (the nn.Linear layout itself);
//                   wgrad:  A = x^T (scratch -> shared), B = dpre [r][n] (scratch).  No shared-memory transposes.
//   * weight gradients accumulate in a per-CTA private buffer (plain read-modify-write, no atomics, deterministic);
//     the only atomics are the volume scatter (red.global.add.v4.f32, 16 per sample).
// Gradient inputs: d rgb (required, or a target image for the fused MSE loss), d depth, d weights, d alpha,
// d input_feat (optional) -- everything `rendering` returns is differentiable as in the reference.
#include "render_frontend.cuh"
#include "mlp_fp32.cuh"

namespace mvsn {

namespace bwd {
// per-CTA scratch (floats)
constexpr int S_PET   = 0;                        // [64][128]  peT[k][r]        (row 63 zero)
constexpr int S_FEATT = S_PET + 64 * 128;         // [64][128]  featT[k][r]      (rows >= 20 zero)
constexpr int S_MODT  = S_FEATT + 64 * 128;       // [128][128] modT[n][r]
constexpr int S_HT    = S_MODT + 128 * 128;       // 6 x [128][128] hT[l][n][r] = h_{l+1}
constexpr int S_FT    = S_HT + 6 * 128 * 128;     // [128][128] fT[k][r]  (feature_linear output)
constexpr int S_DPRE  = S_FT + 128 * 128;         // [128][128] row-major: B operand of the wgrad GEMMs
constexpr int S_DMOD  = S_DPRE + 128 * 128;       // [128][128] d modT[n][r] accumulator
constexpr int SCRATCH = S_DMOD + 128 * 128;
// per-CTA private gradient accumulators (floats); "T" = [k][n]
constexpr int G_W0T   = 0;                        // [64][128]
constexpr int G_W14T  = G_W0T + 64 * 128;         // 4 x [128][128]
constexpr int G_W5PET = G_W14T + 4 * 128 * 128;   // [64][128]
constexpr int G_W5HT  = G_W5PET + 64 * 128;       // [128][128]
constexpr int G_WBT   = G_W5HT + 128 * 128;       // [64][128]   (k < 20 used)
constexpr int G_WFT   = G_WBT + 64 * 128;         // [128][128]
constexpr int G_WVFT  = G_WFT + 128 * 128;        // [128 k][64 j]
constexpr int G_B     = G_WVFT + 128 * 64;        // b0..b5: 6 x [128]
constexpr int G_BB    = G_B + 6 * 128;
constexpr int G_BF    = G_BB + 128;
constexpr int G_WA    = G_BF + 128;
constexpr int G_BV    = G_WA + 128;               // [64]
constexpr int G_WVDT  = G_BV + 64;                // [4][64] (3 used)
constexpr int G_WR    = G_WVDT + 4 * 64;          // [4][64] (3 used)
constexpr int G_BR    = G_WR + 4 * 64;            // [4]: br[0..2], ba
constexpr int GRADS   = G_BR + 4;
// dgrad weight image (floats): B operands [n_out][k_in] of the dgrad GEMMs
constexpr int D_VF = 0;                           // views_linears.0.weight[:, :128]   [64][128]
constexpr int D_F  = D_VF + 64 * 128;             // feature_linear.weight             [128][128]
constexpr int D_5H = D_F + 128 * 128;             // pts_linears.5.weight[:, 63:]      [128][128]
constexpr int D_14 = D_5H + 128 * 128;            // pts_linears.1..4.weight           4 x [128][128]
constexpr int D_B  = D_14 + 4 * 128 * 128;        // pts_bias.weight                   [128][64] (k < 20)
constexpr int DGRAD = D_B + 128 * 64;
}  // namespace bwd

struct BwdIO {
    const float* g_rgb;      // [N,3]   (or null with `target`)
    const float* target;     // [N,3]   fused loss: g_rgb = 2 (rgb - target) / (3 N_total)
    float inv_count;         // 1 / (3 N_total)
    const float* g_depth;    // [N]     optional
    const float* g_weights;  // [N,S]   optional
    const float* g_alpha;    // [N,S]   optional
    const float* g_feat;     // [N,S,20] optional (only the 8 volume channels carry a gradient)
    float* dvol;             // [D,Hp,Wp,8] channels-last, accumulated atomically (null: the volume is frozen)
    float* scratch;          // n_ctas x bwd::SCRATCH
    float* grads;            // n_ctas x bwd::GRADS
    const float* wd;         // dgrad weight image
    float* rgb_out;          // [N,3] optional: the forward result of the recompute
    float* depth_out;        // [N]   optional
    float* loss;             // [1]   optional: += sum (rgb - target)^2 * inv_count
};

namespace {

// fragment <-> memory helpers.  Fragment of thread (ty, tx): rows {4ty+i, 64+4ty+i}, cols {4tx+j, 64+4tx+j}.
__device__ __forceinline__ int frag_row(int ty, int r) { return (r < 4 ? 0 : 64) + ty * 4 + (r & 3); }
__device__ __forceinline__ int frag_col(int tx, int n) { return (n < 4 ? 0 : 64) + tx * 4 + (n & 3); }

// row-major [128][ld] store (shared or global)
__device__ __forceinline__ void frag_store_rm(const float (&acc)[8][8], float* dst, int ld, int tid) {
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float* p = dst + frag_row(ty, r) * ld + tx * 4;
        *reinterpret_cast<float4*>(p) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4*>(p + 64) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
}
// transposed store: dstT[col][row] with 128-float rows (global scratch)
__device__ __forceinline__ void frag_store_T(const float (&acc)[8][8], float* dstT, int tid) {
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        float* p = dstT + frag_col(tx, n) * 128 + ty * 4;
        *reinterpret_cast<float4*>(p) = make_float4(acc[0][n], acc[1][n], acc[2][n], acc[3][n]);
        *reinterpret_cast<float4*>(p + 64) = make_float4(acc[4][n], acc[5][n], acc[6][n], acc[7][n]);
    }
}
__device__ __forceinline__ void frag_load_T(float (&v)[8][8], const float* srcT, int tid) {
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const float* p = srcT + frag_col(tx, n) * 128 + ty * 4;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 64);
        v[0][n] = a.x; v[1][n] = a.y; v[2][n] = a.z; v[3][n] = a.w;
        v[4][n] = b.x; v[5][n] = b.y; v[6][n] = b.z; v[7][n] = b.w;
    }
}
// private accumulator += fragment, row-major [rows][ld] in global memory
template <int MR, int NT>
__device__ __forceinline__ void frag_accumulate(const float (&acc)[MR][NT], float* dst, int ld, int tid) {
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int r = 0; r < MR; ++r) {
        float* p = dst + frag_row(ty, r) * ld + tx * 4;
        float4 a = *reinterpret_cast<float4*>(p);
        a.x += acc[r][0]; a.y += acc[r][1]; a.z += acc[r][2]; a.w += acc[r][3];
        *reinterpret_cast<float4*>(p) = a;
        if constexpr (NT == 8) {
            float4 b = *reinterpret_cast<float4*>(p + 64);
            b.x += acc[r][4]; b.y += acc[r][5]; b.z += acc[r][6]; b.w += acc[r][7];
            *reinterpret_cast<float4*>(p + 64) = b;
        }
    }
}
// bias add (MODE 0) or relu((acc + bias) * mod) (MODE 1) in place; mod read from shared [128][H_LD]
template <int MODE>
__device__ __forceinline__ void epilogue128(float (&acc)[8][8], const float* __restrict__ bias, const float* s_mod, int tid) {
    const int ty = tid >> 4, tx = tid & 15;
    const float4 bl = __ldg(reinterpret_cast<const float4*>(bias + tx * 4));
    const float4 bh = __ldg(reinterpret_cast<const float4*>(bias + 64 + tx * 4));
    const float b[8] = {bl.x, bl.y, bl.z, bl.w, bh.x, bh.y, bh.z, bh.w};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = frag_row(ty, r);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float v = acc[r][n] + b[n];
            if (MODE == 1) v = fmaxf(v * s_mod[row * H_LD + frag_col(tx, n)], 0.f);
            acc[r][n] = v;
        }
    }
}
// [rows][128] dense global -> shared [rows][H_LD]
__device__ __forceinline__ void stage_T(float* s_dst, const float* g_src, int rows, int tid) {
    for (int i = tid; i < rows * 32; i += 256) {
        const int r = i >> 5, c4 = i & 31;
        cp_async16(s_dst + r * H_LD + c4 * 4, g_src + r * 128 + c4 * 4);
    }
    cp_async_commit();
}
// column sums of a row-major shared tile [128][ld], columns [0, ncols): thread c accumulates into dst[c]
__device__ __forceinline__ void colsum_accumulate(const float* s_src, int ld, int ncols, float* dst, int tid) {
    if (tid < ncols) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < TILE_M; ++r) s += s_src[r * ld + tid];
        dst[tid] += s;
    }
}

constexpr int BWD_SMEM_FLOATS = TILE_M * PE_LD + 2 * TILE_M * H_LD + 2 * KCHUNK * 128 + TILE_M * 28;
constexpr size_t BWD_SMEM_BYTES = BWD_SMEM_FLOATS * sizeof(float);
static_assert(TILE_M * PE_LD >= 64 * H_LD, "peT staging [64][H_LD] must fit in the positional-encoding region");

__global__ void __launch_bounds__(256, 1)
render_bwd_kernel(const SceneDev sc, const RenderIO io, const BwdIO bw, const float* __restrict__ wts) {
    extern __shared__ __align__(16) float smem[];
    float* s_pe   = smem;                          // forward: [128][PE_LD] ; backward: peT staged as [64][H_LD]
    float* s_h    = s_pe + TILE_M * PE_LD;         // [128][H_LD]  forward activations ; backward: A operand (row-major)
    float* s_mod  = s_h + TILE_M * H_LD;           // [128][H_LD]  modulation / hv ; backward: A operand (x^T)
    float* s_w    = s_mod + TILE_M * H_LD;         // 2 x [32][128] streamed B chunks
    float* s_misc = s_w + 2 * KCHUNK * 128;
    float* s_dir  = s_misc;                        // [128][4] view direction
    float* s_z    = s_dir + TILE_M * 4;            // [128]
    float* s_sig  = s_z + TILE_M;                  // [128] sigma = relu(alpha_linear)
    float* s_rgb  = s_sig + TILE_M;                // [128][4] r, g, b, alpha
    float* s_T    = s_rgb + TILE_M * 4;            // [128] transmittance in front of the sample
    float* s_g    = s_T + TILE_M;                  // [128][4] d rgb_pre (3), d sigma_pre
    float* s_df   = s_g + TILE_M * 4;              // [128][8] d volume features ; [128][4] spare
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    __shared__ Cams cams;
    load_cams(sc, &cams, tid);

    float* scr = bw.scratch + (size_t)blockIdx.x * bwd::SCRATCH;
    float* G = bw.grads + (size_t)blockIdx.x * bwd::GRADS;
    for (int i = tid; i < bwd::GRADS; i += 256) G[i] = 0.f;
    for (int i = tid; i < 32 * 128; i += 256) scr[bwd::S_FEATT + 32 * 128 + i] = 0.f;      // featT rows 32..63 stay zero
    __syncthreads();

    const int N = io.N, S = io.S;
    const int R = TILE_M / S;                                   // rays per tile (S <= 128, checked by the launcher)
    const int ngroups = (N + R - 1) / R;

    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        // =============================== forward recompute ===========================================
        int r_in = 0, s_idx = 0;
        bool valid = false;
        size_t si = 0;
        if (tid < TILE_M) {
            r_in = tid / S; s_idx = tid - r_in * S;
            const int ray = grp * R + r_in;
            valid = r_in < R && ray < N;
            float pe[3] = {0.f, 0.f, 0.f}, feat[20], dir[3] = {0.f, 0.f, 0.f}, zv = 0.f;
#pragma unroll
            for (int i = 0; i < 20; ++i) feat[i] = 0.f;
            if (valid) {
                si = (size_t)ray * S + s_idx;
                const float px = __ldg(io.pts + si * 3), py = __ldg(io.pts + si * 3 + 1), pz = __ldg(io.pts + si * 3 + 2);
                pe[0] = __ldg(io.ndc + si * 3); pe[1] = __ldg(io.ndc + si * 3 + 1); pe[2] = __ldg(io.ndc + si * 3 + 2);
                zv = __ldg(io.z + si);
                const float dx = __ldg(io.dirs + (size_t)ray * 3), dy = __ldg(io.dirs + (size_t)ray * 3 + 1),
                            dz = __ldg(io.dirs + (size_t)ray * 3 + 2);
                view_dir(cams, dx, dy, dz, dir);
                sample_volume(sc, pe[0], pe[1], pe[2], feat);
#pragma unroll
                for (int v = 0; v < 3; ++v) sample_color(sc, cams, v, px, py, pz, feat + 8 + 4 * v);
            }
            float* pr = s_pe + tid * PE_LD;
            float* peT = scr + bwd::S_PET + tid;
            pr[0] = pe[0]; pr[1] = pe[1]; pr[2] = pe[2];
            peT[0] = pe[0]; peT[128] = pe[1]; peT[256] = pe[2];
            float f = 1.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float sn, cs;
                    sincosf(pe[j] * f, &sn, &cs);
                    pr[3 + 3 * k + j] = sn; pr[33 + 3 * k + j] = cs;
                    peT[(3 + 3 * k + j) * 128] = sn; peT[(33 + 3 * k + j) * 128] = cs;
                }
                f *= 2.f;
            }
            pr[63] = 0.f; peT[63 * 128] = 0.f;
            float* fr = s_h + tid * FEAT_LD;
            float* fT = scr + bwd::S_FEATT + tid;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float v = i < 20 ? feat[i] : 0.f; fr[i] = v; fT[i * 128] = v; }
            s_dir[tid * 4 + 0] = dir[0]; s_dir[tid * 4 + 1] = dir[1]; s_dir[tid * 4 + 2] = dir[2];
            s_z[tid] = zv;
        }
        __syncthreads();
        {
            float acc[8][8];
            zero_acc(acc);                                                     // modulation = pts_bias(feat)
            gemm_pass<128>(acc, s_h, FEAT_LD, 32, wts + w32::WB, s_w, tid);
            epilogue128<0>(acc, wts + w32::BB, nullptr, tid);
            frag_store_rm(acc, s_mod, H_LD, tid);
            frag_store_T(acc, scr + bwd::S_MODT, tid);
            __syncthreads();
            zero_acc(acc);                                                     // layer 0
            gemm_pass<128>(acc, s_pe, PE_LD, 64, wts + w32::W0, s_w, tid);
            epilogue128<1>(acc, wts + w32::B0, s_mod, tid);
            frag_store_rm(acc, s_h, H_LD, tid);
            frag_store_T(acc, scr + bwd::S_HT, tid);
            __syncthreads();
            for (int l = 0; l < 4; ++l) {                                      // layers 1..4
                zero_acc(acc);
                gemm_pass<128>(acc, s_h, H_LD, 128, wts + w32::W1 + l * w32::LSTR, s_w, tid);
                epilogue128<1>(acc, wts + w32::W1 + l * w32::LSTR + 128 * 128, s_mod, tid);
                frag_store_rm(acc, s_h, H_LD, tid);
                frag_store_T(acc, scr + bwd::S_HT + (l + 1) * 16384, tid);
                __syncthreads();
            }
            zero_acc(acc);                                                     // layer 5: [pe, h] -> 128
            gemm_pass<128>(acc, s_pe, PE_LD, 64, wts + w32::W5, s_w, tid);
            gemm_pass<128>(acc, s_h, H_LD, 128, wts + w32::W5 + 64 * 128, s_w, tid);
            epilogue128<1>(acc, wts + w32::B5, s_mod, tid);
            frag_store_rm(acc, s_h, H_LD, tid);
            frag_store_T(acc, scr + bwd::S_HT + 5 * 16384, tid);
            __syncthreads();
            if (tid < TILE_M) {                                                // sigma = relu(alpha_linear(h6))
                const float4* hr = reinterpret_cast<const float4*>(s_h + tid * H_LD);
                const float4* wa = reinterpret_cast<const float4*>(wts + w32::WA);
                float s = 0.f;
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const float4 a = hr[i], b = __ldg(wa + i);
                    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
                }
                s_sig[tid] = fmaxf(s + __ldg(wts + w32::BA), 0.f);
            }
            zero_acc(acc);                                                     // f = feature_linear(h6)
            gemm_pass<128>(acc, s_h, H_LD, 128, wts + w32::WF, s_w, tid);
            epilogue128<0>(acc, wts + w32::BF, nullptr, tid);
            frag_store_rm(acc, s_h, H_LD, tid);
            frag_store_T(acc, scr + bwd::S_FT, tid);
            __syncthreads();
        }
        {
            float acc[8][4];                                                   // hv = relu(views_linear([f, dir])) -> s_mod [128][HV_LD]
            zero_acc(acc);
            gemm_pass<64>(acc, s_h, H_LD, 128, wts + w32::WV, s_w, tid);
            const float4 bv = __ldg(reinterpret_cast<const float4*>(wts + w32::BV + tx * 4));
            const float4 wd0 = __ldg(reinterpret_cast<const float4*>(wts + w32::WVD + 0 * 64 + tx * 4));
            const float4 wd1 = __ldg(reinterpret_cast<const float4*>(wts + w32::WVD + 1 * 64 + tx * 4));
            const float4 wd2 = __ldg(reinterpret_cast<const float4*>(wts + w32::WVD + 2 * 64 + tx * 4));
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int row = frag_row(ty, r);
                const float d0 = s_dir[row * 4], d1 = s_dir[row * 4 + 1], d2 = s_dir[row * 4 + 2];
                float4 o;
                o.x = fmaxf(fmaf(d2, wd2.x, fmaf(d1, wd1.x, fmaf(d0, wd0.x, acc[r][0]))) + bv.x, 0.f);
                o.y = fmaxf(fmaf(d2, wd2.y, fmaf(d1, wd1.y, fmaf(d0, wd0.y, acc[r][1]))) + bv.y, 0.f);
                o.z = fmaxf(fmaf(d2, wd2.z, fmaf(d1, wd1.z, fmaf(d0, wd0.z, acc[r][2]))) + bv.z, 0.f);
                o.w = fmaxf(fmaf(d2, wd2.w, fmaf(d1, wd1.w, fmaf(d0, wd0.w, acc[r][3]))) + bv.w, 0.f);
                *reinterpret_cast<float4*>(s_mod + row * HV_LD + tx * 4) = o;
            }
            __syncthreads();
        }
        if (tid < TILE_M) {                                                    // rgb = sigmoid(rgb_linear(hv)), alpha
            const float4* hr = reinterpret_cast<const float4*>(s_mod + tid * HV_LD);
            float o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4* wr = reinterpret_cast<const float4*>(wts + w32::WR + c * 64);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 a = hr[i], b = __ldg(wr + i);
                    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
                }
                s += __ldg(wts + w32::BR + c);
                o[c] = __fdiv_rn(1.f, 1.f + expf(-s));
            }
            s_rgb[tid * 4 + 0] = o[0]; s_rgb[tid * 4 + 1] = o[1]; s_rgb[tid * 4 + 2] = o[2];
            s_rgb[tid * 4 + 3] = 1.f - expf(-s_sig[tid]);
        }
        __syncthreads();

        // =============================== compositing: forward + reverse scan ===========================
        // one thread per ray (renderer.py:65-92).  f_j = 1 - alpha_j + 1e-10, T_{j+1} = T_j f_j, w_j = alpha_j T_j.
        // d alpha_j = T_j (d w_j - B_j),  B_{j-1} = d w_j alpha_j + f_j B_j  (no division by f_j).
        if (tid < R) {
            const int ray = grp * R + tid, first = tid * S;
            if (ray < N) {
                float cr = 0.f, cg = 0.f, cb = 0.f, dp = 0.f, ac = 0.f, T = 1.f;
                for (int j = first; j < first + S; ++j) {
                    const float a = s_rgb[j * 4 + 3], w = a * T;
                    s_T[j] = T;
                    cr = fmaf(w, s_rgb[j * 4 + 0], cr); cg = fmaf(w, s_rgb[j * 4 + 1], cg); cb = fmaf(w, s_rgb[j * 4 + 2], cb);
                    dp = fmaf(w, s_z[j], dp); ac += w;
                    T *= (1.f - a) + 1e-10f;
                }
                if (sc.white_bkgd) { const float bg = 1.f - ac; cr += bg; cg += bg; cb += bg; }
                if (bw.rgb_out) { bw.rgb_out[(size_t)ray * 3] = cr; bw.rgb_out[(size_t)ray * 3 + 1] = cg; bw.rgb_out[(size_t)ray * 3 + 2] = cb; }
                if (bw.depth_out) bw.depth_out[ray] = dp;
                float g0, g1, g2;
                if (bw.target) {                     // fused img2mse (utils.py: mean((rgb - target)^2))
                    const float e0 = cr - __ldg(bw.target + (size_t)ray * 3), e1 = cg - __ldg(bw.target + (size_t)ray * 3 + 1),
                                e2 = cb - __ldg(bw.target + (size_t)ray * 3 + 2);
                    g0 = 2.f * e0 * bw.inv_count; g1 = 2.f * e1 * bw.inv_count; g2 = 2.f * e2 * bw.inv_count;
                    if (bw.loss) atomicAdd(bw.loss, (e0 * e0 + e1 * e1 + e2 * e2) * bw.inv_count);
                } else {
                    g0 = __ldg(bw.g_rgb + (size_t)ray * 3); g1 = __ldg(bw.g_rgb + (size_t)ray * 3 + 1); g2 = __ldg(bw.g_rgb + (size_t)ray * 3 + 2);
                }
                const float gd = bw.g_depth ? __ldg(bw.g_depth + ray) : 0.f;
                const float gbg = sc.white_bkgd ? (g0 + g1 + g2) : 0.f;
                float B = 0.f;
                for (int j = first + S - 1; j >= first; --j) {
                    const size_t sj = (size_t)ray * S + (j - first);
                    const float a = s_rgb[j * 4 + 3], Tj = s_T[j], w = a * Tj;
                    const float c0 = s_rgb[j * 4], c1 = s_rgb[j * 4 + 1], c2 = s_rgb[j * 4 + 2];
                    float dw = g0 * c0 + g1 * c1 + g2 * c2 + gd * s_z[j] - gbg;
                    if (bw.g_weights) dw += __ldg(bw.g_weights + sj);
                    float da = Tj * (dw - B);
                    if (bw.g_alpha) da += __ldg(bw.g_alpha + sj);
                    B = fmaf(((1.f - a) + 1e-10f), B, dw * a);
                    s_g[j * 4 + 0] = w * g0 * c0 * (1.f - c0);
                    s_g[j * 4 + 1] = w * g1 * c1 * (1.f - c1);
                    s_g[j * 4 + 2] = w * g2 * c2 * (1.f - c2);
                    s_g[j * 4 + 3] = s_sig[j] > 0.f ? da * (1.f - a) : 0.f;
                }
            } else {
                for (int j = first; j < first + S; ++j) { s_g[j * 4] = s_g[j * 4 + 1] = s_g[j * 4 + 2] = s_g[j * 4 + 3] = 0.f; }
            }
        }
        if (tid >= R * S && tid < TILE_M) { s_g[tid * 4] = s_g[tid * 4 + 1] = s_g[tid * 4 + 2] = s_g[tid * 4 + 3] = 0.f; }
        stage_T(s_pe, scr + bwd::S_PET, 64, tid);             // peT -> shared (used by the layer-5 and layer-0 wgrads)
        __syncthreads();

        // =============================== backward through the heads =================================
        // rgb_linear: d Wr, d br, d ba (small reductions over the 128 rows), d hv_pre
        if (tid < 64) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int r = 0; r < TILE_M; ++r) {
                const float h = s_mod[r * HV_LD + tid];
                a0 = fmaf(s_g[r * 4], h, a0); a1 = fmaf(s_g[r * 4 + 1], h, a1); a2 = fmaf(s_g[r * 4 + 2], h, a2);
            }
            G[bwd::G_WR + tid] += a0; G[bwd::G_WR + 64 + tid] += a1; G[bwd::G_WR + 128 + tid] += a2;
        } else if (tid < 68) {
            const int c = tid - 64;
            float a = 0.f;
            for (int r = 0; r < TILE_M; ++r) a += s_g[r * 4 + c];
            G[bwd::G_BR + c] += a;                             // c == 3: d ba = sum d sigma_pre
        }
        {
            // d hv_pre[r][j] = (hv > 0) * sum_c d rgb_pre[r][c] Wr[c][j]   -> s_h (A, lda H_LD) and scratch [128][64] (B)
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wts + w32::WR + tx * 4));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wts + w32::WR + 64 + tx * 4));
            const float4 w2 = __ldg(reinterpret_cast<const float4*>(wts + w32::WR + 128 + tx * 4));
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int row = frag_row(ty, r);
                const float d0 = s_g[row * 4], d1 = s_g[row * 4 + 1], d2 = s_g[row * 4 + 2];
                const float4 hv = *reinterpret_cast<const float4*>(s_mod + row * HV_LD + tx * 4);
                float4 o;
                o.x = hv.x > 0.f ? fmaf(d2, w2.x, fmaf(d1, w1.x, d0 * w0.x)) : 0.f;
                o.y = hv.y > 0.f ? fmaf(d2, w2.y, fmaf(d1, w1.y, d0 * w0.y)) : 0.f;
                o.z = hv.z > 0.f ? fmaf(d2, w2.z, fmaf(d1, w1.z, d0 * w0.z)) : 0.f;
                o.w = hv.w > 0.f ? fmaf(d2, w2.w, fmaf(d1, w1.w, d0 * w0.w)) : 0.f;
                *reinterpret_cast<float4*>(s_h + row * H_LD + tx * 4) = o;
                *reinterpret_cast<float4*>(scr + bwd::S_DPRE + row * 64 + tx * 4) = o;
            }
        }
        __syncthreads();                                        // hv (s_mod) no longer needed; d hv_pre complete
        stage_T(s_mod, scr + bwd::S_FT, 128, tid);              // fT -> A operand of the views wgrad
        if (tid < 64) {                                         // d bv, d Wv[:, 128:131]^T
            float sb = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
            for (int r = 0; r < TILE_M; ++r) {
                const float d = s_h[r * H_LD + tid];
                sb += d; s0 = fmaf(s_dir[r * 4], d, s0); s1 = fmaf(s_dir[r * 4 + 1], d, s1); s2 = fmaf(s_dir[r * 4 + 2], d, s2);
            }
            G[bwd::G_BV + tid] += sb;
            G[bwd::G_WVDT + tid] += s0; G[bwd::G_WVDT + 64 + tid] += s1; G[bwd::G_WVDT + 128 + tid] += s2;
        }
        cp_async_wait<0>();
        __syncthreads();
        {
            float acc[8][4];                                    // d Wv_f^T[k][j] = sum_r fT[k][r] d hv_pre[r][j]
            zero_acc(acc);
            gemm_pass<64>(acc, s_mod, H_LD, 128, scr + bwd::S_DPRE, s_w, tid);
            frag_accumulate(acc, G + bwd::G_WVFT, 64, tid);
        }
        float acc[8][8];
        zero_acc(acc);                                          // d f[r][k] = sum_j d hv_pre[r][j] Wv[j][k]
        gemm_pass<128>(acc, s_h, H_LD, 64, bw.wd + bwd::D_VF, s_w, tid);
        frag_store_rm(acc, s_h, H_LD, tid);
        frag_store_rm(acc, scr + bwd::S_DPRE, 128, tid);
        __syncthreads();
        stage_T(s_mod, scr + bwd::S_HT + 5 * 16384, 128, tid);  // h6T
        colsum_accumulate(s_h, H_LD, 128, G + bwd::G_BF, tid);  // d bf
        cp_async_wait<0>();
        __syncthreads();
        zero_acc(acc);                                          // d Wf^T[k][n] = sum_r h6T[k][r] d f[r][n]
        gemm_pass<128>(acc, s_mod, H_LD, 128, scr + bwd::S_DPRE, s_w, tid);
        frag_accumulate(acc, G + bwd::G_WFT, 128, tid);
        if (tid < TILE_M) {                                     // d wa[k] = sum_r d sigma_pre[r] h6T[k][r]
            float s = 0.f;
            for (int r = 0; r < TILE_M; ++r) s = fmaf(s_g[r * 4 + 3], s_mod[tid * H_LD + r], s);
            G[bwd::G_WA + tid] += s;
        }
        zero_acc(acc);                                          // d h6 = d f Wf + d sigma_pre (x) wa
        gemm_pass<128>(acc, s_h, H_LD, 128, bw.wd + bwd::D_F, s_w, tid);
        {
            const float4 al = __ldg(reinterpret_cast<const float4*>(wts + w32::WA + tx * 4));
            const float4 ah = __ldg(reinterpret_cast<const float4*>(wts + w32::WA + 64 + tx * 4));
            const float wa[8] = {al.x, al.y, al.z, al.w, ah.x, ah.y, ah.z, ah.w};
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float ds = s_g[frag_row(ty, r) * 4 + 3];
#pragma unroll
                for (int n = 0; n < 8; ++n) acc[r][n] = fmaf(ds, wa[n], acc[r][n]);
            }
        }

        // =============================== trunk, layers 5..0 ==========================================
#pragma unroll 1
        for (int l = 5; l >= 0; --l) {
            // element-wise: acc = d h_{l+1}  ->  d g = acc * (h > 0) ; d pre = d g * mod ; d mod += d g * pre, pre = h / mod
            {
                const float* hT = scr + bwd::S_HT + l * 16384;
                const float* mT = scr + bwd::S_MODT;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int col = frag_col(tx, n);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const float4 h4 = *reinterpret_cast<const float4*>(hT + col * 128 + half * 64 + ty * 4);
                        const float4 m4 = *reinterpret_cast<const float4*>(mT + col * 128 + half * 64 + ty * 4);
                        const float hh[4] = {h4.x, h4.y, h4.z, h4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
                        float dm[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = half * 4 + i;
                            const bool on = hh[i] > 0.f;
                            const float dg = on ? acc[r][n] : 0.f;
                            acc[r][n] = dg * mm[i];
                            dm[i] = on ? dg * __fdiv_rn(hh[i], mm[i]) : 0.f;
                        }
                        float4* dmod = reinterpret_cast<float4*>(scr + bwd::S_DMOD + col * 128 + half * 64 + ty * 4);
                        float4 o = make_float4(dm[0], dm[1], dm[2], dm[3]);
                        if (l != 5) { const float4 p = *dmod; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                        *dmod = o;
                    }
                }
            }
            frag_store_rm(acc, s_h, H_LD, tid);
            frag_store_rm(acc, scr + bwd::S_DPRE, 128, tid);
            __syncthreads();
            if (l >= 1) stage_T(s_mod, scr + bwd::S_HT + (l - 1) * 16384, 128, tid);      // x_l^T = h_l^T
            colsum_accumulate(s_h, H_LD, 128, G + bwd::G_B + l * 128, tid);                // d b_l
            cp_async_wait<0>();
            __syncthreads();
            if (l >= 1) {                                       // d W_l^T[k][n] = sum_r h_l^T[k][r] d pre[r][n]
                zero_acc(acc);
                gemm_pass<128>(acc, s_mod, H_LD, 128, scr + bwd::S_DPRE, s_w, tid);
                frag_accumulate(acc, G + (l == 5 ? bwd::G_W5HT : bwd::G_W14T + (l - 1) * 16384), 128, tid);
            }
            if (l == 5 || l == 0) {                             // positional-encoding part: 64-row A
                float a4[4][8];
                zero_acc(a4);
                gemm_pass<128>(a4, s_pe, H_LD, 128, scr + bwd::S_DPRE, s_w, tid);
                frag_accumulate(a4, G + (l == 5 ? bwd::G_W5PET : bwd::G_W0T), 128, tid);
            }
            if (l >= 1) {                                       // d h_l = d pre W_l (h part)
                zero_acc(acc);
                gemm_pass<128>(acc, s_h, H_LD, 128, bw.wd + (l == 5 ? bwd::D_5H : bwd::D_14 + (l - 1) * 16384), s_w, tid);
            }
        }
        // =============================== modulation branch: pts_bias ==================================
        // d mod is complete in scratch (transposed): bring it to row-major -- A operand of d feat (shared), B operand of d Wb
        frag_load_T(acc, scr + bwd::S_DMOD, tid);
        frag_store_rm(acc, s_h, H_LD, tid);
        frag_store_rm(acc, scr + bwd::S_DPRE, 128, tid);
        stage_T(s_mod, scr + bwd::S_FEATT, 64, tid);
        cp_async_wait<0>();
        __syncthreads();
        colsum_accumulate(s_h, H_LD, 128, G + bwd::G_BB, tid);
        {
            float a4[4][8];                                     // d Wb^T[k][n] = sum_r featT[k][r] d mod[r][n]
            zero_acc(a4);
            gemm_pass<128>(a4, s_mod, H_LD, 128, scr + bwd::S_DPRE, s_w, tid);
            frag_accumulate(a4, G + bwd::G_WBT, 128, tid);
        }
        if (bw.dvol) {
            float a[8][4];                                      // d feat[r][k] = sum_n d mod[r][n] Wb[n][k]
            zero_acc(a);
            gemm_pass<64>(a, s_h, H_LD, 128, bw.wd + bwd::D_B, s_w, tid);
            if (tx < 2) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    *reinterpret_cast<float4*>(s_df + frag_row(ty, r) * 8 + tx * 4) = make_float4(a[r][0], a[r][1], a[r][2], a[r][3]);
            }
            __syncthreads();
            if (tid < TILE_M && valid) {
                // trilinear scatter (transpose of utils.index_point_feature, utils.py:357-383): same corner weights
                float g8[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) g8[c] = s_df[tid * 8 + c];
                if (bw.g_feat) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) g8[c] += __ldg(bw.g_feat + si * 20 + c);
                }
                const float nx = __ldg(io.ndc + si * 3), ny = __ldg(io.ndc + si * 3 + 1), nz = __ldg(io.ndc + si * 3 + 2);
                const int W = sc.Wp, H = sc.Hp, D = sc.D;
                const float ix = ((nx * 2.f - 1.f + 1.f) * 0.5f) * (float)(W - 1);
                const float iy = ((ny * 2.f - 1.f + 1.f) * 0.5f) * (float)(H - 1);
                const float iz = ((nz * 2.f - 1.f + 1.f) * 0.5f) * (float)(D - 1);
                const float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
                const float wx[2] = {(x0f + 1.f) - ix, ix - x0f}, wy[2] = {(y0f + 1.f) - iy, iy - y0f}, wz[2] = {(z0f + 1.f) - iz, iz - z0f};
                const int x0 = (int)fminf(fmaxf(x0f, -2.f), (float)W), y0 = (int)fminf(fmaxf(y0f, -2.f), (float)H),
                          z0 = (int)fminf(fmaxf(z0f, -2.f), (float)D);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int x = x0 + (c & 1), y = y0 + ((c >> 1) & 1), z = z0 + (c >> 2);
                    if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H || (unsigned)z >= (unsigned)D) continue;
                    const float wgt = wx[c & 1] * wy[(c >> 1) & 1] * wz[c >> 2];
                    float4* p = reinterpret_cast<float4*>(bw.dvol + (((size_t)z * H + y) * W + x) * 8);
                    atomicAdd(p, make_float4(g8[0] * wgt, g8[1] * wgt, g8[2] * wgt, g8[3] * wgt));
                    atomicAdd(p + 1, make_float4(g8[4] * wgt, g8[5] * wgt, g8[6] * wgt, g8[7] * wgt));
                }
            }
        }
        __syncthreads();
    }
}

// ---- per-CTA accumulators -> the 22 tensors in nn.Linear layout ---------------------------------------------
struct GradOut { float* p[MVSN_N_MLP_TENSORS]; };

__global__ void mlp_grad_reduce_kernel(const float* __restrict__ grads, int n_ctas, GradOut out) {
    // blockIdx.y = tensor index; threads enumerate its elements with the output-row index fastest (coalesced reads)
    const int t = blockIdx.y;
    int rows, cols;             // nn.Linear weight [rows][cols] or bias [rows] (cols = 1)
    switch (t) {
        case 0: rows = 128; cols = 63; break;
        case 2: case 4: case 6: case 8: case 16: rows = 128; cols = 128; break;
        case 10: rows = 128; cols = 191; break;
        case 12: rows = 128; cols = 20; break;
        case 14: rows = 64; cols = 131; break;
        case 18: rows = 1; cols = 128; break;
        case 20: rows = 3; cols = 64; break;
        case 15: rows = 64; cols = 1; break;
        case 19: rows = 1; cols = 1; break;
        case 21: rows = 3; cols = 1; break;
        default: rows = 128; cols = 1; break;          // biases 1,3,5,7,9,11,13,17
    }
    const int total = rows * cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i % rows, k = i / rows;
        int src;
        switch (t) {
            case 0: src = bwd::G_W0T + k * 128 + n; break;
            case 2: case 4: case 6: case 8: src = bwd::G_W14T + (t / 2 - 1) * 16384 + k * 128 + n; break;
            case 10: src = k < 63 ? bwd::G_W5PET + k * 128 + n : bwd::G_W5HT + (k - 63) * 128 + n; break;
            case 12: src = bwd::G_WBT + k * 128 + n; break;
            case 14: src = k < 128 ? bwd::G_WVFT + k * 64 + n : bwd::G_WVDT + (k - 128) * 64 + n; break;
            case 16: src = bwd::G_WFT + k * 128 + n; break;
            case 18: src = bwd::G_WA + k; break;
            case 20: src = bwd::G_WR + n * 64 + k; break;
            case 1: case 3: case 5: case 7: case 9: case 11: src = bwd::G_B + (t / 2) * 128 + n; break;
            case 13: src = bwd::G_BB + n; break;
            case 15: src = bwd::G_BV + n; break;
            case 17: src = bwd::G_BF + n; break;
            case 19: src = bwd::G_BR + 3; break;
            default: src = bwd::G_BR + n; break;        // 21
        }
        float s = 0.f;
        for (int c = 0; c < n_ctas; ++c) s += grads[(size_t)c * bwd::GRADS + src];
        out.p[t][n * cols + k] = s;
    }
}

struct MlpPtrsB { const float* p[MVSN_N_MLP_TENSORS]; };

__global__ void pack_dgrad_kernel(MlpPtrsB w, float* __restrict__ out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < 64 * 128; i += nt) out[bwd::D_VF + i] = w.p[14][(i >> 7) * 131 + (i & 127)];
    for (int i = tid; i < 128 * 128; i += nt) {
        out[bwd::D_F + i] = w.p[16][i];
        out[bwd::D_5H + i] = w.p[10][(i >> 7) * 191 + 63 + (i & 127)];
        for (int l = 1; l <= 4; ++l) out[bwd::D_14 + (l - 1) * 16384 + i] = w.p[2 * l][i];
    }
    for (int i = tid; i < 128 * 64; i += nt) out[bwd::D_B + i] = (i & 63) < 20 ? w.p[12][(i >> 6) * 20 + (i & 63)] : 0.f;
}

// ---- fused Adam (torch.optim.Adam: no weight decay, no amsgrad) ---------------------------------------------
struct AdamScalars { float lr, beta1, beta2, eps, bc1, bc2_sqrt; };

__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const AdamScalars& a) {
    m = fmaf(a.beta1, m, (1.f - a.beta1) * g);
    v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);
    const float denom = __fdiv_rn(sqrtf(v), a.bc2_sqrt) + a.eps;
    return p - __fdiv_rn(a.lr, a.bc1) * __fdiv_rn(m, denom);
}

struct AdamTensors { float* p[MVSN_N_MLP_TENSORS]; const float* g[MVSN_N_MLP_TENSORS]; float* m[MVSN_N_MLP_TENSORS];
                     float* v[MVSN_N_MLP_TENSORS]; int n[MVSN_N_MLP_TENSORS]; int count; };

__global__ void adam_tensors_kernel(AdamTensors t, AdamScalars a) {
    for (int k = blockIdx.y; k < t.count; k += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.n[k]; i += gridDim.x * blockDim.x) {
            float m = t.m[k][i], v = t.v[k][i];
            t.p[k][i] = adam_update(t.p[k][i], t.g[k][i], m, v, a);
            t.m[k][i] = m; t.v[k][i] = v;
        }
}

// volume: the gradient is channels-last [nvox][8] (the scatter target) and is ZEROED here for the next step; parameter and
// moments here are planar [8][nvox] (a checkpoint-layout nn.Parameter); the channels-last case is adam_volume_cl_kernel
__global__ void adam_volume_planar_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                   long long nvox, AdamScalars a) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
        float4* g4 = reinterpret_cast<float4*>(g) + 2 * i;
        const float4 ga = g4[0], gb = g4[1];
        g4[0] = make_float4(0.f, 0.f, 0.f, 0.f); g4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const long long j = c * nvox + i;
            float mm = m[j], vv = v[j];
            p[j] = adam_update(p[j], gg[c], mm, vv, a);
            m[j] = mm; v[j] = vv;
        }
    }
}

// channels-last parameter: parameter, moments and gradient share one layout, so the update is purely element-wise --
// one float4 per thread, consecutive lanes on consecutive 16 bytes (fully coalesced streams)
__global__ void adam_volume_cl_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                      long long n4, AdamScalars a) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 gg = g[i];
        g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pp = p[i], mm = m[i], vv = v[i];
        pp.x = adam_update(pp.x, gg.x, mm.x, vv.x, a); pp.y = adam_update(pp.y, gg.y, mm.y, vv.y, a);
        pp.z = adam_update(pp.z, gg.z, mm.z, vv.z, a); pp.w = adam_update(pp.w, gg.w, mm.w, vv.w, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

static int bwd_grid(int N, int S) {
    const int R = TILE_M / S;
    const int ngroups = (N + R - 1) / R;
    return ngroups < sm_count() ? ngroups : sm_count();
}

}  // namespace

size_t render_backward_workspace_bytes(int N, int S) {
    if (S <= 0 || S > TILE_M || N <= 0) return 0;
    const size_t ctas = (size_t)sm_count();              // sized for the widest launch on this device
    return (ctas * (bwd::SCRATCH + bwd::GRADS) + bwd::DGRAD) * sizeof(float);
}

int launch_render_backward(const SceneDev& sc, const RenderIO& io, const float* wts_fp32, const float* const* mlp_w,
                           const float* g_rgb, const float* target, float inv_count, const float* g_depth,
                           const float* g_weights, const float* g_alpha, const float* g_feat, float* const* grad_mlp,
                           float* dvol, float* rgb_out, float* depth_out, float* loss, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream) {
    MVSN_REQUIRE(io.S <= TILE_M, MVSN_EUNSUPPORTED, "render backward: N_samples=%d > 128 is not implemented", io.S);
    const size_t need = render_backward_workspace_bytes(io.N, io.S);
    MVSN_REQUIRE(workspace && workspace_bytes >= need, MVSN_EWORKSPACE, "render backward: workspace %zu < %zu bytes", workspace_bytes, need);
    MVSN_REQUIRE(aligned16(workspace), MVSN_EALIGN, "render backward: workspace must be 16-byte aligned");
    static bool attr_set[64] = {false};
    int dev = 0;
    MVSN_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        MVSN_CUDA_CHECK(cudaFuncSetAttribute(render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM_BYTES));
        if (dev < 64) attr_set[dev] = true;
    }
    const int grid = bwd_grid(io.N, io.S);
    float* ws = static_cast<float*>(workspace);
    const size_t ctas = (size_t)sm_count();
    BwdIO bw{};
    bw.g_rgb = g_rgb; bw.target = target; bw.inv_count = inv_count; bw.g_depth = g_depth; bw.g_weights = g_weights;
    bw.g_alpha = g_alpha; bw.g_feat = g_feat; bw.dvol = dvol; bw.rgb_out = rgb_out; bw.depth_out = depth_out; bw.loss = loss;
    bw.scratch = ws; bw.grads = ws + ctas * bwd::SCRATCH;
    float* wd = ws + ctas * (bwd::SCRATCH + bwd::GRADS);
    bw.wd = wd;
    MlpPtrsB wp;
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i) wp.p[i] = mlp_w[i];
    pack_dgrad_kernel<<<64, 256, 0, stream>>>(wp, wd);
    MVSN_CUDA_CHECK(cudaGetLastError());
    render_bwd_kernel<<<grid, 256, BWD_SMEM_BYTES, stream>>>(sc, io, bw, wts_fp32);
    MVSN_CUDA_CHECK(cudaGetLastError());
    GradOut go;
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i) go.p[i] = grad_mlp[i];
    mlp_grad_reduce_kernel<<<dim3(16, MVSN_N_MLP_TENSORS), 256, 0, stream>>>(bw.grads, grid, go);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

int launch_adam_tensors(float* const* p, const float* const* g, float* const* m, float* const* v, const int* n, int count,
                        float lr, float beta1, float beta2, float eps, int step, cudaStream_t stream) {
    MVSN_REQUIRE(count >= 0 && count <= MVSN_N_MLP_TENSORS, MVSN_EBADSHAPE, "adam: %d tensors (max %d per call)", count, MVSN_N_MLP_TENSORS);
    if (count == 0) return MVSN_OK;
    AdamTensors t{};
    for (int i = 0; i < count; ++i) { t.p[i] = p[i]; t.g[i] = g[i]; t.m[i] = m[i]; t.v[i] = v[i]; t.n[i] = n[i]; }
    t.count = count;
    const AdamScalars a{lr, beta1, beta2, eps, (float)(1.0 - pow((double)beta1, step)), (float)sqrt(1.0 - pow((double)beta2, step))};
    adam_tensors_kernel<<<dim3(16, count), 256, 0, stream>>>(t, a);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

int launch_adam_volume(float* p, float* g_dhwc, float* m, float* v, long long nvox, int planar, float lr, float beta1,
                       float beta2, float eps, int step, cudaStream_t stream) {
    const AdamScalars a{lr, beta1, beta2, eps, (float)(1.0 - pow((double)beta1, step)), (float)sqrt(1.0 - pow((double)beta2, step))};
    const int grid = sm_count() * 8;
    if (planar) adam_volume_planar_kernel<<<grid, 256, 0, stream>>>(p, g_dhwc, m, v, nvox, a);
    else        adam_volume_cl_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g_dhwc),
                                                                reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), 2 * nvox, a);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // namespace mvsn

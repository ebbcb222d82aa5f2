// K-F: FeatureNet (models.py:688-722) -- eight bias-free 2-D convolutions (3x3 s1 / 5x5 s2), each followed by
// InPlaceABN in TRAIN mode with statistics over ALL views jointly (the reference feeds the V source
// images as one batch, models.py:907-909), then a 1x1 convolution with bias.
//
// Same scheme as the 3-D stack (conv3d.cu): one kernel per layer that applies the previous layer's
// normalisation + leaky-ReLU while loading, convolves, stores the raw result once and accumulates the
// per-channel sum / sum of squares (fp32 per CTA, fp64 across CTAs) for the next layer's load.  A thread owns
// 4 consecutive output pixels along x for CT output channels; input rows arrive as 16-byte loads (two for the
// stride-2 layers), halos come from the neighbouring lanes by shuffle, rows are software-pipelined
// (next row's raw values are requested before the current row's FMAs), weights sit in shared memory as
// [cin][K*K][CT].
#include "conv_common.cuh"

namespace mvsn {

struct Conv2dArgs {
    ActSrc in;                  // x: [V][Cin][Hin][Win] raw (or the plain images for the first layer)
    int V, Cin, Hin, Win;
    const float* w;             // [Cout][Cin][K][K]
    int Cout, Hout, Wout;
    float* out;                 // [V][Cout][Hout][Wout] raw
    double* stats_out;          // [Cout][2]
};

template <int CT, int K, int STRIDE, bool IDENT>
__global__ void __launch_bounds__(128)
conv2d_kernel(const Conv2dArgs a) {
    constexpr int PAD = K / 2;
    constexpr int NV = 4 * STRIDE;                          // centre inputs xin .. xin+NV-1 (one or two float4)
    static_assert((3 * STRIDE + K - 1) - PAD - (NV - 1) == 1, "exactly one right-halo value");
    extern __shared__ __align__(16) float s_w[];            // [Cin][K*K][CT]
    __shared__ float s_sc[kMaxCin], s_sh[kMaxCin];
    __shared__ float s_stat[4 * 2 * CT];            // per-warp partial sums (4 warps), combined in fixed point
    const int tid = threadIdx.x, lane = tid & 31;
    const int cg = blockIdx.y;

    for (int i = tid; i < a.Cin * K * K * CT; i += 128) {
        const int c = i % CT, r = i / CT;                   // r = ci*K*K + tap
        s_w[i] = __ldg(a.w + (size_t)(cg * CT + c) * a.Cin * K * K + r);
    }
    if (!IDENT) load_norm(a.in, a.Cin, s_sc, s_sh, tid, 128);
    
    __syncthreads();

    const int nsx = (a.Wout + 3) >> 2;
    const long long nstrips = (long long)a.V * a.Hout * nsx;
    const long long sid = (long long)blockIdx.x * 128 + tid;
    const bool active = sid < nstrips;
    int v_ = 0, y = 0, sx = 0;
    if (active) { sx = (int)(sid % nsx); long long r = sid / nsx; y = (int)(r % a.Hout); v_ = (int)(r / a.Hout); }
    const int x0 = sx * 4, xin = x0 * STRIDE;
    const bool vec = (a.Win & 3) == 0;
    const size_t iplane = (size_t)a.Hin * a.Win;
    const float* __restrict__ in_v = a.in.x + (size_t)v_ * a.Cin * iplane;

    bool xok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xok[i] = xin + i < a.Win;
    auto row_ptr = [&](int ci, int ky, bool& ok) -> const float* {
        const int yi = y * STRIDE - PAD + ky;
        ok = active && (unsigned)yi < (unsigned)a.Hin;
        return in_v + (size_t)ci * iplane + (long long)yi * a.Win + xin;
    };
    auto fetch = [&](const float* p, int xo, bool ok) -> float4 {     // raw values; invalid taps read nothing
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
            if (vec && xok[xo + 3]) v = __ldg(reinterpret_cast<const float4*>(p + xo));
            else {
                if (xok[xo]) v.x = __ldg(p + xo);
                if (xok[xo + 1]) v.y = __ldg(p + xo + 1);
                if (xok[xo + 2]) v.z = __ldg(p + xo + 2);
                if (xok[xo + 3]) v.w = __ldg(p + xo + 3);
            }
        }
        return v;
    };
    const bool halo_l = lane == 0 && sx > 0;
    const bool halo_r = lane == 31 && sx < nsx - 1 && xin + NV < a.Win;

    float acc[4][CT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[i][c] = 0.f;

    bool ok_n;
    const float* p_n = row_ptr(0, 0, ok_n);
    float4 n0 = fetch(p_n, 0, ok_n), n1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (STRIDE == 2) n1 = fetch(p_n, 4, ok_n);
    for (int ci = 0; ci < a.Cin; ++ci) {
        const float sc0 = IDENT ? 1.f : s_sc[ci], sh0 = IDENT ? 0.f : s_sh[ci];
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const float4 r0 = n0, r1 = n1;
            const bool row_ok = ok_n;
            const float* p = p_n;
            {
                const bool wrap = ky == K - 1;
                p_n = row_ptr(wrap ? ci + 1 : ci, wrap ? 0 : ky + 1, ok_n);
                if (wrap && ci + 1 == a.Cin) ok_n = false;
                n0 = fetch(p_n, 0, ok_n);
                if (STRIDE == 2) n1 = fetch(p_n, 4, ok_n);
            }
            // v[PAD + i] = activated input at xin + i (zero outside the image: zero padding of the ACTIVATED map)
            float v[PAD + NV + 1];
            const float raw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < NV; ++i)
                v[PAD + i] = (IDENT || !(row_ok && xok[i])) ? raw[i] : act(raw[i], sc0, sh0);
            float hl[PAD], hr = 0.f;
#pragma unroll
            for (int j = 0; j < PAD; ++j) {
                hl[j] = 0.f;
                if (halo_l && row_ok) hl[j] = IDENT ? __ldg(p - PAD + j) : act(__ldg(p - PAD + j), sc0, sh0);
            }
            if (halo_r && row_ok) hr = IDENT ? __ldg(p + NV) : act(__ldg(p + NV), sc0, sh0);
#pragma unroll
            for (int j = 0; j < PAD; ++j) {                 // input xin - PAD + j = the left strip's centre value NV - PAD + j
                const float sl = __shfl_up_sync(0xffffffffu, v[PAD + NV - PAD + j], 1);
                v[j] = (lane == 0 || sx == 0) ? hl[j] : sl;
            }
            const float sr = __shfl_down_sync(0xffffffffu, v[PAD], 1);
            v[PAD + NV] = (lane == 31 || sx == nsx - 1) ? hr : sr;

            const float* wrow = s_w + ((ci * K + ky) * K) * CT;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                float wv[CT];
#pragma unroll
                for (int c4 = 0; c4 < CT / 4; ++c4) {
                    const float4 t = *reinterpret_cast<const float4*>(wrow + kx * CT + c4 * 4);
                    wv[c4 * 4] = t.x; wv[c4 * 4 + 1] = t.y; wv[c4 * 4 + 2] = t.z; wv[c4 * 4 + 3] = t.w;
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float in = v[o * STRIDE + kx];
#pragma unroll
                    for (int c = 0; c < CT; ++c) acc[o][c] = fmaf(in, wv[c], acc[o][c]);
                }
            }
        }
    }

    // raw store + batch statistics
    const size_t oplane = (size_t)a.Hout * a.Wout;
    const bool ovec = (a.Wout & 3) == 0;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        float s = 0.f, q = 0.f;
        if (active) {
            float* o = a.out + ((size_t)v_ * a.Cout + cg * CT + c) * oplane + (size_t)y * a.Wout + x0;
            if (ovec) {
                *reinterpret_cast<float4*>(o) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
#pragma unroll
                for (int i = 0; i < 4; ++i) { s += acc[i][c]; q = fmaf(acc[i][c], acc[i][c], q); }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (x0 + i < a.Wout) { o[i] = acc[i][c]; s += acc[i][c]; q = fmaf(acc[i][c], acc[i][c], q); }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            q += __shfl_xor_sync(0xffffffffu, q, off);
        }
        if (lane == 0) { s_stat[(tid >> 5) * 2 * CT + 2 * c] = s; s_stat[(tid >> 5) * 2 * CT + 2 * c + 1] = q; }
    }
    __syncthreads();
    if (tid < 2 * CT)       // four warp partials -> fixed point -> ONE integer atomic per CTA and slot (order-independent)
        atomicAdd(reinterpret_cast<unsigned long long*>(a.stats_out) + 2 * (cg * CT) + tid,
                  stat_fx(s_stat[tid]) + stat_fx(s_stat[2 * CT + tid]) + stat_fx(s_stat[4 * CT + tid]) + stat_fx(s_stat[6 * CT + tid]));
}

// toplayer: feats = W (32x32) . ABN(conv2.2) + b, one thread per pixel, weights broadcast from shared memory
__global__ void __launch_bounds__(128)
toplayer_kernel(ActSrc s, const float* __restrict__ w, const float* __restrict__ bias, int V, long long plane,
                float* __restrict__ out) {
    constexpr int C = 32;
    __shared__ __align__(16) float s_wt[C * C];             // [ci][co]
    __shared__ float sc[C], sh[C], s_b[C];
    for (int i = threadIdx.x; i < C * C; i += 128) { const int co = i / C, ci = i % C; s_wt[ci * C + co] = __ldg(w + i); }
    if (threadIdx.x < C) s_b[threadIdx.x] = __ldg(bias + threadIdx.x);
    load_norm(s, C, sc, sh, threadIdx.x, 128);
    __syncthreads();
    const long long i = blockIdx.x * 128ll + threadIdx.x;
    if (i >= (long long)V * plane) return;
    const long long v = i / plane, px = i - v * plane;
    const float* x = s.x + (size_t)v * C * plane + px;
    float acc[C];
#pragma unroll
    for (int co = 0; co < C; ++co) acc[co] = 0.f;
#pragma unroll 4
    for (int ci = 0; ci < C; ++ci) {
        const float in = act(__ldg(x + (size_t)ci * plane), sc[ci], sh[ci]);
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 t = *reinterpret_cast<const float4*>(s_wt + ci * C + c4 * 4);
            acc[c4 * 4] = fmaf(in, t.x, acc[c4 * 4]);         acc[c4 * 4 + 1] = fmaf(in, t.y, acc[c4 * 4 + 1]);
            acc[c4 * 4 + 2] = fmaf(in, t.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(in, t.w, acc[c4 * 4 + 3]);
        }
    }
    float* o = out + (size_t)v * C * plane + px;
#pragma unroll
    for (int co = 0; co < C; ++co) o[(size_t)co * plane] = acc[co] + s_b[co];
}

template <int CT, int K, int STRIDE, bool IDENT>
static int launch_conv2d_ct(const Conv2dArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)a.Cin * K * K * CT * sizeof(float);
    MVSN_CUDA_CHECK(cudaFuncSetAttribute(conv2d_kernel<CT, K, STRIDE, IDENT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long nstrips = (long long)a.V * a.Hout * ((a.Wout + 3) / 4);
    dim3 grid(cdiv(nstrips, 128), a.Cout / CT);
    conv2d_kernel<CT, K, STRIDE, IDENT><<<grid, 128, smem, st>>>(a);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

// narrow the channel tile until the grid covers the SMs about twice (quarter-resolution layers are small)
template <int K, int STRIDE>
static int launch_conv2d(const Conv2dArgs& a, cudaStream_t st) {
    const long long strips = cdiv((long long)a.V * a.Hout * ((a.Wout + 3) / 4), 128);
    if (a.Cout % 16 == 0 && strips * (a.Cout / 16) >= 2 * sm_count()) return launch_conv2d_ct<16, K, STRIDE, false>(a, st);
    return launch_conv2d_ct<8, K, STRIDE, false>(a, st);
}

}  // namespace mvsn

using namespace mvsn;

static const int kFCin[8]  = {3, 8, 8, 16, 16, 16, 32, 32};
static const int kFCout[8] = {8, 8, 16, 16, 16, 32, 32, 32};
static const int kFLevel[8] = {0, 0, 1, 1, 1, 2, 2, 2};     // output resolution level (H >> level)

static size_t align_up_f(size_t v, size_t a) { return (v + a - 1) / a * a; }
static int half_up(int n) { return (n + 1) / 2; }           // k5 s2 p2: out = floor((n - 1) / 2) + 1

extern "C" {

size_t mvsn_featurenet_workspace_bytes(int V, int H, int W) {
    // statistics block + two ping-pong activation buffers sized for the widest layer (8 ch at full resolution;
    // 16 ch at half and 32 ch at quarter resolution are never larger)
    const size_t act = align_up_f((size_t)V * 8 * H * W * sizeof(float), 256);
    return 4096 + 2 * act;
}

int mvsn_featurenet_forward(const float* const* w, const float* imgs, int V, int H, int W, float* feats,
                            void* workspace, size_t workspace_bytes, void* stream_) {
    return mvsn_featurenet_forward_bn(w, nullptr, MVSN_BN_BATCH, 0.f, imgs, V, H, W, feats, workspace, workspace_bytes, stream_);
}

int mvsn_featurenet_forward_bn(const float* const* w, float* const* running, int bn_mode, float momentum, const float* imgs,
                               int V, int H, int W, float* feats, void* workspace, size_t workspace_bytes, void* stream_) {
    MVSN_RANGE("mvsn_featurenet_forward_bn");
    cudaStream_t st = (cudaStream_t)stream_;
    MVSN_REQUIRE(bn_mode == MVSN_BN_BATCH || bn_mode == MVSN_BN_BATCH_UPDATE || bn_mode == MVSN_BN_RUNNING, MVSN_EBADSHAPE,
                 "mvsn_featurenet_forward_bn: bn_mode %d", bn_mode);
    MVSN_REQUIRE(bn_mode == MVSN_BN_BATCH || running, MVSN_ENULL, "mvsn_featurenet_forward_bn: running statistics are NULL");
    if (bn_mode != MVSN_BN_BATCH)
        for (int i = 0; i < 16; ++i) MVSN_REQUIRE(running[i] != nullptr, MVSN_ENULL, "mvsn_featurenet_forward_bn: running[%d] is NULL", i);
    MVSN_REQUIRE(w && imgs && feats && workspace, MVSN_ENULL, "mvsn_featurenet_forward: NULL argument");
    MVSN_REQUIRE(V > 0 && H >= 4 && W >= 4, MVSN_EBADSHAPE, "mvsn_featurenet_forward: bad shape V=%d H=%d W=%d", V, H, W);
    MVSN_REQUIRE(workspace_bytes >= mvsn_featurenet_workspace_bytes(V, H, W), MVSN_EWORKSPACE,
                 "mvsn_featurenet_forward: workspace too small");
    MVSN_REQUIRE(aligned16(workspace) && aligned16(imgs) && aligned16(feats), MVSN_EALIGN,
                 "mvsn_featurenet_forward: buffers must be 16-byte aligned");
    for (int i = 0; i < MVSN_N_FEATURENET_TENSORS; ++i)
        MVSN_REQUIRE(w[i] != nullptr, MVSN_ENULL, "mvsn_featurenet_forward: weight %d is NULL", i);

    char* p = static_cast<char*>(workspace);
    double* stats = reinterpret_cast<double*>(p);            // 8 layers x 32 ch x 2
    p += 4096;
    const size_t act_bytes = align_up_f((size_t)V * 8 * H * W * sizeof(float), 256);
    float* buf[2] = {reinterpret_cast<float*>(p), reinterpret_cast<float*>(p + act_bytes)};
    MVSN_CUDA_CHECK(cudaMemsetAsync(stats, 0, 8 * 64 * sizeof(double), st));

    int hs[3] = {H, half_up(H), half_up(half_up(H))}, ws[3] = {W, half_up(W), half_up(half_up(W))};
    int rc;
    ActSrc src{imgs, nullptr, nullptr, nullptr, 1.0, nullptr, nullptr};
    BnUpdateArgs upd{};
    upd.momentum = momentum;
    int hin = H, win = W;
    for (int l = 0; l < 8; ++l) {
        const int lv = kFLevel[l];
        Conv2dArgs a;
        a.in = src; a.V = V; a.Cin = kFCin[l]; a.Hin = hin; a.Win = win;
        a.w = w[3 * l]; a.Cout = kFCout[l]; a.Hout = hs[lv]; a.Wout = ws[lv];
        a.out = buf[l & 1]; a.stats_out = stats + (size_t)l * 64;
        if (l == 0) rc = launch_conv2d_ct<8, 3, 1, true>(a, st);
        else if (l == 2 || l == 5) rc = launch_conv2d<5, 2>(a, st);
        else rc = launch_conv2d<3, 1>(a, st);
        if (rc) return rc;
        src.x = a.out; src.stats = a.stats_out; src.gamma = w[3 * l + 1]; src.beta = w[3 * l + 2];
        src.count = (double)V * a.Hout * a.Wout;
        if (bn_mode == MVSN_BN_RUNNING) { src.rmean = running[2 * l]; src.rvar = running[2 * l + 1]; }
        if (bn_mode == MVSN_BN_BATCH_UPDATE) {
            upd.stats[l] = a.stats_out; upd.count[l] = src.count; upd.C[l] = a.Cout;
            upd.rmean[l] = running[2 * l]; upd.rvar[l] = running[2 * l + 1];
        }
        hin = a.Hout; win = a.Wout;
    }
    if (bn_mode == MVSN_BN_BATCH_UPDATE) bn_update_running_kernel<<<8, 64, 0, st>>>(upd);
    const long long plane = (long long)hin * win;
    toplayer_kernel<<<cdiv((long long)V * plane, 128), 128, 0, st>>>(src, w[24], w[25], V, plane, feats);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // extern "C"

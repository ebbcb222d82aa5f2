// C-ABI entry points (include/mvsnerf_b200.h): argument checking, layout helpers, weight
// packing and the render dispatch.  Kernels live in the sibling .cu files.
#include "render_frontend.cuh"

namespace mvsn {

// ------------------------------------------------------------------------------------------
// error channel
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    static int cache[64] = {0};
    if (dev < 64 && cache[dev]) return cache[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev < 64) cache[dev] = n;
    return n;
}

// ------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------
__global__ void pack_images_kernel(const float* __restrict__ src, float4* __restrict__ dst, int V, int HW) {
    const long long n = (long long)V * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i / HW), p = (int)(i - (long long)v * HW);
        const float* s = src + (size_t)v * 3 * HW + p;
        dst[i] = make_float4(s[0], s[HW], s[2 * (size_t)HW], 0.f);
    }
}

// [8][nvox] <-> [nvox][8]
__global__ void vol_to_cl_kernel(const float* __restrict__ src, float4* __restrict__ dst, long long nvox) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __ldg(src + c * nvox + i);
        dst[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
        dst[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}
__global__ void vol_from_cl_kernel(const float4* __restrict__ src, float* __restrict__ dst, long long nvox) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvox; i += (long long)gridDim.x * blockDim.x) {
        float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);
        dst[0 * nvox + i] = a.x; dst[1 * nvox + i] = a.y; dst[2 * nvox + i] = a.z; dst[3 * nvox + i] = a.w;
        dst[4 * nvox + i] = b.x; dst[5 * nvox + i] = b.y; dst[6 * nvox + i] = b.z; dst[7 * nvox + i] = b.w;
    }
}

// rays [n][8] = (origin, direction, near, far) of one camera: d = directions @ c2w[:3,:3]^T, o = c2w[:3,3]
// (data/ray_utils.py:32-53 get_rays + the notebooks' cat with near / far)
__global__ void make_rays_kernel(const float* __restrict__ dirs, const float* __restrict__ c2w, float near, float far, int n,
                                 float4* __restrict__ rays) {
    __shared__ float m[12];
    if (threadIdx.x < 12) m[threadIdx.x] = __ldg(c2w + threadIdx.x);           // rows of [R | t], row-major [.,4]
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = __ldg(dirs + 3 * (size_t)i), y = __ldg(dirs + 3 * (size_t)i + 1), z = __ldg(dirs + 3 * (size_t)i + 2);
        const float dx = fmaf(z, m[2], fmaf(y, m[1], x * m[0]));
        const float dy = fmaf(z, m[6], fmaf(y, m[5], x * m[4]));
        const float dz = fmaf(z, m[10], fmaf(y, m[9], x * m[8]));
        rays[2 * (size_t)i] = make_float4(m[3], m[7], m[11], dx);
        rays[2 * (size_t)i + 1] = make_float4(dy, dz, near, far);
    }
}

// ------------------------------------------------------------------------------------------
// fp32 MLP weight image
// ------------------------------------------------------------------------------------------
struct MlpPtrs { const float* p[MVSN_N_MLP_TENSORS]; };
// indices into the 22-tensor array (see header)
enum { I_W0 = 0, I_B0 = 1, I_WB = 12, I_BB = 13, I_WV = 14, I_BV = 15, I_WF = 16, I_BF = 17,
       I_WA = 18, I_BA = 19, I_WR = 20, I_BR = 21 };

// dst[k][n] (ld = ldn) = src[n][src_col0 + k] for k < kreal, 0 for kreal <= k < kpad
__device__ void transpose_into(float* dst, int ldn, int nout, int kpad, const float* src, int src_ld,
                               int src_col0, int kreal, int tid, int nthreads) {
    for (int i = tid; i < kpad * nout; i += nthreads) {
        const int k = i / nout, n = i - k * nout;
        dst[k * ldn + n] = k < kreal ? src[(size_t)n * src_ld + src_col0 + k] : 0.f;
    }
}
__device__ void copy_into(float* dst, const float* src, int n, int npad, int tid, int nthreads) {
    for (int i = tid; i < npad; i += nthreads) dst[i] = i < n ? src[i] : 0.f;
}

__global__ void pack_mlp_fp32_kernel(MlpPtrs w, float* __restrict__ out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    using namespace w32;
    transpose_into(out + WB, 128, 128, 32, w.p[I_WB], 20, 0, 20, tid, nt);
    copy_into(out + BB, w.p[I_BB], 128, 128, tid, nt);
    transpose_into(out + W0, 128, 128, 64, w.p[I_W0], 63, 0, 63, tid, nt);
    copy_into(out + B0, w.p[I_B0], 128, 128, tid, nt);
    for (int l = 1; l <= 4; ++l) {
        transpose_into(out + W1 + (l - 1) * LSTR, 128, 128, 128, w.p[2 * l], 128, 0, 128, tid, nt);
        copy_into(out + W1 + (l - 1) * LSTR + 128 * 128, w.p[2 * l + 1], 128, 128, tid, nt);
    }
    transpose_into(out + W5, 128, 128, 64, w.p[10], 191, 0, 63, tid, nt);             // pe part
    transpose_into(out + W5 + 64 * 128, 128, 128, 128, w.p[10], 191, 63, 128, tid, nt); // h part
    copy_into(out + B5, w.p[11], 128, 128, tid, nt);
    copy_into(out + WA, w.p[I_WA], 128, 128, tid, nt);
    copy_into(out + BA, w.p[I_BA], 1, 4, tid, nt);
    transpose_into(out + WF, 128, 128, 128, w.p[I_WF], 128, 0, 128, tid, nt);
    copy_into(out + BF, w.p[I_BF], 128, 128, tid, nt);
    transpose_into(out + WV, 64, 64, 128, w.p[I_WV], 131, 0, 128, tid, nt);
    transpose_into(out + WVD, 64, 64, 4, w.p[I_WV], 131, 128, 3, tid, nt);
    copy_into(out + BV, w.p[I_BV], 64, 64, tid, nt);
    // rgb_linear [3][64] is used row-wise (dot products), no transpose
    for (int i = tid; i < 4 * 64; i += nt) out[WR + i] = i < 3 * 64 ? w.p[I_WR][i] : 0.f;
    copy_into(out + BR, w.p[I_BR], 3, 4, tid, nt);
}

static long long* g_trace = nullptr;
long long* debug_trace_buffer() { return g_trace; }     // trace builds only (conv0_tc.cu role profile)

static int make_scene(const mvsn_render_scene* s, SceneDev& d) {
    MVSN_REQUIRE(s != nullptr, MVSN_ENULL, "scene is NULL");
    MVSN_REQUIRE(s->volume_dhwc && s->imgs_hwc4 && s->mlp_packed, MVSN_ENULL, "scene has a NULL buffer");
    MVSN_REQUIRE(s->V == 3, MVSN_EBADSHAPE, "V=%d: the v0 MLP takes 8 + 4*3 feature channels", s->V);
    MVSN_REQUIRE(s->D > 0 && s->Hp > 0 && s->Wp > 0 && s->H > 1 && s->W > 1, MVSN_EBADSHAPE, "bad scene dims");
    MVSN_REQUIRE(aligned16(s->volume_dhwc) && aligned16(s->imgs_hwc4) && aligned16(s->mlp_packed), MVSN_EALIGN,
                 "scene buffers must be 16-byte aligned");
    d.vol = s->volume_dhwc; d.imgs = reinterpret_cast<const float4*>(s->imgs_hwc4);
    d.D = s->D; d.Hp = s->Hp; d.Wp = s->Wp; d.V = s->V; d.H = s->H; d.W = s->W;
    MVSN_REQUIRE(s->w2cs && s->intrinsics, MVSN_ENULL, "scene camera pointers are NULL");
    d.w2cs = s->w2cs; d.intrinsics = s->intrinsics;
    d.white_bkgd = s->white_bkgd;
    return MVSN_OK;
}

static int dispatch_render(const mvsn_render_scene* scene, const SceneDev& sc, const RenderIO& io, bool fast,
                           cudaStream_t stream) {
    switch (scene->mlp_mode) {
        case MVSN_MLP_FP32:
            return launch_render_fp32(sc, io, fast, static_cast<const float*>(scene->mlp_packed), stream);
        case MVSN_MLP_TC_HALF:
            return launch_render_tc(sc, io, fast, scene->mlp_packed, stream);
        case MVSN_MLP_TC_SPLIT:
            return launch_render_tcs(sc, io, fast, scene->mlp_packed, stream);
        case MVSN_MLP_TC_PAIR:
            return launch_render_tc2(sc, io, fast, scene->mlp_packed, stream);
        default:
            set_error("mlp_mode %d is not available in this build", scene->mlp_mode);
            return MVSN_EUNSUPPORTED;
    }
}

}  // namespace mvsn

using namespace mvsn;

extern "C" {

const char* mvsn_last_error(void) { return g_err; }
int mvsn_abi_version(void) { return 1; }
void mvsn_debug_set_trace(long long* device_buffer) { g_trace = device_buffer; }

size_t mvsn_mlp_packed_bytes(int mode) {
    switch (mode) {
        case MVSN_MLP_FP32: return (size_t)w32::TOTAL * sizeof(float);
        case MVSN_MLP_TC_HALF: return mlp_tc_packed_bytes();
        case MVSN_MLP_TC_SPLIT: return mlp_tcs_packed_bytes();
        case MVSN_MLP_TC_PAIR: return mlp_tc2_packed_bytes();
        default: return 0;
    }
}

int mvsn_mlp_pack(const float* const* w, int mode, void* packed, size_t packed_bytes, void* stream) {
    MVSN_RANGE("mvsn_mlp_pack");
    MVSN_REQUIRE(w && packed, MVSN_ENULL, "mvsn_mlp_pack: NULL argument");
    const size_t need = mvsn_mlp_packed_bytes(mode);
    MVSN_REQUIRE(need != 0, MVSN_EUNSUPPORTED, "mvsn_mlp_pack: mode %d not available", mode);
    MVSN_REQUIRE(packed_bytes >= need, MVSN_EWORKSPACE, "mvsn_mlp_pack: need %zu bytes, got %zu", need, packed_bytes);
    MVSN_REQUIRE(aligned16(packed), MVSN_EALIGN, "mvsn_mlp_pack: packed buffer must be 16-byte aligned");
    MlpPtrs p;
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i) {
        MVSN_REQUIRE(w[i] != nullptr, MVSN_ENULL, "mvsn_mlp_pack: tensor %d is NULL", i);
        p.p[i] = w[i];
    }
    if (mode == MVSN_MLP_TC_HALF) return pack_mlp_tc(w, packed, (cudaStream_t)stream);
    if (mode == MVSN_MLP_TC_SPLIT) return pack_mlp_tcs(w, packed, (cudaStream_t)stream);
    if (mode == MVSN_MLP_TC_PAIR) return pack_mlp_tc2(w, packed, (cudaStream_t)stream);
    pack_mlp_fp32_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(p, static_cast<float*>(packed));
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

int mvsn_pack_images(const float* imgs, int V, int H, int W, float* out, void* stream) {
    MVSN_REQUIRE(imgs && out, MVSN_ENULL, "mvsn_pack_images: NULL argument");
    MVSN_REQUIRE(V > 0 && H > 0 && W > 0, MVSN_EBADSHAPE, "mvsn_pack_images: bad shape");
    MVSN_REQUIRE(aligned16(out), MVSN_EALIGN, "mvsn_pack_images: output must be 16-byte aligned");
    const long long n = (long long)V * H * W;
    pack_images_kernel<<<cdiv(n, 256) < 4096 ? cdiv(n, 256) : 4096, 256, 0, (cudaStream_t)stream>>>(
        imgs, reinterpret_cast<float4*>(out), V, H * W);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

int mvsn_volume_to_channels_last(const float* src, int D, int Hp, int Wp, float* dst, void* stream) {
    MVSN_REQUIRE(src && dst, MVSN_ENULL, "mvsn_volume_to_channels_last: NULL argument");
    MVSN_REQUIRE(aligned16(dst), MVSN_EALIGN, "mvsn_volume_to_channels_last: output must be 16-byte aligned");
    const long long n = (long long)D * Hp * Wp;
    MVSN_REQUIRE(n > 0, MVSN_EBADSHAPE, "mvsn_volume_to_channels_last: bad shape");
    vol_to_cl_kernel<<<cdiv(n, 256) < 8192 ? cdiv(n, 256) : 8192, 256, 0, (cudaStream_t)stream>>>(
        src, reinterpret_cast<float4*>(dst), n);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

int mvsn_volume_from_channels_last(const float* src, int D, int Hp, int Wp, float* dst, void* stream) {
    MVSN_REQUIRE(src && dst, MVSN_ENULL, "mvsn_volume_from_channels_last: NULL argument");
    MVSN_REQUIRE(aligned16(src), MVSN_EALIGN, "mvsn_volume_from_channels_last: input must be 16-byte aligned");
    const long long n = (long long)D * Hp * Wp;
    MVSN_REQUIRE(n > 0, MVSN_EBADSHAPE, "mvsn_volume_from_channels_last: bad shape");
    vol_from_cl_kernel<<<cdiv(n, 256) < 8192 ? cdiv(n, 256) : 8192, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(src), dst, n);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

int mvsn_render_samples(const mvsn_render_scene* scene, const float* rays_pts, const float* rays_ndc,
                        const float* z_vals, const float* rays_dir, int N, int S, float* rgb, float* depth,
                        float* weights, float* alpha, float* input_feat, void* stream) {
    MVSN_RANGE("mvsn_render_samples");
    SceneDev sc;
    int rc = make_scene(scene, sc);
    if (rc) return rc;
    MVSN_REQUIRE(N >= 0 && S > 0, MVSN_EBADSHAPE, "mvsn_render_samples: N=%d S=%d", N, S);
    if (N == 0) return MVSN_OK;
    MVSN_REQUIRE(rays_pts && rays_ndc && z_vals && rays_dir && rgb && depth, MVSN_ENULL,
                 "mvsn_render_samples: NULL required pointer");
    MVSN_REQUIRE(!input_feat || aligned16(input_feat), MVSN_EALIGN, "input_feat must be 16-byte aligned");
    RenderIO io{};
    io.pts = rays_pts; io.ndc = rays_ndc; io.z = z_vals; io.dirs = rays_dir;
    io.N = N; io.S = S;
    io.rgb = rgb; io.depth = depth; io.weights = weights; io.alpha = alpha; io.input_feat = input_feat;
    io.trace = g_trace;
    return dispatch_render(scene, sc, io, false, (cudaStream_t)stream);
}

static int render_rays_impl(const mvsn_render_scene* scene, const mvsn_ray_params* rp, const float* rays,
                            const float* t_steps, int N, int S, float* rgb, float* depth, float* weights,
                            float* alpha, float* input_feat, const mvsn_peer_sink* sink, void* stream) {
    MVSN_RANGE("mvsn_render_rays");
    SceneDev sc;
    int rc = make_scene(scene, sc);
    if (rc) return rc;
    MVSN_REQUIRE(rp != nullptr, MVSN_ENULL, "mvsn_render_rays: ray params NULL");
    MVSN_REQUIRE(N >= 0 && S > 0, MVSN_EBADSHAPE, "mvsn_render_rays: N=%d S=%d", N, S);
    if (N == 0) return MVSN_OK;
    MVSN_REQUIRE(rays && t_steps && (sink || (rgb && depth)), MVSN_ENULL, "mvsn_render_rays: NULL required pointer");
    MVSN_REQUIRE(aligned16(rays), MVSN_EALIGN, "rays must be 16-byte aligned");
    MVSN_REQUIRE(!input_feat || aligned16(input_feat), MVSN_EALIGN, "input_feat must be 16-byte aligned");
    RenderIO io{};
    io.rays = rays; io.t_steps = t_steps;
    io.N = N; io.S = S;
    io.rgb = rgb; io.depth = depth; io.weights = weights; io.alpha = alpha; io.input_feat = input_feat;
    // host-side scalars exactly as utils.get_ndc_coordinate forms them (python floats -> fp32)
    io.rg.near = rp->ndc_near;
    io.rg.far_minus_near = (float)((double)rp->ndc_far - (double)rp->ndc_near);
    io.rg.inv_near = (float)(1.0 / (double)rp->ndc_near);
    io.rg.inv_far_minus_inv_near = (float)(1.0 / (double)rp->ndc_far - 1.0 / (double)rp->ndc_near);
    io.rg.pad = rp->pad;
    io.rg.wf = (float)scene->W / 4.0f;     // (inv_scale + 1) / 4, utils.py:139
    io.rg.hf = (float)scene->H / 4.0f;
    io.rg.lindisp = rp->lindisp;
    io.trace = g_trace;
    if (sink) {
        MVSN_REQUIRE(sink->n_peers >= 1 && sink->n_peers <= MVSN_MAX_PEERS, MVSN_EBADSHAPE,
                     "peer sink: n_peers=%d (1..%d)", sink->n_peers, MVSN_MAX_PEERS);
        MVSN_REQUIRE(sink->first_pixel >= 0, MVSN_EBADSHAPE, "peer sink: first_pixel < 0");
        for (int p = 0; p < sink->n_peers; ++p) {
            MVSN_REQUIRE(sink->frame[p] != nullptr, MVSN_ENULL, "peer sink: frame[%d] is NULL", p);
            MVSN_REQUIRE(aligned16(sink->frame[p]), MVSN_EALIGN, "peer sink: frame[%d] must be 16-byte aligned", p);
            io.sink[p] = reinterpret_cast<float4*>(sink->frame[p]);
        }
        io.n_sink = sink->n_peers;
        io.sink_first = sink->first_pixel;
    }
    return dispatch_render(scene, sc, io, true, (cudaStream_t)stream);
}

int mvsn_render_rays(const mvsn_render_scene* scene, const mvsn_ray_params* rp, const float* rays,
                     const float* t_steps, int N, int S, float* rgb, float* depth, float* weights,
                     float* alpha, float* input_feat, void* stream) {
    return render_rays_impl(scene, rp, rays, t_steps, N, S, rgb, depth, weights, alpha, input_feat, nullptr, stream);
}

int mvsn_render_rays_to_peers(const mvsn_render_scene* scene, const mvsn_ray_params* rp, const float* rays,
                              const float* t_steps, int N, int S, const mvsn_peer_sink* sink, float* rgb,
                              float* depth, void* stream) {
    MVSN_REQUIRE(sink != nullptr, MVSN_ENULL, "mvsn_render_rays_to_peers: sink is NULL");
    return render_rays_impl(scene, rp, rays, t_steps, N, S, rgb, depth, nullptr, nullptr, nullptr, sink, stream);
}

int mvsn_make_rays(const float* directions, const float* c2w, float near, float far, int n, float* rays, void* stream) {
    MVSN_RANGE("mvsn_make_rays");
    MVSN_REQUIRE(directions && c2w && rays, MVSN_ENULL, "mvsn_make_rays: NULL argument");
    MVSN_REQUIRE(n >= 0, MVSN_EBADSHAPE, "mvsn_make_rays: n=%d", n);
    MVSN_REQUIRE(aligned16(rays), MVSN_EALIGN, "mvsn_make_rays: rays must be 16-byte aligned");
    if (n == 0) return MVSN_OK;
    make_rays_kernel<<<cdiv(n, 256) < 2048 ? cdiv(n, 256) : 2048, 256, 0, (cudaStream_t)stream>>>(
        directions, c2w, near, far, n, reinterpret_cast<float4*>(rays));
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

// ---- fine-tuning step ------------------------------------------------------------------------------------
size_t mvsn_render_backward_workspace_bytes(int N, int S) { return render_backward_workspace_bytes(N, S); }

int mvsn_render_backward(const mvsn_render_scene* scene, const float* const* mlp_w, const float* rays_pts,
                         const float* rays_ndc, const float* z_vals, const float* rays_dir, int N, int S,
                         const mvsn_render_grads* g, float* const* grad_mlp, float* grad_volume_dhwc, void* workspace,
                         size_t workspace_bytes, void* stream) {
    MVSN_RANGE("mvsn_render_backward");
    SceneDev sc;
    int rc = make_scene(scene, sc);
    if (rc) return rc;
    MVSN_REQUIRE(scene->mlp_mode == MVSN_MLP_FP32, MVSN_EUNSUPPORTED,
                 "mvsn_render_backward: scene->mlp_packed must be the MVSN_MLP_FP32 image (mode %d given)", scene->mlp_mode);
    MVSN_REQUIRE(N >= 0 && S > 0, MVSN_EBADSHAPE, "mvsn_render_backward: N=%d S=%d", N, S);
    MVSN_REQUIRE(g && mlp_w && grad_mlp, MVSN_ENULL, "mvsn_render_backward: NULL argument");
    MVSN_REQUIRE(g->rgb || g->target_rgb, MVSN_ENULL, "mvsn_render_backward: neither g->rgb nor g->target_rgb given");
    for (int i = 0; i < MVSN_N_MLP_TENSORS; ++i)
        MVSN_REQUIRE(mlp_w[i] && grad_mlp[i], MVSN_ENULL, "mvsn_render_backward: tensor %d is NULL", i);
    MVSN_REQUIRE(!grad_volume_dhwc || aligned16(grad_volume_dhwc), MVSN_EALIGN, "grad_volume_dhwc must be 16-byte aligned");
    if (N == 0) return MVSN_OK;
    MVSN_REQUIRE(rays_pts && rays_ndc && z_vals && rays_dir, MVSN_ENULL, "mvsn_render_backward: NULL required pointer");
    RenderIO io{};
    io.pts = rays_pts; io.ndc = rays_ndc; io.z = z_vals; io.dirs = rays_dir;
    io.N = N; io.S = S;
    return launch_render_backward(sc, io, static_cast<const float*>(scene->mlp_packed), mlp_w, g->rgb, g->target_rgb,
                                  g->loss_scale, g->depth, g->weights, g->alpha, g->input_feat, grad_mlp,
                                  grad_volume_dhwc, g->rgb_out, g->depth_out, g->loss_out, workspace, workspace_bytes,
                                  (cudaStream_t)stream);
}

int mvsn_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const int* numel_host, int count, float lr, float beta1, float beta2, float eps, int step,
                   void* stream) {
    MVSN_RANGE("mvsn_adam_step");
    MVSN_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel_host, MVSN_ENULL, "mvsn_adam_step: NULL argument");
    MVSN_REQUIRE(step >= 1, MVSN_EBADSHAPE, "mvsn_adam_step: step=%d (counts from 1)", step);
    return launch_adam_tensors(params, grads, exp_avg, exp_avg_sq, numel_host, count, lr, beta1, beta2, eps, step,
                               (cudaStream_t)stream);
}

int mvsn_adam_step_volume(float* param, float* grad_dhwc, float* exp_avg, float* exp_avg_sq, long long nvox,
                          int planar, float lr, float beta1, float beta2, float eps, int step, void* stream) {
    MVSN_RANGE("mvsn_adam_step_volume");
    MVSN_REQUIRE(param && grad_dhwc && exp_avg && exp_avg_sq, MVSN_ENULL, "mvsn_adam_step_volume: NULL argument");
    MVSN_REQUIRE(nvox > 0 && step >= 1, MVSN_EBADSHAPE, "mvsn_adam_step_volume: nvox=%lld step=%d", nvox, step);
    MVSN_REQUIRE(aligned16(grad_dhwc) && (planar || (aligned16(param) && aligned16(exp_avg) && aligned16(exp_avg_sq))),
                 MVSN_EALIGN, "mvsn_adam_step_volume: buffers must be 16-byte aligned");
    return launch_adam_volume(param, grad_dhwc, exp_avg, exp_avg_sq, nvox, planar, lr, beta1, beta2, eps, step,
                              (cudaStream_t)stream);
}

// ---- exportable frame buffers (CUDA IPC): the one allocation this library makes --------------------------
int mvsn_peer_buffer_create(size_t bytes, void** dev_ptr, unsigned char* handle_host) {
    MVSN_REQUIRE(dev_ptr && handle_host && bytes > 0, MVSN_ENULL, "mvsn_peer_buffer_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MVSN_PEER_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    MVSN_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return MVSN_ECUDA;
    }
    memcpy(handle_host, &h, sizeof(h));
    *dev_ptr = p;
    return MVSN_OK;
}

int mvsn_peer_buffer_open(const unsigned char* handle_host, void** peer_ptr) {
    MVSN_REQUIRE(handle_host && peer_ptr, MVSN_ENULL, "mvsn_peer_buffer_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    MVSN_CUDA_CHECK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MVSN_OK;
}

int mvsn_peer_buffer_close(void* peer_ptr) {
    if (peer_ptr) MVSN_CUDA_CHECK(cudaIpcCloseMemHandle(peer_ptr));
    return MVSN_OK;
}

int mvsn_peer_buffer_destroy(void* dev_ptr) {
    if (dev_ptr) MVSN_CUDA_CHECK(cudaFree(dev_ptr));
    return MVSN_OK;
}

}  // extern "C"

// Per-sample front end of the fused render kernel: everything the reference does between
// "a ray" and "the 86-channel MLP input" (renderer.py:138-165 and callees), as device functions
// shared by the fp32 and the tcgen05 render kernels.
#pragma once
#include "common.cuh"

namespace mvsn {

struct SceneDev {
    const float* vol;      // [D,Hp,Wp,8]
    const float4* imgs;    // [V,H,W] texels (r,g,b,0)
    int D, Hp, Wp;
    int V, H, W;
    const float* w2cs;     // [V,4,4] device
    const float* intrinsics;  // [V,3,3] device
    int white_bkgd;
};

struct Cams {              // staged in shared memory by every render kernel
    float w2c[3][12];
    float K[3][9];
};

__device__ __forceinline__ void load_cams(const SceneDev& sc, Cams* cams, int tid) {
    if (tid < 36) cams->w2c[tid / 12][tid % 12] = __ldg(sc.w2cs + (tid / 12) * 16 + tid % 12);
    else if (tid < 63) { const int i = tid - 36; cams->K[i / 9][i % 9] = __ldg(sc.intrinsics + i); }
}

struct RayGenDev {
    float near, far_minus_near;      // z normalisation of the reference camera (utils.py:128-131)
    float inv_near, inv_far_minus_inv_near;
    float pad, wf, hf;               // utils.py:138-143
    int lindisp;
};

struct RenderIO {
    // signature-compatible inputs (FAST == false)
    const float* pts; const float* ndc; const float* z; const float* dirs;
    // fused-caller inputs (FAST == true)
    const float* rays; const float* t_steps;
    RayGenDev rg;
    int N, S;
    int rays_per_tile;     // tensor-core kernel: chosen by its launcher
    float* rgb; float* depth; float* weights; float* alpha; float* input_feat;
    long long* trace;      // debug timeline (mvsn_debug_set_trace), null in normal operation
    // multi-GPU frame sink (mvsn_render_rays_to_peers): every finished pixel is stored as one 16-byte
    // (r, g, b, depth) texel into EVERY rank's copy of the assembled frame, straight from the compositing
    // epilogue over NVLink peer mappings -- the frame is complete on all ranks when the kernels are, with no
    // gather pass.  n_sink == 0: off.
    float4* sink[MVSN_MAX_PEERS];
    int n_sink;
    long long sink_first;  // frame index of this launch's ray 0
};

// final pixel of ray `ray`: the caller's rgb / depth arrays (either may be null when a sink is given) and the sink
__device__ __forceinline__ void store_pixel(const RenderIO& io, int ray, float r, float g, float b, float d) {
    if (io.rgb) { io.rgb[(size_t)ray * 3 + 0] = r; io.rgb[(size_t)ray * 3 + 1] = g; io.rgb[(size_t)ray * 3 + 2] = b; }
    if (io.depth) io.depth[ray] = d;
    if (io.n_sink > 0) {
        const float4 px = make_float4(r, g, b, d);
        const long long at = io.sink_first + ray;
#pragma unroll 1
        for (int p = 0; p < io.n_sink; ++p) io.sink[p][at] = px;
    }
}

int launch_render_fp32(const SceneDev& sc, const RenderIO& io, bool fast, const float* wts, cudaStream_t stream);
int launch_render_tc(const SceneDev& sc, const RenderIO& io, bool fast, const void* wimg, cudaStream_t stream);
size_t mlp_tc_packed_bytes();
int pack_mlp_tc(const float* const* w, void* packed, cudaStream_t stream);
int launch_render_tcs(const SceneDev& sc, const RenderIO& io, bool fast, const void* wimg, cudaStream_t stream);
size_t mlp_tcs_packed_bytes();
int pack_mlp_tcs(const float* const* w, void* packed, cudaStream_t stream);
int launch_render_tc2(const SceneDev& sc, const RenderIO& io, bool fast, const void* wimg, cudaStream_t stream);
// fine-tuning step (render_bwd.cu)
size_t render_backward_workspace_bytes(int N, int S);
int launch_render_backward(const SceneDev& sc, const RenderIO& io, const float* wts_fp32, const float* const* mlp_w,
                           const float* g_rgb, const float* target, float inv_count, const float* g_depth,
                           const float* g_weights, const float* g_alpha, const float* g_feat, float* const* grad_mlp,
                           float* dvol, float* rgb_out, float* depth_out, float* loss, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream);
int launch_adam_tensors(float* const* p, const float* const* g, float* const* m, float* const* v, const int* n, int count,
                        float lr, float beta1, float beta2, float eps, int step, cudaStream_t stream);
int launch_adam_volume(float* p, float* g_dhwc, float* m, float* v, long long nvox, int planar, float lr, float beta1,
                       float beta2, float eps, int step, cudaStream_t stream);
size_t mlp_tc2_packed_bytes();
int pack_mlp_tc2(const float* const* w, void* packed, cudaStream_t stream);

// cam = R p + t ; pix = K cam ; (u, v) = pix.xy / pix.z / (W-1, H-1)     utils.py:120-127
template <bool PRECISE>
__device__ __forceinline__ float fdiv(float a, float b) { return PRECISE ? __fdiv_rn(a, b) : __fdividef(a, b); }

// PRECISE = IEEE divisions exactly where the reference divides (fp32-parity modes); otherwise
// MUFU.RCP-based divisions (2 ulp), plenty for the 16-bit-operand mode and a fraction of the code.
template <bool PRECISE = true>
__device__ __forceinline__ void project_view(const float* __restrict__ w2c, const float* __restrict__ K,
                                             float px, float py, float pz, float wm1, float hm1,
                                             float& u, float& v, float& zc) {
    float cx = fmaf(pz, w2c[2], fmaf(py, w2c[1], px * w2c[0])) + w2c[3];
    float cy = fmaf(pz, w2c[6], fmaf(py, w2c[5], px * w2c[4])) + w2c[7];
    float cz = fmaf(pz, w2c[10], fmaf(py, w2c[9], px * w2c[8])) + w2c[11];
    float qx = fmaf(cz, K[2], fmaf(cy, K[1], cx * K[0]));
    float qy = fmaf(cz, K[5], fmaf(cy, K[4], cx * K[3]));
    float qz = fmaf(cz, K[8], fmaf(cy, K[7], cx * K[6]));
    if (PRECISE) {
        u = __fdiv_rn(__fdiv_rn(qx, qz), wm1);
        v = __fdiv_rn(__fdiv_rn(qy, qz), hm1);
    } else {
        const float iz = __fdividef(1.f, qz);
        u = qx * iz * __fdividef(1.f, wm1);
        v = qy * iz * __fdividef(1.f, hm1);
    }
    zc = qz;
}

// utils.get_ndc_coordinate for the reference camera (utils.py:112-146)
template <bool PRECISE = true>
__device__ __forceinline__ void ndc_of_point(const SceneDev& sc, const Cams& cams, const RayGenDev& rg,
                                             float px, float py, float pz,
                                             float& nx, float& ny, float& nz) {
    float u, v, zc;
    project_view<PRECISE>(cams.w2c[0], cams.K[0], px, py, pz, (float)(sc.W - 1), (float)(sc.H - 1), u, v, zc);
    if (!rg.lindisp) nz = fdiv<PRECISE>(zc - rg.near, rg.far_minus_near);
    else             nz = fdiv<PRECISE>(fdiv<PRECISE>(1.0f, zc) - rg.inv_near, rg.inv_far_minus_inv_near);
    if (rg.pad > 0.f) {
        float dh = rg.hf + rg.pad * 2.f, dw = rg.wf + rg.pad * 2.f;
        v = __fadd_rn(fdiv<PRECISE>(__fmul_rn(v, rg.hf), dh), fdiv<PRECISE>(rg.pad, dh));
        u = __fadd_rn(fdiv<PRECISE>(__fmul_rn(u, rg.wf), dw), fdiv<PRECISE>(rg.pad, dw));
    }
    nx = u; ny = v;
}

// utils.index_point_feature (utils.py:357-383): trilinear, zeros padding, align_corners=True.
// All sixteen 16-byte loads are issued unconditionally (indices clamped, out-of-volume corners get
// weight 0) so they are in flight together instead of one DRAM latency per corner.
__device__ __forceinline__ void sample_volume(const SceneDev& sc, float nx, float ny, float nz, float* out8) {
    const int W = sc.Wp, H = sc.Hp, D = sc.D;
    float ix = ((nx * 2.f - 1.f + 1.f) * 0.5f) * (float)(W - 1);
    float iy = ((ny * 2.f - 1.f + 1.f) * 0.5f) * (float)(H - 1);
    float iz = ((nz * 2.f - 1.f + 1.f) * 0.5f) * (float)(D - 1);
    float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    float wx[2] = {(x0f + 1.f) - ix, ix - x0f}, wy[2] = {(y0f + 1.f) - iy, iy - y0f}, wz[2] = {(z0f + 1.f) - iz, iz - z0f};
    // clamp before the int conversion so absurd coordinates cannot overflow
    const int x0 = (int)fminf(fmaxf(x0f, -2.f), (float)W);
    const int y0 = (int)fminf(fmaxf(y0f, -2.f), (float)H);
    const int z0 = (int)fminf(fmaxf(z0f, -2.f), (float)D);
    int xo[2], yo[2], zo[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const int x = x0 + d, y = y0 + d, z = z0 + d;
        if ((unsigned)x >= (unsigned)W) wx[d] = 0.f;       // zeros padding: the tap contributes nothing
        if ((unsigned)y >= (unsigned)H) wy[d] = 0.f;
        if ((unsigned)z >= (unsigned)D) wz[d] = 0.f;
        xo[d] = min(max(x, 0), W - 1); yo[d] = min(max(y, 0), H - 1); zo[d] = min(max(z, 0), D - 1);
    }
    float4 va[8], vb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4* p = reinterpret_cast<const float4*>(
            sc.vol + (((size_t)zo[c >> 2] * H + yo[(c >> 1) & 1]) * W + xo[c & 1]) * 8);
        va[c] = __ldg(p); vb[c] = __ldg(p + 1);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) out8[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {                          // accumulation order as aten: x fastest, then y, then z
        const float wgt = wx[c & 1] * wy[(c >> 1) & 1] * wz[c >> 2];
        out8[0] = fmaf(va[c].x, wgt, out8[0]); out8[1] = fmaf(va[c].y, wgt, out8[1]);
        out8[2] = fmaf(va[c].z, wgt, out8[2]); out8[3] = fmaf(va[c].w, wgt, out8[3]);
        out8[4] = fmaf(vb[c].x, wgt, out8[4]); out8[5] = fmaf(vb[c].y, wgt, out8[5]);
        out8[6] = fmaf(vb[c].z, wgt, out8[6]); out8[7] = fmaf(vb[c].w, wgt, out8[7]);
    }
}

// Same arithmetic as sample_volume, one z-plane (8 x 16-byte loads) in flight at a time: for callers that are
// not latency-critical and are short of registers (the producer warps of render_tc2.cu).
__device__ __forceinline__ void sample_volume_2pass(const SceneDev& sc, float nx, float ny, float nz, float* out8) {
    const int W = sc.Wp, H = sc.Hp, D = sc.D;
    float ix = ((nx * 2.f - 1.f + 1.f) * 0.5f) * (float)(W - 1);
    float iy = ((ny * 2.f - 1.f + 1.f) * 0.5f) * (float)(H - 1);
    float iz = ((nz * 2.f - 1.f + 1.f) * 0.5f) * (float)(D - 1);
    float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    float wx[2] = {(x0f + 1.f) - ix, ix - x0f}, wy[2] = {(y0f + 1.f) - iy, iy - y0f}, wz[2] = {(z0f + 1.f) - iz, iz - z0f};
    const int x0 = (int)fminf(fmaxf(x0f, -2.f), (float)W);
    const int y0 = (int)fminf(fmaxf(y0f, -2.f), (float)H);
    const int z0 = (int)fminf(fmaxf(z0f, -2.f), (float)D);
    int xo[2], yo[2], zo[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const int x = x0 + d, y = y0 + d, z = z0 + d;
        if ((unsigned)x >= (unsigned)W) wx[d] = 0.f;
        if ((unsigned)y >= (unsigned)H) wy[d] = 0.f;
        if ((unsigned)z >= (unsigned)D) wz[d] = 0.f;
        xo[d] = min(max(x, 0), W - 1); yo[d] = min(max(y, 0), H - 1); zo[d] = min(max(z, 0), D - 1);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) out8[c] = 0.f;
#pragma unroll 1
    for (int zz = 0; zz < 2; ++zz) {
        float4 va[4], vb[4];
        const int zsel = zz ? zo[1] : zo[0];
        const float wzsel = zz ? wz[1] : wz[0];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4* p = reinterpret_cast<const float4*>(
                sc.vol + (((size_t)zsel * H + yo[(c >> 1) & 1]) * W + xo[c & 1]) * 8);
            va[c] = __ldg(p); vb[c] = __ldg(p + 1);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {                      // same accumulation order: x fastest, then y, then z
            const float wgt = wx[c & 1] * wy[(c >> 1) & 1] * wzsel;
            out8[0] = fmaf(va[c].x, wgt, out8[0]); out8[1] = fmaf(va[c].y, wgt, out8[1]);
            out8[2] = fmaf(va[c].z, wgt, out8[2]); out8[3] = fmaf(va[c].w, wgt, out8[3]);
            out8[4] = fmaf(vb[c].x, wgt, out8[4]); out8[5] = fmaf(vb[c].y, wgt, out8[5]);
            out8[6] = fmaf(vb[c].z, wgt, out8[6]); out8[7] = fmaf(vb[c].w, wgt, out8[7]);
        }
    }
}

// utils.build_color_volume (utils.py:300-332): bilinear, BORDER padding, strict in-bounds mask.
// out4 = (r, g, b, mask)
template <bool PRECISE = true>
__device__ __forceinline__ void sample_color(const SceneDev& sc, const Cams& cams, int v, float px, float py, float pz, float* out4) {
    const int W = sc.W, H = sc.H;
    float u, vv, zc;
    project_view<PRECISE>(cams.w2c[v], cams.K[v], px, py, pz, (float)(W - 1), (float)(H - 1), u, vv, zc);
    float gx = u * 2.f - 1.f, gy = vv * 2.f - 1.f;
    float ix = ((gx + 1.f) * 0.5f) * (float)(W - 1);
    float iy = ((gy + 1.f) * 0.5f) * (float)(H - 1);
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));    // border: clip the source index
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    float x0f = floorf(ix), y0f = floorf(iy);
    float wx1 = ix - x0f, wy1 = iy - y0f, wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
    int x0 = (int)x0f, y0 = (int)y0f;
    if (!(ix == ix)) { x0 = 0; }   // NaN coordinate: keep addresses legal; weights propagate NaN
    if (!(iy == iy)) { y0 = 0; }
    int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);   // the clipped tap always has weight 0
    const float4* img = sc.imgs + (size_t)v * H * W;
    float4 nw = __ldg(img + (size_t)y0 * W + x0), ne = __ldg(img + (size_t)y0 * W + x1);
    float4 sw = __ldg(img + (size_t)y1 * W + x0), se = __ldg(img + (size_t)y1 * W + x1);
    float a = wx0 * wy0, b = wx1 * wy0, c = wx0 * wy1, d = wx1 * wy1;
    out4[0] = fmaf(se.x, d, fmaf(sw.x, c, fmaf(ne.x, b, nw.x * a)));
    out4[1] = fmaf(se.y, d, fmaf(sw.y, c, fmaf(ne.y, b, nw.y * a)));
    out4[2] = fmaf(se.z, d, fmaf(sw.z, c, fmaf(ne.z, b, nw.z * a)));
    out4[3] = (gx > -1.f && gx < 1.f && gy > -1.f && gy < 1.f) ? 1.f : 0.f;
}

// gen_dir_feature (renderer.py:111-122,142-147): unit direction in the reference camera frame
template <bool PRECISE = true>
__device__ __forceinline__ void view_dir(const Cams& cams, float dx, float dy, float dz, float* out3) {
    if (PRECISE) {
        float n = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        dx = __fdiv_rn(dx, n); dy = __fdiv_rn(dy, n); dz = __fdiv_rn(dz, n);
    } else {
        const float rn = rsqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        dx *= rn; dy *= rn; dz *= rn;
    }
    const float* R = cams.w2c[0];
    out3[0] = fmaf(dz, R[2], fmaf(dy, R[1], dx * R[0]));
    out3[1] = fmaf(dz, R[6], fmaf(dy, R[5], dx * R[4]));
    out3[2] = fmaf(dz, R[10], fmaf(dy, R[9], dx * R[8]));
}

}  // namespace mvsn

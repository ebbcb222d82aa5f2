// K-A: plane-sweep cost volume (MVSNet.build_volume_costvar_img, models.py:839-893, with
// utils.homo_warp, utils.py:580-630).
//
// One pass, one thread per voxel (x fastest, so every channel plane is written with fully
// coalesced 128-byte stores): the homography of the two source views is evaluated in registers,
// the 32-channel feature maps and the quarter-resolution RGB are gathered from channel-QUAD staging
// copies ([C/4][h][w][4]: one 16-byte load brings 4 channels of a tap, and neighbouring lanes still read
// neighbouring pixels = contiguous bytes; 8.6 MB total, L2 resident), and mean/variance over the
// visible views are formed on the fly.  None of the reference's 600 MB temporaries (warped volumes, sum, sum of squares,
// the replicated reference volume, the sampling grids) exists.  HBM traffic = the 41-channel
// result (+ the optional masks): 176 B / voxel.
#include "common.cuh"

namespace mvsn {

struct CostArgs {
    const float4* small;    // [V][h][w] (r,g,b,0) quarter-resolution normalised images
    const float4* feats;    // [V][8][h][w] channel quads of the FeatureNet output
    const float* proj;      // [V][3][4] device
    const float* depths;
    int V, h, w, D, pad;
    float* cost;            // [41][D][hp][wp]
    float* masks;           // [V][D][hp][wp] or null
};

// F.interpolate(imgs, (h, w), mode='bilinear', align_corners=False)  (models.py:859), written as (r,g,b,0) texels
__global__ void downsample_images_kernel(const float* __restrict__ imgs, float4* __restrict__ out,
                                         int V, int H, int W, int h, int w) {
    const int n = V * h * w;
    const float sy = (float)H / (float)h, sx = (float)W / (float)w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = i / (h * w), r = i - v * h * w, y = r / w, x = r - y * w;
        float fy = fmaxf(sy * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)x + 0.5f) - 0.5f, 0.f);
        int y0 = (int)fy, x0 = (int)fx;
        int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        float ly = fy - (float)y0, lx = fx - (float)x0;
        float c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float* p = imgs + (size_t)(v * 3 + k) * H * W;
            float top = (1.f - lx) * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1];
            float bot = (1.f - lx) * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1];
            c[k] = (1.f - ly) * top + ly * bot;
        }
        out[i] = make_float4(c[0], c[1], c[2], 0.f);
    }
}

// feats [V][32][h][w] (reference layout) -> [V][8][h][w][4]
__global__ void feats_to_quads_kernel(const float* __restrict__ feats, float4* __restrict__ out, int V, int hw) {
    const long long n = (long long)V * 8 * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long vq = i / hw;
        const int px = (int)(i - vq * hw);
        const float* p = feats + (size_t)vq * 4 * hw + px;
        out[i] = make_float4(__ldg(p), __ldg(p + hw), __ldg(p + 2 * (size_t)hw), __ldg(p + 3 * (size_t)hw));
    }
}

struct Taps {              // bilinear taps of one source view for one voxel (zeros padding)
    int off[4];            // pixel index (y*w + x) of nw, ne, sw, se (clamped; weight 0 when outside)
    float wgt[4];
    float mask;
};

__device__ __forceinline__ Taps make_taps(const float* __restrict__ P, float xr, float yr, int h, int w) {
    // q = R (x, y, 1)^T + T / depth           utils.py:612 ; P[3], P[7], P[11] already hold T / depth of this plane
    float q0 = fmaf(P[1], yr, P[0] * xr) + P[2] + P[3];
    float q1 = fmaf(P[5], yr, P[4] * xr) + P[6] + P[7];
    float q2 = fmaf(P[9], yr, P[8] * xr) + P[10] + P[11];
    float u = __fdiv_rn(q0, q2), v = __fdiv_rn(q1, q2);                       // :617
    float gx = __fdiv_rn(u, (float)(((double)w - 1.0) / 2.0)) - 1.f;           // :619-620
    float gy = __fdiv_rn(v, (float)(((double)h - 1.0) / 2.0)) - 1.f;
    Taps t;
    t.mask = (gx > -1.f && gx < 1.f && gy > -1.f && gy < 1.f) ? 1.f : 0.f;     // models.py:874-877
    float ix = ((gx + 1.f) * 0.5f) * (float)(w - 1);
    float iy = ((gy + 1.f) * 0.5f) * (float)(h - 1);
    float x0f = floorf(ix), y0f = floorf(iy);
    float wx1 = ix - x0f, wy1 = iy - y0f, wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
    int x0 = (int)fminf(fmaxf(x0f, -2.f), (float)w), y0 = (int)fminf(fmaxf(y0f, -2.f), (float)h);
    if (!(ix == ix) || !(iy == iy)) { x0 = -2; y0 = -2; }                       // NaN: no contribution
    const bool vx0 = (unsigned)x0 < (unsigned)w, vx1 = (unsigned)(x0 + 1) < (unsigned)w;
    const bool vy0 = (unsigned)y0 < (unsigned)h, vy1 = (unsigned)(y0 + 1) < (unsigned)h;
    // every tap is loaded unconditionally (clamped address) with weight 0 when it falls outside:
    // the eight 16-byte loads per channel group are then in flight together
    const int xc0 = min(max(x0, 0), w - 1), xc1 = min(max(x0 + 1, 0), w - 1);
    const int yc0 = min(max(y0, 0), h - 1), yc1 = min(max(y0 + 1, 0), h - 1);
    t.off[0] = yc0 * w + xc0; t.off[1] = yc0 * w + xc1; t.off[2] = yc1 * w + xc0; t.off[3] = yc1 * w + xc1;
    t.wgt[0] = (vx0 && vy0) ? wx0 * wy0 : 0.f; t.wgt[1] = (vx1 && vy0) ? wx1 * wy0 : 0.f;
    t.wgt[2] = (vx0 && vy1) ? wx0 * wy1 : 0.f; t.wgt[3] = (vx1 && vy1) ? wx1 * wy1 : 0.f;
    return t;
}

// Thread per voxel, x fastest, grid = (plane tiles, depth planes).  For one tap the 32 lanes of a warp read
// neighbouring source pixels (the homography is locally affine) = contiguous 16-byte texels of the channel-quad
// copies, and every output channel plane is written with fully coalesced stores.
// 124 registers, no spills (at 3 blocks per SM the compiler spills 11 values that the quad loop reloads every
// iteration: 24 % of the L1 wavefronts; measured equal within noise, so the spill-free form is kept)
__global__ void __launch_bounds__(256, 2)
cost_volume_kernel(const CostArgs a) {
    const int hp = a.h + 2 * a.pad, wp = a.w + 2 * a.pad;
    const long long plane = (long long)hp * wp, nvox = plane * a.D;
    const int hw = a.h * a.w;
    // grid = (plane tiles, D): one voxel per thread, the depth index is the block's y coordinate (no 64-bit division)
    const int d = blockIdx.y;
    // the block's depth plane is fixed, so the translation column is divided by the depth once per block (same IEEE
    // division, same operands as the per-voxel form: bit-identical) instead of six times per voxel
    __shared__ float s_proj[36];
    if (threadIdx.x < 36) {
        float p = __ldg(a.proj + threadIdx.x);
        if ((threadIdx.x & 3) == 3) p = __fdiv_rn(p, __ldg(a.depths + d));
        s_proj[threadIdx.x] = p;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < (int)plane) {
        const long long i = (long long)d * plane + r;
        const int yp = r / wp, xp = r - yp * wp;
        const int y = yp - a.pad, x = xp - a.pad;
        const bool interior = (unsigned)y < (unsigned)a.h && (unsigned)x < (unsigned)a.w;
        const int ref_off = interior ? y * a.w + x : 0;
        float* out = a.cost + i;

        Taps t[2];
        float nvis = 1.f;
        if (a.masks) a.masks[i] = 1.f;
#pragma unroll
        for (int v = 1; v < 3; ++v) {
            t[v - 1] = make_taps(s_proj + 12 * v, (float)x, (float)y, a.h, a.w);
            nvis += t[v - 1].mask;
            if (a.masks) a.masks[(size_t)v * nvox + i] = t[v - 1].mask;
        }
        // 1 / (number of visible views), models.py:889: the count is 1, 2 or 3, so the IEEE quotient is one of three constants
        const float inv_n = nvis == 1.f ? 1.f : (nvis == 2.f ? 0.5f : 0.333333343267440796f);

        // the four taps of a 4-channel texel, combined per channel in the reference's tap order (nw, ne, sw, se)
        auto gather4 = [&](const float4* __restrict__ p, const Taps& tp) -> float4 {
            const float4 a0 = __ldg(p + tp.off[0]), a1 = __ldg(p + tp.off[1]);
            const float4 a2 = __ldg(p + tp.off[2]), a3 = __ldg(p + tp.off[3]);
            float4 r;
            r.x = fmaf(a3.x, tp.wgt[3], fmaf(a2.x, tp.wgt[2], fmaf(a1.x, tp.wgt[1], a0.x * tp.wgt[0])));
            r.y = fmaf(a3.y, tp.wgt[3], fmaf(a2.y, tp.wgt[2], fmaf(a1.y, tp.wgt[1], a0.y * tp.wgt[0])));
            r.z = fmaf(a3.z, tp.wgt[3], fmaf(a2.z, tp.wgt[2], fmaf(a1.z, tp.wgt[1], a0.z * tp.wgt[0])));
            r.w = fmaf(a3.w, tp.wgt[3], fmaf(a2.w, tp.wgt[2], fmaf(a1.w, tp.wgt[1], a0.w * tp.wgt[0])));
            return r;
        };
        // channels 0:9 -- reference RGB (interior only; the border is defined as zero, F5) and warped source RGB
        {
            const float4 rr = interior ? __ldg(a.small + ref_off) : make_float4(0.f, 0.f, 0.f, 0.f);
            out[0] = rr.x; out[(size_t)nvox] = rr.y; out[(size_t)2 * nvox] = rr.z;
#pragma unroll
            for (int v = 1; v < 3; ++v) {
                const float4 c = gather4(a.small + (size_t)v * hw, t[v - 1]);                 // models.py:872
                out[(size_t)(3 * v) * nvox] = c.x; out[(size_t)(3 * v + 1) * nvox] = c.y; out[(size_t)(3 * v + 2) * nvox] = c.z;
            }
        }
        // channels 9:41 -- variance of the 32 feature channels over {ref, warped src 1, 2}, four channels at a time
#pragma unroll 2
        for (int q = 0; q < 8; ++q) {
            const float4 rv4 = interior ? __ldg(a.feats + (size_t)q * hw + ref_off) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 w1 = gather4(a.feats + (size_t)(8 + q) * hw, t[0]);
            const float4 w2 = gather4(a.feats + (size_t)(16 + q) * hw, t[1]);
            const float rv[4] = {rv4.x, rv4.y, rv4.z, rv4.w};
            const float wa[4] = {w1.x, w1.y, w1.z, w1.w}, wb[4] = {w2.x, w2.y, w2.z, w2.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float s1 = rv[k], s2 = __fmul_rn(rv[k], rv[k]);
                s1 = __fadd_rn(s1, wa[k]); s2 = __fadd_rn(s2, __fmul_rn(wa[k], wa[k]));
                s1 = __fadd_rn(s1, wb[k]); s2 = __fadd_rn(s2, __fmul_rn(wb[k], wb[k]));
                const float m = __fmul_rn(s1, inv_n);
                out[(size_t)(9 + 4 * q + k) * nvox] = __fsub_rn(__fmul_rn(s2, inv_n), __fmul_rn(m, m));
            }
        }
    }
}

}  // namespace mvsn

using namespace mvsn;

extern "C" {

size_t mvsn_cost_volume_workspace_bytes(int V, int h, int w) {
    // quarter-resolution (r,g,b,0) texels + channel-quad copy of the feature maps
    return (size_t)V * h * w * 4 * sizeof(float) + (size_t)V * 32 * h * w * sizeof(float);
}

int mvsn_build_cost_volume(const float* imgs, const float* feats, const float* proj, const float* depths,
                           int V, int H, int W, int D, int pad, float* cost, float* in_masks,
                           void* workspace, size_t workspace_bytes, void* stream_) {
    MVSN_RANGE("mvsn_build_cost_volume");
    cudaStream_t stream = (cudaStream_t)stream_;
    MVSN_REQUIRE(imgs && feats && proj && depths && cost && workspace, MVSN_ENULL, "mvsn_build_cost_volume: NULL argument");
    MVSN_REQUIRE(V == 3, MVSN_EBADSHAPE, "mvsn_build_cost_volume: V=%d (the cost volume is 9+32 channels = 3 views)", V);
    MVSN_REQUIRE(H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0 && D > 0 && pad >= 0, MVSN_EBADSHAPE,
                 "mvsn_build_cost_volume: H=%d W=%d must be positive multiples of 4, D=%d, pad=%d", H, W, D, pad);
    const int h = H / 4, w = W / 4;
    MVSN_REQUIRE(workspace_bytes >= mvsn_cost_volume_workspace_bytes(V, h, w), MVSN_EWORKSPACE,
                 "mvsn_build_cost_volume: workspace too small");
    MVSN_REQUIRE(aligned16(workspace), MVSN_EALIGN, "mvsn_build_cost_volume: workspace must be 16-byte aligned");
    MVSN_REQUIRE(D <= 65535 && (long long)(h + 2 * pad) * (w + 2 * pad) < (1ll << 31), MVSN_EBADSHAPE,
                 "mvsn_build_cost_volume: D=%d or the padded plane is too large", D);
    float4* small = static_cast<float4*>(workspace);
    float4* featq = small + (size_t)V * h * w;
    downsample_images_kernel<<<cdiv((long long)V * h * w, 256), 256, 0, stream>>>(imgs, small, V, H, W, h, w);
    feats_to_quads_kernel<<<cdiv((long long)V * 8 * h * w, 256), 256, 0, stream>>>(feats, featq, V, h * w);
    CostArgs a;
    a.small = small; a.feats = featq; a.depths = depths;
    a.V = V; a.h = h; a.w = w; a.D = D; a.pad = pad; a.cost = cost; a.masks = in_masks;
    a.proj = proj;
    dim3 grid(cdiv((long long)(h + 2 * pad) * (w + 2 * pad), 256), D);
    cost_volume_kernel<<<grid, 256, 0, stream>>>(a);
    MVSN_CUDA_CHECK(cudaGetLastError());
    return MVSN_OK;
}

}  // extern "C"

/*
 * mvsnerf_b200 -- C ABI of the B200-native MVSNeRF render hot path.
 *
 * The reference (apchenstu/mvsnerf) has no FFI layer: the hot path sits behind
 * plain Python callables.  Each entry point below replaces the
 * arithmetic of one of those callables; the Python mirror of the reference's
 * call signatures lives in mvsnerf_b200/backend.py and binds this header with
 * ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all tensors are dense fp32 unless stated, layouts are written as C arrays;
 *   - the caller owns every buffer (including workspaces); the library allocates
 *     nothing and keeps no state between calls;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     the call returns without synchronising;
 *   - return value: 0 on success, a negative MVSN_E* code otherwise; the text of
 *     the last error on the calling thread is available from mvsn_last_error();
 *     nothing ever throws across this boundary.
 */
#ifndef MVSNERF_B200_H
#define MVSNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVSN_OK          0
#define MVSN_EBADSHAPE  -1   /* illegal size (e.g. padded volume dims not divisible by 8)   */
#define MVSN_EALIGN     -2   /* pointer not 16-byte aligned where the kernel needs it       */
#define MVSN_ECUDA      -3   /* a CUDA call failed; see mvsn_last_error()                   */
#define MVSN_ENULL      -4   /* required pointer is NULL                                    */
#define MVSN_EWORKSPACE -5   /* workspace too small                                         */
#define MVSN_EUNSUPPORTED -6 /* feature/mode not available                                  */

/* arithmetic modes of the per-sample MLP (the GEMMs inside the render kernel) */
#define MVSN_MLP_FP32        0  /* fp32 FFMA, parity <= 1e-4 RGB Linf (north-star fp32 gate)        */
#define MVSN_MLP_TC_HALF     1  /* tcgen05 kind::f16 operands, fp32 accumulate, gate 5e-3           */
#define MVSN_MLP_TC_SPLIT    2  /* tcgen05, 2-term fp16 operand split (3 MMAs), fp32-grade, 1e-4    */
#define MVSN_MLP_TC_PAIR     3  /* as TC_HALF, on CTA pairs (cta_group::2): weights resident in shared     */
                                /* memory, activations in tensor memory, views/feature layers folded     */
                                /* (TC modes take any N_samples; rays are tiled 32 at a time)           */

#define MVSN_N_MLP_TENSORS   22 /* network_fn_state_dict, reference models.py:145-222 / SURVEY App. B */
#define MVSN_N_COSTREG_TENSORS 30 /* 10 x (conv weight, bn gamma, bn beta), models.py:725-769          */
#define MVSN_N_FEATURENET_TENSORS 26 /* 8 x (conv weight, bn gamma, bn beta) + toplayer (weight, bias)   */
#define MVSN_VOL_CH          8
#define MVSN_MAX_PEERS       16 /* ranks of one NVLink domain a frame sink can address              */
#define MVSN_COST_CH         41
#define MVSN_FEAT_CH         32

const char* mvsn_last_error(void);
int mvsn_abi_version(void);

/* ---------------------------------------------------------------------------------------
 * MLP weights  (replaces: nn.Linear parameter reads in Renderer_ours.forward, models.py:194-222)
 *
 * w[22]: device pointers in this order (PyTorch [out,in] row-major, exactly the tensors of
 *   network_fn_state_dict):  pts_linears.{0..5}.weight/.bias interleaved (w0,b0,...,w5,b5),
 *   pts_bias.weight, pts_bias.bias, views_linears.0.weight, .bias, feature_linear.weight, .bias,
 *   alpha_linear.weight, .bias, rgb_linear.weight, .bias.
 * Re-call after every optimiser step when fine-tuning (cost: one small kernel).
 * ------------------------------------------------------------------------------------- */
size_t mvsn_mlp_packed_bytes(int mode);
int mvsn_mlp_pack(const float* const* w_host_array_of_device_ptrs, int mode,
                  void* packed, size_t packed_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Layout helpers
 *   images  [V,3,H,W] (planar, un-normalised [0,1])  ->  [V,H,W,4] (r,g,b,0) texel-interleaved
 *   volume  [8,D,Hp,Wp] (reference layout)           ->  [D,Hp,Wp,8] channels-last
 * The render kernel reads only the interleaved forms (one 16/32-byte sector per tap).
 * ------------------------------------------------------------------------------------- */
int mvsn_pack_images(const float* imgs, int V, int H, int W, float* imgs_hwc4, void* stream);
int mvsn_volume_to_channels_last(const float* vol_cdhw, int D, int Hp, int Wp,
                                 float* vol_dhwc, void* stream);
int mvsn_volume_from_channels_last(const float* vol_dhwc, int D, int Hp, int Wp,
                                   float* vol_cdhw, void* stream);

/* ---------------------------------------------------------------------------------------
 * Rendering  (replaces: renderer.rendering, renderer.py:138-165, and everything it calls:
 *   gen_dir_feature :111-122, gen_pts_feats :124-136, utils.index_point_feature utils.py:357-383,
 *   utils.build_color_volume utils.py:300-332, run_network_mvs renderer.py:42-63,
 *   Embedder.embed models.py:47-51, Renderer_ours.forward models.py:194-222,
 *   raw2outputs/raw2alpha renderer.py:18-26,65-92)
 * ------------------------------------------------------------------------------------- */
typedef struct mvsn_render_scene {
    const float* volume_dhwc;   /* [D,Hp,Wp,8] channels-last encoding volume                    */
    int D, Hp, Wp;
    const float* imgs_hwc4;     /* [V,H,W,4] source images, from mvsn_pack_images               */
    int V, H, W;                /* V must be 3 (feat_dim = 8 + 4V = 20, train_mvs_nerf_pl.py:38) */
    const float* w2cs;          /* [V,4,4] device: world -> camera, pose_source['w2cs']           */
    const float* intrinsics;    /* [V,3,3] device: full-resolution K, pose_source['intrinsics']   */
    const void* mlp_packed;     /* from mvsn_mlp_pack                                             */
    int mlp_mode;               /* MVSN_MLP_*                                                     */
    int white_bkgd;             /* renderer.py:91-92                                              */
} mvsn_render_scene;

/* Signature-compatible entry: the caller has already run ray_marcher and get_ndc_coordinate.
 *   rays_pts [N,S,3] world points, rays_ndc [N,S,3] volume coords in [0,1], z_vals [N,S],
 *   rays_dir [N,3] (un-normalised).
 * Outputs: rgb [N,3], depth [N] required; weights [N,S], alpha [N,S], input_feat [N,S,20]
 * optional (NULL to skip).  */
int mvsn_render_samples(const mvsn_render_scene* scene,
                        const float* rays_pts, const float* rays_ndc, const float* z_vals,
                        const float* rays_dir, int N, int S,
                        float* rgb, float* depth, float* weights, float* alpha, float* input_feat,
                        void* stream);

/* Fused-caller entry: also replaces data/ray_utils.ray_marcher (data/ray_utils.py:152-197,
 * perturb = 0) and utils.get_ndc_coordinate (utils.py:112-146) for the reference camera.
 *   rays [N,8] = (origin, direction, near, far); t_steps [S] = linspace(0,1,S) as the caller's
 *   framework computes it (kept as an input so z_vals match it bit for bit);
 *   ndc_near/ndc_far: the source views' near_far; pad: cost-volume padding in feature pixels
 *   (callers pass pad * imgScale_test); lindisp as in the reference. */
typedef struct mvsn_ray_params {
    float ndc_near, ndc_far;
    float pad;
    int lindisp;
} mvsn_ray_params;

int mvsn_render_rays(const mvsn_render_scene* scene, const mvsn_ray_params* rp,
                     const float* rays, const float* t_steps, int N, int S,
                     float* rgb, float* depth, float* weights, float* alpha, float* input_feat,
                     void* stream);

/* Ray generation for one camera (replaces: data/ray_utils.get_rays, data/ray_utils.py:32-53, and the notebooks'
 * `torch.cat([rays_o, rays_d, near, far])`): directions [n,3] = get_ray_directions(H, W, focal) in camera coordinates
 * (resident on the device, they depend on the intrinsics only), c2w = the first three rows of the camera-to-world matrix,
 * row-major with row stride 4 ([3,4] or [4,4]), on the device.  rays [n,8] = (c2w[:3,3], directions @ c2w[:3,:3]^T,
 * near, far): the input of mvsn_render_rays, so a frame needs 48 bytes of host input. */
int mvsn_make_rays(const float* directions, const float* c2w, float near, float far, int n, float* rays, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fine-tuning step  (replaces: autograd through renderer.rendering + torch.optim.Adam in
 *   train_mvs_nerf_finetuning_pl.py:140-189 -- gradients of the 22 MLP tensors, models.py:145-222, and of
 *   RefVolume.feat_volume, models.py:935-950 -- SURVEY.md 8(f) row 2).
 *
 * mvsn_render_backward re-evaluates the samples in fp32 and back-propagates the given output gradients in one
 * kernel: reverse compositing scan, MLP dgrad/wgrad, trilinear scatter-add into the volume gradient.
 *   scene->mlp_packed must be the MVSN_MLP_FP32 image (scene->mlp_mode == MVSN_MLP_FP32);
 *   mlp_w[22]: the live nn.Linear tensors (device pointers, order as mvsn_mlp_pack);
 *   inputs as mvsn_render_samples; N_samples <= 128;
 *   g: output gradients.  Either g->rgb [N,3] (d loss / d rgb_map) or g->target_rgb [N,3] with g->loss_scale =
 *      1 / (3 * N_total): the img2mse loss and its gradient are then formed in the kernel (g->loss_out, if set, is
 *      ACCUMULATED with this call's share of the loss; g->rgb_out / g->depth_out receive the forward result).
 *      g->depth [N], g->weights [N,S], g->alpha [N,S], g->input_feat [N,S,20] optional (NULL = zero).
 *   grad_mlp[22]: OVERWRITTEN with d loss / d tensor, nn.Linear layouts;
 *   grad_volume_dhwc: [D,Hp,Wp,8] channels-last, ACCUMULATED (atomics); NULL = volume frozen;
 *   workspace: mvsn_render_backward_workspace_bytes(N, S) bytes, 16-byte aligned.
 * mvsn_adam_step / mvsn_adam_step_volume: torch.optim.Adam arithmetic (betas, eps, bias correction by `step` >= 1,
 *   no weight decay / amsgrad).  The volume variant reads the channels-last gradient, zeroes it for the next step,
 *   and updates a parameter (and moments) stored channels-last (planar = 0) or planar [8][nvox] (planar = 1).
 * ------------------------------------------------------------------------------------- */
typedef struct mvsn_render_grads {
    const float* rgb;
    const float* target_rgb;
    float loss_scale;
    const float* depth;
    const float* weights;
    const float* alpha;
    const float* input_feat;
    float* rgb_out;
    float* depth_out;
    float* loss_out;
} mvsn_render_grads;

size_t mvsn_render_backward_workspace_bytes(int N, int S);
int mvsn_render_backward(const mvsn_render_scene* scene, const float* const* mlp_w,
                         const float* rays_pts, const float* rays_ndc, const float* z_vals, const float* rays_dir,
                         int N, int S, const mvsn_render_grads* g, float* const* grad_mlp, float* grad_volume_dhwc,
                         void* workspace, size_t workspace_bytes, void* stream);
int mvsn_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const int* numel_host, int count, float lr, float beta1, float beta2, float eps, int step,
                   void* stream);
int mvsn_adam_step_volume(float* param, float* grad_dhwc, float* exp_avg, float* exp_avg_sq, long long nvox,
                          int planar, float lr, float beta1, float beta2, float eps, int step, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU frame assembly (replaces: nothing in the reference -- its DDP flag is dead code, SURVEY.md 2.3;
 * this is the north star's "rays shard across the GPUs of one box, the rendered image is gathered at the end",
 * SURVEY.md 8(e), done from the render kernel's epilogue instead of by a gather pass).
 *
 * A frame is [n_pixels][4] fp32 texels (r, g, b, depth).  Every rank owns one copy, allocated with
 * mvsn_peer_buffer_create (the ONE place this library allocates: the memory must be exportable to the other
 * processes), exports its 64-byte handle, and maps the other ranks' copies with mvsn_peer_buffer_open.
 * mvsn_render_rays_to_peers then renders this rank's band of rays and stores each finished pixel into all
 * `n_peers` copies (NVLink peer stores, 16 bytes per pixel per peer).  The frame is complete on every rank once
 * every rank's launch has completed (the caller orders that with any stream-level barrier).
 * rgb / depth may be NULL here (the sink is then the only output).
 * ------------------------------------------------------------------------------------- */
#define MVSN_PEER_HANDLE_BYTES 64
int mvsn_peer_buffer_create(size_t bytes, void** dev_ptr, unsigned char handle_host[MVSN_PEER_HANDLE_BYTES]);
int mvsn_peer_buffer_open(const unsigned char handle_host[MVSN_PEER_HANDLE_BYTES], void** peer_ptr);
int mvsn_peer_buffer_close(void* peer_ptr);
int mvsn_peer_buffer_destroy(void* dev_ptr);

typedef struct mvsn_peer_sink {
    float* frame[MVSN_MAX_PEERS]; /* frame[r]: rank r's copy, [n_pixels,4] fp32, 16-byte aligned (own or peer-mapped) */
    int n_peers;
    long long first_pixel;        /* index in the frame of this call's ray 0                                       */
} mvsn_peer_sink;

int mvsn_render_rays_to_peers(const mvsn_render_scene* scene, const mvsn_ray_params* rp,
                              const float* rays, const float* t_steps, int N, int S,
                              const mvsn_peer_sink* sink, float* rgb, float* depth, void* stream);

/* ---------------------------------------------------------------------------------------
 * Cost volume  (replaces: MVSNet.build_volume_costvar_img, models.py:839-893, with
 *   utils.homo_warp utils.py:580-630 and the F.interpolate at models.py:859)
 *   imgs   [V,3,H,W]  ImageNet-normalised source images (H = 4h, W = 4w)
 *   feats  [V,32,h,w] FeatureNet output
 *   proj   [V,3,4]    src_proj @ inv(ref_proj) in feature space (row 0 unused)
 *   depths [D]
 * Outputs (reference layouts): cost [41,D,h+2pad,w+2pad]; in_masks [V,D,h+2pad,w+2pad] or NULL.
 * The never-written pad border of channels 0:3 is defined as zero.
 * workspace: mvsn_cost_volume_workspace_bytes(V,h,w) bytes.
 * ------------------------------------------------------------------------------------- */
size_t mvsn_cost_volume_workspace_bytes(int V, int h, int w);
int mvsn_build_cost_volume(const float* imgs, const float* feats, const float* proj,
                           const float* depths, int V, int H, int W, int D, int pad,
                           float* cost, float* in_masks, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ---------------------------------------------------------------------------------------
 * Feature extraction  (replaces: FeatureNet.forward + ConvBnReLU + InPlaceABN in train mode,
 *   models.py:661-672,688-722; called from MVSNet.forward models.py:907-909).
 *   w[26]: device pointers, for each of conv0.0, conv0.1, conv1.0, conv1.1, conv1.2, conv2.0, conv2.1,
 *          conv2.2 in that order: (conv weight [Cout,Cin,k,k], bn gamma, bn beta); then toplayer.weight
 *          [32,32,1,1] and toplayer.bias [32].
 *   imgs [V,3,H,W] (ImageNet-normalised) -> feats [V,32,ceil(H/4),ceil(W/4)].  Batch statistics are taken
 *   over all V views jointly, as the reference batches them (B*V images through one BN).
 * ------------------------------------------------------------------------------------- */
size_t mvsn_featurenet_workspace_bytes(int V, int H, int W);
int mvsn_featurenet_forward(const float* const* w_host_array_of_device_ptrs, const float* imgs,
                            int V, int H, int W, float* feats,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Cost regularisation  (replaces: CostRegNet.forward + ConvBnReLU3D + InPlaceABN in train mode,
 *   models.py:674-685,725-769).
 *   w[30]: device pointers, for each of conv0..conv6, conv7, conv9, conv11 in that order:
 *          (conv weight, bn gamma, bn beta).  conv weight layouts as in the checkpoint:
 *          Conv3d [Cout,Cin,3,3,3]; ConvTranspose3d [Cin,Cout,3,3,3].
 *   cost [41,D,Hp,Wp] -> volume_dhwc [D,Hp,Wp,8] (channels-last; use
 *   mvsn_volume_from_channels_last for the reference layout).  D, Hp, Wp must be divisible by 8.
 *   Batch statistics (every shipped caller runs MVSNet.train()); see mvsn_costreg_forward_bn for eval mode.
 * ------------------------------------------------------------------------------------- */
size_t mvsn_costreg_workspace_bytes(int D, int Hp, int Wp);
int mvsn_costreg_forward(const float* const* w_host_array_of_device_ptrs, const float* cost,
                         int D, int Hp, int Wp, float* volume_dhwc,
                         void* workspace, size_t workspace_bytes, void* stream);

/* BatchNorm mode of the two encoder entry points (InPlaceABN, models.py:661-685; `self.training` dispatch,
 * SURVEY.md 8(b)):
 *   MVSN_BN_BATCH         batch statistics, running statistics untouched (the plain entry points above);
 *   MVSN_BN_BATCH_UPDATE  batch statistics AND the train-mode side effect of F.batch_norm: running_mean / running_var
 *                         are updated in place with `momentum` (variance unbiased) -- what MVSNet.train()(...) does
 *                         to the buffers save_ckpt later serialises;
 *   MVSN_BN_RUNNING       eval mode: normalise with running_mean / running_var.
 * running[2 L]: for each BN layer in the order of w: (running_mean, running_var); L = 8 (FeatureNet), 10 (CostRegNet). */
#define MVSN_BN_BATCH         0
#define MVSN_BN_BATCH_UPDATE  1
#define MVSN_BN_RUNNING       2
#define MVSN_CONV0_FFMA       0x100 /* diagnostic flag, OR-ed into bn_mode of mvsn_costreg_forward_bn: run the first layer
                                       (41 -> 8 channels) on the fp32 FFMA kernel instead of the tcgen05 kernel */
int mvsn_featurenet_forward_bn(const float* const* w_host_array_of_device_ptrs, float* const* running, int bn_mode,
                               float momentum, const float* imgs, int V, int H, int W, float* feats,
                               void* workspace, size_t workspace_bytes, void* stream);
int mvsn_costreg_forward_bn(const float* const* w_host_array_of_device_ptrs, float* const* running, int bn_mode,
                            float momentum, const float* cost, int D, int Hp, int Wp, float* volume_dhwc,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Diagnostic: one 128 x N x K fp16 GEMM (fp32 accumulate) through the same tcgen05 building blocks
 * the fused render kernel uses.  A [128,K], B [N,K] fp16 row-major; Bc [N,16] optional extra K-step
 * (D += A[:,16:32] * Bc^T, the "bias" step); D [128,N] fp32.  Used by tests only.
 * ------------------------------------------------------------------------------------- */
/* Debug: when non-NULL, CTA 0 of the tensor-core render kernel records a clock64 timeline of its
 * pipeline events into this device buffer (8 roles x 1024 entries).  NULL (default) disables it. */
void mvsn_debug_set_trace(long long* device_buffer);

int mvsn_selftest_umma(const void* A, const void* B, const void* Bc, int N, int K, float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVSNERF_B200_H */

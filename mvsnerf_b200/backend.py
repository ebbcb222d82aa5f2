"""Host-side mirror of the reference's Python surface for the render hot path.

Same names, argument meaning and return tuples as apchenstu/mvsnerf:

    create_nerf_mvs(args, ...)                      models.py:569-654
    MVSNet(imgs, proj_mats, near_far, pad, ...)     models.py:771-932
    MVSNeRF / network_fn                            models.py:540-567 (Renderer_ours, :145-222)
    RefVolume                                       models.py:935-950
    rendering(args, pose_ref, rays_pts, ...)        renderer.py:138-165

so the reference's per-scene fine-tuning script (train_mvs_nerf_finetuning_pl.py) and the render notebooks run with
`from mvsnerf_b200.backend import create_nerf_mvs, rendering, RefVolume`.  What does NOT carry over: end-to-end
training of the encoder (train_mvs_nerf_pl.py back-propagates through MVSNet): the encoding-volume kernels are
forward-only, `MVSNet.forward` returns a volume without a graph and says so with a warning when called under
autograd, and `create_nerf_mvs` therefore leaves the encoder's parameters out of `grad_vars`.  The modules keep the
reference's parameter names, so `ckpts/mvsnerf-v0.tar` (and fine-tuned checkpoints) load with
`load_state_dict(strict=True)`.  All arithmetic of the path runs in libmvsnerf_b200.so (hand
written sm_100a CUDA, bound through ctypes); PyTorch here only owns device memory, streams and
parameters.  There is no CPU path: CPU tensors are rejected with a RuntimeError.

`render_rays` is the fused-caller entry (ray marching + NDC conversion also in the kernel) that
replaces the notebooks' per-chunk loop of ray_marcher -> get_ndc_coordinate -> rendering.
"""
from __future__ import annotations

import ctypes as C
import warnings
import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as _lib

N_DEPTH_PLANES = 128     # models.py:914


# --------------------------------------------------------------------------------------------
# parameter containers with the reference's state_dict keys
# --------------------------------------------------------------------------------------------
class InPlaceABN(nn.Module):
    """BatchNorm + leaky-ReLU(0.01) with the affine weight used as |gamma|+eps (inplace_abn).

    A parameter container inside FeatureNet / CostRegNet (the CUDA kernels apply it while loading the next
    layer's input); forward() is the plain PyTorch statement of the op for stand-alone use."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, slope=0.01):
        super().__init__()
        self.eps, self.momentum, self.slope = eps, momentum, slope
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        y = F.batch_norm(x, self.running_mean, self.running_var, self.weight.abs() + self.eps, self.bias,
                         self.training, self.momentum, self.eps)
        return F.leaky_relu(y, self.slope)


class ConvBnReLU(nn.Module):
    def __init__(self, cin, cout, k=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False)
        self.bn = InPlaceABN(cout)

    def forward(self, x):
        return self.bn(self.conv(x))


class ConvBnReLU3D(nn.Module):
    """Parameter container (conv.weight, bn.*); evaluated by mvsn_costreg_forward."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn = InPlaceABN(cout)


class FeatureNet(nn.Module):
    """models.py:688-722 parameter layout; forward = one C-ABI call (mvsn_featurenet_forward):
    eight conv + train-mode InPlaceABN layers (statistics over all views jointly) and the 1x1 toplayer."""

    def __init__(self):
        super().__init__()
        self.conv0 = nn.Sequential(ConvBnReLU(3, 8), ConvBnReLU(8, 8))
        self.conv1 = nn.Sequential(ConvBnReLU(8, 16, 5, 2, 2), ConvBnReLU(16, 16), ConvBnReLU(16, 16))
        self.conv2 = nn.Sequential(ConvBnReLU(16, 32, 5, 2, 2), ConvBnReLU(32, 32), ConvBnReLU(32, 32))
        self.toplayer = nn.Conv2d(32, 32, 1)

    def weight_list(self):
        out = []
        for block in (self.conv0, self.conv1, self.conv2):
            for m in block:
                out += [m.conv.weight, m.bn.weight, m.bn.bias]
        return out + [self.toplayer.weight, self.toplayer.bias]

    def bn_modules(self):
        return [m.bn for block in (self.conv0, self.conv1, self.conv2) for m in block]

    def forward(self, x):
        """x [V,3,H,W] -> [V,32,ceil(H/4),ceil(W/4)].  BatchNorm dispatches on `self.training` like InPlaceABN
        (models.py:661-672): train = batch statistics over the V views (+ running-statistics update), eval = running."""
        lib = _lib.load()
        x = _lib.dev_f32(x.detach(), "FeatureNet input")
        V, C, H, W = x.shape
        if C != 3:
            raise RuntimeError(f"FeatureNet expects [V,3,H,W] images, got {tuple(x.shape)}")
        ws_bytes = lib.mvsn_featurenet_workspace_bytes(V, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        feats = torch.empty(V, 32, (H + 3) // 4, (W + 3) // 4, dtype=torch.float32, device=x.device)
        weights = [_lib.dev_f32(w.detach(), "FeatureNet weight") for w in self.weight_list()]
        running, mode, momentum = _bn_call_args(self.bn_modules(), self.training)
        with torch.cuda.device(x.device):
            _lib.check(lib.mvsn_featurenet_forward_bn(_lib.ptr_array(weights), _lib.ptr_array(running), mode, momentum,
                                                      _lib.ptr(x), V, H, W, _lib.ptr(feats), _lib.ptr(ws), ws_bytes,
                                                      _lib.stream_ptr()), "mvsn_featurenet_forward_bn")
        return feats


def _bn_call_args(bns, training):
    """(running-statistics pointers, MVSN_BN_* mode, momentum) for a stack of InPlaceABN parameter containers."""
    running = []
    for bn in bns:
        for t in (bn.running_mean, bn.running_var):
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("BatchNorm running statistics must be contiguous CUDA fp32 buffers")
            running.append(t)
    if not training:
        return running, _lib.BN_RUNNING, 0.0
    for bn in bns:
        bn.num_batches_tracked += 1                     # as nn.BatchNorm / InPlaceABN do in train mode
    return running, _lib.BN_BATCH_UPDATE, float(bns[0].momentum)


class CostRegNet(nn.Module):
    """models.py:725-769 parameter layout; forward = one C-ABI call."""

    LAYERS = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv9", "conv11")

    def __init__(self, in_channels=41):
        super().__init__()
        self.conv0 = ConvBnReLU3D(in_channels, 8)
        self.conv1 = ConvBnReLU3D(8, 16, stride=2)
        self.conv2 = ConvBnReLU3D(16, 16)
        self.conv3 = ConvBnReLU3D(16, 32, stride=2)
        self.conv4 = ConvBnReLU3D(32, 32)
        self.conv5 = ConvBnReLU3D(32, 64, stride=2)
        self.conv6 = ConvBnReLU3D(64, 64)
        for name, cin, cout in (("conv7", 64, 32), ("conv9", 32, 16), ("conv11", 16, 8)):
            setattr(self, name, nn.Sequential(
                nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False),
                InPlaceABN(cout)))

    def weight_list(self):
        out = []
        for name in self.LAYERS:
            m = getattr(self, name)
            conv, bn = (m.conv, m.bn) if isinstance(m, ConvBnReLU3D) else (m[0], m[1])
            out += [conv.weight, bn.weight, bn.bias]
        return out

    def bn_modules(self):
        out = []
        for name in self.LAYERS:
            m = getattr(self, name)
            out.append(m.bn if isinstance(m, ConvBnReLU3D) else m[1])
        return out

    def forward(self, cost):
        """cost [1,41,D,Hp,Wp] (reference layout) -> [1,8,D,Hp,Wp] (channels-last memory).  BatchNorm dispatches on
        `self.training` (models.py:674-685): batch statistics (+ running update) in train mode, running in eval."""
        lib = _lib.load()
        cost = _lib.dev_f32(cost, "cost volume")
        _, _, D, Hp, Wp = cost.shape
        ws_bytes = lib.mvsn_costreg_workspace_bytes(D, Hp, Wp)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cost.device)
        vol = torch.empty(D, Hp, Wp, 8, dtype=torch.float32, device=cost.device)
        weights = [_lib.dev_f32(w.detach(), "CostRegNet weight") for w in self.weight_list()]
        running, mode, momentum = _bn_call_args(self.bn_modules(), self.training)
        with torch.cuda.device(cost.device):
            _lib.check(lib.mvsn_costreg_forward_bn(_lib.ptr_array(weights), _lib.ptr_array(running), mode, momentum,
                                                   _lib.ptr(cost), D, Hp, Wp, _lib.ptr(vol), _lib.ptr(ws), ws_bytes,
                                                   _lib.stream_ptr()), "mvsn_costreg_forward_bn")
        return vol.permute(3, 0, 1, 2).unsqueeze(0)


class FrozenEncoderWarning(UserWarning):
    """MVSNet.forward was called under autograd with trainable encoder parameters (see its message)."""


class MVSNet(nn.Module):
    """Encoding-volume builder with the reference's call signature (models.py:771-932).  `.train()` / `.eval()` select
    batch-statistics or running-statistics BatchNorm exactly as for the reference module (every shipped caller runs
    `.train()` first; eval mode gives very different volumes with the shipped checkpoint, App. D)."""

    def __init__(self):
        super().__init__()
        self.feature = FeatureNet()
        self.cost_reg_2 = CostRegNet(32 + 9)
        self.N_importance = 0
        self.chunk = 1024

    def build_volume_costvar_img(self, imgs, feats, proj_mats, depth_values, pad=0):
        """models.py:839-893.  imgs [1,V,3,H,W], feats [1,V,32,h,w], proj_mats [1,V,3,4], depth_values [1,D]
        -> (img_feat [1,41,D,h',w'], in_masks [1,V,D,h',w'])."""
        lib = _lib.load()
        B, V, C, h, w = feats.shape
        if B != 1:
            raise RuntimeError("MVSNet: batch size must be 1 (as in every reference call site)")
        H, W = imgs.shape[-2:]
        D = depth_values.shape[-1]
        dev = feats.device
        imgs_c = _lib.dev_f32(imgs.reshape(V, 3, H, W), "imgs")
        feats_c = _lib.dev_f32(feats.reshape(V, C, h, w).detach(), "feats")
        proj_c = _lib.dev_f32(proj_mats.reshape(V, 3, 4).to(dev), "proj_mats")
        depth_c = _lib.dev_f32(depth_values.reshape(D).to(dev), "depth_values")
        hp, wp = h + 2 * pad, w + 2 * pad
        cost = torch.empty(1, 41, D, hp, wp, dtype=torch.float32, device=dev)
        masks = torch.empty(1, V, D, hp, wp, dtype=torch.float32, device=dev)
        ws_bytes = lib.mvsn_cost_volume_workspace_bytes(V, h, w)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mvsn_build_cost_volume(_lib.ptr(imgs_c), _lib.ptr(feats_c), _lib.ptr(proj_c),
                                                  _lib.ptr(depth_c), V, H, W, D, int(pad), _lib.ptr(cost),
                                                  _lib.ptr(masks), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                       "mvsn_build_cost_volume")
        return cost, masks

    def forward(self, imgs, proj_mats, near_far, pad=0, return_color=False, lindisp=False):
        if not imgs.is_cuda:
            raise RuntimeError("MVSNet: inputs must be CUDA tensors; mvsnerf_b200 has no CPU path")
        if torch.is_grad_enabled() and (imgs.requires_grad or any(p.requires_grad for p in self.parameters())):
            # train_mvs_nerf_pl.py:113 trains the encoder through this call; the kernels here are forward-only
            warnings.warn("mvsnerf_b200.MVSNet.forward runs forward-only CUDA kernels: the returned encoding volume "
                          "carries no autograd graph, so FeatureNet / CostRegNet receive no gradients (the encoder is "
                          "frozen). Call it under torch.no_grad() (as train_mvs_nerf_finetuning_pl.py:63 does) or "
                          "freeze the module to silence this.", FrozenEncoderWarning, stacklevel=2)
        B, V, _, H, W = imgs.shape
        feats = self.feature(imgs.reshape(B * V, 3, H, W))
        feats_l = feats.view(B, V, *feats.shape[1:])
        t = torch.linspace(0.0, 1.0, steps=N_DEPTH_PLANES, device=imgs.device, dtype=imgs.dtype)
        near, far = near_far
        if not lindisp:
            depth_values = near * (1.0 - t) + far * t
        else:
            depth_values = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
        depth_values = depth_values.unsqueeze(0)
        cost, in_masks = self.build_volume_costvar_img(imgs, feats_l, proj_mats, depth_values, pad=pad)
        if return_color:
            feats_l = torch.cat((cost[:, :V * 3].view(B, V, 3, *cost.shape[2:]), in_masks.unsqueeze(2)), dim=2)
        volume = self.cost_reg_2(cost)
        return volume, feats_l, depth_values


class _RendererV0(nn.Module):
    """Parameter layout of Renderer_ours (models.py:145-222) for net_type 'v0'."""

    def __init__(self, D=6, W=128, input_ch=63, input_ch_views=3, input_ch_feat=20, skips=(4,)):
        super().__init__()
        if (D, W, input_ch, input_ch_views, input_ch_feat, tuple(skips)) != (6, 128, 63, 3, 20, (4,)):
            raise RuntimeError("the CUDA render kernel is built for the v0 network of ckpts/mvsnerf-v0.tar: "
                               "netdepth 6, netwidth 128, 63-ch positional encoding, 3-ch view dir, 20 features")
        self.skips = tuple(skips)
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W + input_ch if (i in self.skips) else W, W) for i in range(D - 1)])
        self.pts_bias = nn.Linear(input_ch_feat, W)
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)

    def trunk(self, pts, feats):
        mod = self.pts_bias(feats)
        h = pts
        for i, layer in enumerate(self.pts_linears):
            h = F.relu(layer(h) * mod)
            if i in self.skips:
                h = torch.cat([pts, h], -1)
        return h

    def forward_alpha(self, x):
        pts, feats = x[..., :63], x[..., 63:]
        return torch.relu(self.alpha_linear(self.trunk(pts, feats)))

    def forward(self, x):
        pts, feats, dirs = x[..., :63], x[..., 63:83], x[..., 83:]
        h = self.trunk(pts, feats)
        sigma = torch.relu(self.alpha_linear(h))
        hv = F.relu(self.views_linears[0](torch.cat([self.feature_linear(h), dirs], -1)))
        return torch.cat([torch.sigmoid(self.rgb_linear(hv)), sigma], -1)


class MVSNeRF(nn.Module):
    """models.py:540-567.  `.nerf.*` parameter names as in network_fn_state_dict.  The PyTorch
    forward below exists for the non-hot-path callers (alpha-only queries); `rendering` never uses
    it -- it hands the parameters to the fused kernel."""

    def __init__(self, D=6, W=128, input_ch_pts=63, input_ch_views=3, input_ch_feat=20, skips=(4,), net_type="v0"):
        super().__init__()
        if net_type != "v0":
            raise RuntimeError(f"net_type {net_type!r}: only 'v0' (the shipped checkpoint) is implemented")
        self.nerf = _RendererV0(D, W, input_ch_pts, input_ch_views, input_ch_feat, skips)
        self._packed = {}

    def forward(self, x):
        return self.nerf(x)

    def forward_alpha(self, x):
        return self.nerf.forward_alpha(x)

    def ordered_params(self):
        n = self.nerf
        out = []
        for layer in n.pts_linears:
            out += [layer.weight, layer.bias]
        out += [n.pts_bias.weight, n.pts_bias.bias, n.views_linears[0].weight, n.views_linears[0].bias,
                n.feature_linear.weight, n.feature_linear.bias, n.alpha_linear.weight, n.alpha_linear.bias,
                n.rgb_linear.weight, n.rgb_linear.bias]
        return out

    def packed(self, mode=_lib.MLP_FP32):
        """Kernel-layout weight image; re-packed whenever a parameter was modified in place."""
        lib = _lib.load()
        params = self.ordered_params()
        dev = params[0].device
        if not params[0].is_cuda:
            raise RuntimeError("network_fn must live on a CUDA device; mvsnerf_b200 has no CPU path")
        key = (mode, dev, tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        hit = self._packed.get(mode)
        if hit is not None and hit[0] == key:
            return hit[1]
        nbytes = lib.mvsn_mlp_packed_bytes(mode)
        if nbytes == 0:
            raise RuntimeError(f"MLP mode {mode} is not available in this build of libmvsnerf_b200")
        buf = hit[1] if hit is not None and hit[1].numel() == nbytes and hit[1].device == dev else \
            torch.empty(nbytes, dtype=torch.uint8, device=dev)
        srcs = [_lib.dev_f32(p.detach(), "MLP parameter") for p in params]
        with torch.cuda.device(dev):
            _lib.check(lib.mvsn_mlp_pack(_lib.ptr_array(srcs), mode, _lib.ptr(buf), nbytes, _lib.stream_ptr()),
                       "mvsn_mlp_pack")
        self._packed[mode] = (key, buf)
        return buf


class RefVolume(nn.Module):
    """models.py:935-950: the encoding volume as a parameter (per-scene fine-tuning)."""

    def __init__(self, volume):
        super().__init__()
        self.feat_volume = nn.Parameter(volume)

    def forward(self, ray_coordinate_ref):
        H, W = ray_coordinate_ref.shape[-3:-1]
        grid = ray_coordinate_ref.view(-1, 1, H, W, 3).to(self.feat_volume.device) * 2 - 1.0
        f = F.grid_sample(self.feat_volume, grid, align_corners=True, mode="bilinear")
        return f[:, :, 0].permute(2, 3, 0, 1).squeeze()


# --------------------------------------------------------------------------------------------
# scene-constant device state (packed images, channels-last volume), cached per tensor version
# --------------------------------------------------------------------------------------------
_cache = {}


cache_stats = {"hit": 0, "miss": 0}


def _cached(kind, t, build):
    """One entry per kind, keyed on the IDENTITY of the caller's tensor object (held by weakref)
    and its in-place version counter -- never on data_ptr, which the caching allocator recycles.
    `t` must be the object the CALLER holds (a Parameter, a checkpoint tensor), not a view made here."""
    hit = _cache.get(kind)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        cache_stats["hit"] += 1
        return hit[2]
    cache_stats["miss"] += 1
    val = build()
    _cache[kind] = (weakref.ref(t), t._version, val)
    return val


def clear_cache():
    _cache.clear()


def _volume_channels_last(volume_feature):
    """Accepts a tensor [1,8,D,Hp,Wp] (any strides) or a RefVolume; returns ([D,Hp,Wp,8] fp32, dims)."""
    owner = volume_feature.feat_volume if isinstance(volume_feature, nn.Module) else volume_feature
    vol = owner.detach()          # a NEW tensor object every call: the cache below is keyed on `owner`
    if vol.dim() != 5 or vol.shape[0] != 1 or vol.shape[1] != 8:
        raise RuntimeError(f"encoding volume must be [1,8,D,H,W], got {tuple(vol.shape)}")
    if not vol.is_cuda:
        raise RuntimeError("encoding volume must be a CUDA tensor; mvsnerf_b200 has no CPU path")
    _, _, D, Hp, Wp = vol.shape
    cl = vol[0].permute(1, 2, 3, 0)
    if vol.dtype == torch.float32 and cl.is_contiguous():
        return cl, (D, Hp, Wp)                         # MVSNet.forward output: zero-copy

    def build():
        lib = _lib.load()
        src = _lib.dev_f32(vol[0], "volume")
        dst = torch.empty(D, Hp, Wp, 8, dtype=torch.float32, device=vol.device)
        with torch.cuda.device(vol.device):
            _lib.check(lib.mvsn_volume_to_channels_last(_lib.ptr(src), D, Hp, Wp, _lib.ptr(dst), _lib.stream_ptr()),
                       "mvsn_volume_to_channels_last")
        return dst

    return _cached("volume", owner, build), (D, Hp, Wp)


def _images_packed(imgs):
    """imgs [1,V,3,H,W] un-normalised -> [V,H,W,4]."""
    if imgs.dim() != 5 or imgs.shape[0] != 1 or imgs.shape[2] != 3:
        raise RuntimeError(f"imgs must be [1,V,3,H,W], got {tuple(imgs.shape)}")
    if not imgs.is_cuda:
        raise RuntimeError("imgs must be a CUDA tensor; mvsnerf_b200 has no CPU path")
    _, V, _, H, W = imgs.shape

    def build():
        lib = _lib.load()
        src = _lib.dev_f32(imgs[0].detach(), "imgs")
        dst = torch.empty(V, H, W, 4, dtype=torch.float32, device=imgs.device)
        with torch.cuda.device(imgs.device):
            _lib.check(lib.mvsn_pack_images(_lib.ptr(src), V, H, W, _lib.ptr(dst), _lib.stream_ptr()),
                       "mvsn_pack_images")
        return dst

    return _cached("imgs", imgs, build), (V, H, W)


def _make_scene(pose_ref, volume_feature, imgs, network_fn, white_bkgd, mode):
    vol, (D, Hp, Wp) = _volume_channels_last(volume_feature)
    im, (V, H, W) = _images_packed(imgs)
    if V != 3:
        raise RuntimeError(f"{V} source views: the v0 network takes exactly 3 (feat_dim = 8 + 3*4)")
    sc = _lib.RenderScene()
    sc.volume_dhwc, sc.D, sc.Hp, sc.Wp = vol.data_ptr(), D, Hp, Wp
    sc.imgs_hwc4, sc.V, sc.H, sc.W = im.data_ptr(), V, H, W
    w2cs = _lib.dev_f32(pose_ref["w2cs"].detach(), "pose_ref['w2cs']")               # [V,4,4]
    intr = _lib.dev_f32(pose_ref["intrinsics"].detach(), "pose_ref['intrinsics']")   # [V,3,3]
    if tuple(w2cs.shape) != (3, 4, 4) or tuple(intr.shape) != (3, 3, 3):
        raise RuntimeError(f"pose_ref: expected w2cs [3,4,4] and intrinsics [3,3,3], got "
                           f"{tuple(w2cs.shape)} and {tuple(intr.shape)}")
    sc.w2cs, sc.intrinsics = w2cs.data_ptr(), intr.data_ptr()
    packed = network_fn.packed(mode)
    sc.mlp_packed, sc.mlp_mode, sc.white_bkgd = packed.data_ptr(), mode, int(bool(white_bkgd))
    keep = (vol, im, packed, w2cs, intr)          # keep the device buffers alive for the duration of the call
    return sc, keep


# Default arithmetic of `rendering` / `render_rays`: the fp32-grade tensor-core mode (2-term fp16 operand split,
# RGB Linf 1.2e-5 vs the reference on every pixel of the 512x640 config -- the same as the FFMA kernel, several
# times faster).  MLP_FP32 (FFMA) and MLP_TC_HALF / MLP_TC_PAIR (5e-3 tier) are selected with `mlp_mode=`.
DEFAULT_MLP_MODE = _lib.MLP_TC_SPLIT


def rendering(args, pose_ref, rays_pts, rays_ndc, depth_candidates, rays_o, rays_dir,
              volume_feature=None, imgs=None, network_fn=None, img_feat=None, network_query_fn=None,
              white_bkgd=False, **kwargs):
    """Drop-in for renderer.rendering (renderer.py:138-165).  Returns
    (rgb_map [N,3], input_feat [N,S,20], weights [N,S], depth_map [N], alpha [N,S], {}).

    Extra keyword arguments the reference's callers pass (perturb, N_importance, network_fine,
    use_viewdirs, raw_noise_std, NDC_local) are accepted and ignored, as the reference does.
    `mlp_mode=` selects the GEMM arithmetic (default: DEFAULT_MLP_MODE, the fp32-grade tensor mode); `want_aux=False` skips the three
    per-sample outputs (they are returned as None)."""
    if pose_ref is None or img_feat is not None or getattr(args, "use_color_volume", False):
        raise RuntimeError("rendering: only the pose_ref / image-gather branch of the reference is implemented "
                           "(use_color_volume=False, img_feat=None) -- the branch every shipped config uses")
    mode = kwargs.pop("mlp_mode", DEFAULT_MLP_MODE)
    want_aux = kwargs.pop("want_aux", True)
    N, S = rays_pts.shape[:2]
    z = depth_candidates.expand(N, S) if depth_candidates.shape != (N, S) else depth_candidates
    vol_t = volume_feature.feat_volume if isinstance(volume_feature, nn.Module) else volume_feature
    if torch.is_grad_enabled() and (vol_t.requires_grad or any(p.requires_grad for p in network_fn.parameters())):
        # training step (fine-tuning, train_mvs_nerf_finetuning_pl.py:164): same kernel forward, gradients for
        # the MLP parameters and the encoding volume (see _RenderSamplesFn)
        params = network_fn.ordered_params()
        rgb, feat, weights, depth, alpha = _RenderSamplesFn.apply(
            rays_pts, rays_ndc, z, rays_dir, vol_t, imgs, pose_ref["w2cs"], pose_ref["intrinsics"],
            bool(white_bkgd), mode, network_fn, volume_feature, *params)
        return rgb, feat, weights, depth, alpha, {}
    rgb, feat, weights, depth, alpha = _render_samples_kernel(
        pose_ref, rays_pts, rays_ndc, z, rays_dir, volume_feature, imgs, network_fn, white_bkgd, mode, want_aux)
    return rgb, feat, weights, depth, alpha, {}


def _render_samples_kernel(pose_ref, rays_pts, rays_ndc, z, rays_dir, volume_feature, imgs, network_fn, white_bkgd,
                           mode, want_aux=True):
    """One mvsn_render_samples launch (no autograd graph)."""
    lib = _lib.load()
    N, S = rays_pts.shape[:2]
    dev = rays_pts.device
    pts = _lib.dev_f32(rays_pts.detach(), "rays_pts")
    ndc = _lib.dev_f32(rays_ndc.detach(), "rays_ndc")
    z = _lib.dev_f32(z.detach(), "depth_candidates")
    dirs = _lib.dev_f32(rays_dir.detach(), "rays_dir")
    sc, keep = _make_scene(pose_ref, volume_feature, imgs, network_fn, white_bkgd, mode)
    rgb = torch.empty(N, 3, dtype=torch.float32, device=dev)
    depth = torch.empty(N, dtype=torch.float32, device=dev)
    feat = weights = alpha = None
    if want_aux:
        feat = torch.empty(N, S, 20, dtype=torch.float32, device=dev)
        weights = torch.empty(N, S, dtype=torch.float32, device=dev)
        alpha = torch.empty(N, S, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mvsn_render_samples(C.byref(sc), _lib.ptr(pts), _lib.ptr(ndc), _lib.ptr(z), _lib.ptr(dirs),
                                           N, S, _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(weights), _lib.ptr(alpha),
                                           _lib.ptr(feat), _lib.stream_ptr()), "mvsn_render_samples")
    del keep
    return rgb, feat, weights, depth, alpha


def _render_samples_torch(pts, ndc, z, rays_dir, vol, imgs, w2cs, intrinsics, nerf, white_bkgd):
    """Differentiable PyTorch statement of renderer.rendering (renderer.py:138-165) -- used ONLY inside
    _RenderSamplesFn.backward to obtain gradients; forward values always come from the CUDA kernel."""
    N, S = pts.shape[:2]
    _, V, _, H, W = imgs.shape
    grid = (ndc * 2 - 1.0).view(1, 1, N, S, 3)
    vfeat = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0, :, 0].permute(1, 2, 0)
    inv_scale = torch.tensor([W - 1.0, H - 1.0], device=pts.device)
    cols = []
    for v in range(V):                                                           # utils.py:300-332
        cam = pts.reshape(-1, 3) @ w2cs[v, :3, :3].t() + w2cs[v, :3, 3].view(1, 3)
        pix = cam @ intrinsics[v].t()
        g = ((pix[:, :2] / pix[:, 2:3] / inv_scale) * 2 - 1.0).view(1, N, S, 2)
        c = F.grid_sample(imgs[0, v:v + 1], g, mode="bilinear", padding_mode="border", align_corners=True)[0].permute(1, 2, 0)
        m = ((g[0] > -1.0) & (g[0] < 1.0)).all(-1, keepdim=True).to(c.dtype)
        cols += [c, m]
    feat = torch.cat([vfeat] + cols, -1)
    d = rays_dir / rays_dir.norm(dim=-1, keepdim=True)
    d = d @ w2cs[0, :3, :3].t()                                                  # renderer.py:111-122
    raw = nerf(torch.cat([_embed(ndc), feat, d[:, None].expand(-1, S, -1)], -1))
    alpha = 1.0 - torch.exp(-raw[..., 3])                                        # renderer.py:18-26
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb = (weights.unsqueeze(-1) * raw[..., :3]).sum(-2)
    depth = (weights * z).sum(-1)
    if white_bkgd:
        rgb = rgb + (1.0 - weights.sum(-1, keepdim=True))
    return rgb, feat, weights, depth, alpha


class _RenderSamplesFn(torch.autograd.Function):
    """Training-step form of `rendering` (interim).

    forward: the fused CUDA kernel, exactly as in inference (no graph, nothing per-sample kept).
    backward: gradients w.r.t. the 22 MLP tensors and the encoding volume by re-evaluating the chunk with
    PyTorch ops under autograd (_render_samples_torch) -- activation memory exists only during backward.
    A hand-written backward kernel (MLP dgrad/wgrad + trilinear scatter) is the planned replacement; the
    interface (what is differentiable, what re-packs after an optimiser step) will not change."""

    @staticmethod
    def forward(ctx, pts, ndc, z, rays_dir, vol, imgs, w2cs, intrinsics, white_bkgd, mode, network_fn, volume_feature,
                *params):
        pose = {"w2cs": w2cs, "intrinsics": intrinsics}
        out = _render_samples_kernel(pose, pts, ndc, z, rays_dir, volume_feature, imgs, network_fn, white_bkgd, mode)
        ctx.save_for_backward(pts, ndc, z, rays_dir, vol, imgs, w2cs, intrinsics, *params)
        ctx.white_bkgd, ctx.network_fn, ctx.volume_feature = white_bkgd, network_fn, volume_feature
        return out

    @staticmethod
    def backward(ctx, g_rgb, g_feat, g_weights, g_depth, g_alpha):
        pts, ndc, z, rays_dir, vol, imgs, w2cs, intrinsics, *params = ctx.saved_tensors
        S = pts.shape[1]
        if S <= 128 and BACKWARD_IMPL == "kernel":
            # the hand-written backward kernel (csrc/render_bwd.cu): recompute + dgrad/wgrad + volume scatter
            need_vol = ctx.needs_input_grad[4]
            grads = {"rgb": g_rgb, "depth": g_depth, "weights": g_weights, "alpha": g_alpha, "input_feat": g_feat}
            g_params, dvol, _, _ = render_backward(
                {"w2cs": w2cs, "intrinsics": intrinsics}, pts, ndc, z, rays_dir, ctx.volume_feature, imgs, ctx.network_fn,
                ctx.white_bkgd, grads=grads, want_volume_grad=need_vol)
            g_vol = dvol.permute(3, 0, 1, 2).unsqueeze(0) if need_vol else None
            g_params = [g if need else None for g, need in zip(g_params, ctx.needs_input_grad[12:])]
            return (None, None, None, None, g_vol, None, None, None, None, None, None, None, *g_params)
        with torch.enable_grad():
            vol_g = vol.detach().requires_grad_(ctx.needs_input_grad[4])
            # evaluate through a functional copy of the module so the user's parameters are not touched
            leaves = [p.detach().requires_grad_(need) for p, need in zip(params, ctx.needs_input_grad[12:])]
            names = [n for n, _ in _ordered_named_params(ctx.network_fn)]
            nerf = lambda x: torch.func.functional_call(ctx.network_fn, dict(zip(names, leaves)), (x,))
            outs = _render_samples_torch(pts.detach(), ndc.detach(), z.detach(), rays_dir.detach(), vol_g, imgs.detach(),
                                         w2cs.detach(), intrinsics.detach(), nerf, ctx.white_bkgd)
            pairs = [(o, g) for o, g in zip(outs, (g_rgb, g_feat, g_weights, g_depth, g_alpha))
                     if g is not None and o.requires_grad]
            wrt = [t for t in [vol_g] + leaves if t.requires_grad]
            grads = torch.autograd.grad([o for o, _ in pairs], wrt, [g for _, g in pairs], allow_unused=True) if wrt and pairs else []
        it = iter(grads)
        g_vol = next(it) if vol_g.requires_grad and grads else None
        g_params = [next(it) if (leaf.requires_grad and grads) else None for leaf in leaves]
        return (None, None, None, None, g_vol, None, None, None, None, None, None, None, *g_params)


# "kernel": csrc/render_bwd.cu (N_samples <= 128); "torch": the PyTorch-recompute backward above (kept as the
# independent statement the gradient tests compare against, and for N_samples > 128)
BACKWARD_IMPL = "kernel"
_bwd_workspace = {}


def _backward_workspace(dev, N, S):
    lib = _lib.load()
    need = lib.mvsn_render_backward_workspace_bytes(int(N), int(S))
    if need == 0:
        raise RuntimeError(f"render backward: unsupported shape N={N}, N_samples={S} (N_samples <= 128)")
    ws = _bwd_workspace.get(dev)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _bwd_workspace[dev] = ws
    return ws, need


def render_backward(pose_ref, rays_pts, rays_ndc, z_vals, rays_dir, volume_feature, imgs, network_fn, white_bkgd=False,
                    grads=None, target_rgb=None, n_total=None, want_volume_grad=True, grad_volume=None, grad_mlp=None,
                    want_forward=False, loss_out=None):
    """One mvsn_render_backward launch.  Either `grads` (dict with 'rgb' and optionally 'depth', 'weights', 'alpha',
    'input_feat': d loss / d output of `rendering`) or `target_rgb` [N,3] (img2mse formed in the kernel, normalised by
    3 * n_total).  Returns (grad_mlp[22] in ordered_params() order, grad_volume [D,Hp,Wp,8] channels-last or None,
    rgb [N,3] or None, depth [N] or None).  `grad_volume` (accumulated into) and `grad_mlp` (overwritten) may be passed
    to reuse buffers."""
    lib = _lib.load()
    N, S = rays_pts.shape[:2]
    dev = rays_pts.device
    pts = _lib.dev_f32(rays_pts.detach(), "rays_pts")
    ndc = _lib.dev_f32(rays_ndc.detach(), "rays_ndc")
    z = _lib.dev_f32((z_vals.expand(N, S) if z_vals.shape != (N, S) else z_vals).detach(), "depth_candidates")
    dirs = _lib.dev_f32(rays_dir.detach(), "rays_dir")
    sc, keep = _make_scene(pose_ref, volume_feature, imgs, network_fn, white_bkgd, _lib.MLP_FP32)
    params = [_lib.dev_f32(p.detach(), "MLP parameter") for p in network_fn.ordered_params()]
    if grad_mlp is None:
        grad_mlp = [torch.empty_like(p) for p in params]
    if want_volume_grad and grad_volume is None:
        grad_volume = torch.zeros(sc.D, sc.Hp, sc.Wp, 8, dtype=torch.float32, device=dev)
    g = _lib.RenderGrads()
    held = []

    def opt(t, shape, name):
        if t is None:
            return None
        t = _lib.dev_f32(t.detach(), name)
        if tuple(t.shape) != tuple(shape):
            t = t.expand(shape).contiguous()
        held.append(t)
        return t.data_ptr()
    if target_rgb is not None:
        g.target_rgb = opt(target_rgb, (N, 3), "target_rgb")
        g.loss_scale = 1.0 / (3.0 * float(N if n_total is None else n_total))
    else:
        if grads is None or grads.get("rgb") is None:
            grads = dict(grads or {})
            grads["rgb"] = torch.zeros(N, 3, dtype=torch.float32, device=dev)
        g.rgb = opt(grads["rgb"], (N, 3), "grad rgb")
    if grads is not None:
        g.depth = opt(grads.get("depth"), (N,), "grad depth")
        g.weights = opt(grads.get("weights"), (N, S), "grad weights")
        g.alpha = opt(grads.get("alpha"), (N, S), "grad alpha")
        g.input_feat = opt(grads.get("input_feat"), (N, S, 20), "grad input_feat")
    rgb = depth = None
    if want_forward:
        rgb = torch.empty(N, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        g.rgb_out, g.depth_out = rgb.data_ptr(), depth.data_ptr()
    if loss_out is not None:
        g.loss_out = loss_out.data_ptr()
    ws, ws_bytes = _backward_workspace(dev, N, S)
    with torch.cuda.device(dev):
        _lib.check(lib.mvsn_render_backward(C.byref(sc), _lib.ptr_array(params), _lib.ptr(pts), _lib.ptr(ndc), _lib.ptr(z),
                                            _lib.ptr(dirs), N, S, C.byref(g), _lib.ptr_array(grad_mlp),
                                            _lib.ptr(grad_volume) if want_volume_grad else None, _lib.ptr(ws), ws_bytes,
                                            _lib.stream_ptr()), "mvsn_render_backward")
    del keep, held
    return grad_mlp, (grad_volume if want_volume_grad else None), rgb, depth


class FineTuner:
    """The reference's per-scene fine-tuning step (train_mvs_nerf_finetuning_pl.py:140-189: rendering -> img2mse ->
    Adam over the MLP and RefVolume.feat_volume) as four launches and no autograd graph:

        mvsn_mlp_pack (fp32 image) -> mvsn_render_backward (forward recompute + loss + all gradients)
        -> mvsn_adam_step (22 MLP tensors) -> mvsn_adam_step_volume (volume; also zeroes its gradient buffer)

    The parameters stay the caller's nn.Parameters (updated in place, version counters bumped), so checkpoints, the
    render entry points and scene_io see them as after a torch.optim.Adam step with the same hyper-parameters."""

    def __init__(self, network_fn, volume, imgs, pose_ref, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, white_bkgd=False):
        self.network_fn, self.volume, self.imgs, self.pose_ref = network_fn, volume, imgs, pose_ref
        self.lr, self.betas, self.eps, self.white_bkgd = float(lr), (float(betas[0]), float(betas[1])), float(eps), bool(white_bkgd)
        self.step_count = 0
        self.params = network_fn.ordered_params()
        dev = self.params[0].device
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.g = [torch.zeros_like(p) for p in self.params]
        fv = volume.feat_volume
        if fv.dim() != 5 or fv.shape[0] != 1 or fv.shape[1] != 8 or not fv.is_cuda or fv.dtype != torch.float32:
            raise RuntimeError("FineTuner: volume.feat_volume must be a CUDA fp32 tensor [1,8,D,H,W]")
        _, _, D, Hp, Wp = fv.shape
        self.nvox = D * Hp * Wp
        cl = fv.detach()[0].permute(1, 2, 3, 0)
        if cl.is_contiguous():
            self.planar = 0                                   # channels-last storage (what MVSNet.forward returns)
        elif fv.is_contiguous():
            self.planar = 1                                   # checkpoint layout [8][nvox]
        else:
            raise RuntimeError("FineTuner: feat_volume must be contiguous either planar or channels-last")
        self.vol_m = torch.zeros_like(fv.detach())
        self.vol_v = torch.zeros_like(fv.detach())
        self.vol_g = torch.zeros(D, Hp, Wp, 8, dtype=torch.float32, device=dev)   # zeroed again by every Adam step
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._numel = (C.c_int * len(self.params))(*[p.numel() for p in self.params])

    def step(self, rays_pts, rays_ndc, z_vals, rays_dir, target_rgb, lr=None, want_forward=False):
        """One optimisation step on a batch.  Returns (loss [1] device tensor -- img2mse of this batch BEFORE the
        update, as the reference logs it -- and (rgb, depth) of the forward pass when `want_forward`)."""
        lib = _lib.load()
        lr = self.lr if lr is None else float(lr)
        self.step_count += 1
        self.loss.zero_()
        _, _, rgb, depth = render_backward(self.pose_ref, rays_pts, rays_ndc, z_vals, rays_dir, self.volume, self.imgs,
                                           self.network_fn, self.white_bkgd, target_rgb=target_rgb, want_volume_grad=True,
                                           grad_volume=self.vol_g, grad_mlp=self.g, want_forward=want_forward,
                                           loss_out=self.loss)
        dev = self.params[0].device
        fv = self.volume.feat_volume
        with torch.cuda.device(dev):
            _lib.check(lib.mvsn_adam_step(_lib.ptr_array([p.detach() for p in self.params]), _lib.ptr_array(self.g),
                                          _lib.ptr_array(self.m), _lib.ptr_array(self.v), self._numel, len(self.params),
                                          lr, self.betas[0], self.betas[1], self.eps, self.step_count, _lib.stream_ptr()),
                       "mvsn_adam_step")
            _lib.check(lib.mvsn_adam_step_volume(_lib.ptr(fv.detach()), _lib.ptr(self.vol_g), _lib.ptr(self.vol_m),
                                                 _lib.ptr(self.vol_v), self.nvox, self.planar, lr, self.betas[0],
                                                 self.betas[1], self.eps, self.step_count, _lib.stream_ptr()),
                       "mvsn_adam_step_volume")
        # the parameters changed behind PyTorch's back: bump their version counters (weight-image / volume caches)
        torch.autograd.graph.increment_version([p for p in self.params] + [fv])
        return self.loss, (rgb, depth)


def ray_marcher(rays, N_samples=64, lindisp=False, perturb=0):
    """data/ray_utils.py:152-197 (without the bbox branch): sample points along rays [N,8] = (o, d, near, far).
    Returns (xyz [N,S,3], rays_o [N,3], rays_d [N,3], z_vals [N,S]).  Host-side mirror for the training callers
    (perturb > 0 draws torch.rand like the reference); inference uses the fused render_rays entry instead."""
    n = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    z_steps = torch.linspace(0, 1, N_samples, device=rays.device)
    if not lindisp:
        z_vals = near * (1 - z_steps) + far * z_steps
    else:
        z_vals = 1 / (1 / near * (1 - z_steps) + 1 / far * z_steps)
    z_vals = z_vals.expand(n, N_samples)
    if perturb > 0:
        mid = 0.5 * (z_vals[:, :-1] + z_vals[:, 1:])
        upper = torch.cat([mid, z_vals[:, -1:]], -1)
        lower = torch.cat([z_vals[:, :1], mid], -1)
        z_vals = lower + (upper - lower) * (perturb * torch.rand(z_vals.shape, device=rays.device))
    xyz = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z_vals.unsqueeze(2)
    return xyz, rays_o, rays_d, z_vals


def get_ndc_coordinate(w2c_ref, intrinsic_ref, point_samples, inv_scale, near=2, far=6, pad=0, lindisp=False):
    """utils.py:112-146 (projection branch): world points [N,S,3] -> volume coordinates in [0,1]."""
    n, s = point_samples.shape[:2]
    p = point_samples.reshape(-1, 3)
    p = torch.matmul(p, w2c_ref[:3, :3].t()) + w2c_ref[:3, 3:].reshape(1, 3)
    q = p @ intrinsic_ref.t()
    q[:, :2] = (q[:, :2] / q[:, -1:] + 0.0) / inv_scale.reshape(1, 2)
    if not lindisp:
        q[:, 2] = (q[:, 2] - near) / (far - near)
    else:
        q[:, 2] = (1.0 / q[:, 2] - 1.0 / near) / (1.0 / far - 1.0 / near)
    if pad > 0:
        w_feat, h_feat = (inv_scale + 1) / 4.0
        q[:, 1] = q[:, 1] * h_feat / (h_feat + pad * 2) + pad / (h_feat + pad * 2)
        q[:, 0] = q[:, 0] * w_feat / (w_feat + pad * 2) + pad / (w_feat + pad * 2)
    return q.view(n, s, 3)


def finetune_step_timing(dev, weights_npz, steps=20, warmup=5, batch=1024, n_samples=128):
    """bench.py's BASELINE-config-3 entry: fine-tuning steps on a Blender-shaped scene (800x800, pad 0, white_bkgd,
    near_far [2, 6]; encoding volume 8x128x200x200), 1024 rays x 128 samples per step, perturb = 1 -- the fused
    FineTuner step and, beside it, the same step through `rendering` under autograd + torch.optim.Adam."""
    from . import synthetic
    fn, mvs = MVSNeRF().to(dev), MVSNet().to(dev).train()
    load_weights_npz(fn, mvs, weights_npz)
    sc = synthetic.make_scene(800, 800, pad=0, seed=3, near_far=(2.0, 6.0))
    d = sc.to(dev)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=0)
    rays_all = synthetic.scene_rays(sc).to(dev)
    target_all = d.imgs_raw[0, 0].permute(1, 2, 0).reshape(-1, 3).contiguous()
    inv_scale = torch.tensor([sc.W - 1.0, sc.H - 1.0], device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)

    def batch_of():
        idx = torch.randint(0, rays_all.shape[0], (batch,), device=dev, generator=gen)
        rays, tgt = rays_all[idx], target_all[idx]
        xyz, _, rays_d, z = ray_marcher(rays, N_samples=n_samples, perturb=1.0)
        ndc = get_ndc_coordinate(d.pose_source["w2cs"][0], d.pose_source["intrinsics"][0], xyz, inv_scale,
                                 near=sc.near_far[0], far=sc.near_far[1], pad=0)
        return xyz, ndc, z, rays_d, tgt

    def timed(fn_step):
        for _ in range(warmup):
            fn_step(*batch_of())
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            b = batch_of()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn_step(*b); e.record(); e.synchronize()
            ts.append(a.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2]

    out = {"rays": batch, "n_samples": n_samples, "volume": list(vol.shape), "white_bkgd": True}
    volume = RefVolume(vol.detach().clone()).to(dev)
    tuner = FineTuner(fn, volume, d.imgs_raw, d.pose_source, lr=5e-4, white_bkgd=True)
    losses = []
    out["fused_ms"] = timed(lambda xyz, ndc, z, rd, tgt: losses.append(tuner.step(xyz, ndc, z, rd, tgt)[0].clone()))
    out["fused_loss_first_last"] = [float(losses[0]), float(losses[-1])]
    # the reference's own step shape: rendering under autograd (kernel forward + kernel backward) + torch.optim.Adam
    fn2 = MVSNeRF().to(dev)
    load_weights_npz(fn2, None, weights_npz)
    volume2 = RefVolume(vol.detach().clone()).to(dev)
    opt = torch.optim.Adam(list(fn2.parameters()) + list(volume2.parameters()), lr=5e-4, betas=(0.9, 0.999))
    from types import SimpleNamespace
    args = SimpleNamespace(use_color_volume=False)

    def autograd_step(xyz, ndc, z, rd, tgt):
        rgb = rendering(args, d.pose_source, xyz, ndc, z, None, rd, volume2, d.imgs_raw, network_fn=fn2, white_bkgd=True,
                        want_aux=True)[0]
        loss = torch.mean((rgb - tgt) ** 2)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    out["autograd_adam_ms"] = timed(autograd_step)
    out["note"] = ("median CUDA-event ms per step incl. ray marching (torch ops) ; fused = FineTuner.step (mvsn_render_backward "
                   "+ mvsn_adam_step + mvsn_adam_step_volume); autograd = backend.rendering under autograd (kernel backward) "
                   "+ torch.optim.Adam; README.md's reference figure is ~15 min for 10k such steps")
    return out


def _ordered_named_params(network_fn):
    """(name, parameter) in MVSNeRF.ordered_params() order."""
    by_id = {id(p): n for n, p in network_fn.named_parameters()}
    return [(by_id[id(p)], p) for p in network_fn.ordered_params()]


_tsteps = {}


def render_rays(rays, volume_feature, imgs, pose_ref, network_fn, near_far, pad, N_samples=128,
                white_bkgd=False, lindisp=False, mlp_mode=None, out=None, sink=None):
    """Fused-caller entry: one launch renders all `rays` [N,8] = (o, d, near, far).

    Replaces the notebooks' per-chunk loop `ray_marcher -> get_ndc_coordinate -> rendering`
    (renderer_video.ipynb "DTU video rendering"; data/ray_utils.py:152-197, utils.py:112-146) with
    perturb = 0.  `near_far` / `pad` are the arguments the reference passes to get_ndc_coordinate
    (near_far of the source views, pad * imgScale_test).  Returns (rgb [N,3], depth [N]).

    `sink` (a lib.PeerSink from distributed.PeerFrame.sink): the kernel epilogue additionally stores every pixel
    as (r, g, b, depth) into all ranks' copies of the assembled frame (mvsn_render_rays_to_peers); with a sink and
    no `out`, nothing else is written and (None, None) is returned."""
    lib = _lib.load()
    mode = DEFAULT_MLP_MODE if mlp_mode is None else mlp_mode
    rays = _lib.dev_f32(rays, "rays")
    N = rays.shape[0]
    dev = rays.device
    S = int(N_samples)
    tk = (S, dev)
    if tk not in _tsteps:
        _tsteps[tk] = torch.linspace(0, 1, S, device=dev)          # data/ray_utils.py:175
    sc, keep = _make_scene(pose_ref, volume_feature, imgs, network_fn, white_bkgd, mode)
    rp = _lib.RayParams(float(near_far[0]), float(near_far[1]), float(pad), int(bool(lindisp)))
    if out is not None:
        rgb, depth = out
    elif sink is not None:
        rgb = depth = None
    else:
        rgb = torch.empty(N, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if sink is not None:
            _lib.check(lib.mvsn_render_rays_to_peers(C.byref(sc), C.byref(rp), _lib.ptr(rays), _lib.ptr(_tsteps[tk]), N, S,
                                                     C.byref(sink), _lib.ptr(rgb), _lib.ptr(depth), _lib.stream_ptr()),
                       "mvsn_render_rays_to_peers")
        else:
            _lib.check(lib.mvsn_render_rays(C.byref(sc), C.byref(rp), _lib.ptr(rays), _lib.ptr(_tsteps[tk]), N, S,
                                            _lib.ptr(rgb), _lib.ptr(depth), None, None, None, _lib.stream_ptr()),
                       "mvsn_render_rays")
    del keep
    return rgb, depth


def get_rays(directions, c2w, near, far, out=None):
    """data/ray_utils.py:32-53 (get_rays) + the notebooks' `cat([rays_o, rays_d, near, far])` in one launch:
    directions [H,W,3] or [n,3] (camera frame, on the device), c2w [3,4] / [4,4] on the device -> rays [n,8]."""
    lib = _lib.load()
    d = _lib.dev_f32(directions.reshape(-1, 3), "directions")
    m = _lib.dev_f32(c2w, "c2w")
    if m.dim() != 2 or m.shape[1] != 4 or m.shape[0] < 3 or m.device != d.device:
        raise RuntimeError(f"get_rays: c2w must be [3,4] or [4,4] on {d.device}, got {tuple(m.shape)} on {m.device}")
    n = d.shape[0]
    rays = torch.empty(n, 8, dtype=torch.float32, device=d.device) if out is None else out
    with torch.cuda.device(d.device):
        _lib.check(lib.mvsn_make_rays(_lib.ptr(d), _lib.ptr(m), float(near), float(far), n, _lib.ptr(rays), _lib.stream_ptr()),
                   "mvsn_make_rays")
    return rays


class HostFrameRenderer:
    """Host-buffer entry: rays arrive in (pinned) host memory, pixels are returned in host memory.

    This is the call the notebooks' frame loop makes in effect (`rgb.cpu()` / `depth_pred.cpu()` per
    chunk, renderer_video.ipynb DTU cell) collapsed to one H2D copy, one launch, one D2H copy."""

    def __init__(self, n_rays, device):
        self.n, self.device = int(n_rays), torch.device(device)
        self.rays_dev = torch.empty(self.n, 8, dtype=torch.float32, device=self.device)
        self.rgb_dev = torch.empty(self.n, 3, dtype=torch.float32, device=self.device)
        self.depth_dev = torch.empty(self.n, dtype=torch.float32, device=self.device)
        self.rgb_host = torch.empty(self.n, 3, dtype=torch.float32).pin_memory()
        self.depth_host = torch.empty(self.n, dtype=torch.float32).pin_memory()
        self.h2d_bytes = self.n * 8 * 4
        self.d2h_bytes = self.n * 4 * 4

    def render(self, rays_host, volume_feature, imgs, pose_ref, network_fn, near_far, pad, after_launch=None, **kw):
        """`after_launch`: optional callable enqueued between the launch and the read-back (multi-GPU frame assembly:
        distributed.PeerFrame.complete, or an all-gather); `sink=` is forwarded to render_rays."""
        if rays_host.is_cuda or tuple(rays_host.shape) != (self.n, 8):
            raise RuntimeError(f"HostFrameRenderer: expected a host tensor [{self.n}, 8]")
        self.rays_dev.copy_(rays_host, non_blocking=True)
        render_rays(self.rays_dev, volume_feature, imgs, pose_ref, network_fn, near_far, pad,
                    out=(self.rgb_dev, self.depth_dev), **kw)
        if after_launch is not None:
            after_launch()
        self.rgb_host.copy_(self.rgb_dev, non_blocking=True)
        self.depth_host.copy_(self.depth_dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.rgb_host, self.depth_host

    def render_camera(self, c2w_host, directions_dev, volume_feature, imgs, pose_ref, network_fn, near_far, pad,
                      ray_near_far=None, after_launch=None, **kw):
        """The notebooks' frame loop as it is actually fed (renderer_video.ipynb: one `c2w` per frame, `get_rays` on the
        device): the host input of a frame is the camera pose (pinned [4,4] or [3,4] tensor, 48-64 bytes H2D), the rays
        are generated on the device (mvsn_make_rays) from the resident `directions_dev` [H*W,3], pixels come back to
        pinned host memory."""
        if c2w_host.is_cuda or directions_dev.reshape(-1, 3).shape[0] != self.n:
            raise RuntimeError(f"HostFrameRenderer.render_camera: expected a host c2w and {self.n} device directions")
        if not hasattr(self, "c2w_dev"):
            self.c2w_dev = torch.empty(4, 4, dtype=torch.float32, device=self.device)
        rows = c2w_host.shape[0]
        self.c2w_dev[:rows].copy_(c2w_host, non_blocking=True)
        nf = near_far if ray_near_far is None else ray_near_far
        get_rays(directions_dev, self.c2w_dev, nf[0], nf[1], out=self.rays_dev)
        render_rays(self.rays_dev, volume_feature, imgs, pose_ref, network_fn, near_far, pad,
                    out=(self.rgb_dev, self.depth_dev), **kw)
        if after_launch is not None:
            after_launch()
        self.rgb_host.copy_(self.rgb_dev, non_blocking=True)
        self.depth_host.copy_(self.depth_dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.rgb_host, self.depth_host


# --------------------------------------------------------------------------------------------
# factory (models.py:569-654)
# --------------------------------------------------------------------------------------------
def _embed(x, n_freqs=10):
    """models.py:47-51 (used only by the PyTorch query function below)."""
    freqs = 2.0 ** torch.arange(n_freqs, dtype=x.dtype, device=x.device)
    s = (x.unsqueeze(-2) * freqs.view(-1, 1)).reshape(*x.shape[:-1], -1)
    return torch.cat([x, torch.sin(s), torch.cos(s)], -1)


def _network_query(pts, viewdirs, rays_feats, network_fn, netchunk=1024):
    """run_network_mvs (renderer.py:42-63) for the callers outside `rendering` (alpha-only queries)."""
    x = _embed(pts)
    if rays_feats is not None:
        x = torch.cat([x, rays_feats], -1)
    if viewdirs is not None:
        if viewdirs.dim() != 3:
            viewdirs = viewdirs[:, None].expand(-1, x.shape[1], -1)
        x = torch.cat([x, viewdirs], -1)
    fn = network_fn.forward_alpha if viewdirs is None else network_fn
    return torch.cat([fn(x[i:i + netchunk]) for i in range(0, x.shape[0], netchunk)], 0)


def create_nerf_mvs(args, pts_embedder=True, use_mvs=False, dir_embedder=True, device=None):
    """Same contract as models.create_nerf_mvs: returns
    (render_kwargs_train, render_kwargs_test, start, grad_vars)."""
    if not pts_embedder or dir_embedder:
        raise RuntimeError("create_nerf_mvs: the fused kernel implements pts_embedder=True, dir_embedder=False "
                           "(the combination every shipped caller uses)")
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("create_nerf_mvs: no CUDA device; mvsnerf_b200 has no CPU path")
        device = torch.device("cuda", torch.cuda.current_device())
    model = MVSNeRF(D=args.netdepth, W=args.netwidth, input_ch_pts=args.pts_dim * (1 + 2 * args.multires),
                    input_ch_views=args.dir_dim, input_ch_feat=args.feat_dim, net_type=args.net_type).to(device)
    grad_vars = list(model.parameters())
    if getattr(args, "N_importance", 0) > 0:
        raise RuntimeError("N_importance > 0 (hierarchical sampling) is outside the hot path (SURVEY.md 2.1)")
    encoding_net = None
    if use_mvs:
        encoding_net = MVSNet().to(device)
        # models.py:622 adds the encoder's parameters here; this encoder is forward-only (no gradients reach it,
        # see MVSNet.forward), so handing them to the optimiser would only pretend to train them.
    ckpt_path = getattr(args, "ckpt", None)
    if ckpt_path is not None and ckpt_path != "None":
        ckpt = torch.load(ckpt_path, map_location=device, weights_only=False)
        if use_mvs:
            encoding_net.load_state_dict(ckpt["network_mvs_state_dict"])
        model.load_state_dict(ckpt["network_fn_state_dict"])
    render_kwargs_train = {
        "network_query_fn": lambda pts, viewdirs, rays_feats, network_fn: _network_query(
            pts, viewdirs, rays_feats, network_fn, args.netchunk),
        "perturb": args.perturb, "N_importance": args.N_importance, "network_fine": None,
        "N_samples": args.N_samples, "network_fn": model, "network_mvs": encoding_net,
        "use_viewdirs": args.use_viewdirs, "white_bkgd": args.white_bkgd, "raw_noise_std": args.raw_noise_std,
    }
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test["perturb"] = False
    return render_kwargs_train, render_kwargs_test, 0, grad_vars


def load_weights_npz(model_fn: MVSNeRF | None, model_mvs: MVSNet | None, path: str):
    """Load the `mlp/` and `mvs/` tensors of tests/golden/mvsnerf_v0_weights.npz (an export of
    ckpts/mvsnerf-v0.tar that travels with the repository)."""
    import numpy as np
    z = np.load(path)
    if model_fn is not None:
        model_fn.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("mlp/")})
    if model_mvs is not None:
        model_mvs.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("mvs/")})

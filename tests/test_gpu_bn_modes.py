"""GPU: the `self.training` dispatch of the encoder's BatchNorm ("implement both"; InPlaceABN,
models.py:661-685).  eval mode = running statistics (pinned against the live reference's MVSNet.eval() in
tests/test_oracle_pins.py through the oracle's eval_mode); train mode = batch statistics PLUS the in-place update of
running_mean / running_var / num_batches_tracked that F.batch_norm(training=True) performs in the reference."""
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import mvsnerf_oracle as orc
from mvsnerf_b200 import backend, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
WPATH = os.path.join(GOLDEN, "mvsnerf_v0_weights.npz")


def test_eval_mode_uses_running_statistics(weights):
    sc = synthetic.make_scene(64, 96, pad=4, seed=12)
    d = sc.to(DEV)
    mvs = backend.MVSNet().to(DEV)
    backend.load_weights_npz(None, mvs, WPATH)
    before = {k: v.clone() for k, v in mvs.state_dict().items()}
    with torch.no_grad():
        vol_eval, feats, _ = mvs.eval()(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    for k, v in mvs.state_dict().items():                       # eval mode touches no buffer
        assert torch.equal(v, before[k]), k
    ref = orc.encode_volume(sc.imgs_norm, sc.proj_mats, sc.near_far, sc.pad, weights, eval_mode=True)
    assert (vol_eval.cpu() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    f_ref = orc.feature_net(sc.imgs_norm[0], weights, eval_mode=True)
    assert (feats[0].cpu() - f_ref).abs().max().item() <= 1e-4 * f_ref.abs().max().item()
    # and it is a different function from train mode with this checkpoint (volume Linf ~7)
    with torch.no_grad():
        vol_train, _, _ = mvs.train()(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    assert (vol_train - vol_eval).abs().max().item() > 0.1


def test_train_mode_updates_running_statistics_like_batch_norm(weights):
    sc = synthetic.make_scene(64, 96, pad=4, seed=13)
    d = sc.to(DEV)
    mvs = backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(None, mvs, WPATH)
    rec = {}
    ref = orc.encode_volume(sc.imgs_norm, sc.proj_mats, sc.near_far, sc.pad, weights, record=rec)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    assert (vol.cpu() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    sd = mvs.state_dict()
    assert len(rec) == 18
    for name, (mean, var_unbiased) in rec.items():
        key = name[len("mvs/"):]
        want_m = 0.9 * weights[name + ".running_mean"] + 0.1 * mean
        want_v = 0.9 * weights[name + ".running_var"] + 0.1 * var_unbiased
        assert (sd[key + ".running_mean"].cpu() - want_m).abs().max().item() <= 1e-4 * max(1.0, want_m.abs().max().item()), name
        assert (sd[key + ".running_var"].cpu() - want_v).abs().max().item() <= 1e-4 * max(1.0, want_v.abs().max().item()), name
        assert int(sd[key + ".num_batches_tracked"]) == int(weights[name + ".num_batches_tracked"]) + 1


def test_both_modes_against_the_reference_fixture(golden_bn_modes):
    """The same sequence the fixture recorded from the unmodified reference (tests/golden/make_golden_bn.py): a
    train-mode forward from the shipped state, the running statistics it leaves, then an eval-mode forward."""
    import collections
    g = golden_bn_modes
    Scene = collections.namedtuple("Scene", "imgs proj nf pad")
    sc = Scene(g["imgs_norm"].to(DEV), g["proj_mats"].to(DEV), g["near_far"].tolist(), int(g["HW_pad"][2]))
    idx = g["vox_idx"].to(DEV)
    mvs = backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(None, mvs, WPATH)
    with torch.no_grad():
        vol, _, _ = mvs(sc.imgs, sc.proj, sc.nf, pad=sc.pad)
    want = g["volume_train_sub"].to(DEV)
    assert (vol[0].reshape(8, -1)[:, idx] - want).abs().max().item() <= 2e-4 * max(1.0, want.abs().max().item())
    n = 0
    for k, v in mvs.state_dict().items():
        if not k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            continue
        ref = g["stats/" + k].to(DEV)
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(ref), k
        else:
            assert (v - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), k
        n += 1
    assert n == 54
    with torch.no_grad():
        vol_e, feats, _ = mvs.eval()(sc.imgs, sc.proj, sc.nf, pad=sc.pad)
    want = g["volume_eval_sub"].to(DEV)
    assert (vol_e[0].reshape(8, -1)[:, idx] - want).abs().max().item() <= 2e-4 * max(1.0, want.abs().max().item())
    chsum = vol_e[0].double().sum((1, 2, 3)).cpu()
    assert torch.allclose(chsum, g["volume_eval_chsum"], rtol=1e-4, atol=0.5)
    fe = g["feats_eval"].to(DEV)
    assert (feats[0] - fe).abs().max().item() <= 1e-4 * fe.abs().max().item()


@pytest.mark.parametrize("D,Hp,Wp", [(16, 24, 40), (24, 16, 64), (16, 40, 72), (32, 48, 56)])
def test_conv0_tensor_core_kernel_vs_ffma_kernel(D, Hp, Wp):
    """conv0 on tcgen05 (csrc/conv0_tc.cu: TMA-staged tiles, 2-term fp16 split, shifted-tap epilogue) against the
    round-1 FFMA kernel through the whole CostRegNet, on volumes whose sizes are not multiples of the 6 x 10 x 30 brick
    (partial bricks in every direction), and bit-identical from run to run.  (Volumes whose coarsest level has only a
    voxel or two are left out: BatchNorm over one voxel divides by sqrt(eps) and turns 1e-6 into 1e-3.)"""
    import ctypes as C
    from mvsnerf_b200 import lib
    L = lib.load()
    mvs = backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(None, mvs, WPATH)
    reg = mvs.cost_reg_2
    g = torch.Generator(device=DEV).manual_seed(D * 1000 + Wp)
    cost = torch.randn(1, 41, D, Hp, Wp, device=DEV, generator=g) * torch.linspace(0.05, 30.0, 41, device=DEV).view(1, 41, 1, 1, 1)
    weights = [w.detach().contiguous() for w in reg.weight_list()]
    ws_bytes = L.mvsn_costreg_workspace_bytes(D, Hp, Wp)
    outs = []
    for flags in (lib.BN_BATCH, lib.BN_BATCH, lib.BN_BATCH | lib.CONV0_FFMA):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
        vol = torch.empty(D, Hp, Wp, 8, device=DEV)
        lib.check(L.mvsn_costreg_forward_bn(lib.ptr_array(weights), None, flags, 0.0, lib.ptr(cost), D, Hp, Wp, lib.ptr(vol),
                                            lib.ptr(ws), ws_bytes, lib.stream_ptr()), "mvsn_costreg_forward_bn")
        outs.append(vol)
    assert torch.equal(outs[0], outs[1])                                   # deterministic
    scale = outs[2].abs().max().item()
    assert (outs[0] - outs[2]).abs().max().item() <= 2e-5 * scale, ((outs[0] - outs[2]).abs().max().item(), scale)

"""CPU: the C-ABI library builds/loads and exports every declared symbol; host-side mirror logic
(state_dict compatibility with the reference checkpoint, loud failures without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT
from mvsnerf_b200 import backend, lib, synthetic


@pytest.fixture(scope="module")
def built():
    from mvsnerf_b200 import build
    return build.build_library()


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "mvsnerf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mvsn_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(lib.EXPORTS)
    dll = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(dll, name), name
    L = lib.load()
    assert L.mvsn_abi_version() == 1
    assert L.mvsn_mlp_packed_bytes(lib.MLP_FP32) > 126788 * 4
    assert L.mvsn_costreg_workspace_bytes(128, 176, 208) > 400e6
    assert L.mvsn_cost_volume_workspace_bytes(3, 128, 160) == 3 * 3 * 128 * 160 * 4
    assert L.mvsn_featurenet_workspace_bytes(3, 512, 640) == 4096 + 2 * 3 * 8 * 512 * 640 * 4


def test_argument_errors_do_not_need_a_gpu(built):
    L = lib.load()
    rc = L.mvsn_pack_images(None, 3, 4, 4, None, None)
    assert rc == -4 and b"NULL" in L.mvsn_last_error()
    rc = L.mvsn_costreg_forward(None, None, 128, 20, 24, None, None, 0, None)
    assert rc == -4
    rc = L.mvsn_featurenet_forward(None, None, 3, 32, 32, None, None, 0, None)
    assert rc == -4 and b"featurenet" in L.mvsn_last_error()
    assert L.mvsn_mlp_packed_bytes(99) == 0


def test_state_dict_keys_match_reference_checkpoint():
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))
    fn, mvs = backend.MVSNeRF(), backend.MVSNet()
    assert sorted(fn.state_dict().keys()) == sorted(k[4:] for k in z.files if k.startswith("mlp/"))
    assert sorted(mvs.state_dict().keys()) == sorted(k[4:] for k in z.files if k.startswith("mvs/"))
    backend.load_weights_npz(fn, mvs, os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))   # strict load
    assert len(fn.ordered_params()) == lib.N_MLP_TENSORS
    assert len(mvs.cost_reg_2.weight_list()) == lib.N_COSTREG_TENSORS
    assert len(mvs.feature.weight_list()) == lib.N_FEATURENET_TENSORS
    assert mvs.feature.weight_list()[6].shape == (16, 8, 5, 5) and mvs.feature.weight_list()[24].shape == (32, 32, 1, 1)
    assert sum(p.numel() for p in fn.parameters()) == 126788


def test_pytorch_mirror_of_mlp_matches_oracle(weights):
    """MVSNeRF.forward (kept for the alpha-only callers) is the same function as the oracle MLP."""
    from oracle import mvsnerf_oracle as orc
    fn = backend.MVSNeRF()
    backend.load_weights_npz(fn, None, os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))
    x = torch.rand(50, 7, 86, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        assert (fn(x) - orc.mlp(x, weights)).abs().max() < 1e-6


def test_no_silent_cpu_fallback():
    fn, mvs = backend.MVSNeRF(), backend.MVSNet().train()
    sc = synthetic.make_scene(32, 32, pad=4, seed=0)
    with pytest.raises(RuntimeError):
        mvs(sc.imgs_norm, sc.proj_mats, sc.near_far, pad=4)
    with pytest.raises(RuntimeError):
        fn.packed()
    with pytest.raises(RuntimeError):
        backend.render_rays(synthetic.scene_rays(sc), torch.zeros(1, 8, 128, 16, 16), sc.imgs_raw, sc.pose_source,
                            fn, sc.near_far, 4.0)
    with pytest.raises(RuntimeError):
        backend.MVSNet().eval()(sc.imgs_norm, sc.proj_mats, sc.near_far)


def test_synthetic_scene_contract():
    sc = synthetic.make_scene(64, 96, pad=4, seed=3)
    assert sc.imgs_norm.shape == (1, 3, 3, 64, 96) and sc.proj_mats.shape == (1, 3, 3, 4)
    assert torch.allclose(sc.proj_mats[0, 0], torch.eye(4)[:3])
    rays = synthetic.scene_rays(sc)
    assert rays.shape == (64 * 96, 8) and float(rays[0, 6]) == pytest.approx(sc.near_far[0])
    sc2 = synthetic.make_scene(64, 96, pad=4, seed=3)
    assert torch.equal(sc.imgs_raw, sc2.imgs_raw)

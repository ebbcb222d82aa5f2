"""Generate tests/golden/bn_modes_64x96_pad4.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_bn.py

The reference's BatchNorm (InPlaceABN, models.py:661-685) has two behaviours selected by `self.training`:
train mode normalises with batch statistics AND updates running_mean / running_var / num_batches_tracked in place;
eval mode normalises with the running statistics.  This fixture records both on one seeded scene, in the order a
user meets them: one train-mode forward from the shipped initial state (the statistics it leaves behind are recorded),
then an eval-mode forward with those statistics.  The weights are the ones in mvsnerf_v0_weights.npz (same seed).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from mvsnerf_b200 import synthetic  # noqa: E402
from make_golden import scene_inputs  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    R = ref_shim.build_reference()
    shipped = np.load(os.path.join(HERE, "mvsnerf_v0_weights.npz"))
    for k, v in R.mvsnet.state_dict().items():
        assert np.array_equal(shipped["mvs/" + k], v.detach().cpu().numpy()), k     # same initial state as the fixture

    sc = synthetic.make_scene(64, 96, pad=4, seed=12)
    m = R.mvsnet
    with torch.no_grad():
        m.train()
        vol_train, _, _ = m(sc.imgs_norm, sc.proj_mats, sc.near_far, pad=sc.pad)
        stats = {k: v.detach().clone() for k, v in m.state_dict().items()
                 if k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
        m.eval()
        B, V, _, H, W = sc.imgs_norm.shape
        feats_eval = m.feature(sc.imgs_norm.reshape(B * V, 3, H, W))
        vol_eval, _, _ = m(sc.imgs_norm, sc.proj_mats, sc.near_far, pad=sc.pad)
        m.train()
    for k, v in m.state_dict().items():                                               # eval mode left them alone
        if k in stats:
            assert torch.equal(v, stats[k]), k

    g = torch.Generator().manual_seed(11)
    nvox = vol_train[0, 0].numel()
    vox_idx = torch.randperm(nvox, generator=g)[:8192]
    out = scene_inputs(sc)
    out.update(vox_idx=vox_idx.numpy(),
               volume_train_sub=vol_train[0].reshape(8, -1)[:, vox_idx].numpy(),
               volume_train_chsum=vol_train[0].double().sum((1, 2, 3)).numpy(),
               volume_eval_sub=vol_eval[0].reshape(8, -1)[:, vox_idx].numpy(),
               volume_eval_chsum=vol_eval[0].double().sum((1, 2, 3)).numpy(),
               feats_eval=feats_eval.numpy())
    for k, v in stats.items():
        out["stats/" + k] = v.numpy()
    path = os.path.join(HERE, "bn_modes_64x96_pad4.npz")
    np.savez_compressed(path, **out)
    print("bn modes:", len(stats), "statistics tensors;", os.path.getsize(path) // 1024, "KiB;",
          "train/eval volume Linf", (vol_train - vol_eval).abs().max().item())


if __name__ == "__main__":
    main()

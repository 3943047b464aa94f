"""GPU: the reference's evaluation loop (eval.py:183-366) on this package — examples/eval_from_folder.py."""
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_eval_loop_on_a_reference_format_folder_with_a_subdivided_checkpoint(tmp_path, golden_dir):
    import eval_from_folder as EF
    import train_from_folder as TF
    from gomavatar_b200 import dataset_io as IO
    from gomavatar_b200.lpips import LPIPS, seeded_random_trunk
    from oracle import losses as OL
    dev = torch.device("cuda:0")
    data = str(tmp_path / "subject")
    torch.manual_seed(0)
    TF.write_synthetic_subject(data, 2000, 64, 5, dev)
    ck = os.path.join(data, "checkpoints")
    TF.train(data, 30, 64, dev, ckpt_dir=ck, save_freq=30, lr=5e-3, subdivide_iters=(10,))
    heads = np.load(os.path.join(golden_dir, "golden_lpips.npz"))
    lp = LPIPS(seeded_random_trunk(0), [heads[f"lin{k}"] for k in range(5)]).to(dev)
    summary, per_frame, save_dir = EF.evaluate(data, os.path.join(ck, "iter_30.pt"), 64, dev, data, batch=2, n_subdivisions=1, lpips=lp)
    # result files in the reference's layout and format (eval.py:131-143, :365)
    names = sorted(os.listdir(save_dir))
    assert len(names) == 5 and all(n.endswith(".png") for n in names)
    res = np.load(os.path.join(data, "eval", "metric_train.npy"), allow_pickle=True).item()
    assert sorted(res) == ["lpips", "mse", "psnr", "ssim"] and all(len(res[k]) == 5 for k in res)
    assert res["psnr"] == per_frame["psnr"] and abs(summary["psnr"] - np.mean(res["psnr"])) < 1e-9
    assert all(np.isfinite(v) for k in res for v in res[k]) and summary["psnr"] > 5 and 0 < summary["ssim"] <= 1
    # every number equals the reference's per-frame recipe on the PNG that was written and the dataset's target:
    # to_8b_image both, / 255, mse -> psnr, skimage SSIM (oracle restatement), eval.py:101-128
    ds = IO.Dataset(data, bgcolor=[0.0, 0.0, 0.0], target_size=[64, 64])
    for i, name in enumerate(ds.framelist):
        pred = np.asarray(Image.open(os.path.join(save_dir, name + ".png"))).astype(np.float64) / 255.0
        gt = (255.0 * np.clip(ds[i]["target_rgbs"], 0.0, 1.0)).astype(np.uint8).astype(np.float64) / 255.0
        mse = np.mean((pred - gt) ** 2)
        assert abs(res["mse"][i] - mse) <= 1e-12 and abs(res["psnr"][i] + 10 * np.log(mse) / np.log(10)) <= 1e-9
        assert abs(res["ssim"][i] - OL.ssim(pred, gt)) <= 1e-9

"""CPU: the oracle's restatement of the reference's reconstruction metrics and image transforms, pinned to (a) golden values
recorded from the UNMODIFIED /root/reference/train/train_utils.py (oracle/make_golden_metrics.py -> tests/golden/metrics_*.npz)
and (b) torchvision's own transforms for the image pipeline."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import metrics as M
from oracle.make_golden_metrics import inputs

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "metrics_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_metrics_oracle_matches_reference_goldens(path):
    g = np.load(path)
    N, C, H, W, seed = (int(v) for v in g["shape"])
    a, b = inputs(N, C, H, W, seed)
    assert abs(float(M.pearson(a, b)) - float(g["pcc"])) < 1e-6
    assert abs(float(M.pearson(a.double(), b.double())) - float(g["pcc64"])) < 1e-12
    assert abs(float(M.ssim(a, b)) - float(g["ssim"])) < 1e-6
    assert abs(float(torch.nn.functional.mse_loss(a, b)) - float(g["mse"])) < 1e-7


def test_image_pipeline_oracle_matches_torchvision_and_scipy():
    """ToTensor + Normalize + horizontal flip against torchvision, integer shift against scipy.ndimage.shift(order=0,
    mode='nearest') as RandomShift calls it (data_preprocessing/data_loader.py:213-216)."""
    from scipy.ndimage import shift as nd_shift
    from torchvision import transforms as T

    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (4, 20, 24, 3), generator=g, dtype=torch.uint8)
    mean, std = (0.5, 0.4, 0.3), (0.5, 0.25, 0.2)
    flip = torch.tensor([0, 1, 1, 0], dtype=torch.int32)
    sh = torch.tensor([[0, 0], [2, -3], [-5, 1], [4, 4]], dtype=torch.int32)
    got = M.image_pipeline(u8, flip, sh, mean, std)
    for n in range(4):
        img = u8[n].numpy()
        img = nd_shift(img, [int(sh[n, 0]), int(sh[n, 1]), 0], prefilter=False, order=0, mode="nearest")
        t = torch.from_numpy(img).permute(2, 0, 1).float() / 255.0         # ToTensor
        if flip[n]:
            t = T.functional.hflip(t)
        t = T.Normalize(mean, std)(t)
        assert torch.allclose(got[n], t, atol=1e-6), n
    grey = torch.randint(0, 256, (2, 8, 8, 1), generator=g, dtype=torch.uint8)
    out = M.image_pipeline(grey, None, None, (0.5,) * 3, (0.5,) * 3)
    assert torch.equal(out[:, 0], out[:, 1]) and torch.equal(out[:, 0], out[:, 2])      # GreyToColor: replicated plane
    assert torch.allclose(out[:, 0], (grey[..., 0].float() / 255 - 0.5) / 0.5)

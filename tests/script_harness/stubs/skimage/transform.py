"""skimage.transform.resize(image, (H, W)): uint8 input is converted to float in [0, 1] (img_as_float), output is float64
H x W x C, bilinear with anti-aliasing -- the behaviour the reference's Rescale relies on. Implemented with torch."""
import numpy as np
import torch
import torch.nn.functional as F


def resize(image, output_shape, **kwargs):
    a = np.asarray(image)
    if a.dtype == np.uint8:
        a = a.astype(np.float64) / 255.0
    else:
        a = a.astype(np.float64)
    squeeze = a.ndim == 2
    if squeeze:
        a = a[:, :, None]
    t = torch.from_numpy(a).permute(2, 0, 1)[None]
    t = F.interpolate(t, size=tuple(int(v) for v in output_shape[:2]), mode="bilinear", align_corners=False, antialias=True)
    out = t[0].permute(1, 2, 0).numpy()
    return out[:, :, 0] if squeeze else out

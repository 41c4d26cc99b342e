"""matplotlib.image.imread via PIL: H x W x C uint8 for JPEG (what matplotlib returns for non-PNG files)."""
import numpy as np
from PIL import Image


def imread(fname, format=None):
    with Image.open(fname) as im:
        return np.asarray(im)

"""`configs` package of the drop-in surface. This repository ships ONE module of it, `configs.models_config` (the
architecture constants `models/vae_gan.py` reads); the reference's other config modules -- `configs.gan_config`,
`configs.wae_config`, `configs.data_config`, `configs.inference_config`, `configs.vae_gan_config`, which its unchanged
train / inference scripts import (/root/reference/train/train_wae_stage1.py:20, inference/inference_gan.py:19-20) -- are
NOT rebuilt here. So that `PYTHONPATH=<this repo>:<reference>` works, the package path is extended with every other
`configs/` directory on sys.path: `configs.models_config` resolves here (first entry), everything else in the reference.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

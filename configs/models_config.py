"""Architecture constants read by ``models/vae_gan.py`` at module-construction time.

Same attribute names as the reference's ``configs/models_config.py`` (it is imported as
``import configs.models_config as config`` and read attribute-by-attribute, reference vae_gan.py:8,18-19,74,79,107,
112-119,146-160), so scripts that tweak ``config.<name>`` keep working. The two resolution presets the reference
ships (active 100x100 / latent 512 block at models_config.py:13-21, commented 64x64 / latent 128 block at :23-31) are
both available through ``use_resolution``; the module-level defaults are the reference's active values.
"""

_COMMON = dict(kernel_size=5, stride=2, padding=2, dropout=0.7,
               encoder_channels=[64, 128, 256], discrim_channels=[32, 128, 256, 256, 512], fc_output=1024)

PRESETS = {
    100: dict(image_size=100, fc_input=13, fc_input_gan=7, fc_output_gan=256, stride_gan=2, latent_dim=512,
              output_pad_dec=[False, True, True], decoder_channels=[256, 128, 64, 3]),
    64: dict(image_size=64, fc_input=8, fc_input_gan=8, fc_output_gan=512, stride_gan=1, latent_dim=128,
             output_pad_dec=[True, True, True], decoder_channels=[256, 128, 32, 3]),
}


def use_resolution(size):
    """Switch every architecture constant to the preset for ``size`` (64 or 100). Affects modules built afterwards."""
    if size not in PRESETS:
        raise ValueError(f"no preset for image_size={size}; known: {sorted(PRESETS)}")
    g = globals()
    g.update({k: (list(v) if isinstance(v, list) else v) for k, v in _COMMON.items()})
    g.update({k: (list(v) if isinstance(v, list) else v) for k, v in PRESETS[size].items()})
    return size


import os as _os

# The reference selects the resolution by editing this file (active block :13-21 = 100x100 / latent 512; commented block
# :23-31 = 64x64 / latent 128). Here the default is the reference's active block, and FMRI_MODELS_CONFIG=64 selects the other
# one for an unchanged script (the WAE scripts' defaults, configs/wae_config.py:6-7, are the 64x64 ones).
use_resolution(int(_os.environ.get("FMRI_MODELS_CONFIG", "100")))

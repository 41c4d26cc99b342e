"""Stand-in for nibabel (absent from this image). The reference imports it at the top of data_preprocessing/data_loader.py;
only the raw-NIfTI loader Bold5000Dataloader (not used by the train / inference scripts) calls it."""
__version__ = "0.0-stub"


def load(path):
    raise RuntimeError("nibabel stub: raw NIfTI loading is outside the script harness")

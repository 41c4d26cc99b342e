"""Stand-in for h5py (absent from this image); imported by the reference's ROI extraction only."""
__version__ = "0.0-stub"


def File(*a, **k):
    raise RuntimeError("h5py stub: HDF5 files are outside the script harness")

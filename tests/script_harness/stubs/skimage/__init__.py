"""Stand-in for scikit-image (absent from this image): only skimage.transform.resize, which the reference's Rescale
transform calls (data_preprocessing/data_loader.py:113-131). Test infrastructure only."""
__version__ = "0.0-stub"

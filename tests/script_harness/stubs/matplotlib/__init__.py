"""Minimal stand-in for matplotlib (absent from this image) so that the reference's UNCHANGED scripts import and run:
every plotting call is accepted and ignored; matplotlib.image.imread really reads the file (PIL). Test infrastructure only
(tests/script_harness). Importing it also registers the harness's exit report: how many kernels of libfmri_b200.so the
process launched (FMRI_HARNESS_REPORT=<json path>), which is how the tests know the script ran on the product path."""
import atexit
import json
import os
import sys

__version__ = "0.0-stub"


def use(*a, **k):
    pass


def _report():
    path = os.environ.get("FMRI_HARNESS_REPORT")
    if not path:
        return
    rep = dict(lib_loaded=False, launches=0, models_file=None)
    lib = sys.modules.get("thesis_fmri_reconstruction_b200.lib")
    if lib is not None and getattr(lib, "_lib", None) is not None:
        rep["lib_loaded"] = True
        rep["launches"] = int(lib.launch_count())
    m = sys.modules.get("models.vae_gan")
    rep["models_file"] = getattr(m, "__file__", None)
    c = sys.modules.get("configs.models_config")
    rep["models_config_file"] = getattr(c, "__file__", None)
    with open(path, "w") as f:
        json.dump(rep, f)


atexit.register(_report)

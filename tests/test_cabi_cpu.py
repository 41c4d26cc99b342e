"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/fmri_b200.h declares, reports its ABI version, and refuses CPU tensors loudly (no fallback path)."""
import ctypes
import os
import re

import pytest
import torch

from thesis_fmri_reconstruction_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_present_and_loads():
    assert os.path.exists(L.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    L.load()


def test_every_header_symbol_is_exported():
    names = L.header_functions()
    assert len(names) >= 40, names
    c = ctypes.CDLL(L.LIB_PATH)
    missing = [n for n in names if not hasattr(c, n)]
    assert not missing, missing


def test_abi_version_matches_header():
    src = open(L.HEADER_PATH).read()
    want = int(re.search(r"#define\s+FMRI_ABI_VERSION\s+(\d+)", src).group(1))
    assert L.load().fmri_version() == want


def test_no_torch_types_in_the_abi():
    src = open(L.HEADER_PATH).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # comments may cite torch call sites; declarations may not use torch types
    assert "at::" not in code and "torch" not in code.lower() and "#include <torch" not in src
    assert 'extern "C"' in src


def test_tensor_path_query_without_gpu_is_clean():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.load().fmri_tensor_path_available() == 0


def test_cpu_tensors_are_refused():
    x = torch.zeros(4, 8)
    with pytest.raises(L.FmriError):
        L.colsum(x, 4, 8, torch.zeros(8))


def test_conv_geometry_helper():
    # Conv2d: OH = (H-1)/s + 1 ; ConvTranspose2d: OH = 2H - 1 + output_pad   (vae_gan.py:18-20, 46-53)
    for H, s, want in ((64, 2, 32), (25, 2, 13), (13, 2, 7), (64, 1, 64), (100, 2, 50)):
        d = L.conv_desc(1, H, H, 32, 32, s, False, 0, torch.float32)
        assert L.conv_out_hw(d) == (want, want)
    for H, op, want in ((8, 1, 16), (13, 0, 25), (25, 1, 50), (50, 1, 100)):
        d = L.conv_desc(1, H, H, 32, 32, 2, True, op, torch.float32)
        assert L.conv_out_hw(d) == (want, want)


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "thesis_fmri_reconstruction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") and f != "smoke.py":
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
    for f in ("vae_gan.py",):
        p = os.path.join(ROOT, "models", f)
        if os.path.exists(p):
            assert not re.search(r"^\s*(from|import)\s+oracle", open(p).read(), flags=re.M)


def test_epoch_end_schedules_match_torch_schedulers():
    """hp.epoch_end_vgan / epoch_end_wae against torch.optim.lr_scheduler (ExponentialLR, StepLR) and the margin /
    equilibrium / lambda_mse rules of train/train_vgan_stage1.py:446-457."""
    import torch

    from thesis_fmri_reconstruction_b200 import hp as H

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.RMSprop([p], lr=1e-4)
    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.98)
    lr, h = {"encoder.": 1e-4}, dict(H.HP_VGAN)
    for _ in range(5):
        opt.step(); sch.step()
        H.epoch_end_vgan(lr, h, decay_lr=0.98, decay_margin=1.5, decay_equilibrium=0.9, decay_mse=3.0)
    assert abs(lr["encoder."] - opt.param_groups[0]["lr"]) < 1e-18
    assert h["equilibrium"] >= h["margin"] and h["lambda_mse"] <= 1.0
    m, e = H.HP_VGAN["margin"], H.HP_VGAN["equilibrium"]
    for _ in range(5):
        m *= 1.5; e *= 0.9
        if m > e:
            e = m
    assert abs(h["margin"] - m) < 1e-12 and abs(h["equilibrium"] - e) < 1e-12
    opt = torch.optim.Adam([p], lr=1e-3)
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=30, gamma=0.5)
    lr = {"decoder.": 1e-3}
    for epoch in range(1, 95):
        opt.step(); sch.step()
        H.epoch_end_wae(lr, epoch)
        assert abs(lr["decoder."] - opt.param_groups[0]["lr"]) < 1e-15, epoch


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No CPU / Python fallback: with the shared object absent, the first use raises FmriError naming the build command."""
    import importlib

    from thesis_fmri_reconstruction_b200 import lib as L

    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "libfmri_b200.so"))
    monkeypatch.setattr(L, "_lib", None)
    try:
        with pytest.raises(L.FmriError) as e:
            L.load()
        assert "build" in str(e.value)
    finally:
        monkeypatch.undo()
        importlib.reload(L) if L._lib is None and False else None


def test_models_refuse_cpu_execution():
    """The drop-in nn.Modules are parameter containers over the CUDA library: a forward on CPU tensors must raise, never
    silently compute with torch ops."""
    import torch

    import configs.models_config as mc
    from thesis_fmri_reconstruction_b200 import lib as L

    mc.use_resolution(64)
    from models.vae_gan import Decoder, Encoder

    with pytest.raises(L.FmriError):
        Encoder(channel_in=3, z_size=128)(torch.zeros(2, 3, 64, 64))
    with pytest.raises(L.FmriError):
        Decoder(z_size=128, size=256)(torch.zeros(2, 128))

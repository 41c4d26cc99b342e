"""The training step's left edge: batches arrive from the host, the step wants normalised fp32 NCHW images
(and fp32 voxel vectors) resident in HBM.

Reference (all under /root/reference): the DataLoader workers decode, crop and resize on the CPU, then apply per image
`RandomHorizontalFlip -> ToTensor (/255) -> GreyToColor -> Normalize(mean, std)` (train/train_vgan_stage1.py:161-171;
data_preprocessing/data_loader.py:374-400) or `RandomShift -> SampleToTensor -> Normalization` for the BOLD5000 stimuli
(train/train_vgan_stage2.py:160-168; data_loader.py:93-111, 187-217), and the loop does `x.to(device)` synchronously
(train_vgan_stage1.py:322). Here the host ships the decoded uint8 pixels (4x fewer bytes than fp32) from pinned memory on a
copy stream while the previous step computes, and ONE kernel (fmri_image_pipeline) does /255, grey -> RGB, flip, shift
and normalisation on the device.

    pipe = DevicePipeline(mean=gan_cfg.mean, std=gan_cfg.std)
    for x, extras in Prefetcher(batches, pipe):       # batches yields (uint8 [B,H,W,C] tensor, dict of other host tensors)
        trainer.step(x, ...)
"""
from __future__ import annotations

import torch

from . import lib as L


class DevicePipeline:
    """uint8 [B, H, W, C] (C = 1 or 3) device tensor -> fp32 [B, 3, H, W], normalised, optionally flipped / shifted.
    The random draws use the torch RNG calls of the reference transforms' semantics (p = 0.5 flip per image;
    integer shifts uniform in [-max_shift, max_shift] per axis), made once per batch on the host."""

    def __init__(self, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), random_flip=False, max_shift=0, generator=None):
        self.mean, self.std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self.random_flip, self.max_shift, self.gen = bool(random_flip), int(max_shift), generator

    def draw(self, B):
        """Per-image augmentation parameters (host int32 tensors, or None): flip [B], shift_yx [B, 2]."""
        flip = shift = None
        if self.random_flip:
            flip = (torch.rand(B, generator=self.gen) < 0.5).to(torch.int32)
        if self.max_shift > 0:
            shift = torch.randint(-self.max_shift, self.max_shift + 1, (B, 2), generator=self.gen, dtype=torch.int32)
        return flip, shift

    def __call__(self, u8, flip=None, shift_yx=None, out=None):
        if not u8.is_cuda or u8.dtype != torch.uint8 or u8.dim() != 4:
            raise L.FmriError("DevicePipeline expects a uint8 [B, H, W, C] tensor on the device")
        B, H, W, _ = u8.shape
        if out is None:
            out = torch.empty(B, 3, H, W, dtype=torch.float32, device=u8.device)
        f = flip.to(u8.device, torch.int32, non_blocking=True).contiguous() if flip is not None else None
        s = shift_yx.to(u8.device, torch.int32, non_blocking=True).contiguous() if shift_yx is not None else None
        L.image_pipeline(u8.contiguous(), f, s, self.mean, self.std, out)
        return out


class Prefetcher:
    """Double-buffered host -> device feed: while step i computes, batch i+1 is copied from pinned memory on a copy stream
    and run through the DevicePipeline on that stream. Iterating yields (images fp32 [B,3,H,W] on the device, extras dict on
    the device). `batches` yields (uint8 image tensor [B,H,W,C] on the host, dict of further host tensors, e.g. 'fmri')."""

    def __init__(self, batches, pipeline, device="cuda"):
        self.it, self.pipe, self.dev = iter(batches), pipeline, torch.device(device)
        self.stream = torch.cuda.Stream()
        self._next = None
        self._stage()

    def _pin(self, t):
        return t if t.is_pinned() else t.pin_memory()

    def _stage(self):
        try:
            u8, extras = next(self.it)
        except StopIteration:
            self._next = None
            return
        flip, shift = self.pipe.draw(u8.shape[0])
        with torch.cuda.stream(self.stream):
            d8 = self._pin(u8).to(self.dev, non_blocking=True)
            x = self.pipe(d8, flip, shift)
            ex = {k: self._pin(v).to(self.dev, non_blocking=True) for k, v in (extras or {}).items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (x, ex, ev, d8)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        x, ex, ev, d8 = self._next
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for t in [x, d8] + list(ex.values()):
            t.record_stream(cur)
        self._stage()          # the next batch's copy + pipeline overlap this batch's compute
        return x, ex

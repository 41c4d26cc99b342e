"""Compact, order-stable summaries of tensors for the golden fixtures (TEST INFRASTRUCTURE ONLY).

A fixture stores, per tensor: shape, L2 norm, sum, the first 8 elements and 8 elements at a fixed stride of the
flattened tensor -- enough to pin every parameter gradient of a 33.7 M-parameter model in a few hundred KB.
"""
from __future__ import annotations

import numpy as np
import torch


def summarize(t):
    t = t.detach().double().reshape(-1)
    n = t.numel()
    stride = max(1, n // 8)
    idx = (torch.arange(8) * stride + stride // 2).clamp_max(n - 1)
    return np.concatenate([[float(n), float(t.norm()), float(t.sum())], t[:8].numpy() if n >= 8 else
                           np.pad(t.numpy(), (0, 8 - n)), t[idx].numpy()]).astype(np.float64)


def summarize_dict(d, prefix):
    return {prefix + k: summarize(v) for k, v in d.items()}


def summary_error(got, want):
    """Worst deviation between two summaries: |d norm| / norm, |d sum| / (norm sqrt(n)), and the sampled elements'
    max |d| relative to max(largest sampled |element|, RMS of the tensor)."""
    assert got[0] == want[0], f"numel {got[0]} != {want[0]}"
    scale = max(abs(want[1]), 1e-300)
    n = max(want[0], 1.0)
    rms = scale / np.sqrt(n)
    e_norm = abs(got[1] - want[1]) / scale
    e_sum = abs(got[2] - want[2]) / (scale * np.sqrt(n))
    e_el = np.max(np.abs(got[3:] - want[3:])) / max(np.max(np.abs(want[3:])), rms)
    return max(e_norm, e_sum, e_el)

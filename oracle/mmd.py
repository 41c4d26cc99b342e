"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU statement of the WAE-MMD latent penalty.

PARITY UNPINNED: the reference ships no MMD code at all (its WAE scripts use the latent discriminator,
train/train_wae_stage1.py:271-311). The north star names an MMD loss, so the kernels fmri_mmd_imq_{fwd,bwd} are an
extension checked against THIS restatement of the published estimator: Tolstikhin, Bousquet, Gelly, Schoelkopf,
"Wasserstein Auto-Encoders" (ICLR 2018), Algorithm 2 / the `mmd_penalty` of the authors' implementation that the
repositories cited in the reference's README.md:287,293 follow:

    k(a, b) = sum_s C_s / (C_s + |a - b|^2),   C_s = 2 * z_dim * sigma2 * s,   s in {.1, .2, .5, 1, 2, 5, 10}
    MMD = [sum_{i != j} k(q_i, q_j) + sum_{i != j} k(p_i, p_j)] / (B (B - 1)) - 2 / B^2 * sum_{i, j} k(q_i, p_j)
"""
import torch

SCALES = (0.1, 0.2, 0.5, 1.0, 2.0, 5.0, 10.0)


def mmd_imq(zq, zp, sigma2=1.0):
    """zq: encoded latents [B, Z] (differentiable), zp: prior samples [B, Z]. Returns the scalar unbiased estimate."""
    B, Z = zq.shape
    dqq = (zq[:, None, :] - zq[None, :, :]).pow(2).sum(-1)
    dpp = (zp[:, None, :] - zp[None, :, :]).pow(2).sum(-1)
    dqp = (zq[:, None, :] - zp[None, :, :]).pow(2).sum(-1)
    off = 1.0 - torch.eye(B, dtype=zq.dtype)
    stat = zq.new_zeros(())
    for s in SCALES:
        C = 2.0 * Z * sigma2 * s
        stat = stat + ((C / (C + dqq) + C / (C + dpp)) * off).sum() / (B * B - B) - 2.0 * (C / (C + dqp)).sum() / (B * B)
    return stat

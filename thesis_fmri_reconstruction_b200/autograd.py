"""torch.autograd bridge between the nn.Module drop-in surface (models/vae_gan.py) and the kernel library.

One autograd Function per sub-network call (Encoder, CognitiveEncoder, Decoder, Discriminator "REC" / "GAN",
WaeDiscriminator) plus the small loss Functions. Each Function's forward/backward is nets.py's forward()/backward() on
raw buffers, i.e. launches of libfmri_b200.so kernels; nothing here dispatches to ATen convolution / GEMM / cuDNN /
cuBLAS, and CPU tensors are refused (no fallback).

Contract kept for the reference's unchanged training scripts:
  * parameters stay ordinary fp32 nn.Parameters in the reference layouts; gradients are returned to autograd, so
    .grad accumulation across several backward() calls, zero_grad(), requires_grad toggles and grad clamping all work;
  * backward(retain_graph=True) may be called repeatedly on one graph (the saved state is never freed or mutated);
  * no Parameter is saved with save_for_backward: the tensor-core operands are bf16 *packs* made at forward time, so an
    optimizer.step() between two backward sweeps (train_vgan_stage1.py:412-432) neither trips autograd's version check
    nor changes the weights a later sweep differentiates through -- the torch-1.4 semantics the scripts were written for.
"""
from __future__ import annotations

import torch

from . import lib as L
from . import nets as NN
from .nets import BF16, F32

_COMPUTE_DTYPE = BF16


def set_compute_dtype(dtype):
    """torch.bfloat16 (default): tcgen05 tensor path; torch.float32: exact CUDA-core path (slow; parity checks)."""
    global _COMPUTE_DTYPE
    if dtype not in (BF16, F32):
        raise ValueError("compute dtype must be torch.bfloat16 or torch.float32")
    _COMPUTE_DTYPE = dtype


def compute_dtype():
    return _COMPUTE_DTYPE


def _cuda_f32(t, what):
    if not t.is_cuda:
        raise L.FmriError(f"{what} is on {t.device}: the sm_100a kernel library has no CPU path (move the model and data to cuda)")
    return t.detach().to(F32).contiguous()


class NetHost:
    """Mixin for nn.Modules whose arithmetic runs through a nets.*Net: parameter / buffer dictionaries keyed by the
    reference's state_dict names, the kernel-side network object per compute dtype, and bf16 pack refresh keyed on the
    parameters' version counters."""

    def _host_init(self):
        self.__dict__["_nets"] = {}
        self.__dict__["_pack_key"] = {}

    def _named(self):
        return list(self.named_parameters())

    def _bufs(self):
        return {k: v for k, v in self.named_buffers()}

    def _make_net(self, adt):  # pragma: no cover - overridden
        raise NotImplementedError

    def _net(self, P):
        adt = _COMPUTE_DTYPE
        net = self._nets.get(adt)
        if net is None:
            net = self._nets[adt] = self._make_net(adt)
        key = tuple((p.data_ptr(), p._version) for p in P.values())
        if self._pack_key.get(adt) != key:
            net.refresh(P, inplace=False)  # new pack tensors: graphs of earlier forwards keep the packs they were built with
            self._pack_key[adt] = key
        return net

    def _flush_nbt(self, nbt):
        bufs = self._bufs()
        for k, n in nbt.items():
            bufs[k] += n


def _params_of(mod):
    named = mod._named()
    names = [k for k, _ in named]
    params = [p for _, p in named]
    return names, params


# ====================================================================================================== encoders
class _EncoderFn(torch.autograd.Function):
    """Encoder / CognitiveEncoder forward: input -> (mu, logvar)."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        names, _ = _params_of(mod)
        P = dict(zip(names, (p.detach() for p in params)))
        for p in P.values():
            if not p.is_cuda:
                raise L.FmriError("model parameters are on the CPU: the kernel library has no CPU path")
        net = mod._net(P)
        nbt = {}
        xin = _cuda_f32(x, "encoder input")
        ycat, c = net.forward(P, mod._bufs(), xin, mod.training, 1, nbt)
        mod._flush_nbt(nbt)
        ctx.net, ctx.c, ctx.P, ctx.names, ctx.z = net, c, P, names, net.z
        ctx.set_materialize_grads(False)  # an unused logvar (WAE) must reach backward as None, not as zeros
        z, B = net.z, xin.shape[0]
        mu = torch.empty(B, z, dtype=F32, device=ycat.device)
        lv = torch.empty(B, z, dtype=F32, device=ycat.device)
        L.cast2d(ycat[:, :z], 2 * z, mu, z, B, z)
        L.cast2d(ycat[:, z:], 2 * z, lv, z, B, z)
        return mu, lv

    @staticmethod
    def backward(ctx, dmu, dlv):
        net, z = ctx.net, ctx.z
        needs = ctx.needs_input_grad[2:]
        B = ctx.c.B
        dycat = torch.empty(B, 2 * z, dtype=net.adt, device=ctx.P[ctx.names[0]].device)
        if dmu is None:
            dycat[:, :z].zero_()
        else:
            L.cast2d(dmu.contiguous(), z, dycat[:, :z], 2 * z, B, z)
        if dlv is None:
            dycat[:, z:].zero_()
        else:
            L.cast2d(dlv.contiguous(), z, dycat[:, z:], 2 * z, B, z)
        G = {n: torch.empty_like(ctx.P[n]) for n in ctx.names}
        if any(needs):
            net.backward(ctx.P, ctx.c, dycat, G, False, True, dlv is not None)
            NN.join_side()   # weight gradients were produced on the side stream (nets.WGRAD_SIDE_STREAM)
        if dlv is None:  # WAE: logvar unused -> its head receives no gradient (Adam must skip it, train_wae_stage1.py:296)
            for n in ("l_var.weight", "l_var.bias"):
                G[n] = None
        return (None, None) + tuple(G[n] if need else None for n, need in zip(ctx.names, needs))


# ====================================================================================================== standalone blocks
class _BlockFn(torch.autograd.Function):
    """EncoderBlock.forward(ten, out) / DecoderBlock.forward(ten) called on their own (vae_gan.py:23-35, 56-60)."""

    @staticmethod
    def forward(ctx, mod, want_raw, x, *params):
        names, _ = _params_of(mod)
        P = dict(zip(names, (p.detach() for p in params)))
        net = mod._net(P)
        nbt = {}
        y, raw, c = net.forward(P, mod._bufs(), _cuda_f32(x, "block input"), mod.training, want_raw, nbt)
        mod._flush_nbt(nbt)
        ctx.net, ctx.c, ctx.P, ctx.names, ctx.want_raw = net, c, P, names, want_raw
        ctx.set_materialize_grads(False)
        if want_raw:
            return y, raw
        return y

    @staticmethod
    def backward(ctx, dy, draw=None):
        needs = ctx.needs_input_grad[3:]
        need_dw = any(needs)
        G = {n: torch.zeros_like(ctx.P[n]) for n in ctx.names}
        f = lambda t: None if t is None else t.to(F32).contiguous()
        if dy is None and draw is None:
            return (None, None, None) + tuple(None for _ in needs)
        dx = ctx.net.backward(ctx.P, ctx.c, f(dy), f(draw), G, need_dw, ctx.needs_input_grad[2])
        NN.join_side()
        return (None, None, dx) + tuple(G[n] if need else None for n, need in zip(ctx.names, needs))


def run_block(mod, x, want_raw):
    _on_cuda(mod, x)
    _, params = _params_of(mod)
    return _BlockFn.apply(mod, bool(want_raw), x, *params)


# ====================================================================================================== decoder
class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, zin, *params):
        names, _ = _params_of(mod)
        P = dict(zip(names, (p.detach() for p in params)))
        net = mod._net(P)
        nbt = {}
        img, c = net.forward(P, mod._bufs(), _cuda_f32(zin, "decoder input"), mod.training, 1, nbt)
        mod._flush_nbt(nbt)
        # the saved state must not hold the RETURNED tensor object (output -> grad_fn -> ctx -> output is a reference cycle that
        # only the cyclic GC frees, pinning every decoder call's activations): keep a detached alias for the tanh backward
        c.img = img.detach()
        ctx.net, ctx.c, ctx.P, ctx.names = net, c, P, names
        return img

    @staticmethod
    def backward(ctx, dimg):
        needs = ctx.needs_input_grad[2:]
        need_dw = any(needs)
        need_dz = ctx.needs_input_grad[1]
        G = {n: torch.empty_like(ctx.P[n]) for n in ctx.names} if need_dw else None
        dz = ctx.net.backward(ctx.P, ctx.c, 1.0, dimg.to(F32).contiguous(), 0.0, None, G, False, need_dw, need_dz)
        NN.join_side()
        return (None, dz) + tuple(G[n] if need else None for n, need in zip(ctx.names, needs))


# ====================================================================================================== discriminator
def _slices_needed(needs3):
    idx = [i for i, n in enumerate(needs3) if n]
    return (min(idx), max(idx) + 1) if idx else None


class _DiscriminatorFn(torch.autograd.Function):
    """Discriminator.forward(ten_orig, ten_predicted, ten_sampled, mode): "REC" -> raw conv output of block 3 flattened
    NCHW [3B, C*h*w] fp32; otherwise -> sigmoid class score [3B, 1]."""

    @staticmethod
    def forward(ctx, mod, mode, xo, xp, xs, *params):
        names, _ = _params_of(mod)
        P = dict(zip(names, (p.detach() for p in params)))
        net = mod._net(P)
        nbt = {}
        imgs = [_cuda_f32(t, "discriminator input") for t in (xo, xp, xs)]
        rec = mode == "REC"
        raw3, p, c = net.forward(P, mod._bufs(), imgs, mod.training, 1, not rec, nbt, 1)
        mod._flush_nbt(nbt)
        ctx.net, ctx.c, ctx.P, ctx.names, ctx.rec = net, c, P, names, rec
        N = c.N
        if rec:
            _, h, w, Ct = raw3.shape
            out = torch.empty(N, Ct * h * w, dtype=F32, device=raw3.device)
            L.nhwc_to_nchw(raw3, out, N, Ct, h, w)  # layer_ten.view(len, -1) of an NCHW tensor (vae_gan.py:173)
            ctx.tap_shape = (h, w, Ct)
            return out
        return p.view(N, 1)

    @staticmethod
    def backward(ctx, g):
        net, c, P = ctx.net, ctx.c, ctx.P
        needs = ctx.needs_input_grad[5:]
        need_dw = any(needs)
        sl = _slices_needed(ctx.needs_input_grad[2:5])
        G = {n: torch.zeros_like(P[n]) for n in ctx.names} if need_dw else None
        N = c.N
        if ctx.rec:
            h, w, Ct = ctx.tap_shape
            draw3 = torch.empty(N, h, w, Ct, dtype=net.adt, device=g.device)
            L.nchw_to_nhwc(g.to(F32).contiguous(), draw3, N, Ct, h, w)
            dimg = net.backward_rec(P, c, draw3, G, False, need_dw, sl)
        else:
            dimg = net.backward_gan(P, c, g.to(F32).contiguous().view(-1), G, False, need_dw, sl)
        NN.join_side()
        gi = [None, None, None]
        if sl is not None:
            for i in range(sl[0], sl[1]):
                if ctx.needs_input_grad[2 + i]:
                    gi[i] = dimg[(i - sl[0]) * c.Bs:(i - sl[0] + 1) * c.Bs]
        return (None, None, gi[0], gi[1], gi[2]) + tuple(G[n] if need else None for n, need in zip(ctx.names, needs))


# ====================================================================================================== WAE discriminator
class _WaeDiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, zin, *params):
        names, _ = _params_of(mod)
        P = dict(zip(names, (p.detach() for p in params)))
        net = mod._net(P)
        p, c = net.forward(P, _cuda_f32(zin, "WAE discriminator input"))
        ctx.net, ctx.c, ctx.P, ctx.names = net, c, P, names
        return p.view(-1, 1)

    @staticmethod
    def backward(ctx, gp):
        needs = ctx.needs_input_grad[2:]
        need_dw = any(needs)
        G = {n: torch.zeros_like(ctx.P[n]) for n in ctx.names} if need_dw else None
        dz = ctx.net.backward(ctx.P, ctx.c, gp.to(F32).contiguous().view(-1), G, False, need_dw, ctx.needs_input_grad[1])
        return (None, dz) + tuple(G[n] if need else None for n, need in zip(ctx.names, needs))


# ====================================================================================================== losses
class _ReparamFn(torch.autograd.Function):
    """z = eps * exp(0.5 * logvar) + mu (vae_gan.py:266-269); eps is drawn by the caller with the reference's RNG call."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu_c, lv_c, eps_c = (_cuda_f32(t, "reparameterize input") for t in (mu, logvar, eps))
        B, Zd = mu_c.shape
        z = torch.empty_like(mu_c)
        L.reparam_kl_fwd(mu_c, lv_c, eps_c, z, None, B, Zd)
        ctx.save_for_backward(mu_c, lv_c, eps_c)
        return z

    @staticmethod
    def backward(ctx, gz):
        mu, lv, eps = ctx.saved_tensors
        B, Zd = mu.shape
        dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
        L.reparam_kl_bwd(mu, lv, eps, gz.to(F32).contiguous(), None, dmu, dlv, B, Zd, gkl_const=0.0)
        return dmu, dlv, None


class _KlFn(torch.autograd.Function):
    """kl[b] = -0.5 * sum_j (1 + logvar - mu^2 - exp(logvar)) (vae_gan.py:310)."""

    @staticmethod
    def forward(ctx, mu, logvar):
        mu_c, lv_c = _cuda_f32(mu, "kl input"), _cuda_f32(logvar, "kl input")
        B, Zd = mu_c.shape
        kl = torch.empty(B, dtype=F32, device=mu_c.device)
        L.reparam_kl_fwd(mu_c, lv_c, None, None, kl, B, Zd)
        ctx.save_for_backward(mu_c, lv_c)
        return kl

    @staticmethod
    def backward(ctx, gkl):
        mu, lv = ctx.saved_tensors
        B, Zd = mu.shape
        dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
        L.reparam_kl_bwd(mu, lv, None, None, gkl.to(F32).contiguous(), dmu, dlv, B, Zd)
        return dmu, dlv


class _RowSqDiffFn(torch.autograd.Function):
    """out[b] = scale * sum_j (a[b,j] - b[b,j])^2: feature-matching MSE (vae_gan.py:313, scale 0.5)."""

    @staticmethod
    def forward(ctx, a, b, scale):
        a_c, b_c = _cuda_f32(a, "mse input"), _cuda_f32(b, "mse input")
        rows, Fd = a_c.shape[0], a_c[0].numel()
        out = torch.empty(rows, dtype=F32, device=a_c.device)
        L.rowsqdiff_fwd(a_c, b_c, out, rows, Fd, scale)
        ctx.save_for_backward(a_c, b_c)
        ctx.scale, ctx.shape = scale, a.shape
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        rows, Fd = a.shape[0], a[0].numel()
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        if da is not None or db is not None:
            L.rowsqdiff_bwd(a, b, g.to(F32).contiguous(), da, db, rows, Fd, ctx.scale)
        return da, db, None


class _BceFn(torch.autograd.Function):
    """-scale * log(p + 1e-3) (positive) or -scale * log(1 - p + 1e-3) (vae_gan.py:316-318; train_wae_stage1.py:281-303)."""

    @staticmethod
    def forward(ctx, p, positive, scale):
        p_c = _cuda_f32(p, "bce input")
        out = torch.empty_like(p_c)
        L.bce_fwd(p_c, out, p_c.numel(), positive, scale)
        ctx.save_for_backward(p_c)
        ctx.positive, ctx.scale = positive, scale
        return out

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        dp = torch.empty_like(p)
        L.bce_bwd(p, g.to(F32).contiguous(), dp, p.numel(), ctx.positive, ctx.scale)
        return dp, None, None


class _MmdImqFn(torch.autograd.Function):
    """WAE-MMD penalty with the inverse-multiquadratic kernel (fmri_mmd_imq_{fwd,bwd}). EXTENSION: the reference has no
    MMD; pinned against oracle/mmd.py only. Differentiable in the encoded latents zq; zp are prior samples."""

    @staticmethod
    def forward(ctx, zq, zp, sigma2, lam):
        q, p = _cuda_f32(zq, "mmd latents"), _cuda_f32(zp, "mmd prior samples")
        if q.dim() != 2 or q.shape != p.shape:
            raise L.FmriError("mmd_imq expects two [B, Z] tensors of the same shape")
        B, Z = q.shape
        out = torch.empty(1, dtype=F32, device=q.device)
        ws = torch.empty(3, dtype=torch.float64, device=q.device)
        L.mmd_imq_fwd(q, p, B, Z, sigma2, lam, out, ws)
        ctx.save_for_backward(q, p)
        ctx.sigma2, ctx.lam = sigma2, lam
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        q, p = ctx.saved_tensors
        B, Z = q.shape
        dq = torch.empty_like(q)
        L.mmd_imq_bwd(q, p, B, Z, ctx.sigma2, ctx.lam, dq)
        return dq * g.to(F32), None, None, None


def mmd_imq(zq, zp, sigma2=1.0, lam=1.0):
    """lam * MMD_IMQ(zq, zp): the WAE-MMD latent penalty (Tolstikhin et al.), sigma2 = the prior's variance."""
    return _MmdImqFn.apply(zq, zp, float(sigma2), float(lam))


def reparameterize(mu, logvar, eps):
    return _ReparamFn.apply(mu, logvar, eps)


def kl_divergence(mu, logvar):
    return _KlFn.apply(mu, logvar)


def row_sq_diff(a, b, scale=0.5):
    return _RowSqDiffFn.apply(a, b, scale)


def bce(p, positive, scale=1.0):
    return _BceFn.apply(p, positive, scale)


def _on_cuda(mod, *inputs):
    """No CPU path: refuse before anything is allocated, with the library's own error type."""
    for t in inputs:
        if t is not None and not t.is_cuda:
            raise L.FmriError(f"{type(mod).__name__} input is on {t.device}: the sm_100a kernel library has no CPU path "
                              "(move the model and data to cuda)")
    for p in mod.parameters():
        if not p.is_cuda:
            raise L.FmriError(f"{type(mod).__name__} parameters are on {p.device}: the sm_100a kernel library has no CPU "
                              "path (call .to('cuda'))")
        break


def run_encoder(mod, x):
    _on_cuda(mod, x)
    _, params = _params_of(mod)
    return _EncoderFn.apply(mod, x, *params)


def run_decoder(mod, z):
    _on_cuda(mod, z)
    _, params = _params_of(mod)
    return _DecoderFn.apply(mod, z, *params)


def run_discriminator(mod, mode, xo, xp, xs):
    _on_cuda(mod, xo, xp, xs)
    _, params = _params_of(mod)
    return _DiscriminatorFn.apply(mod, mode, xo, xp, xs, *params)


def run_wae_discriminator(mod, z):
    _on_cuda(mod, z)
    _, params = _params_of(mod)
    return _WaeDiscriminatorFn.apply(mod, z, *params)

"""Parameter / buffer tables of the sub-networks and their initialisation, keyed by the reference's state_dict names.

Shapes: /root/reference/models/vae_gan.py:63-85 (Encoder), :99-123 (Decoder), :135-161 (Discriminator), :190-207
(CognitiveEncoder), :499-521 (WaeDiscriminator). Initialisation: VaeGan.init_parameters (:252-264) -- every Conv /
ConvTranspose / Linear weight ~ U(-s, s) with s = 1 / sqrt(prod(shape[1:])) / sqrt(3), biases 0, BatchNorm gamma 1,
beta 0, running_mean 0, running_var 1 -- and WaeDiscriminator's N(0, 0.0099999) / zero bias (:522-525).
(The drop-in nn.Modules in models/vae_gan.py construct torch.nn layers in the reference's order instead, so that a
seeded construction consumes the RNG exactly like the reference; this table form feeds the fused engine directly.)
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from .hp import NUM_VOXELS


def encoder_table(cfg, z):
    ps, bns, cin = [], [], 3
    for i, c in enumerate(cfg["encoder_channels"]):
        ps += [(f"conv.{i}.conv.weight", (c, cin, 5, 5)), (f"conv.{i}.bn.weight", (c,)), (f"conv.{i}.bn.bias", (c,))]
        bns.append((f"conv.{i}.bn.", c))
        cin = c
    fo, fi = cfg["fc_output"], cfg["fc_input"] ** 2 * cin
    ps += [("fc.0.weight", (fo, fi)), ("fc.1.weight", (fo,)), ("fc.1.bias", (fo,)), ("l_mu.weight", (z, fo)),
           ("l_mu.bias", (z,)), ("l_var.weight", (z, fo)), ("l_var.bias", (z,))]
    return ps, bns + [("fc.1.", fo)]


def decoder_table(cfg, z, size=256):
    dc = cfg["decoder_channels"]
    fo = cfg["fc_input"] ** 2 * size
    ps = [("fc.0.weight", (fo, z)), ("fc.1.weight", (fo,)), ("fc.1.bias", (fo,))]
    bns = [("fc.1.", fo)]
    for i, (ci, co) in enumerate([(size, size), (size, dc[1]), (dc[1], dc[2])]):
        ps += [(f"conv.{i}.conv.weight", (ci, co, 5, 5)), (f"conv.{i}.bn.weight", (co,)), (f"conv.{i}.bn.bias", (co,))]
        bns.append((f"conv.{i}.bn.", co))
    return ps + [("conv.3.0.weight", (dc[3], dc[2], 5, 5)), ("conv.3.0.bias", (dc[3],))], bns


def discriminator_table(cfg):
    ch = cfg["discrim_channels"]
    ps, bns = [("conv.0.0.weight", (ch[0], 3, 5, 5)), ("conv.0.0.bias", (ch[0],))], []
    for i in (1, 2, 3):
        ps += [(f"conv.{i}.conv.weight", (ch[i], ch[i - 1], 5, 5)), (f"conv.{i}.bn.weight", (ch[i],)),
               (f"conv.{i}.bn.bias", (ch[i],))]
        bns.append((f"conv.{i}.bn.", ch[i]))
    fo, fi = cfg["fc_output_gan"], cfg["fc_input_gan"] ** 2 * ch[3]
    ps += [("fc.0.weight", (fo, fi)), ("fc.1.weight", (fo,)), ("fc.1.bias", (fo,)), ("fc.3.weight", (1, fo)),
           ("fc.3.bias", (1,))]
    return ps, bns + [("fc.1.", fo)]


def cognitive_encoder_table(z, input_size=NUM_VOXELS):
    return ([("fc1.0.weight", (1024, input_size)), ("fc1.1.weight", (1024,)), ("fc1.1.bias", (1024,)),
             ("l_mu.weight", (z, 1024)), ("l_mu.bias", (z,)), ("l_var.weight", (z, 1024)), ("l_var.bias", (z,))],
            [("fc1.1.", 1024)])


def wae_discriminator_table(z, dim_h=512):
    ps = []
    for i, (o, k) in zip((0, 2, 4, 6, 8), [(dim_h, z), (dim_h, dim_h), (dim_h, dim_h), (dim_h, dim_h), (1, dim_h)]):
        ps += [(f"main.{i}.weight", (o, k)), (f"main.{i}.bias", (o,))]
    return ps, []


def init_net(prefix, table, gen, normal_std=None):
    ps, bns = table
    P, S = OrderedDict(), OrderedDict()
    for name, shape in ps:
        if len(shape) >= 2:
            if normal_std is not None:
                t = torch.randn(shape, generator=gen) * normal_std
            else:
                s = 1.0 / math.sqrt(math.prod(shape[1:])) / math.sqrt(3.0)
                t = (torch.rand(shape, generator=gen) * 2 - 1) * s
        elif any(name.startswith(b) for b, _ in bns) and name.endswith("weight"):
            t = torch.ones(shape)
        else:
            t = torch.zeros(shape)
        P[prefix + name] = t
    for b, c in bns:
        S[prefix + b + "running_mean"] = torch.zeros(c)
        S[prefix + b + "running_var"] = torch.ones(c)
        S[prefix + b + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return P, S


def init_vaegan(cfg, z, seed=12345):
    gen = torch.Generator().manual_seed(seed)
    P, S = OrderedDict(), OrderedDict()
    for pre, tab in (("encoder.", encoder_table(cfg, z)), ("decoder.", decoder_table(cfg, z)),
                     ("discriminator.", discriminator_table(cfg))):
        p, s = init_net(pre, tab, gen)
        P.update(p)
        S.update(s)
    return P, S


def init_waegan(cfg, z, seed=12345):
    gen = torch.Generator().manual_seed(seed)
    P, S = OrderedDict(), OrderedDict()
    for pre, tab, std in (("encoder.", encoder_table(cfg, z), None), ("decoder.", decoder_table(cfg, z), None),
                          ("discriminator.", wae_discriminator_table(z), 0.0099999)):
        p, s = init_net(pre, tab, gen, std)
        P.update(p)
        S.update(s)
    return P, S


def init_cognitive(cfg, z, seed=12345, with_teacher=True, input_size=NUM_VOXELS):
    """CognitiveEncoder (torch default Linear init, its init_parameters is commented out in the reference,
    models/vae_gan.py:208-222) + Decoder + Discriminator (+ the teacher's visual Encoder for Stage II)."""
    gen = torch.Generator().manual_seed(seed)
    P, S = OrderedDict(), OrderedDict()
    p, s = init_net("encoder.", cognitive_encoder_table(z, input_size), gen)
    for k, v in p.items():
        if v.dim() == 2:
            b = 1.0 / math.sqrt(v.shape[1])
            p[k] = (torch.rand(v.shape, generator=gen) * 2 - 1) * b
    P.update(p)
    S.update(s)
    tabs = [("decoder.", decoder_table(cfg, z)), ("discriminator.", discriminator_table(cfg))]
    if with_teacher:
        tabs.append(("teacher_net.encoder.", encoder_table(cfg, z)))
    for pre, tab in tabs:
        p, s = init_net(pre, tab, gen)
        P.update(p)
        S.update(s)
    return P, S


def init_dual_stage3(cfg, z, seed=12345, input_size=NUM_VOXELS):
    """init_cognitive() with the teacher's visual Encoder plus the latent WaeDiscriminator of WaeGanCognitive
    (models/vae_gan.py:542: keeps its own N(0, 0.0099999) init, :522-525) under latent_discriminator.*."""
    P, S = init_cognitive(cfg, z, seed, True, input_size)
    gen = torch.Generator().manual_seed(seed + 1)
    p, s = init_net("latent_discriminator.", wae_discriminator_table(z), gen, 0.0099999)
    P.update(p)
    S.update(s)
    return P, S


def init_cognitive_wae(cfg, z, seed=12345, input_size=NUM_VOXELS):
    """WaeGanCognitive (models/vae_gan.py:532-546): CognitiveEncoder + Decoder + latent WaeDiscriminator (its own
    N(0, 0.0099999) init) + the Stage-I teacher's visual Encoder (train_wae_stage2.py:195-203)."""
    P, S = init_cognitive(cfg, z, seed, True, input_size)
    P = OrderedDict((k, v) for k, v in P.items() if not k.startswith("discriminator."))
    S = OrderedDict((k, v) for k, v in S.items() if not k.startswith("discriminator."))
    gen = torch.Generator().manual_seed(seed + 2)
    p, s = init_net("discriminator.", wae_discriminator_table(z), gen, 0.0099999)
    P.update(p)
    S.update(s)
    return P, S

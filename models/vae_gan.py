"""Drop-in replacement of the reference's models/vae_gan.py: same nn.Module class names, constructors, forward
signatures, attributes and state_dict keys (/root/reference/models/vae_gan.py), with every module's arithmetic executed
by the hand-written sm_100a kernels of libfmri_b200.so through torch.autograd Functions
(thesis_fmri_reconstruction_b200/autograd.py). train_vgan_stage{1,2,3}.py, train_wae_stage{1,2,3}.py and
inference_gan.py import from here unchanged:

    from models.vae_gan import VaeGan, WaeGan, CognitiveEncoder, Encoder, Decoder, VaeGanCognitive, Discriminator, \
        WaeGanCognitive, DCGan

The torch.nn layer objects below (Conv2d, ConvTranspose2d, Linear, BatchNorm) are parameter CONTAINERS only -- built in
the reference's order so that state_dict keys, parameters() order and seeded initialisation are identical -- their
ATen forward is never called: Encoder / Decoder / Discriminator / CognitiveEncoder / WaeDiscriminator each run as one
kernel-library network. CUDA only: CPU tensors raise (there is no fallback path).
Architecture constants are read from configs.models_config at construction time, as in the reference (vae_gan.py:8).
"""
import numpy
import torch
import torch.nn as nn

import configs.models_config as config
from thesis_fmri_reconstruction_b200 import autograd as _ag
from thesis_fmri_reconstruction_b200 import nets as _nets
from thesis_fmri_reconstruction_b200.hp import cfg_from_module as _cfg
from thesis_fmri_reconstruction_b200.lib import FmriError


class EncoderBlock(nn.Module, _ag.NetHost):
    """Conv2d(k5, s2, p2, no bias) -> BatchNorm2d(momentum=0.9) -> ReLU (reference vae_gan.py:11-35). Inside an Encoder /
    Discriminator the block runs in that network's fused pipeline; called on its own, forward(ten, out=False, t=False) does
    what the reference's does: relu(bn(conv(ten))), and with out=True also the raw conv output (the feature tap, :25-30)."""

    def __init__(self, channel_in, channel_out):
        super(EncoderBlock, self).__init__()
        self._host_init()
        self.conv = nn.Conv2d(in_channels=channel_in, out_channels=channel_out, kernel_size=config.kernel_size,
                              padding=config.padding, stride=config.stride, bias=False)
        self.bn = nn.BatchNorm2d(num_features=channel_out, momentum=0.9)
        self.__dict__["_io"] = (channel_in, channel_out)

    def _make_net(self, adt):
        return _nets.BlockNet(self._io[0], self._io[1], False, 0, adt)

    def forward(self, ten, out=False, t=False):
        return _ag.run_block(self, ten, out)


class DecoderBlock(nn.Module, _ag.NetHost):
    """ConvTranspose2d(k5, s2, p2, output_padding=out) -> BatchNorm2d(0.9) -> ReLU (reference vae_gan.py:38-60); callable on
    its own like the reference's (inside a Decoder it runs in the fused pipeline)."""

    def __init__(self, channel_in, channel_out, out=False):
        super(DecoderBlock, self).__init__()
        self._host_init()
        self.__dict__["_io"] = (channel_in, channel_out, 1 if out else 0)
        if out:
            self.conv = nn.ConvTranspose2d(channel_in, channel_out, kernel_size=config.kernel_size,
                                           padding=config.padding, stride=config.stride, output_padding=1, bias=False)
        else:
            self.conv = nn.ConvTranspose2d(channel_in, channel_out, kernel_size=config.kernel_size,
                                           padding=config.padding, stride=config.stride, bias=False)
        self.bn = nn.BatchNorm2d(channel_out, momentum=0.9)

    def _make_net(self, adt):
        return _nets.BlockNet(self._io[0], self._io[1], True, self._io[2], adt)

    def forward(self, ten):
        return _ag.run_block(self, ten, False)


class Encoder(nn.Module, _ag.NetHost):
    """Visual encoder (reference vae_gan.py:63-96): forward(ten) -> (mu, logvar)."""

    def __init__(self, channel_in=3, z_size=128):
        super(Encoder, self).__init__()
        self._host_init()
        self.size = channel_in
        layers_list = []
        for i in range(3):
            layers_list.append(EncoderBlock(channel_in=self.size, channel_out=config.encoder_channels[i]))
            self.size = config.encoder_channels[i]
        self.conv = nn.Sequential(*layers_list)
        self.fc = nn.Sequential(nn.Linear(in_features=config.fc_input * config.fc_input * self.size,
                                          out_features=config.fc_output, bias=False),
                                nn.BatchNorm1d(num_features=config.fc_output, momentum=0.9),
                                nn.ReLU(True))
        self.l_mu = nn.Linear(in_features=config.fc_output, out_features=z_size)
        self.l_var = nn.Linear(in_features=config.fc_output, out_features=z_size)
        self.__dict__["_z"], self.__dict__["_cfg"] = z_size, _cfg(config)

    def _make_net(self, adt):
        return _nets.EncoderNet(self._cfg, self._z, adt)

    def forward(self, ten):
        return _ag.run_encoder(self, ten)


class Decoder(nn.Module, _ag.NetHost):
    """Decoder (reference vae_gan.py:99-132): forward(ten [B, z]) -> image [B, 3, H, W] in (-1, 1)."""

    def __init__(self, z_size, size):
        super(Decoder, self).__init__()
        self._host_init()
        self.fc = nn.Sequential(nn.Linear(in_features=z_size, out_features=config.fc_input * config.fc_input * size,
                                          bias=False),
                                nn.BatchNorm1d(num_features=config.fc_input * config.fc_input * size, momentum=0.9),
                                nn.ReLU(True))
        self.size = size
        layers_list = []
        layers_list.append(DecoderBlock(channel_in=self.size, channel_out=self.size, out=config.output_pad_dec[0]))
        layers_list.append(DecoderBlock(channel_in=self.size, channel_out=config.decoder_channels[1],
                                        out=config.output_pad_dec[1]))
        self.size = config.decoder_channels[1]
        layers_list.append(DecoderBlock(channel_in=self.size, channel_out=config.decoder_channels[2],
                                        out=config.output_pad_dec[2]))
        self.size = config.decoder_channels[2]
        layers_list.append(nn.Sequential(
            nn.Conv2d(in_channels=self.size, out_channels=config.decoder_channels[3], kernel_size=5, stride=1,
                      padding=2),
            nn.Tanh()))
        self.conv = nn.Sequential(*layers_list)
        self.__dict__["_z"], self.__dict__["_size0"], self.__dict__["_cfg"] = z_size, size, _cfg(config)

    def _make_net(self, adt):
        return _nets.DecoderNet(self._cfg, self._z, adt, self._size0)

    def forward(self, ten):
        return _ag.run_decoder(self, ten)


class Discriminator(nn.Module, _ag.NetHost):
    """Image discriminator (reference vae_gan.py:135-187): forward(ten_orig, ten_predicted, ten_sampled, mode='REC').
    'REC' -> raw conv output of block `recon_level` flattened [3B, C*h*w]; anything else -> sigmoid score [3B, 1]."""

    def __init__(self, channel_in=3, recon_level=3):
        super(Discriminator, self).__init__()
        self._host_init()
        self.size = channel_in
        self.recon_levl = recon_level
        self.conv = nn.ModuleList()
        self.conv.append(nn.Sequential(
            nn.Conv2d(in_channels=3, out_channels=config.discrim_channels[0], kernel_size=5, stride=config.stride_gan,
                      padding=2),
            nn.ReLU(inplace=True)))
        self.size = config.discrim_channels[0]
        self.conv.append(EncoderBlock(channel_in=self.size, channel_out=config.discrim_channels[1]))
        self.size = config.discrim_channels[1]
        self.conv.append(EncoderBlock(channel_in=self.size, channel_out=config.discrim_channels[2]))
        self.size = config.discrim_channels[2]
        self.conv.append(EncoderBlock(channel_in=self.size, channel_out=config.discrim_channels[3]))
        self.fc = nn.Sequential(
            nn.Linear(in_features=config.fc_input_gan * config.fc_input_gan * self.size,
                      out_features=config.fc_output_gan, bias=False),
            nn.BatchNorm1d(num_features=config.fc_output_gan, momentum=0.9),
            nn.ReLU(inplace=True),
            nn.Linear(in_features=config.fc_output_gan, out_features=1),
        )
        self.__dict__["_cfg"] = _cfg(config)

    def _make_net(self, adt):
        return _nets.DiscriminatorNet(self._cfg, adt, self.recon_levl)

    def forward(self, ten_orig, ten_predicted, ten_sampled, mode='REC'):
        return _ag.run_discriminator(self, "REC" if mode == "REC" else "GAN", ten_orig, ten_predicted, ten_sampled)


class CognitiveEncoder(nn.Module, _ag.NetHost):
    """fMRI encoder (reference vae_gan.py:190-232): forward(ten [B, input_size]) -> (mu, logvar). Default torch init."""

    def __init__(self, input_size, z_size=128, channel_in=3):
        super(CognitiveEncoder, self).__init__()
        self._host_init()
        self.size = channel_in
        self.fc1 = nn.Sequential(nn.Linear(in_features=input_size, out_features=1024, bias=False),
                                 nn.BatchNorm1d(num_features=1024, momentum=0.9),
                                 nn.ReLU(True))
        self.l_mu = nn.Linear(in_features=1024, out_features=z_size)
        self.l_var = nn.Linear(in_features=1024, out_features=z_size)
        self.__dict__["_z"], self.__dict__["_v"] = z_size, input_size

    def _make_net(self, adt):
        return _nets.CognitiveEncoderNet(self._v, self._z, adt)

    def forward(self, ten):
        return _ag.run_encoder(self, ten)


def _init_parameters(model):
    """VaeGan.init_parameters / WaeGan.init_parameters (reference vae_gan.py:252-264, 452-464)."""
    for m in model.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            if hasattr(m, "weight") and m.weight is not None and m.weight.requires_grad:
                scale = 1.0 / numpy.sqrt(numpy.prod(m.weight.shape[1:]))
                scale /= numpy.sqrt(3)
                nn.init.uniform_(m.weight, -scale, scale)
            if hasattr(m, "bias") and m.bias is not None and m.bias.requires_grad:
                nn.init.constant_(m.bias, 0.0)


def _reparameterize(mu, logvar):
    """reference vae_gan.py:266-269: eps drawn on the latent's device with the same RNG call, then z = eps*std + mu."""
    eps = logvar.data.new(logvar.size()).normal_()
    return _ag.reparameterize(mu, logvar, eps)


def _loss(x, x_tilde, disc_layer_original, disc_layer_predicted, disc_layer_sampled, disc_class_original,
          disc_class_predicted, disc_class_sampled, mus, variances):
    """VaeGan.loss == VaeGanCognitive.loss (reference vae_gan.py:302-320, 411-432): per-sample tensors, callers sum them."""
    # reconstruction error: logging only in the default 'vae-gan' mode -- plain elementwise torch
    nle = 0.5 * (x.view(len(x), -1) - x_tilde.view(len(x_tilde), -1)) ** 2
    kl = _ag.kl_divergence(mus, variances)
    mse = _ag.row_sq_diff(disc_layer_original, disc_layer_predicted, 0.5)
    bce_dis_original = _ag.bce(disc_class_original, True)
    bce_dis_predicted = _ag.bce(disc_class_predicted, False)
    bce_dis_sampled = _ag.bce(disc_class_sampled, False)
    return nle, kl, mse, bce_dis_original, bce_dis_predicted, bce_dis_sampled


class VaeGan(nn.Module):
    """Stage-I VAE/GAN (reference vae_gan.py:235-320)."""

    def __init__(self, device, z_size=128, recon_level=3):
        super(VaeGan, self).__init__()
        self.z_size = z_size
        self.encoder = Encoder(z_size=self.z_size).to(device)
        self.decoder = Decoder(z_size=self.z_size, size=self.encoder.size).to(device)
        self.discriminator = Discriminator(channel_in=3, recon_level=recon_level).to(device)
        self.init_parameters()
        self.device = device

    def init_parameters(self):
        _init_parameters(self)

    def reparameterize(self, mu, logvar):
        return _reparameterize(mu, logvar)

    def forward(self, x, gen_size=10):
        if x is not None:
            x = x.to(self.device)
        if self.training:
            mus, log_variances = self.encoder(x)
            z = self.reparameterize(mus, log_variances)
            x_tilde = self.decoder(z)
            # prior sample drawn on the CPU and moved, as in the reference (vae_gan.py:281) -- same RNG stream
            z_p = torch.randn(len(x), self.z_size).to(self.device).requires_grad_(True)
            x_p = self.decoder(z_p)
            disc_layer = self.discriminator(x, x_tilde, x_p, "REC")
            disc_class = self.discriminator(x, x_tilde, x_p, "GAN")
            return x_tilde, disc_class, disc_layer, mus, log_variances
        else:
            if x is None:
                z_p = torch.randn(gen_size, self.z_size).to(self.device)
                return self.decoder(z_p)
            mus, log_variances = self.encoder(x)
            z = self.reparameterize(mus, log_variances)
            return self.decoder(z)

    loss = staticmethod(_loss)


class VaeGanCognitive(nn.Module):
    """Dual-VAE/GAN of Stages II / III (reference vae_gan.py:323-432)."""

    def __init__(self, device, encoder, decoder, discriminator, z_size=128, recon_level=3, teacher_net=None, stage=1,
                 mode='vae'):
        super(VaeGanCognitive, self).__init__()
        self.device = device
        self.z_size = z_size
        self.encoder = encoder
        self.decoder = decoder
        self.discriminator = discriminator
        self.teacher_net = teacher_net
        self.stage = stage
        self.mode = mode

    def reparameterize(self, mu, logvar):
        return _reparameterize(mu, logvar)

    def forward(self, sample, gen_size=10):
        if sample is not None:
            x = sample['fmri'].to(self.device)
            gt_x = sample['image'].to(self.device)
            if self.training:
                if self.mode == 'vae':
                    mus, log_variances = self.encoder(x)
                    z = self.reparameterize(mus, log_variances)
                    x_tilde = self.decoder(z)
                    if self.teacher_net is not None and self.stage == 2:
                        for param in self.teacher_net.encoder.parameters():
                            param.requires_grad = False
                        mu_teacher, logvar_teacher = self.teacher_net.encoder(gt_x)
                        z_teacher = self.reparameterize(mu_teacher, logvar_teacher)
                        gt_x = self.decoder(z_teacher)
                elif self.mode == 'wae':
                    mus, log_variances = self.encoder(x)
                    x_tilde = self.decoder(mus)
                    mu_teacher, logvar_teacher = self.teacher_net.encoder(gt_x)
                    gt_x = self.decoder(mu_teacher)
                z_p = torch.randn(len(x), self.z_size).to(self.device).requires_grad_(True)
                x_p = self.decoder(z_p)
                disc_layer = self.discriminator(gt_x, x_tilde, x_p, "REC")
                disc_class = self.discriminator(gt_x, x_tilde, x_p, "GAN")
                return gt_x, x_tilde, disc_class, disc_layer, mus, log_variances
            else:
                mus, log_variances = self.encoder(x)
                z = self.reparameterize(mus, log_variances)
                return self.decoder(z)
        else:
            z_p = torch.randn(gen_size, self.z_size).to(self.device)
            return self.decoder(z_p)

    loss = staticmethod(_loss)


class WaeDiscriminator(nn.Module, _ag.NetHost):
    """Latent discriminator of the WAE (reference vae_gan.py:499-529): forward(x [N, z]) -> [N, 1]."""

    def __init__(self, z_size=128, dim_h=512):
        super(WaeDiscriminator, self).__init__()
        self._host_init()
        self.n_z = z_size
        self.dim_h = dim_h
        self.main = nn.Sequential(
            nn.Linear(self.n_z, self.dim_h), nn.ReLU(True),
            nn.Linear(self.dim_h, self.dim_h), nn.ReLU(True),
            nn.Linear(self.dim_h, self.dim_h), nn.ReLU(True),
            nn.Linear(self.dim_h, self.dim_h), nn.ReLU(True),
            nn.Linear(self.dim_h, 1), nn.Sigmoid())
        for m in self.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(0.0, 0.0099999)
                m.bias.data.zero_()

    def _make_net(self, adt):
        return _nets.WaeDiscriminatorNet(self.n_z, adt, self.dim_h)

    def forward(self, x):
        return _ag.run_wae_discriminator(self, x)


class WaeGan(nn.Module):
    """WAE with GAN penalty, Stage I (reference vae_gan.py:435-496). The train scripts drive .encoder / .decoder /
    .discriminator directly; forward's train branch is unused by design in the reference (vae_gan.py:473)."""

    def __init__(self, device, z_size=128):
        super(WaeGan, self).__init__()
        self.z_size = z_size
        self.encoder = Encoder(z_size=self.z_size).to(device)
        self.decoder = Decoder(z_size=self.z_size, size=self.encoder.size).to(device)
        self.discriminator = WaeDiscriminator(z_size=self.z_size).to(device)
        self.init_parameters()
        self.device = device

    def init_parameters(self):
        _init_parameters(self)

    def forward(self, x, gen_size=10):
        if x is not None:
            x = x.to(self.device)
        if self.training:
            raise FmriError("WaeGan.forward(train) is unused in the reference (its own call signature is inconsistent, "
                            "vae_gan.py:473-481); use .encoder / .decoder / .discriminator as train_wae_stage1.py does")
        if x is None:
            raise FmriError("WaeGan.forward(None) dereferences x in the reference (vae_gan.py:488)")
        mus, log_variances = self.encoder(x)
        return self.decoder(mus)


class WaeGanCognitive(nn.Module):
    """WAE/GAN for Stages II / III (reference vae_gan.py:532-578); freezes the decoder at construction (:545-546)."""

    def __init__(self, device, encoder, decoder, z_size=128, recon_level=3):
        super(WaeGanCognitive, self).__init__()
        self.z_size = z_size
        self.encoder = encoder
        self.discriminator = WaeDiscriminator(z_size=self.z_size).to(device)
        self.device = device
        self.decoder = decoder
        for param in self.decoder.parameters():
            param.requires_grad = False

    def reparameterize(self, mu, logvar):
        return _reparameterize(mu, logvar)

    def forward(self, x, gen_size=10):
        if x is not None:
            x = x.to(self.device)
        if self.training:
            raise FmriError("WaeGanCognitive.forward(train) is unused in the reference (vae_gan.py:563-573 calls the "
                            "1-argument discriminator with 3); the train scripts drive the sub-modules directly")
        if x is None:
            raise FmriError("WaeGanCognitive.forward(None) dereferences x in the reference (vae_gan.py:576)")
        mus, log_variances = self.encoder(x)
        return self.decoder(mus)


class DCGan(nn.Module):
    """Container used by experiments/exp_dcgan_stage*.py (reference vae_gan.py:581-622)."""

    def __init__(self, device, decoder, discriminator, z_size=128, recon_level=3):
        super(DCGan, self).__init__()
        self.device = device
        self.z_size = z_size
        self.decoder = decoder
        self.discriminator = discriminator

    def reparameterize(self, mu, logvar):
        return _reparameterize(mu, logvar)

    def forward(self, sample, gen_size=10):
        if sample is not None:
            gt_x = sample.to(self.device)
            if self.training:
                z_p = torch.randn(len(gt_x), self.z_size).to(self.device).requires_grad_(True)
                x_tilde = self.decoder(z_p)
                disc_layer = self.discriminator(gt_x, x_tilde, x_tilde, "REC")
                disc_class = self.discriminator(gt_x, x_tilde, x_tilde, "GAN")
                return gt_x, x_tilde, disc_class, disc_layer
            z_p = torch.randn(gt_x.shape[0], self.z_size).to(self.device)
            return self.decoder(z_p)
        z_p = torch.randn(gen_size, self.z_size).to(self.device)
        return self.decoder(z_p)


class WaeDecoder(nn.Module):
    """Dead variant in the reference (vae_gan.py:625-655, commented out at :447); out of scope."""

    def __init__(self, z_size, size):
        super(WaeDecoder, self).__init__()
        raise FmriError("WaeDecoder is an unused variant of the reference and is outside the accelerated hot path")


class ResNetEncoder(nn.Module):
    """Needs a pretrained torchvision ResNet-152 download (reference vae_gan.py:658-); out of scope."""

    def __init__(self, z_size=128, fc_hidden1=1024, fc_hidden2=768, drop_p=0.3):
        super(ResNetEncoder, self).__init__()
        raise FmriError("ResNetEncoder is an unused variant of the reference and is outside the accelerated hot path")

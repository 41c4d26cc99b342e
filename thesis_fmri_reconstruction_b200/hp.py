"""Hyper-parameters and architecture presets of the reference, restated (no import of the reference or the oracle).

/root/reference/configs/models_config.py:3-31 (architecture: the active block is 100x100 / latent 512, the commented
block 64x64 / latent 128), configs/gan_config.py:17-32, configs/wae_config.py:16-30, configs/data_config.py:62-73.
"""
CFG64 = dict(image_size=64, fc_input=8, fc_output=1024, fc_input_gan=8, fc_output_gan=512, stride_gan=1,
             latent_dim=128, output_pad_dec=[True, True, True], encoder_channels=[64, 128, 256],
             decoder_channels=[256, 128, 32, 3], discrim_channels=[32, 128, 256, 256, 512])
CFG100 = dict(image_size=100, fc_input=13, fc_output=1024, fc_input_gan=7, fc_output_gan=256, stride_gan=2,
              latent_dim=512, output_pad_dec=[False, True, True], encoder_channels=[64, 128, 256],
              decoder_channels=[256, 128, 64, 3], discrim_channels=[32, 128, 256, 256, 512])
NUM_VOXELS = 3620

# train_vgan_stage1.py:275-283 (RMSprop alpha 0.9, eps 1e-8), gan_config.py:18,25,30,31
HP_VGAN = dict(lr=1e-4, alpha=0.9, eps=1e-8, lambda_mse=1e-6, margin=0.35, equilibrium=0.68)
# train_wae_stage1.py:221-224 (Adam betas (0.5, 0.999); the discriminator runs at lr / 2), wae_config.py:17
HP_WAE = dict(lr=1e-4, beta1=0.5, beta2=0.999, eps=1e-8)
# train_wae_stage2.py:237-243, train_wae_stage3.py (hard-coded: encoder / decoder 1e-3, latent discriminator 5e-4)
HP_WAE23 = dict(lr=1e-3, lr_dis=5e-4, beta1=0.5, beta2=0.999, eps=1e-8)


def cfg_from_module(mc):
    """Architecture dict from a configs.models_config-style module (attribute names of the reference)."""
    return dict(image_size=mc.image_size, fc_input=mc.fc_input, fc_output=mc.fc_output, fc_input_gan=mc.fc_input_gan,
                fc_output_gan=mc.fc_output_gan, stride_gan=mc.stride_gan, latent_dim=mc.latent_dim,
                output_pad_dec=list(mc.output_pad_dec), encoder_channels=list(mc.encoder_channels),
                decoder_channels=list(mc.decoder_channels), discrim_channels=list(mc.discrim_channels))


def epoch_end_vgan(lr, hp, decay_lr=0.98, decay_margin=1.0, decay_equilibrium=1.0, decay_mse=1.0):
    """End-of-epoch schedule of the VAE/GAN scripts (train_vgan_stage1.py:446-457; defaults gan_config.py:26-31):
    ExponentialLR(gamma=decay_lr) on every optimizer, margin / equilibrium decay with equilibrium >= margin. lambda_mse is NOT
    changed: the reference only scales a local variable that its loss never reads (hp["lambda_mse_local"] mirrors it). `lr`: {bucket prefix: learning rate}; `hp`: the trainer's hyper-parameter dict. Both are updated in place."""
    for k in lr:
        lr[k] *= decay_lr
    hp["margin"] *= decay_margin
    hp["equilibrium"] *= decay_equilibrium
    if hp["margin"] > hp["equilibrium"]:
        hp["equilibrium"] = hp["margin"]
    # The script decays a LOCAL copy of lambda_mse (:456-458) while its loss mix keeps reading the constant args.lambda_mse
    # (:372): the weight the step uses never changes. The decayed copy is kept under its own key, for logging only.
    hp["lambda_mse_local"] = min(1.0, hp.get("lambda_mse_local", hp["lambda_mse"]) * decay_mse)


def epoch_end_wae(lr, epoch, step_size=30, decay_lr=0.5):
    """StepLR(step_size, gamma=decay_lr) of the WAE scripts (train_wae_stage1.py:226-228, 334-336; wae_config.py:21-22;
    stages II / III hard-code StepLR(30, 0.5), train_wae_stage2.py:241-243). `epoch` = number of finished epochs (1-based)."""
    if epoch > 0 and epoch % step_size == 0:
        for k in lr:
            lr[k] *= decay_lr

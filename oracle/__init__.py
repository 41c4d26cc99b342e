"""CPU oracle for the VAE/GAN / WAE/GAN training step.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain torch-CPU functional code (no nn.Module, fp32 or fp64), the arithmetic of the
reference's training hot path: models/vae_gan.py (Encoder, Decoder, Discriminator, CognitiveEncoder,
WaeDiscriminator, VaeGan.loss) and the update logic of train/train_vgan_stage{1,2,3}.py and train/train_wae_stage1.py.
Every function cites the reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it, and only as
the checker / the timed CPU baseline.  The product path (thesis_fmri_reconstruction_b200, models/) never imports it.

Parity pin: the reference ships no tests or golden vectors for this path, so the oracle is
pinned by running the reference's own modules in the build container: oracle/make_golden.py imports
/root/reference/models/vae_gan.py unmodified, runs the reference step sequence, and writes small fixtures to
tests/golden/; tests/test_oracle_cpu.py checks this restatement against those fixtures.
"""

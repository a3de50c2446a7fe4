"""downgan_b200 — B200-native WGAN-GP training iteration for DoWnGAN.

Only what the hot path needs: ``networks.Generator`` /
``networks.Critic`` (drop-ins for the reference modules), ``GAN.wasserstein``
(drop-in trainer), ``config.hyperparams``, and the C-ABI extension under
``csrc/`` bound by ``_lib``.  Importing the package does not load CUDA.
"""
__version__ = "0.1.0"

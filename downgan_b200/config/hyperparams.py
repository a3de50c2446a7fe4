"""Hyper-parameters of the WGAN-GP iteration, same names and defaults as the
reference's ``DoWnGAN/config/hyperparams.py:16-27`` (module-level globals that
the trainer reads at call time, so callers may patch them the same way)."""

# Hyper params
gp_lambda = 10
critic_iterations = 5
batch_size = 32
gamma = 0.01
content_lambda = 5
lr = 0.00025

# Run configuration parameters
epochs = 1000
print_every = 250
save_every = 250
use_cuda = True

# Frequency separation is dead code in the reference (freq_sep = False, hyperparams.py:31)
freq_sep = False

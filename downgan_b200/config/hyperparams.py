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

# Frequency separation parameters (hyperparams.py:30-35).  The reference builds `low = nn.AvgPool2d(filter_size, stride=1,
# padding=0)` and `rf = nn.ReplicationPad2d(padding)` here; this path applies the same filter in one kernel (dg_lowpass).
# freq_sep = False upstream: GAN/wasserstein_fs.py is not reachable from train.py; WassersteinGANFS below mirrors it anyway.
freq_sep = False
filter_size = 5
padding = filter_size // 2

# Metrics of the per-batch metric pass (hyperparams.py:38-43 `metrics_to_calculate`).  MAE, MSE and Wass are computed by
# dg_metrics; "MSSSIM" needs the third-party pytorch_msssim package (absent here) and is not provided.
metrics_to_calculate = ("MAE", "MSE", "Wass")

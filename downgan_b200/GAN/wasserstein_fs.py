"""Import path of the reference's frequency-separation trainer (``DoWnGAN/GAN/wasserstein_fs.py:15``)."""
from .wasserstein import WassersteinGANFS  # noqa: F401

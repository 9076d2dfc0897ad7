"""vcagan_b200 -- B200-native (sm_100a) implementation of the VCA-GAN generator / visual-front / discriminator hot
path.  Importing the package loads libvcagan_b200.so; there is no CPU or library fallback."""
from ._lib import lib, VcaError, LIB_PATH  # noqa: F401
from .ops import cfg, set_precision, manual_seed, set_rng_rank  # noqa: F401
from . import ops, models  # noqa: F401

__all__ = ["lib", "VcaError", "cfg", "set_precision", "manual_seed", "ops", "models"]

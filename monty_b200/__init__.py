"""monty_b200 -- B200-native multi-stream RNG, samplers and log-densities with
the API of mrc-ide/monty's `monty_rng_*` / `monty_random_*` functions.

The Python functions here mirror R/random.R of the reference one to one and
call the C-ABI of include/monty_b200.h through ctypes, in the same way the
reference's R functions call its cpp11 layer.  All arithmetic happens in CUDA
kernels (monty_b200/csrc); nothing falls back to the CPU.
"""
from ._lib import MontyError  # noqa: F401
from .random import *  # noqa: F401,F403
from .random import __all__ as _random_all
from . import density  # noqa: F401

__all__ = list(_random_all) + ["MontyError", "density"]
__version__ = "0.1.0"

"""sapiogenesis_b200 -- B200-native per-tick world step of edgecase963/Sapiogenesis.

Python host code and PyTorch tensors (device memory, streams) call hand-written sm_100a CUDA kernels through the
C-ABI of include/sapio.h (libsapio_b200.so). There is no CPU path in this package.
"""
from .world import World, SapioParams, default_params  # noqa: F401

__version__ = "0.1.0"

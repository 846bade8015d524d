"""nlmc_b200 -- B200-native (sm_100a CUDA) implementation of the Monte Carlo hot path of
usra-riacs/Nonlocal-Monte-Carlo behind the reference's own class API.

    from nlmc_b200 import NMC, NPT, APT_preprocessor, APT_ICM      # same names as the reference
"""
from .apt_ICM import APT_ICM
from .apt_preprocessor import APT_preprocessor
from .nmc import NMC
from .npt import NPT

__all__ = ["NMC", "NPT", "APT_preprocessor", "APT_ICM"]

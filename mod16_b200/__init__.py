'''
mod16_b200 -- B200-native backend for the MOD16 evapotranspiration forward model.

Drop-in for the hot path of arthur-e/MOD16: ``MOD16(params)``, ``MOD16._et``,
``MOD16._evapotranspiration`` and ``MOD16.evapotranspiration`` keep the reference's
signatures and run as one fused CUDA kernel (sm_100a) behind the C ABI declared in
``include/mod16_b200.h``.  There is no CPU implementation in this package.
'''
from .model import MOD16, PFT_VALID  # noqa: F401
from .components import (  # noqa: F401
    latent_heat_vaporization, psychrometric_constant, radiation_net, svp, svp_slope)

STEFAN_BOLTZMANN = 5.67e-8
SPECIFIC_HEAT_CAPACITY_AIR = 1013
MOL_WEIGHT_WET_DRY_RATIO_AIR = 0.622

__version__ = '0.1.0'

"""spiking_fullsubnet_b200 -- B200-native (sm_100a) GSN hot path of Spiking-FullSubNet.

Drop-in model classes (select them from a recipe TOML with
`[model] path = "spiking_fullsubnet_b200.SpikingFullSubNet"`), backed by libgsn_b200.so
(C ABI: include/gsn_b200.h).  See DESIGN.md / INTEGRATION.md.
"""
from .modeling import (CirmGSN, GSUCell, GSULayer, MemoryState, Separator, SequenceModel,  # noqa: F401
                       SpikingFullSubNet, StackedGSU, SubbandModel, SubBandSequenceModel,
                       efficient_spiking_neuron)
from . import losses  # noqa: F401  (freq_MAE / mag_MAE / SISNRLoss of the recipes)

__version__ = "0.1.0"

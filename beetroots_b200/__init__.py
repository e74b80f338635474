"""beetroots_b200: B200-native (sm_100a) implementation of the beetroots sampler step
(PMALA + MTM over an N-pixel x D-parameter map) behind the reference's Python API.

Importing the package never touches CUDA; using ``B200Sampler`` / ``DevicePosterior`` loads
``libbeetroots_b200.so`` and raises if it is missing (there is no CPU fallback).
"""
from .modelling import (L22DiscreteGradSpatialPrior, MixingModelsLikelihood, NeuralNetworkApprox, Posterior,  # noqa: F401
                        SmoothIndicatorPrior)
from .network import DenseMLPWeights, synthetic_weights  # noqa: F401
from .device import MLP_BF16_TC, MLP_FP32, DevicePosterior  # noqa: F401
from .sampler import B200Sampler, MySamplerParams  # noqa: F401
from .saver import ChainSaver  # noqa: F401

__version__ = "0.1.0"

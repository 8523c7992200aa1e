"""Drop-in for ``inference.gp`` on the GpRegressor hot path (reference inference/gp/__init__.py)."""
from inference_tools_b200.gp.regression import GpRegressor
from inference_tools_b200.gp.optimisation import GpOptimiser
from inference_tools_b200.gp.inversion import GpLinearInverter
from inference_tools_b200.gp.acquisition import ExpectedImprovement, UpperConfidenceBound, MaxVariance
from inference_tools_b200.gp.mean import ConstantMean, LinearMean, QuadraticMean
from inference_tools_b200.gp.covariance import (
    SquaredExponential,
    RationalQuadratic,
    WhiteNoise,
    HeteroscedasticNoise,
    ChangePoint,
)

__all__ = [
    "GpRegressor",
    "GpOptimiser",
    "GpLinearInverter",
    "ExpectedImprovement",
    "UpperConfidenceBound",
    "MaxVariance",
    "ConstantMean",
    "LinearMean",
    "QuadraticMean",
    "SquaredExponential",
    "RationalQuadratic",
    "WhiteNoise",
    "HeteroscedasticNoise",
    "ChangePoint",
]

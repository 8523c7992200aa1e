"""Drop-in for ``inference.gp`` on the Gaussian-process hot path: the same public names as the reference package
(reference inference/gp/__init__.py), each backed by the CUDA engine in ``libgpb200.so``."""
from inference_tools_b200.gp import acquisition as _acq
from inference_tools_b200.gp import covariance as _cov
from inference_tools_b200.gp import inversion as _inv
from inference_tools_b200.gp import mean as _mean
from inference_tools_b200.gp import optimisation as _opt
from inference_tools_b200.gp import regression as _reg

# public name -> defining module; the groups follow the layers of the engine, not the reference's import order
_EXPORTS = {
    _reg: ("GpRegressor",),
    _opt: ("GpOptimiser",),
    _inv: ("GpLinearInverter",),
    _acq: ("ExpectedImprovement", "UpperConfidenceBound", "MaxVariance"),
    _mean: ("ConstantMean", "LinearMean", "QuadraticMean"),
    _cov: ("SquaredExponential", "RationalQuadratic", "WhiteNoise", "HeteroscedasticNoise", "ChangePoint"),
}
__all__ = []
for _module, _names in _EXPORTS.items():
    for _name in _names:
        globals()[_name] = getattr(_module, _name)
        __all__.append(_name)
del _module, _names, _name

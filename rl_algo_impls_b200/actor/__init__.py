from .categorical import MaskedCategorical
from .gaussian import GaussianDistribution
from .gridnet import GridnetDistribution, ValueDependentMask

__all__ = ["MaskedCategorical", "GaussianDistribution", "GridnetDistribution", "ValueDependentMask"]

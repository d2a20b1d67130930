from .dist_mnist_problem import DistMNISTProblem
from .dist_dense_problem import DistDensityProblem
from .dist_online_dense_problem import DistOnlineDensityProblem

__all__ = ["DistMNISTProblem", "DistDensityProblem", "DistOnlineDensityProblem"]

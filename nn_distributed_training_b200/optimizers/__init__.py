from .dinno import DiNNO
from .dsgd import DSGD
from .dsgt import DSGT

ALGORITHMS = {"dinno": DiNNO, "dsgd": DSGD, "dsgt": DSGT}


def build_optimizer(problem, device, opt_conf):
    """Factory keyed by ``alg_name`` (runners: experiments/dist_mnist_ex.py:195-202)."""
    try:
        cls = ALGORITHMS[opt_conf["alg_name"]]
    except KeyError:
        raise NameError("Unknown distributed opt algorithm.")
    return cls(problem, device, opt_conf)

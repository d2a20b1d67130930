"""CLI shim: `python dist_mnist_ex.py <config.yaml>` (same invocation as the reference's experiments/dist_mnist_ex.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nn_distributed_training_b200.experiments.dist_mnist_ex import main  # noqa: E402

if __name__ == "__main__":
    main(sys.argv)

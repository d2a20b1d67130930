"""Packaging of nn_distributed_training_b200 (the reference's setup.py names no packages, which is why
`pip install` of it fails on its flat layout — DESIGN.md §6).  The sm_100a extension is compiled in-tree by
`python -m nn_distributed_training_b200.ops.build` (or `__graft_entry__.build()`), not by setuptools, so the
built `.so` sits next to the sources it was built from.

    pip install --no-build-isolation --no-deps -e .
    python -m nn_distributed_training_b200.ops.build
"""
from setuptools import find_packages, setup

setup(
    name="nn_distributed_training_b200",
    version="0.1.0",
    description="Decentralized neural-network training (DiNNO / DSGD / DSGT) for NVIDIA B200",
    packages=find_packages(include=["nn_distributed_training_b200", "nn_distributed_training_b200.*"]),
    package_data={"nn_distributed_training_b200.ops": ["csrc/*", "_C*.so"]},
    python_requires=">=3.10",
    install_requires=["torch", "numpy", "networkx", "scipy", "pyyaml", "pillow"],
)

from . import graph_generation  # noqa: F401

from .results import load_results, summarize_run, rounds_to_threshold, plot_run
from . import animations  # noqa: E402,F401
from . import figures  # noqa: E402,F401

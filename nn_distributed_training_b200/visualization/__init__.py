from .results import load_results, summarize_run, rounds_to_threshold, plot_run

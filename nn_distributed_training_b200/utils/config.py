"""Typed view of the experiment YAML files.

The reference passes ``yaml.safe_load`` dicts straight down and fails with ``KeyError``
on anything missing.  Here every key the code reads is declared with its type
and (where the reference had none) a default; unknown optimizer/metric names fail at load time
with the offending path, and the validated dicts that flow down keep the reference's layout
(``experiment`` / ``problem_configs.<key>.optimizer_config``) so existing YAMLs load unchanged.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Iterable, Optional

import yaml

REQUIRED = object()

ALGS = ("dinno", "dsgd", "dsgt")
MNIST_METRICS = ("forward_pass_count", "validation_loss", "consensus_error", "top1_accuracy",
                 "current_epoch", "validation_as_vector")
DENSITY_METRICS = ("forward_pass_count", "validation_loss", "consensus_error", "mesh_grid_density",
                   "current_epoch", "train_loss_moving_average", "current_position", "current_graph")

OPT_SCHEMA = {
    "dinno": {"rho_init": REQUIRED, "rho_scaling": 1.0, "outer_iterations": REQUIRED,
              "primal_iterations": REQUIRED, "primal_optimizer": "adam", "persistant_primal_opt": False,
              "primal_lr_start": REQUIRED, "primal_lr_finish": None, "lr_decay_type": "constant",
              "profile": False},
    "dsgd": {"alpha0": REQUIRED, "mu": REQUIRED, "outer_iterations": REQUIRED, "profile": False},
    "dsgt": {"alpha": REQUIRED, "init_grads": True, "outer_iterations": REQUIRED, "profile": False},
}
# framework extensions accepted in every optimizer_config
OPT_EXTRA = ("mixing_order", "update_graph", "consensus_backend", "persistent_follows_schedule",
             "checkpoint_every", "checkpoint_dir", "resume")


class ConfigError(ValueError):
    pass


def _fill(d: Dict[str, Any], schema: Dict[str, Any], path: str, extra: Iterable[str] = ()) -> Dict[str, Any]:
    out = dict(d)
    for k, dflt in schema.items():
        if k not in out:
            if dflt is REQUIRED:
                raise ConfigError(f"missing required key {path}.{k}")
            out[k] = copy.deepcopy(dflt)
    return out


def validate_optimizer(conf: Dict[str, Any], path: str = "optimizer_config") -> Dict[str, Any]:
    if "alg_name" not in conf:
        raise ConfigError(f"missing required key {path}.alg_name")
    alg = conf["alg_name"]
    if alg == "cadmm":   # the README's older name for DiNNO (README.md:33,156)
        alg = "dinno"
    if alg not in ALGS:
        raise ConfigError(f"{path}.alg_name: unknown algorithm {alg!r} (expected one of {ALGS})")
    c = dict(conf, alg_name=alg)
    if alg == "dinno":
        # accept the stale schema of dist_dense_v2.yaml (`primal_lr`)
        if "primal_lr" in c and "primal_lr_start" not in c:
            c["primal_lr_start"] = c["primal_lr"]
        if "rho" in c and "rho_init" not in c:
            c["rho_init"] = c["rho"]
    c = _fill(c, OPT_SCHEMA[alg], path)
    if alg == "dinno":
        if c["primal_lr_finish"] is None:
            c["primal_lr_finish"] = c["primal_lr_start"]
        if c["lr_decay_type"] not in ("constant", "linear", "log"):
            raise ConfigError(f"{path}.lr_decay_type: {c['lr_decay_type']!r}")
        if c["primal_optimizer"] not in ("adam", "sgd", "adamw"):
            raise ConfigError(f"{path}.primal_optimizer: {c['primal_optimizer']!r}")
    if int(c["outer_iterations"]) <= 0:
        raise ConfigError(f"{path}.outer_iterations must be positive")
    return c


# extension keys of a problem config (absent from the reference's schema) and their legal values
PROBLEM_CHOICES = {
    "input_pipeline": ("auto", "resident", "staged", "host"),
    "host_gather": ("gpu_pull", "cpu_loader"),
    "host_pull_driver": ("graph", "runner"),
    "host_loss": ("mirror", "memcpy"),
}


def _check_extensions(c: Dict[str, Any], path: str) -> None:
    for k, legal in PROBLEM_CHOICES.items():
        if k in c and c[k] not in legal:
            raise ConfigError(f"{path}.{k} must be one of {'|'.join(legal)} (got {c[k]!r})")
    if "samples_per_cta" in c and not (c["samples_per_cta"] == 0 or 4 <= int(c["samples_per_cta"]) <= 8):
        raise ConfigError(f"{path}.samples_per_cta must be 0 (automatic) or 4..8")
    f = c.get("fault_injection")
    if f is not None:
        if not isinstance(f, dict) or not 0.0 <= float(f.get("link_drop_prob", 0.0)) <= 1.0:
            raise ConfigError(f"{path}.fault_injection needs link_drop_prob in [0, 1]")


def validate_problem(conf: Dict[str, Any], path: str, kind: str) -> Dict[str, Any]:
    c = _fill(conf, {"problem_name": REQUIRED, "train_batch_size": REQUIRED, "val_batch_size": REQUIRED,
                     "metrics": REQUIRED, "metrics_config": REQUIRED, "optimizer_config": REQUIRED,
                     "verbose_evals": True}, path)
    _check_extensions(c, path)
    allowed = MNIST_METRICS if kind == "mnist" else DENSITY_METRICS
    for m in c["metrics"]:
        if m not in allowed:
            raise ConfigError(f"{path}.metrics: unknown metric {m!r} for a {kind} problem")
    mc = dict(c["metrics_config"])
    if "evaluate_frequency" not in mc:
        raise ConfigError(f"missing required key {path}.metrics_config.evaluate_frequency")
    if kind == "online_density":
        mc.setdefault("tloss_decay", 0.2)
        mc.setdefault("mesh_only_at_end", True)
        c = _fill(c, {"comm_radius": REQUIRED, "dynamic_graph": True, "save_models": False}, path)
    c["metrics_config"] = mc
    c["optimizer_config"] = validate_optimizer(c["optimizer_config"], path + ".optimizer_config")
    return c


SOLO_DEFAULT = {"train_solo": False, "optimizer": "adam", "lr": 0.005, "epochs": 1,
                "train_batch_size": 100, "val_batch_size": 100, "verbose": True}


def validate_experiment(conf: Dict[str, Any], kind: str) -> Dict[str, Any]:
    """``kind``: mnist | mnist_scaling | density | online_density."""
    if "experiment" not in conf:
        raise ConfigError("missing top-level key `experiment`")
    out = copy.deepcopy(conf)
    exp = _fill(out["experiment"], {"name": REQUIRED, "output_metadir": REQUIRED, "writeout": True,
                                     "use_cuda": True, "loss": REQUIRED, "model": REQUIRED}, "experiment")
    if kind in ("mnist", "density"):
        exp = _fill(exp, {"graph": REQUIRED}, "experiment")
    if kind in ("mnist", "density", "online_density"):
        exp["individual_training"] = _fill(exp.get("individual_training", {}), SOLO_DEFAULT,
                                           "experiment.individual_training")
    if kind == "mnist":
        exp = _fill(exp, {"data_dir": "../data/", "data_split_type": "random"}, "experiment")
        if exp["data_split_type"] not in ("random", "hetero"):
            raise ConfigError("experiment.data_split_type must be random|hetero")
    if kind == "mnist_scaling":
        exp = _fill(exp, {"data_dir": "../data/", "scaling": REQUIRED}, "experiment")
    if kind.startswith("mnist"):
        exp.setdefault("data_source", "auto")
        if exp["data_source"] not in ("auto", "mnist", "synthetic", "synthetic_hard"):
            raise ConfigError("experiment.data_source must be auto|mnist|synthetic|synthetic_hard")
        if exp.get("dtype", "float32") not in ("float32", "float64"):
            raise ConfigError("experiment.dtype must be float32|float64")
    if kind in ("density", "online_density"):
        exp = _fill(exp, {"data": REQUIRED}, "experiment")
        if kind == "online_density":
            exp.setdefault("seed", 0)
    out["experiment"] = exp
    pk = "mnist" if kind.startswith("mnist") else kind
    if kind == "mnist_scaling":
        out["problem"] = validate_problem(dict(out.get("problem", {}), problem_name=out.get("problem", {}).get("problem_name", "trial")),
                                          "problem", pk)
    else:
        if not out.get("problem_configs"):
            raise ConfigError("missing top-level key `problem_configs`")
        out["problem_configs"] = {k: validate_problem(v, f"problem_configs.{k}", pk)
                                  for k, v in out["problem_configs"].items()}
    return out


def load_experiment(yaml_pth: str, kind: str) -> Dict[str, Any]:
    with open(yaml_pth) as f:
        raw = yaml.safe_load(f)
    return validate_experiment(raw, kind)

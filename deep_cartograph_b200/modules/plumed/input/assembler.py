"""PLUMED export of a linear CV: the COMBINE lines the reference's assembler writes
from ``LinearCalculator.get_cv_parameters()`` (``modules/plumed/input/assembler.py:333-379``, command
text as ``modules/plumed/command.py:357-419``), so that the weights and normalisations produced on the
GPU can drive enhanced sampling.  Only this block of the assembler is mirrored; the rest of the PLUMED
input (features, biases, printing) is the reference's MD-side code and out of scope.

    feat_i      = (x_i - features_norm_mean_i) * (1 / features_norm_range_i)     [if a normalisation mode is set]
    <cv>_j      = sum_i weights[i, j] * feat_i
    norm_<cv>_j = (<cv>_j - (min_j + max_j)/2) * 2 / (max_j - min_j)

``evaluate_combine`` applies such lines to a feature table in float64 -- the arithmetic PLUMED's COMBINE
performs -- and is what the N3 test compares with the projected CSV (the reference's own check:
``tests/test_deep_cartograph.py:209-258``, tolerance 1e-2).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np


def combine(command_label: str, arguments: Sequence[str], coefficients=None, parameters=None, powers=None,
            periodic: bool = False) -> str:
    """``label: COMBINE ARG=... COEFFICIENTS=... PARAMETERS=... PERIODIC=NO`` (C = sum c_i (x_i - a_i)^p_i)."""
    cmd = command_label + ": COMBINE ARG=" + ",".join(arguments)
    if coefficients is not None:
        cmd += " COEFFICIENTS=" + ",".join(f"{float(c):.17g}" for c in coefficients)
    if parameters is not None:
        cmd += " PARAMETERS=" + ",".join(f"{float(p):.17g}" for p in parameters)
    if powers is not None:
        cmd += " POWERS=" + ",".join(f"{float(p):.10g}" for p in powers)
    cmd += " PERIODIC=YES" if periodic else " PERIODIC=NO"
    return cmd + "\n"


def validate_linear_cv(cv_params: Dict) -> None:
    for key in ("cv_name", "features_norm_mode", "features_norm_mean", "features_norm_range", "cv_stats", "weights"):
        if key not in cv_params:
            raise ValueError(f"Linear CV parameters must contain '{key}'")
    if cv_params["weights"] is None:
        raise ValueError("Linear CV has no weights")


def add_linear_cv(cv_params: Dict, features_list: List[str]):
    """Returns (input text, labels of the normalised CV components)."""
    validate_linear_cv(cv_params)
    text = ""
    mode = cv_params["features_norm_mode"]
    mean, rng = cv_params["features_norm_mean"], cv_params["features_norm_range"]
    W = np.asarray(cv_params["weights"])
    if mode is not None:
        text += "\n# Normalized features\n"
        feats = []
        for i, feature in enumerate(features_list):
            text += combine(f"feat_{i}", [feature], [1 / rng[i]], [mean[i]])
            feats.append(f"feat_{i}")
    else:
        feats = list(features_list)
    text += "\n# Collective variable\n"
    cv_labels = []
    for j in range(W.shape[1]):
        name = f"{cv_params['cv_name']}_{j}"
        text += combine(name, feats, W[:, j])
        cv_labels.append(name)
    st = cv_params["cv_stats"]
    offset = (np.asarray(st["min"]) + np.asarray(st["max"])) / 2
    scale = 2 / (np.asarray(st["max"]) - np.asarray(st["min"]))
    text += "\n# Normalized Collective variable\n"
    out = []
    for j in range(W.shape[1]):
        name = f"norm_{cv_params['cv_name']}_{j}"
        text += combine(name, [cv_labels[j]], [scale[j]], [offset[j]])
        out.append(name)
    return text, out


def evaluate_combine(text: str, table: Dict[str, np.ndarray], wanted: Optional[List[str]] = None) -> Dict[str, np.ndarray]:
    """Float64 evaluation of the COMBINE lines of ``text`` on the columns of ``table`` (PERIODIC=NO)."""
    values = {k: np.asarray(v, dtype=np.float64) for k, v in table.items()}
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        label, rest = line.split(":", 1)
        tokens = rest.split()
        if tokens[0] != "COMBINE":
            raise ValueError(f"not a COMBINE line: {line}")
        kv = dict(t.split("=", 1) for t in tokens[1:])
        args = kv["ARG"].split(",")
        coef = [float(c) for c in kv["COEFFICIENTS"].split(",")] if "COEFFICIENTS" in kv else [1.0] * len(args)
        par = [float(c) for c in kv["PARAMETERS"].split(",")] if "PARAMETERS" in kv else [0.0] * len(args)
        pw = [float(c) for c in kv["POWERS"].split(",")] if "POWERS" in kv else [1.0] * len(args)
        acc = 0.0
        for a, c, p0, p in zip(args, coef, par, pw):
            term = values[a] - p0
            acc = acc + c * (term if p == 1.0 else term ** p)
        values[label.strip()] = acc
    return {k: values[k] for k in (wanted or values)}

"""PLUMED colvars text -> float32 feature matrix (boundary feeder).

Keeps the semantics the hot path depends on (reference ``modules/plumed/colvars.py:17-60,
254-320, 322-473``): float32 values, ``iloc[start:stop:stride]`` per file BEFORE concatenation,
drop of ``time / *labels / *bias / *walker`` columns, ``features_list`` selects AND orders the
columns, files are concatenated in order into ONE time series.  Cross-topology feature-name
translation is out of scope (needs MDAnalysis): all files must share feature names."""
from __future__ import annotations

import logging
import re
from typing import List, Optional, Union

import numpy as np
import pandas as pd

logger = logging.getLogger(__name__)

_DROP = re.compile(r"labels|time|bias|walker")


def read_column_names(colvars_path: str) -> List[str]:
    with open(colvars_path) as fh:
        head = fh.readline().split()
    if len(head) < 3 or head[0] != "#!" or head[1] != "FIELDS":
        raise ValueError(f"{colvars_path} is not a PLUMED colvars file (missing '#! FIELDS' header)")
    return head[2:]


def read_colvars(colvars_path: str) -> pd.DataFrame:
    names = read_column_names(colvars_path)
    return pd.read_csv(colvars_path, sep=r"\s+", dtype=np.float32, comment="#", header=None, names=names)


def create_dataframe_from_files(colvars_paths: Union[List[str], str],
                                topology_paths=None, reference_topology=None,
                                features_list: Optional[List[str]] = None,
                                file_label: Optional[str] = None,
                                start: int = 0, stop: Optional[int] = None, stride: int = 1,
                                **_ignored) -> pd.DataFrame:
    if isinstance(colvars_paths, str):
        colvars_paths = [colvars_paths]
    frames = []
    for idx, path in enumerate(colvars_paths):
        df = read_colvars(path).iloc[start:stop:stride, :]
        if df.isna().any().any():
            raise ValueError(f"Clean your data! NaNs found in {path}")
        df = df[[c for c in df.columns if not _DROP.search(c)]]
        if features_list:
            missing = set(features_list) - set(df.columns)
            if missing:
                raise ValueError(f"Features {missing} not found in {path}.")
            df = df[list(features_list)]
        if file_label:
            df = df.assign(**{file_label: idx})
        frames.append(df)
    if not features_list:
        for i, df in enumerate(frames[1:], 1):
            if not df.columns.equals(frames[0].columns):
                raise ValueError(f"Column names in {colvars_paths[i]} do not match those in "
                                 f"{colvars_paths[0]}; provide a features_list.")
    out = pd.concat(frames, ignore_index=True)
    if out.empty:
        raise ValueError("The resulting dataframe is empty.")
    return out

"""PLUMED colvars text -> float32 feature matrix (boundary feeder).

Keeps the semantics the hot path depends on (reference ``modules/plumed/colvars.py:17-60,
254-320, 322-473``): float32 values, ``iloc[start:stop:stride]`` per file BEFORE concatenation,
drop of ``time / *labels / *bias / *walker`` columns, ``features_list`` selects AND orders the
columns, files are concatenated in order into ONE time series.  Cross-topology feature-name
translation is out of scope (needs MDAnalysis): all files must share feature names."""
from __future__ import annotations

import json
import logging
import os
import re
from typing import List, Optional, Tuple, Union

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


# ---- binary sidecar ---------------------------------------------------------------
# At C3 size the colvars TEXT is ~0.5 TB and parsing it dwarfs every GPU pass.  A sidecar holds the
# same table as raw float32 (row-major, every column of the text file, in file order) next to a
# small JSON header; it is memory-mapped, sliced and column-selected with exactly the semantics of
# the text path above, and feeds `load_training_tensor`'s chunked host->device streaming without a
# DataFrame in between.  `write_sidecar` is the one-time conversion.

SIDECAR_SUFFIX = ".f32"


def sidecar_paths(colvars_path: str) -> Tuple[str, str]:
    return colvars_path + SIDECAR_SUFFIX, colvars_path + SIDECAR_SUFFIX + ".json"


def write_sidecar(colvars_path: str, chunk_rows: int = 1 << 18) -> str:
    """Convert a colvars text file to its binary sidecar (streamed, `chunk_rows` rows at a time).
    Values are parsed exactly as `read_colvars` does (pandas float32)."""
    names = read_column_names(colvars_path)
    data_path, meta_path = sidecar_paths(colvars_path)
    rows = 0
    with open(data_path + ".tmp", "wb") as out:
        for chunk in pd.read_csv(colvars_path, sep=r"\s+", dtype=np.float32, comment="#", header=None,
                                 names=names, chunksize=chunk_rows):
            arr = np.ascontiguousarray(chunk.to_numpy(dtype=np.float32))
            out.write(arr.tobytes())
            rows += arr.shape[0]
    os.replace(data_path + ".tmp", data_path)
    st = os.stat(colvars_path)
    with open(meta_path, "w") as fh:
        json.dump({"columns": names, "rows": rows, "dtype": "float32",
                   "source_size": st.st_size, "source_mtime_ns": st.st_mtime_ns}, fh)
    return data_path


def has_fresh_sidecar(colvars_path: str) -> bool:
    """A sidecar exists and was made from the text file as it is now (size and mtime recorded)."""
    data_path, meta_path = sidecar_paths(colvars_path)
    if not (os.path.exists(data_path) and os.path.exists(meta_path)):
        return False
    try:
        with open(meta_path) as fh:
            meta = json.load(fh)
    except (OSError, ValueError):
        return False
    if os.path.exists(colvars_path):
        st = os.stat(colvars_path)
        if st.st_size != meta.get("source_size") or st.st_mtime_ns != meta.get("source_mtime_ns"):
            return False
    return os.path.getsize(data_path) == 4 * meta["rows"] * len(meta["columns"])


def open_sidecar(colvars_path: str) -> Tuple[np.memmap, List[str]]:
    data_path, meta_path = sidecar_paths(colvars_path)
    with open(meta_path) as fh:
        meta = json.load(fh)
    cols = list(meta["columns"])
    mm = np.memmap(data_path, dtype=np.float32, mode="r", shape=(int(meta["rows"]), len(cols)))
    return mm, cols


def create_matrix_from_sidecars(colvars_paths: Union[List[str], str],
                                features_list: Optional[List[str]] = None,
                                start: int = 0, stop: Optional[int] = None, stride: int = 1,
                                out: Optional[np.ndarray] = None, chunk_rows: int = 1 << 16,
                                allocator=None, **_ignored) -> Tuple[np.ndarray, List[str], np.ndarray]:
    """Same result as `create_dataframe_from_files(..., file_label='traj_label')` -- per-file
    `[start:stop:stride]`, drop of time / labels / bias / walker columns, `features_list` selects and
    orders, files concatenated in order -- as (float32 matrix, column names, file index per row),
    gathered from the memory-mapped sidecars in row chunks.  `out` may be a preallocated (e.g.
    pinned) buffer of at least the result's shape; `allocator(rows, features)` may provide one once
    the shape is known (the calculators allocate pinned memory this way: no second host copy)."""
    if isinstance(colvars_paths, str):
        colvars_paths = [colvars_paths]
    plans = []
    names_ref = None
    for path in colvars_paths:
        mm, cols = open_sidecar(path)
        kept = [c for c in cols if not _DROP.search(c)]
        if features_list:
            missing = set(features_list) - set(kept)
            if missing:
                raise ValueError(f"Features {missing} not found in {path}.")
            names = list(features_list)
        else:
            names = kept
            if names_ref is not None and names != names_ref:
                raise ValueError(f"Column names in {path} do not match those in {colvars_paths[0]}; "
                                 "provide a features_list.")
        names_ref = names_ref or names
        idx = np.asarray([cols.index(c) for c in names], dtype=np.int64)
        rows = np.arange(mm.shape[0])[start:stop:stride]
        plans.append((path, mm, idx, rows))
    total = sum(len(p[3]) for p in plans)
    if total == 0:
        raise ValueError("The resulting dataframe is empty.")
    f = len(names_ref)
    if out is None and allocator is not None:
        out = allocator(total, f)
    if out is None:
        out = np.empty((total, f), dtype=np.float32)
    elif out.shape[0] < total or out.shape[1] != f or out.dtype != np.float32:
        raise ValueError("out must be a float32 array of at least (rows, features)")
    labels = np.empty(total, dtype=np.int64)
    o = 0
    for file_idx, (path, mm, idx, rows) in enumerate(plans):
        contiguous_cols = bool(len(idx) and np.all(np.diff(idx) == 1))
        for r0 in range(0, len(rows), chunk_rows):
            rr = rows[r0:r0 + chunk_rows]
            block = mm[rr[0]:rr[-1] + 1:stride] if stride >= 1 else mm[rr]
            # like the text path: a NaN in ANY column of the selected rows is an error, also in
            # columns that are dropped or not selected afterwards
            if np.isnan(block).any():
                raise ValueError(f"Clean your data! NaNs found in {path}")
            block = block[:, idx[0]:idx[-1] + 1] if contiguous_cols else block[:, idx]
            out[o:o + len(rr)] = block
            o += len(rr)
        labels[o - len(rows):o] = file_idx
    return out[:total], names_ref, labels

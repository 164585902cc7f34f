"""CV calculators of the hot path: PCA, TICA, hTICA (linear) on the B200 kernels.

Host-side mirror of the reference's ``modules/cv_learning/cv_calculator.py``: same class names,
constructor / method protocol, attribute names, error behaviour and ``model.zip`` layout
(``CVCalculator`` :23, ``LinearCalculator`` :749, ``PCACalculator`` :2174, ``TICACalculator``
:2217, ``HTICACalculator`` :2269, ``cv_calculators_map`` :2952).  The arithmetic runs on the
device through ``deep_cartograph_b200.ops``; nothing here falls back to the CPU.

Differences a maintainer should know (all documented in DESIGN.md):
  * ``training_data`` lives on the GPU and stays RAW; standardisation is fused into the
    covariance and projection kernels (the reference standardises in place at load, :799-804).
    ``standardized_training_data()`` returns the reference's tensor on demand.
  * ``project_data`` does not mutate its input (the reference does, :953-956).
  * ``load_training_tensor`` accepts an in-memory matrix (synthetic benchmarks, sharded runs).
"""
from __future__ import annotations

import copy
import json
import logging
import os
import shutil
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import pandas as pd
import torch

from ... import linalg, ops
from ...parallel import FrameShards
from ..common import unzip_files, zip_files
from ..plumed import colvars as colvars_io
from ..plumed.colvars import create_dataframe_from_files

logger = logging.getLogger(__name__)


def _device(configuration: Dict) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("deep_cartograph_b200 needs a CUDA device (B200); there is no CPU fallback")
    idx = (configuration.get("backend") or {}).get("device")
    if idx is None:
        idx = torch.cuda.current_device()
    return torch.device("cuda", int(idx))


class CVCalculator:
    """Base class (reference cv_calculator.py:23-745, hot-path subset)."""

    def __init__(self, configuration: Optional[Dict] = None, output_path: Optional[str] = None):
        self.configuration: Dict = copy.deepcopy(configuration) if configuration is not None else {}
        self.architecture_config: Dict = self.configuration.get("architecture", {})
        self.training_reading_settings: Dict = self.configuration.get("input_colvars", {}) or {}
        self.feats_norm_mode = self.configuration.get("features_normalization", None)
        self.backend: Dict = self.configuration.get("backend") or {}

        self.ref_topology_path: Optional[str] = None
        self.training_data: Optional[torch.Tensor] = None      # CUDA, float32, RAW features
        self.training_data_labels: Optional[np.ndarray] = None
        self.projection_data_labels: Optional[np.ndarray] = None
        self.shards: Optional[FrameShards] = None              # set for multi-GPU frame sharding

        self.features_ref_labels: List[str] = []
        self.features_stats: Dict[str, np.ndarray] = {}
        self.features_norm_mean: Optional[np.ndarray] = None
        self.features_norm_range: Optional[np.ndarray] = None
        self.num_features: int = 0
        self.num_frames: int = 0                                # global number of frames

        self.cv = None
        self.cv_dimension: int = self.configuration.get("dimension")
        self.cv_labels: List[str] = []
        self.cv_name: Optional[str] = None
        self.cv_range: List[Tuple[float, float]] = []
        self.eigenvalues: Optional[np.ndarray] = None

        self.parent_output_path: Optional[str] = output_path
        self.temp_model_path: Optional[str] = None
        self._dev_norm: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    def __del__(self):
        try:
            if self.temp_model_path and os.path.exists(self.temp_model_path):
                shutil.rmtree(self.temp_model_path)
        except Exception:
            pass

    # ---- model.zip ------------------------------------------------------------------------
    @classmethod
    def load(cls, model_path: str, output_path: str):
        """Factory: restore the right calculator from a ``model.zip`` (reference :92-149)."""
        if not os.path.exists(model_path):
            raise FileNotFoundError(f"Model file not found: {model_path}")
        temp_model_path = os.path.join(output_path, "model")
        unzip_files(model_path, output_path)
        metadata_path = os.path.join(temp_model_path, "metadata.json")
        cv_name = None
        if os.path.exists(metadata_path):
            with open(metadata_path) as fh:
                cv_name = json.load(fh).get("cv_name")
        if not cv_name:
            raise ValueError("Could not determine the CV name from the model file.")
        klass = cv_calculators_map.get(cv_name)
        if not klass:
            raise TypeError(f"Unknown CV calculator name: {cv_name}")
        inst = klass(output_path=output_path)
        inst._load_from_folder(temp_model_path)
        inst.temp_model_path = temp_model_path
        return inst

    def _load_from_folder(self, folder_path: str):
        with open(os.path.join(folder_path, "metadata.json")) as fh:
            meta = json.load(fh)
        self.cv_dimension = meta.get("cv_dimension")
        self.cv_name = meta.get("cv_name")
        self.set_labels()
        self.model_output_folder = os.path.join(self.parent_output_path, self.cv_name, "model")
        if os.path.exists(self.model_output_folder):
            shutil.rmtree(self.model_output_folder)
        shutil.copytree(folder_path, self.model_output_folder)
        with open(os.path.join(self.model_output_folder, "features_labels.txt")) as fh:
            self.features_ref_labels = fh.read().strip().split("\n")
        self.num_features = len(self.features_ref_labels)
        ref_top = os.path.join(self.model_output_folder, "ref_topology.pdb")
        self.ref_topology_path = ref_top if os.path.exists(ref_top) else None

    def create_output_folders(self):
        self.output_path = Path(self.parent_output_path) / self.cv_name
        self.output_path.mkdir(parents=True, exist_ok=True)
        self.training_output_folder = self.output_path / "training"
        self.training_output_folder.mkdir(parents=True, exist_ok=True)
        self.model_output_folder = self.output_path / "model"
        self.model_output_folder.mkdir(parents=True, exist_ok=True)

    def save_model(self):
        """Files common to all calculators (reference :436-452)."""
        with open(os.path.join(self.model_output_folder, "metadata.json"), "w") as fh:
            json.dump({"cv_name": self.cv_name, "cv_dimension": self.cv_dimension}, fh)
        np.savetxt(os.path.join(self.model_output_folder, "features_labels.txt"),
                   self.features_ref_labels, fmt="%s")
        if self.ref_topology_path is not None and os.path.exists(self.ref_topology_path):
            shutil.copyfile(self.ref_topology_path, os.path.join(self.model_output_folder, "ref_topology.pdb"))

    # ---- data -----------------------------------------------------------------------------
    def load_training_data(self, train_colvars_paths: List[str],
                           train_topology_paths: Optional[List[str]] = None,
                           ref_topology_path: Optional[str] = None,
                           features_list: Optional[List[str]] = None):
        """Colvars files -> device matrix + statistics (reference :248-300)."""
        self.ref_topology_path = ref_topology_path
        if train_topology_paths is not None and self.ref_topology_path is None:
            self.ref_topology_path = train_topology_paths[0]
        paths = [train_colvars_paths] if isinstance(train_colvars_paths, str) else list(train_colvars_paths)
        if paths and all(colvars_io.has_fresh_sidecar(p) for p in paths):
            # binary sidecars: memory-mapped float32 tables, no text parsing, gathered
            # straight into a pinned buffer that load_training_tensor streams to the device
            logger.info("Reading training data from binary colvars sidecars...")
            holder = {}

            def pinned(rows, feats):
                try:
                    holder["t"] = torch.empty((rows, feats), dtype=torch.float32, pin_memory=torch.cuda.is_available())
                except RuntimeError:
                    holder["t"] = torch.empty((rows, feats), dtype=torch.float32)
                return holder["t"].numpy()

            Xh, names, labels = colvars_io.create_matrix_from_sidecars(paths, features_list=features_list,
                                                                       allocator=pinned,
                                                                       **self.training_reading_settings)
            X = holder["t"][:Xh.shape[0]]
            self.load_training_tensor(X, names, labels)
            return
        logger.info("Reading training data from colvars files...")
        df = create_dataframe_from_files(colvars_paths=train_colvars_paths,
                                         topology_paths=train_topology_paths,
                                         reference_topology=self.ref_topology_path,
                                         features_list=features_list, file_label="traj_label",
                                         **self.training_reading_settings)
        labels = df.pop("traj_label").to_numpy()
        X = torch.from_numpy(np.ascontiguousarray(df.to_numpy(dtype=np.float32)))
        self.load_training_tensor(X, df.columns.tolist(), labels)

    def load_training_tensor(self, X: torch.Tensor, features_labels: Optional[List[str]] = None,
                             traj_labels: Optional[np.ndarray] = None,
                             shards: Optional[FrameShards] = None):
        """In-memory entry point: ``X`` is this rank's (frames x features) float32 shard
        (CPU or CUDA).  With ``shards`` the statistics are merged across ranks and the lag
        halo is exchanged later."""
        dev = _device(self.configuration)
        if X.dtype != torch.float32:
            X = X.to(torch.float32)
        lag = int(self.configuration.get("lag_time") or 0)
        n_in, f_in = X.shape
        halo = lag if (shards is not None and lag > 0) else 0
        ld = (f_in + 3) // 4 * 4
        aligned = (X.device == dev and X.dim() == 2 and X.stride(1) == 1 and X.stride(0) % 4 == 0
                   and X.data_ptr() % 16 == 0)
        # a resident shard that is the leading view of a buffer with `lag` spare rows receives its halo
        # in place (FrameShards.with_halo): no copy (at C3 a shard is up to 99 GB)
        base = X._base
        room = (halo == 0 or (base is not None and base.dim() == 2 and base.stride(1) == 1
                              and base.stride(0) == X.stride(0) and base.data_ptr() == X.data_ptr()
                              and base.shape[0] >= n_in + halo and base.shape[1] >= f_in))
        self.shards = shards
        self._spec = None
        st = None
        if not (aligned and room):
            # HBM layout: rows padded to a multiple of 4 floats so every row starts 16-byte aligned
            # (16-byte loads in every kernel; e.g. 4950 features -> row stride 4952), plus spare rows
            # after the shard so the lag halo is received in place (no second copy)
            buf = torch.empty((n_in + halo, ld), dtype=torch.float32, device=dev)
            self.training_data = buf[:n_in, :f_in]
            if X.device.type == "cpu" and self.backend.get("overlap_h2d", True):
                st = self._stream_in(X)
            else:
                self.training_data.copy_(X, non_blocking=True)
        else:
            self.training_data = X                       # already resident with 16-byte aligned rows
        self.training_data_labels = traj_labels
        n, f = self.training_data.shape
        self.features_ref_labels = list(features_labels) if features_labels is not None else [f"f{i}" for i in range(f)]
        self.num_features = f
        logger.info(f"Number of features: {self.num_features}")

        if st is None:
            st = ops.column_stats(self.training_data)
        if shards is not None:
            st = shards.merge_stats(st)
        self.num_frames = int(st["n"])
        nn = self.num_frames
        std = torch.sqrt(st["m2"] / (nn - 1)) if nn > 1 else torch.full_like(st["m2"], float("nan"))
        # the reference's statistics are float32 (pandas on float32 columns)
        self.features_stats = {
            "mean": st["mean"].to(torch.float32).cpu().numpy(),
            "std": std.to(torch.float32).cpu().numpy(),
            "min": st["min"].cpu().numpy(),
            "max": st["max"].cpu().numpy(),
        }
        self.features_norm_mean, self.features_norm_range = self.prepare_normalization()
        self._dev_norm = None
        self._dev_bounds = None

    # ---- host -> device streaming with speculative sums ------------------------------------------
    def _sums_plan(self):
        """(lag, block, want_st) of the covariance sums ``compute_cv`` will ask for, or None."""
        return None

    def _provisional_norm(self, st: dict):
        """Device-side (mean, range) of the configured mode from the statistics of the rows seen so
        far (float32, |range| < 1e-8 -> 1).  Only used to condition the speculative sums."""
        n = st["n"]
        mean = st["mean"].to(torch.float32)
        if self.feats_norm_mode is None:
            return torch.zeros_like(mean), torch.ones_like(mean)
        if self.feats_norm_mode == "mean_std":
            m, r = mean, torch.sqrt(st["m2"] / max(n - 1, 1)).to(torch.float32)
        elif self.feats_norm_mode == "min_max_range1":
            m, r = st["min"], st["max"] - st["min"]
        elif self.feats_norm_mode == "min_max_range2":
            m, r = (st["min"] + st["max"]) / 2, (st["max"] - st["min"]) / 2
        else:
            raise ValueError(f"Normalization mode {self.feats_norm_mode} not recognized.")
        return m.contiguous(), torch.where(r.abs() < 1e-8, torch.ones_like(r), r).contiguous()

    def _stream_in(self, X: torch.Tensor) -> dict:
        """Copy the host matrix to the device in row chunks on a copy stream; as each chunk lands the
        compute stream takes its column statistics and -- when the calculator knows which sums it
        will need -- accumulates the lagged covariance sums of the rows seen so far under
        PROVISIONAL standardisation parameters (from a strided row sample).  The exact sums follow
        from the provisional ones by an FP64 affine correction once the global statistics are known
        (``linalg.restandardize_sums``), so the contraction overlaps the PCIe transfer instead of
        following it.  Returns the merged column statistics of this shard."""
        data = self.training_data
        n, f = data.shape
        dev = data.device
        plan = self._sums_plan() if self.backend.get("speculative_sums", True) else None
        chunk_bytes = int(self.backend.get("h2d_chunk_bytes", 256 << 20))
        rows = max(1, min(n, chunk_bytes // max(1, 4 * f)))
        if plan is not None:
            rows = max(rows, 4 * plan[0] + 1)
        bounds = [(c0, min(n, c0 + rows)) for c0 in range(0, n, rows)]
        if len(bounds) > 1 and bounds[-1][1] - bounds[-1][0] <= (plan[0] if plan else 0):
            bounds[-2] = (bounds[-2][0], n)                   # no chunk shorter than the lag
            bounds.pop()
        main = torch.cuda.current_stream(dev)
        copier = torch.cuda.Stream(dev)
        copier.wait_stream(main)                              # the buffer was allocated on `main`
        events = []
        with torch.cuda.stream(copier):
            for c0, c1 in bounds:
                data[c0:c1].copy_(X[c0:c1], non_blocking=True)
                e = torch.cuda.Event()
                e.record(copier)
                events.append(e)
        parts, acc, norm0, start = [], None, None, 0
        lo = hi = None
        engine = self.backend.get("cov_engine")
        if plan is not None and self.feats_norm_mode is not None:
            # provisional parameters: statistics of ~2048 evenly spaced rows, taken on the host before
            # any byte has moved (spans the whole series, so slow drifts do not bias it the way the
            # first chunk's statistics would)
            idx = torch.linspace(0, n - 1, steps=min(n, 2048)).round().to(torch.int64)
            smp = X.index_select(0, idx).to(torch.float64)
            m = smp.shape[0]
            mu = smp.mean(dim=0)
            norm0 = None if self.feats_norm_mode is None else self._provisional_norm({
                "n": m, "mean": mu.to(dev), "m2": ((smp - mu) ** 2).sum(dim=0).to(dev),
                "min": smp.min(dim=0).values.to(torch.float32).to(dev),
                "max": smp.max(dim=0).values.to(torch.float32).to(dev)})
        for (c0, c1), e in zip(bounds, events):
            main.wait_event(e)
            parts.append(ops.column_stats(data[c0:c1]))
            if plan is None:
                continue
            lag, block, want_st = plan
            if c1 - start <= lag:
                continue
            # column bounds of every row seen so far (a superset of this chunk's rows)
            lo = parts[-1]["min"] if lo is None else torch.minimum(lo, parts[-1]["min"])
            hi = parts[-1]["max"] if hi is None else torch.maximum(hi, parts[-1]["max"])
            s = ops.lagged_covariance(data[start:c1], lag, norm0[0] if norm0 else None,
                                      norm0[1] if norm0 else None, block=block, engine=engine, want_st=want_st,
                                      xmin=lo, xmax=hi)
            s.pop("clamped", None)
            if acc is None:
                acc = s
            else:
                for k in ("S0", "St", "a", "b"):
                    if s.get(k) is not None:
                        acc[k] += s[k]
                acc["M"] += s["M"]
            start = c1 - lag                                  # the last `lag` rows pair with the next chunk
        if plan is not None and acc is not None:
            self._spec = {"plan": plan, "norm0": norm0, "sums": acc, "rows_done": start + plan[0]}
        return ops.merge_column_stats(parts)

    def prepare_normalization(self) -> Tuple[np.ndarray, np.ndarray]:
        """(mean, range) per normalisation mode; |range| < 1e-8 -> 1 (reference :308-363)."""
        s = self.features_stats
        if self.feats_norm_mode is None:
            means = np.zeros(len(s["mean"]))
            ranges = np.ones(len(s["mean"]))
        elif self.feats_norm_mode == "mean_std":
            means, ranges = s["mean"], s["std"]
        elif self.feats_norm_mode == "min_max_range1":
            means, ranges = s["min"], s["max"] - s["min"]
        elif self.feats_norm_mode == "min_max_range2":
            means, ranges = (s["min"] + s["max"]) / 2, (s["max"] - s["min"]) / 2
        else:
            logger.error(f"Normalization mode {self.feats_norm_mode} not recognized. Exiting...")
            raise ValueError(f"Normalization mode {self.feats_norm_mode} not recognized.")
        ranges = np.array(ranges, copy=True)
        small = np.abs(ranges) < 1e-8
        for i in np.flatnonzero(small):
            logger.warning(f"Range for feature {i} is close to zero. Setting it to 1.0.")
        ranges[small] = 1.0
        return np.array(means, copy=True), ranges

    def _norm_on_device(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.features_norm_mean is None or self.features_norm_range is None:
            raise ValueError("Feature normalization parameters have not been computed. Cannot normalize data.")
        if self._dev_norm is None:
            dev = self.training_data.device if self.training_data is not None else _device(self.configuration)
            self._dev_norm = (torch.tensor(self.features_norm_mean, dtype=torch.float32, device=dev),
                              torch.tensor(self.features_norm_range, dtype=torch.float32, device=dev))
        return self._dev_norm

    def _bounds_on_device(self):
        """(min, max) of every feature over ALL frames (all shards): the column bounds the exact integer
        covariance engine scales its fixed point by."""
        if getattr(self, "_dev_bounds", None) is None:
            dev = self.training_data.device
            self._dev_bounds = (torch.tensor(self.features_stats["min"], dtype=torch.float32, device=dev),
                                torch.tensor(self.features_stats["max"], dtype=torch.float32, device=dev))
        return self._dev_bounds

    def standardized_training_data(self) -> torch.Tensor:
        """The tensor the reference holds in ``training_data`` after load (:799-804)."""
        m, r = self._norm_on_device()
        return ops.standardize_(self.training_data.clone(), m, r)

    def cv_ready(self) -> bool:
        return self.cv is not None

    # ---- main flow (reference :366-414) ------------------------------------------------------
    def run(self, cv_dimension: Union[int, None] = None) -> Union[pd.DataFrame, None]:
        if self.training_data is None:
            logger.error("Training data not loaded. Cannot compute CV.")
            return None
        self.create_output_folders()
        if cv_dimension:
            self.cv_dimension = cv_dimension
        self.compute_cv()
        self.set_labels()
        projection_df = None
        if self.cv is not None:
            projection = self.normalize_cv()          # one fused pass gives P and its min/max
            self.save_model()
            projection_df = pd.DataFrame(projection.cpu().numpy(), columns=self.cv_labels)
        return projection_df

    def compute_cv(self):
        raise NotImplementedError

    def set_labels(self):
        self.cv_labels = [f"{cv_components_map[self.cv_name]} {i + 1}" for i in range(self.cv_dimension)]

    def get_labels(self) -> List[str]:
        return self.cv_labels

    def get_cv_dimension(self) -> int:
        return self.cv_dimension

    def get_range(self):
        return self.cv_range

    def project_colvars(self, colvars_paths: Union[List[str], str],
                        topology_paths: Union[List[str], str, None] = None) -> Union[pd.DataFrame, None]:
        """Colvars file(s) -> projected DataFrame (reference :478-526)."""
        df = create_dataframe_from_files(colvars_paths=colvars_paths, topology_paths=topology_paths,
                                         reference_topology=self.ref_topology_path,
                                         features_list=self.features_ref_labels, file_label="traj_label")
        self.projection_data_labels = df.pop("traj_label").to_numpy()
        X = torch.from_numpy(np.ascontiguousarray(df.to_numpy(dtype=np.float32)))
        P = self.project_data(X)
        return pd.DataFrame(P.cpu().numpy(), columns=self.cv_labels)


class LinearCalculator(CVCalculator):
    """Linear CVs: weights F x d + feature / CV normalisation (reference :749-1047)."""

    def __init__(self, configuration: Optional[Dict] = None, output_path: Optional[str] = None):
        super().__init__(configuration, output_path)
        self.cv: Optional[np.ndarray] = None
        self.weights_path: Optional[str] = None
        self.cv_stats: Dict[str, np.ndarray] = {}
        self.cv_norm_mean: Optional[np.ndarray] = None
        self.cv_norm_range: Optional[np.ndarray] = None

    def _load_from_folder(self, folder_path: str):
        super()._load_from_folder(folder_path)
        f = self.model_output_folder
        self.cv = np.load(os.path.join(f, "cv_weights.npy"))
        self.cv_norm_mean = np.load(os.path.join(f, "cv_norm_mean.npy"))
        self.cv_norm_range = np.load(os.path.join(f, "cv_norm_range.npy"))
        self.features_norm_mean = np.load(os.path.join(f, "features_norm_mean.npy"))
        self.features_norm_range = np.load(os.path.join(f, "features_norm_range.npy"))

    def normalize_data(self, data: torch.Tensor, normalizing_mean: torch.Tensor,
                       normalizing_range: torch.Tensor) -> torch.Tensor:
        """In-place ``(data - mean) / range`` on the device (reference :806-837)."""
        dev = data.device if data.is_cuda else _device(self.configuration)
        out = data if data.is_cuda else data.to(dev)
        ops.standardize_(out, normalizing_mean.to(dev, torch.float32), normalizing_range.to(dev, torch.float32))
        if not data.is_cuda:
            data.copy_(out.cpu())
        return data

    def save_weights(self, weights_path: str):
        np.save(weights_path, self.cv)

    def save_model(self):
        """model.zip with the reference's member names and float32 dtypes (:853-892)."""
        super().save_model()
        if self.cv is None:
            raise ValueError("No Linear CV weights to save. Please compute the CV before saving the model.")
        if self.cv_norm_mean is None or self.cv_norm_range is None:
            raise ValueError("CV normalization parameters have not been computed. Cannot save model.")
        if self.features_norm_mean is None or self.features_norm_range is None:
            raise ValueError("Features normalization parameters have not been computed. Cannot save model.")
        f = self.model_output_folder
        self.save_weights(os.path.join(f, "cv_weights.npy"))
        np.save(os.path.join(f, "cv_norm_mean.npy"), self.cv_norm_mean)
        np.save(os.path.join(f, "cv_norm_range.npy"), self.cv_norm_range)
        np.save(os.path.join(f, "features_norm_mean.npy"), self.features_norm_mean)
        np.save(os.path.join(f, "features_norm_range.npy"), self.features_norm_range)
        model_path = os.path.join(self.output_path, "model.zip")
        zip_files(model_path, str(f))
        shutil.rmtree(f)
        logger.info(f"Model saved to {model_path}")

    def get_cv_parameters(self) -> Dict:
        return {"cv_name": self.cv_name, "cv_dimension": self.cv_dimension,
                "features_norm_mode": self.feats_norm_mode,
                "features_norm_mean": self.features_norm_mean,
                "features_norm_range": self.features_norm_range,
                "cv_stats": self.cv_stats, "weights": self.cv}

    def get_cv_type(self) -> str:
        return "linear"

    def _weights_on(self, dev) -> torch.Tensor:
        return torch.tensor(np.asarray(self.cv), dtype=torch.float32, device=dev)

    def project_data(self, data: torch.Tensor, normalize_data: bool = True) -> torch.Tensor:
        """Project onto the normalised CV space (reference :918-972): optional feature
        standardisation, ``@ W``, then ``(P - cv_norm_mean) / cv_norm_range``.
        Returns a tensor on the device of ``data``."""
        if self.cv is None:
            logger.error("CV has not been computed. Cannot project data.")
            raise ValueError("CV has not been computed. Cannot project data.")
        if normalize_data and (self.features_norm_mean is None or self.features_norm_range is None):
            logger.error("Feature normalization parameters have not been computed. Cannot normalize data.")
            raise ValueError("Feature normalization parameters have not been computed. Cannot normalize data.")
        if self.cv_norm_mean is None or self.cv_norm_range is None:
            logger.error("CV normalization parameters have not been computed. Cannot normalize projected data.")
            raise ValueError("CV normalization parameters have not been computed. Cannot normalize projected data.")
        dev = data.device if data.is_cuda else _device(self.configuration)
        X = data.to(dev, torch.float32).contiguous()
        mean = rng = None
        if normalize_data:
            mean = torch.tensor(self.features_norm_mean, dtype=torch.float32, device=dev)
            rng = torch.tensor(self.features_norm_range, dtype=torch.float32, device=dev)
        P, _, _ = ops.project(X, self._weights_on(dev), mean, rng, minmax=False)
        ops.standardize_(P, torch.tensor(self.cv_norm_mean, dtype=torch.float32, device=dev),
                         torch.tensor(self.cv_norm_range, dtype=torch.float32, device=dev))
        return P if data.is_cuda else P.cpu()

    def normalize_cv(self) -> torch.Tensor:
        """CV min-max normalisation from the training projection (reference :974-991) and the
        normalised training projection itself (reference run(), :403-406) from ONE fused pass
        over the raw training data.  Returns the projected, normalised training shard."""
        if self.training_data is None:
            raise ValueError("Training data not loaded. Cannot compute CV statistics for normalization.")
        dev = self.training_data.device
        mean, rng = self._norm_on_device()
        P, pmin, pmax = ops.project(self.training_data, self._weights_on(dev), mean, rng, minmax=True)
        if self.shards is not None:
            pmin, pmax = self.shards.allreduce_minmax(pmin, pmax)
        mn = pmin.cpu().numpy()
        mx = pmax.cpu().numpy()
        self.cv_stats = {"min": mn, "max": mx}
        self.cv_norm_mean = (mx + mn) / 2
        self.cv_norm_range = (mx - mn) / 2
        ops.standardize_(P, torch.tensor(self.cv_norm_mean, dtype=torch.float32, device=dev),
                         torch.tensor(self.cv_norm_range, dtype=torch.float32, device=dev))
        return P

    # ---- shared by TICA / hTICA / PCA ---------------------------------------------------------
    def _kernel_norm(self):
        """(mean, range) handed to the fused kernels, or (None, None) when ``features_normalization``
        is None: raw features are then contracted as they are, and the ``auto`` engine stays on
        3xTF32 (FP16 pieces overflow at |x| >= 65504 and lose their low piece to subnormals for tiny
        features)."""
        if self.feats_norm_mode is None:
            return None, None
        return self._norm_on_device()

    def _lagged_sums(self, lag: int, block: int = 0, want_st: bool = True) -> dict:
        """Raw FP64 sums over the time-lagged pairs of this rank's shard (+ halo), summed over
        ranks.  Pairs: x_t = Z[:N-lag], x_lag = Z[lag:] on the CONCATENATED series (files are
        one time series, reference :278-300, 2247)."""
        mean, rng = self._kernel_norm()
        X = self.training_data
        if self.shards is not None:
            X = self.shards.with_halo(X, lag)
        engine = self.backend.get("cov_engine")
        s = self._take_speculative_sums(X, lag, block, want_st, mean, rng, engine)
        if s is None:
            xmin, xmax = self._bounds_on_device()
            s = ops.lagged_covariance(X, lag, mean, rng, block=block, engine=engine, want_st=want_st,
                                      xmin=xmin, xmax=xmax)
        s.pop("clamped", None)
        if self.shards is not None:
            s = self.shards.allreduce_sums(s, m_total=(self.num_frames - lag) if self.num_frames else None)
        # ALWAYS rebuild the lower triangle from the upper one after the reduction: a rank that kept
        # its speculative sums contributes a full symmetric S0, a rank that recomputed contributes the
        # upper triangle only (the kernel skips tiles below the diagonal) -- the upper triangle of the
        # sum is right in both cases, the lower one only if every rank made the same choice.
        s["S0"] = ops.symmetrize_upper(s["S0"])
        return s

    def _take_speculative_sums(self, X, lag, block, want_st, mean, rng, engine):
        """Exact sums of this rank's rows from the ones accumulated during the host->device copy,
        or None (no speculation, other parameters, or provisional statistics too far from the final
        ones: the split-precision contraction is accurate relative to sum z'^2 = M (1 + delta^2), so
        a provisional centre more than one standard deviation off, or a scale off by more than 2x,
        would cost accuracy -- then the sums are recomputed on the resident matrix)."""
        spec, self._spec = getattr(self, "_spec", None), None
        if spec is None or spec["plan"] != (lag, block, want_st):
            return None
        norm0 = spec["norm0"]
        if (norm0 is None) != (mean is None):
            return None
        if norm0 is not None:
            m0, r0 = norm0
            be = ((m0.double() - mean.double()) / rng.double()).abs().max()
            al = r0.double() / rng.double()
            worst = torch.stack([be, al.max(), 1.0 / al.min()]).tolist()
            if not (worst[0] <= 1.0 and worst[1] <= 2.0 and worst[2] <= 2.0):
                logger.debug("Speculative sums discarded (provisional statistics off by %.2f sigma)" % worst[0])
                return None
        else:
            m0 = r0 = None
        s = spec["sums"]
        done = spec["rows_done"]                               # pairs t < done - lag are in `s`
        if X.shape[0] > done:                                  # lag halo of the next shard arrived since
            t = ops.lagged_covariance(X[done - lag:], lag, m0, r0, block=block, engine=engine, want_st=want_st)
            t.pop("clamped", None)
            for k in ("S0", "St", "a", "b"):
                if t.get(k) is not None:
                    s[k] += t[k]
            s["M"] += t["M"]
        if norm0 is None:
            return s                                           # raw features: the sums are already exact
        # restandardize_sums needs the full symmetric S0 (outside the diagonal blocks of a block-mode
        # pass the kernel leaves zeros and the consumers never look)
        s["S0"] = ops.symmetrize_upper(s["S0"])
        return linalg.restandardize_sums(s, m0, r0, mean, rng)


class PCACalculator(LinearCalculator):
    """PCA (reference :2174-2215): sklearn ``PCA(n_components).fit`` == eigh of the covariance
    of the standardised data; first weight of every component made non-negative."""

    def __init__(self, configuration: Optional[Dict] = None, output_path: Optional[str] = None):
        super().__init__(configuration, output_path)
        self.cv_name = "pca"

    def _sums_plan(self):
        return (0, 0, False)

    def compute_cv(self):
        if self.training_data is None:
            logger.error("No training data available to compute PCA.")
            return
        s = self._lagged_sums(0, want_st=False)            # lag 0: plain Gram over all rows
        evals, W = linalg.pca_from_sums(s["S0"], s["a"], s["M"], self.cv_dimension)
        self.eigenvalues = evals.cpu().numpy()
        self.cv = W.to(torch.float32).cpu().numpy()


class TICACalculator(LinearCalculator):
    """TICA (reference :2217-2267; mlcolvar TICA.compute with remove_average=True, reg 1e-6)."""

    def __init__(self, configuration: Optional[Dict] = None, output_path: Optional[str] = None):
        super().__init__(configuration, output_path)
        self.cv_name = "tica"

    def _sums_plan(self):
        lag = self.configuration.get("lag_time")
        return (int(lag), 0, True) if lag else None

    def compute_cv(self):
        lag = self.configuration.get("lag_time")
        try:
            s = self._lagged_sums(lag)
            evals, V = linalg.tica_from_sums(s["S0"], s["St"], s["a"], s["b"], s["M"], self.cv_dimension)
        except Exception as e:   # reference :2259-2264: log, leave self.cv unset
            logger.error(f"TICA could not be computed. Error message: {e}")
            return
        self.eigenvalues = evals.cpu().numpy()
        self.cv = V.to(torch.float32).cpu().numpy()


class HTICACalculator(LinearCalculator):
    """Hierarchical TICA (reference :2269-2384)."""

    def __init__(self, configuration: Optional[Dict] = None, output_path: Optional[str] = None):
        super().__init__(configuration, output_path)
        self.cv_name = "htica"
        self.num_subspaces = self.configuration.get("num_subspaces")
        self.subspaces_dimension = self.configuration.get("subspaces_dimension")

    def _sums_plan(self):
        lag = self.configuration.get("lag_time")
        if not lag or self.training_data is None or not self.num_subspaces:
            return None
        F = self.training_data.shape[1]
        if F // self.num_subspaces == 0:
            return None
        full = F <= int(self.backend.get("htica_full_gram_max_features", 2048))
        return (int(lag), 0 if full else F // self.num_subspaces, True)

    def compute_cv(self):
        lag = self.configuration.get("lag_time")
        F = self.num_features
        chunks = linalg.htica_chunks(F, self.num_subspaces)
        if not chunks:
            logger.error(f"Number of subspaces {self.num_subspaces} is larger than number of features {F}. Exiting...")
            return
        full_max = int(self.backend.get("htica_full_gram_max_features", 2048))
        try:
            if F <= full_max:
                # one data pass: full Gram; level 2 = T1^T S T1 (exact)
                s = self._lagged_sums(lag)
                W, _, V2 = linalg.htica_from_full_sums(s["S0"], s["St"], s["a"], s["b"], s["M"],
                                                       self.num_subspaces, self.subspaces_dimension,
                                                       self.cv_dimension)
            else:
                # block-diagonal level 1 (1/num_subspaces of the MMA work), then a second pass
                # over the level-1 projections for the small level-2 problem
                width = F // self.num_subspaces
                s = self._lagged_sums(lag, block=width)
                T1 = linalg.htica_level1(s["S0"], s["St"], s["a"], s["b"], s["M"], chunks,
                                         self.subspaces_dimension)
                W = self._htica_level2(T1, lag)
        except Exception as e:   # reference :2352-2357, 2376-2381
            logger.error(f"TICA could not be computed. Error message: {e}")
            return
        self.cv = W.to(torch.float32).cpu().numpy()

    def _htica_level2(self, T1: torch.Tensor, lag: int) -> torch.Tensor:
        mean, rng = self._norm_on_device()
        T1f = T1.to(torch.float32)
        # T1 is block diagonal: block b projects only its own columns of X, so the level-1
        # projections are ONE pass over X in total (each call reads one column block)
        chunks = linalg.htica_chunks(self.num_features, self.num_subspaces)
        X = self.training_data
        width, sdim = chunks[0][1] - chunks[0][0], int(self.subspaces_dimension)
        if (sdim <= 16 and width <= 1018 and X.data_ptr() % 16 == 0 and X.stride(0) % 4 == 0
                and X.stride(1) == 1):
            # compact (F, s) weights: row i = the weights of feature i inside its own block
            Wc = torch.zeros((self.num_features, sdim), dtype=torch.float32, device=T1f.device)
            c = 0
            for (s0, e0) in chunks:
                w = min(sdim, e0 - s0)
                Wc[s0:e0, :w] = T1f[s0:e0, c:c + w]
                c += w
            P = ops.project_blocks(X, Wc, width, mean, rng)
        else:
            parts = []
            c = 0
            for (s0, e0) in chunks:
                w = min(sdim, e0 - s0)
                Pc, _, _ = ops.project(X[:, s0:e0], T1f[s0:e0, c:c + w].contiguous(),
                                       mean[s0:e0], rng[s0:e0], minmax=False)
                parts.append(Pc)
                c += w
            P = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        if self.shards is not None:
            P = self.shards.with_halo(P, lag)
        s = ops.lagged_covariance(P, lag, engine=self.backend.get("cov_engine"))
        s.pop("clamped", None)
        if self.shards is not None:
            s = self.shards.allreduce_sums(s)
        S0 = ops.symmetrize_upper(s["S0"])
        _, V2 = linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], self.cv_dimension)
        return T1 @ V2


class DeepTICACalculator(CVCalculator):
    """DeepTICA (reference ``NonLinear`` + ``DeepTICACalculator``, cv_calculator.py:1038-1891,
    2517-2611): a feed-forward network trained on time-lagged pairs with the loss
    ``-sum lambda_i^2`` of the Cholesky-reduced eigenvalues of the minibatch C0 / C_tau (mlcolvar
    ``DeepTICA.training_step``), whose correlation sums run in ``dcg_ticacov_f32``.

    The reference drives mlcolvar through a lightning ``Trainer``; here the loop is plain PyTorch
    with the same knobs (``architecture.encoder``, ``training.general`` / ``early_stopping`` /
    ``optimizer`` / ``model_to_save``, ``tica_regularization``) and the same decisions: ``num_tries``
    seeded trainings (seed + try), random train/validation split by ``lengths``, validation every
    ``check_val_every_n_epoch`` epochs, early stopping on ``valid_loss`` (patience, min_delta), best
    or last weights, a try rejected when its score is below -d (impossible for autocorrelations).
    The matrix stays resident on the device and minibatches are gathered there (raw features;
    ``norm_in`` is part of the model, reference :1366-1374).  The linear TICA read-out is fitted on
    the whole training set (the reference keeps the last minibatch's), then the outputs are min-max
    normalised to [-1, 1] (:1735-1754).  Training results are seed-dependent in the reference too
    (unpinned); what is pinned is the arithmetic of the loss (tests) and the model graph."""

    def __init__(self, configuration: Optional[Dict] = None, output_path: Optional[str] = None):
        super().__init__(configuration, output_path)
        self.cv_name = "deep_tica"
        tr = self.configuration.get("training") or {}
        g = tr.get("general") or {}
        self.num_tries = int(g.get("num_tries", 10))
        self.seed = int(g.get("seed", 42))
        self.lengths = list(g.get("lengths", [0.8, 0.2]))
        self.batch_size = int(g.get("batch_size", 32))
        self.max_epochs = int(g.get("max_epochs", 1000))
        self.shuffle = bool(g.get("shuffle", False))
        self.random_split = bool(g.get("random_split", True))
        self.check_val_every = max(1, int(g.get("check_val_every_n_epoch", 10)))
        es = tr.get("early_stopping") or {}
        self.patience = int(es.get("patience", 20))
        self.min_delta = float(es.get("min_delta", 1e-5))
        opt = tr.get("optimizer") or {}
        self.opt_name = opt.get("name", "Adam")
        self.opt_kwargs = dict(opt.get("kwargs") or {"lr": 1e-4, "weight_decay": 0.0})
        self.model_to_save = tr.get("model_to_save", "best")
        enc = (self.architecture_config.get("encoder") or {})
        self.hidden_layers = list(enc.get("layers", [64, 32, 16]))
        act = enc.get("activation", "leaky_relu")
        self.activation = (act[0] if isinstance(act, (list, tuple)) and act else act) or "leaky_relu"
        self.reg = float(self.configuration.get("tica_regularization", 1e-6))
        self.best_score: Optional[float] = None
        self.tries_report: List[Dict] = []

    def get_cv_type(self) -> str:
        return "non-linear"

    # ---- training ---------------------------------------------------------------------------------
    def _split(self, M: int, gen: torch.Generator):
        n_train = int(M * self.lengths[0])
        if self.random_split:
            perm = torch.randperm(M, generator=gen)
        else:
            perm = torch.arange(M)
        return perm[:n_train], perm[n_train:]

    def _epoch_batches(self, idx: torch.Tensor, gen: torch.Generator, shuffle: bool):
        if shuffle:
            idx = idx[torch.randperm(idx.numel(), generator=gen)]
        bs = self.batch_size
        for s0 in range(0, idx.numel(), bs):
            yield idx[s0:s0 + bs]

    def train(self) -> bool:
        from .deep_tica import DeepTICA
        from .deep_tica import allreduce_gradients_
        X = self.training_data
        dev = X.device
        lag = int(self.configuration.get("lag_time") or 1)
        sh = self.shards
        if sh is not None:
            # frame-sharded training (exact all-reduce): own rows + lag halo; every rank
            # runs the same number of equally sized local minibatches (the loss has collectives),
            # so all ranks use the smallest local pair count
            X = sh.with_halo(X, lag)
            M_loc = X.shape[0] - lag
            mt = torch.tensor([float(M_loc)], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(mt, op=torch.distributed.ReduceOp.MIN, group=sh.group)
            M = int(mt.item())
        else:
            M = X.shape[0] - lag
        if M < 4:
            logger.error("Not enough frames to build time-lagged pairs.")
            return False
        mean, rng = self._norm_on_device()
        d = self.cv_dimension
        layers = [self.num_features] + self.hidden_layers + [d]
        n_train = int(M * self.lengths[0])
        if self.batch_size >= n_train:                       # reference check_batch_size (:1297-1310)
            self.batch_size = 1 << max(0, (n_train.bit_length() - 1))
            logger.warning(f"Batch size larger than the training set; using {self.batch_size}")
        best = None
        for attempt in range(1, self.num_tries + 1):
            seed = self.seed + attempt
            torch.manual_seed(seed)
            gen = torch.Generator().manual_seed(seed)
            model = DeepTICA(layers, mean, rng, activation=self.activation, reg=self.reg).to(dev)
            opt = getattr(torch.optim, self.opt_name)(model.nn.parameters(), **self.opt_kwargs)
            tr_idx, va_idx = self._split(M, gen)
            tr_idx, va_idx = tr_idx.to(dev), va_idx.to(dev)
            best_val, best_state, bad, last_val = float("inf"), None, 0, None
            history = []
            for epoch in range(self.max_epochs):
                model.train()
                tot, nb = torch.zeros((), dtype=torch.float64, device=dev), 0      # no host read per minibatch
                for b in self._epoch_batches(tr_idx, gen, self.shuffle):
                    if b.numel() <= d + 1:
                        continue
                    loss, _ = model.loss_indexed(X, b, lag, shards=sh)
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    allreduce_gradients_(model.nn, sh)
                    opt.step()
                    tot += loss.detach(); nb += 1
                if (epoch + 1) % self.check_val_every:
                    continue
                model.eval()
                with torch.no_grad():
                    vt, vn = 0.0, 0
                    for b in self._epoch_batches(va_idx, gen, False):
                        if b.numel() <= d + 1:
                            continue
                        vl, _ = model.loss_indexed(X, b, lag, shards=sh)
                        vt += float(vl); vn += 1
                last_val = vt / max(vn, 1)
                history.append({"epoch": epoch + 1, "train_loss": float(tot) / max(nb, 1), "valid_loss": last_val})
                if last_val < best_val - self.min_delta:
                    best_val, bad = last_val, 0
                    best_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
                else:
                    bad += 1
                    if bad >= self.patience:
                        break
            if last_val is None:                             # never validated: validate once
                model.eval()
                with torch.no_grad():
                    last_val = float(model.loss_indexed(X, va_idx, lag, shards=sh)[0]) if va_idx.numel() > d + 1 else float("nan")
                best_val = last_val
            score = best_val if self.model_to_save == "best" else last_val
            ok = bool(np.isfinite(score)) and score >= -float(d) - 1e-6      # reference :1624-1626
            self.tries_report.append({"try": attempt, "seed": seed, "score": score, "accepted": ok,
                                      "epochs": len(history) * self.check_val_every})
            if not ok:
                logger.warning(f"DeepTICA try {attempt}: score {score} rejected")
                continue
            if self.model_to_save == "best" and best_state is not None:
                model.load_state_dict(best_state)
            if best is None or score < best[0]:
                best = (score, model)
        if best is None:
            logger.error("DeepTICA training failed in every try.")
            return False
        self.best_score, self.cv = best
        self.cv.eval()
        return True

    def compute_cv(self):
        if self.training_data is None:
            logger.error("No training data available to train DeepTICA.")
            return
        if not self.train():
            self.cv = None
            return
        # linear TICA read-out from the whole training set, streamed in batches (FP64 raw sums)
        X, lag = self.training_data, int(self.configuration.get("lag_time") or 1)
        if self.shards is not None:
            X = self.shards.with_halo(X, lag)
        M = X.shape[0] - lag
        acc = None
        with torch.no_grad():
            for s0 in range(0, M, 1 << 16):
                e0 = min(M, s0 + (1 << 16))
                s = ops.ticacov_sums(self.cv.features(X[s0:e0]), self.cv.features(X[s0 + lag:e0 + lag]))["flat"]
                acc = s if acc is None else acc + s
            if self.shards is not None:                          # global sums and pair count
                acc = self.shards.allreduce_sum_(torch.cat([acc, torch.tensor([float(M)], dtype=torch.float64,
                                                                                device=acc.device)]))
                M = int(round(float(acc[-1].item())))
                acc = acc[:-1]
            d_, o_ = self.cv_dimension, 2 + self.cv_dimension
            acc = {"swf": acc[2:o_], "sff": acc[o_:o_ + d_ * d_].view(d_, d_),
                   "sfg": acc[o_ + d_ * d_:o_ + 2 * d_ * d_].view(d_, d_), "slg": acc[o_ + 2 * d_ * d_ + d_:]}
            evals, V = linalg.tica_from_sums(acc["sff"], acc["sfg"], acc["swf"], acc["slg"], M,
                                             self.cv_dimension, self.reg)
            self.cv.tica_mean.copy_((acc["swf"] / M).to(torch.float32))
            self.cv.tica_evecs.copy_(V.to(torch.float32))
        self.eigenvalues = evals.cpu().numpy()
        logger.info(f"DeepTICA eigenvalues: {self.eigenvalues}")

    # ---- projection (reference :1735-1754, 1842-1891) ------------------------------------------
    @torch.no_grad()
    def _forward(self, X: torch.Tensor) -> torch.Tensor:
        out = torch.empty((X.shape[0], self.cv_dimension), dtype=torch.float32, device=X.device)
        for s0 in range(0, X.shape[0], 1 << 16):
            out[s0:s0 + (1 << 16)] = self.cv(X[s0:s0 + (1 << 16)])
        return out

    def normalize_cv(self) -> torch.Tensor:
        self.cv.out_mean.zero_()
        self.cv.out_range.fill_(1.0)
        P = self._forward(self.training_data)
        mn, mx = P.min(dim=0).values, P.max(dim=0).values
        if self.shards is not None:
            mn, mx = self.shards.allreduce_minmax(mn, mx)
        self.cv.out_mean.copy_((mx + mn) / 2)
        rg = (mx - mn) / 2
        self.cv.out_range.copy_(torch.where(rg.abs() < 1e-8, torch.ones_like(rg), rg))
        self.cv_stats = {"min": mn.cpu().numpy(), "max": mx.cpu().numpy()}
        return (P - self.cv.out_mean) / self.cv.out_range

    def project_data(self, data: torch.Tensor, normalize_data: bool = True) -> torch.Tensor:
        if self.cv is None:
            raise ValueError("No model to project with. Train or load the CV first.")
        dev = next(self.cv.buffers()).device
        return self._forward(data.to(dev, torch.float32))

    def get_cv_parameters(self) -> Dict:
        return {"cv_name": self.cv_name, "cv_dimension": self.cv_dimension,
                "features_norm_mode": self.feats_norm_mode, "weights_path": getattr(self, "weights_path", None)}

    def save_model(self):
        """model.zip: metadata, feature labels and the TorchScript module ``cv_weights.pt``
        (reference :1773-1806; loader :1147-1152)."""
        super().save_model()
        f = self.model_output_folder
        self.weights_path = os.path.join(f, "cv_weights.pt")
        example = self.training_data[:2].detach()
        torch.jit.trace(self.cv, example).save(self.weights_path)
        model_path = os.path.join(self.output_path, "model.zip")
        zip_files(model_path, str(f))
        shutil.rmtree(f)
        logger.info(f"Model saved to {model_path}")

    def _load_from_folder(self, folder_path: str):
        super()._load_from_folder(folder_path)
        self.cv = torch.jit.load(os.path.join(self.model_output_folder, "cv_weights.pt"),
                                 map_location=_device(self.configuration))


cv_calculators_map = {
    "pca": PCACalculator,
    "tica": TICACalculator,
    "htica": HTICACalculator,
    "deep_tica": DeepTICACalculator,
}

cv_names_map = {"pca": "PCA", "tica": "TICA", "htica": "HTICA", "deep_tica": "DeepTICA"}

cv_components_map = {"pca": "PC", "tica": "TIC", "htica": "HTIC", "deep_tica": "DeepTIC"}

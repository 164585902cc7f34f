"""train_colvars step API on the B200 backend.

Same signature, return value and on-disk layout as the reference's
``tools/train_colvars/train_colvars.py:20-155`` / ``train_colvars_workflow.py:24-411`` for the
CVs of the hot path (pca, tica, htica, deep_tica).  The free-energy surfaces of the projection
(reference train_colvars_workflow.py:146-182) are computed on the device and saved as ``fes*.npy``
(no figure is drawn: matplotlib is not part of this image); plots, PLUMED input
export and sensitivity analysis are outside the hot path.
"""
from __future__ import annotations

import logging
import os
import sys
import time
from pathlib import Path
from typing import Dict, List, Optional, Union

from ...modules.common import files_exist, merge_configurations, validate_configuration
from ...modules.cv_learning.cv_calculator import cv_calculators_map, cv_names_map
from ...modules.figures import figures
from ...yaml_schemas.train_colvars import TrainColvarsSchema

logger = logging.getLogger(__name__)


class TrainColvarsWorkflow:
    """Reference ``TrainColvarsWorkflow`` (train_colvars_workflow.py:24-411), hot-path subset."""

    def __init__(self, configuration: Dict, train_colvars_paths: Union[str, List[str]],
                 train_topology_paths: Optional[List[str]] = None,
                 trajectory_names: Optional[List[str]] = None,
                 ref_topology_path: Optional[str] = None,
                 features_list: Optional[List[str]] = None,
                 cv_dimension: Optional[int] = None, cvs: Optional[List[str]] = None,
                 frames_per_sample: Optional[int] = 1, output_folder: str = "train_colvars"):
        if isinstance(train_colvars_paths, str):
            train_colvars_paths = [train_colvars_paths]
        self.output_folder = output_folder
        os.makedirs(output_folder, exist_ok=True)
        self.configuration = validate_configuration(configuration, TrainColvarsSchema, output_folder)
        self.train_colvars_paths = train_colvars_paths
        self.train_topology_paths = train_topology_paths
        self.ref_topology_path = ref_topology_path
        self.features_list = features_list
        self.cv_dimension = cv_dimension
        self.frames_per_sample = frames_per_sample
        self.cvs_list = cvs if cvs is not None else self.configuration["cvs"]
        self.trajectory_names = trajectory_names or [Path(p).stem for p in train_colvars_paths]
        for p in train_colvars_paths:          # reference :104-121
            if not os.path.isfile(p):
                logger.error(f"Colvars file not found: {p}")
                sys.exit(1)
        if train_topology_paths and len(train_topology_paths) != len(train_colvars_paths):
            logger.error("The number of topology files must match the number of colvars files.")
            sys.exit(1)

    def get_output_cv_model_path(self, cv_name: str) -> str:
        return os.path.join(self.output_folder, cv_name, "model.zip")

    def get_output_cv_trajectories(self, cv_name: str) -> List[str]:
        base = os.path.join(self.output_folder, cv_name, "traj_data")
        return [os.path.join(base, name, "projected_trajectory.csv") for name in self.trajectory_names]

    def get_output_paths(self) -> Dict:
        return {cv: {"output_folder": os.path.join(self.output_folder, cv),
                     "model_path": self.get_output_cv_model_path(cv),
                     "traj_paths": self.get_output_cv_trajectories(cv)} for cv in self.cvs_list}

    def workflow_finished(self) -> bool:
        return all(files_exist(self.get_output_cv_model_path(cv)) and
                   files_exist(*self.get_output_cv_trajectories(cv)) for cv in self.cvs_list)

    def create_fes_plots(self, data, cv_name: str, cv_labels: List[str], output_folder: str):
        """Reference ``create_fes_plots`` (train_colvars_workflow.py:146-182): one 1-D FES per CV
        component with 100 blocks, one 2-D FES per pair of components with 1 block."""
        import numpy as np
        settings = (self.configuration.get("figures") or {}).get("fes") or {}
        cv_type = cv_names_map.get(cv_name, cv_name)
        X = data.to_numpy(dtype=np.float32)
        d = X.shape[1]
        for i in range(d):
            figures.plot_fes(X[:, i], [cv_labels[i]], settings,
                             os.path.join(output_folder, f"fes_{cv_type}_{i + 1}"), num_blocks=100)
        for i in range(d - 1):
            for j in range(i + 1, d):
                figures.plot_fes(X[:, [i, j]], [cv_labels[i], cv_labels[j]], settings,
                                 os.path.join(output_folder, f"fes_{cv_type}_{i + 1}_{j + 1}"), num_blocks=1)

    def run(self) -> Dict:
        if self.workflow_finished():
            logger.info("Skipping collective variable computation.")
            logger.info("All collective variables have already been computed.")
            return self.get_output_paths()
        logger.info(f"Collective variables to compute: {self.cvs_list}")
        for cv_name in self.cvs_list:
            if cv_name not in cv_calculators_map:
                logger.warning(f"{cv_name} is outside the B200 hot path (pca, tica, htica). Skipping this CV.")
                continue
            cv_output_folder = os.path.join(self.output_folder, cv_name)
            merged = merge_configurations(self.configuration["common"], self.configuration.get(cv_name, {}))
            calc = cv_calculators_map[cv_name](configuration=merged, output_path=self.output_folder)
            calc.load_training_data(train_colvars_paths=self.train_colvars_paths,
                                    train_topology_paths=self.train_topology_paths,
                                    ref_topology_path=self.ref_topology_path,
                                    features_list=self.features_list)
            projected = calc.run(self.cv_dimension)
            self.cv_dimension = calc.get_cv_dimension()
            if projected is None:
                logger.warning(f"Projected colvars dataframe is empty for {cv_name}. Skipping this CV.")
                continue
            projected["traj_label"] = calc.training_data_labels
            for traj_index, traj_name in enumerate(self.trajectory_names):
                out = os.path.join(cv_output_folder, "traj_data", traj_name)
                os.makedirs(out, exist_ok=True)
                df_i = projected[projected["traj_label"] == traj_index].drop("traj_label", axis=1)
                df_i.to_csv(os.path.join(out, "projected_trajectory.csv"), index=False, float_format="%.4f")
                if ((self.configuration.get("figures") or {}).get("fes") or {}).get("compute", False):
                    self.create_fes_plots(df_i, cv_name, calc.get_labels(), out)
        return self.get_output_paths()


def train_colvars(configuration: Dict, train_colvars_paths: Union[str, List[str]],
                  train_topologies: Optional[List[str]] = None,
                  trajectory_names: Optional[List[str]] = None,
                  val_colvars_paths: Optional[Union[str, List[str]]] = None,
                  val_topologies: Optional[List[str]] = None,
                  sup_topologies: Optional[List[str]] = None,
                  sup_traj_names: Optional[List[str]] = None,
                  waypoint_structures: Optional[List[str]] = None,
                  reference_topology: Optional[str] = None,
                  features_list: Optional[List[str]] = None,
                  dimension: Optional[int] = None, cvs: Optional[List[str]] = None,
                  frames_per_sample: Optional[int] = 1,
                  output_folder: str = "train_colvars") -> Dict[str, Dict]:
    """Train / compute the collective variables and project the training data.

    Returns ``{cv: {'output_folder', 'model_path', 'traj_paths'}}`` exactly as the reference.
    ``val_*``, ``sup_*`` and ``waypoint_structures`` only feed non-linear CVs / PLUMED export in
    the reference and are accepted for signature compatibility."""
    start = time.time()
    os.makedirs(output_folder, exist_ok=True)
    wf = TrainColvarsWorkflow(configuration=configuration, train_colvars_paths=train_colvars_paths,
                              train_topology_paths=train_topologies, trajectory_names=trajectory_names,
                              ref_topology_path=reference_topology, features_list=features_list,
                              cv_dimension=dimension, cvs=cvs, frames_per_sample=frames_per_sample,
                              output_folder=output_folder)
    out = wf.run()
    logger.info("Elapsed time (Train colvars): %s", time.strftime("%H h %M min %S s", time.gmtime(time.time() - start)))
    return out

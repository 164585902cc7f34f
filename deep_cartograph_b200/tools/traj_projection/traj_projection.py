"""traj_projection step API on the B200 backend: load ``model.zip``, project colvars, write CSV.

Reference: ``tools/traj_projection/traj_projection.py:19`` / ``traj_projection_workflow.py:199-298``
-> ``CVCalculator.load`` + ``project_colvars``.  The pure "projection of every frame" pass."""
from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import Dict, List, Optional

from ...modules.cv_learning.cv_calculator import CVCalculator

logger = logging.getLogger(__name__)


def traj_projection(configuration: Optional[Dict], colvars_paths: List[str],
                    trajectories: Optional[List[str]] = None, topologies: Optional[List[str]] = None,
                    model_paths: Optional[List[str]] = None, output_folder: str = "traj_projection",
                    **_ignored) -> Dict[str, Dict]:
    """Returns ``{cv_name: {'model_path', 'traj_paths': [csv per colvars file]}}``."""
    os.makedirs(output_folder, exist_ok=True)
    out: Dict[str, Dict] = {}
    for model_path in model_paths or []:
        calc = CVCalculator.load(model_path, output_folder)
        cv_name = calc.cv_name
        traj_paths = []
        for i, colvars in enumerate(colvars_paths):
            name = Path(trajectories[i]).stem if trajectories else Path(colvars).stem
            folder = os.path.join(output_folder, cv_name, name)
            os.makedirs(folder, exist_ok=True)
            csv = os.path.join(folder, "projected_trajectory.csv")
            if calc.ref_topology_path is None:
                logger.warning("Reference topology not set. Make sure the colvars file matches the training data.")
            df = calc.project_colvars(colvars, topologies[i] if topologies else None)
            df.to_csv(csv, index=False, float_format="%.4f")
            traj_paths.append(csv)
        out[cv_name] = {"model_path": model_path, "traj_paths": traj_paths}
    return out

"""Configuration contract of the traj_cluster step (reference ``yaml_schemas/traj_cluster.py:18-47``).

The reference default algorithm is ``hierarchical``; this backend accelerates ``kmeans`` only and
raises for the others (HDBSCAN / agglomerative are out of scope)."""
from typing import Dict, List, Literal, Optional, Union

from pydantic import BaseModel, ConfigDict


class TrajClusterSchema(BaseModel):
    model_config = ConfigDict(extra="allow")
    run: bool = True
    output_structures: Optional[Literal["centroids", "all"]] = "centroids"
    algorithm: Literal["kmeans", "hdbscan", "hierarchical"] = "hierarchical"
    opt_num_clusters: bool = True
    search_interval: List[int] = [3, 10]
    num_clusters: int = 10
    linkage: str = "complete"
    n_init: int = 20
    min_cluster_size: int = 5
    max_cluster_size: Union[int, None] = None
    min_samples: int = 3
    cluster_selection_epsilon: float = 0
    cluster_selection_method: Literal["eom", "leaf"] = "eom"
    figures: Dict = {}

from .traj_cluster import traj_cluster, TrajClusterWorkflow  # noqa: F401

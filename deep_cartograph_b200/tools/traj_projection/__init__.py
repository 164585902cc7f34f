from .traj_projection import traj_projection  # noqa: F401

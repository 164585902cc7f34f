from .train_colvars import train_colvars  # noqa: F401
from .traj_cluster import traj_cluster  # noqa: F401
from .traj_projection import traj_projection  # noqa: F401

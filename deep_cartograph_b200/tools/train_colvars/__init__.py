from .train_colvars import train_colvars, TrainColvarsWorkflow  # noqa: F401

from . import figures  # noqa: F401

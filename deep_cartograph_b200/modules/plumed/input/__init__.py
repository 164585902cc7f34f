from . import assembler  # noqa: F401

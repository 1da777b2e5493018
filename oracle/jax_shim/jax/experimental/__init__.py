from . import loops  # noqa: F401

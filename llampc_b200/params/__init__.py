from .orca import ORCA  # noqa: F401

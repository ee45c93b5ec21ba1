from .dynamic import Dynamic  # noqa: F401

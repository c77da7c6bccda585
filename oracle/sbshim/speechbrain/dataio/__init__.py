from . import dataio  # noqa: F401

from . import natsorted  # noqa: F401

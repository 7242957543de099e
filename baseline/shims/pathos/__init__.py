"""Stand-in for pathos: pathos.multiprocessing.Pool is multiprocess.Pool, the class pathos re-exports."""
from . import multiprocessing  # noqa: F401

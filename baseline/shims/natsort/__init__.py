"""Stand-in for natsort: natsorted with numeric-aware keys (see ../README.md)."""
import re


def _natural_key(x):
    if isinstance(x, (tuple, list)):
        return tuple(_natural_key(v) for v in x)
    if isinstance(x, (int, float)):
        return (0, x, "")
    parts = re.split(r"(\d+)", str(x))
    return tuple((0, int(p), "") if p.isdigit() else (1, 0, p) for p in parts)


def natsorted(seq, key=None, reverse=False, alg=None):
    if key is None:
        return sorted(seq, key=_natural_key, reverse=reverse)
    return sorted(seq, key=lambda v: _natural_key(key(v)), reverse=reverse)


from . import natsort  # noqa: E402,F401  (`from natsort import natsort; natsort.natsorted(...)`, INDEL_Processing.py:17)

"""Stand-in for python-Levenshtein: distance(a, b) = unit-cost edit distance (see ../README.md)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_lev.so")


def build(force=False):
    src = os.path.join(_HERE, "_lev.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


def _load():
    try:
        lib = ctypes.CDLL(build() if not os.path.exists(_SO) else _SO)
        lib.lev_distance.restype = ctypes.c_int
        lib.lev_distance.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        return lib
    except Exception:
        return None


_lib = _load()
compiled = _lib is not None


def _py_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def distance(a, b):
    if _lib is None:
        return _py_distance(a, b)
    x = a.encode("utf-8") if isinstance(a, str) else bytes(a)
    y = b.encode("utf-8") if isinstance(b, str) else bytes(b)
    if len(x) != len(a) or len(y) != len(b):        # non-ASCII text: count code points, not bytes
        return _py_distance(a, b)
    return _lib.lev_distance(x, len(x), y, len(y))

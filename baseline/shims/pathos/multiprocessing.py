import os


class _SerialPool:
    """In-process pool with the two methods the reference uses (fixture generation: same results, no pickling)."""

    def __init__(self, *a, **k):
        pass

    def starmap(self, fn, it):
        return [fn(*args) for args in it]

    def map(self, fn, it):
        return [fn(a) for a in it]


def Pool(*args, **kwargs):
    if os.environ.get("SCAR_SHIM_SERIAL_POOL") == "1":
        return _SerialPool(*args, **kwargs)
    import multiprocess
    return multiprocess.Pool(*args, **kwargs)

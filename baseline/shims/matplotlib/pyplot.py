from unittest import mock as _mock

_m = _mock.MagicMock()


def __getattr__(name):
    return getattr(_m, name)

from unittest import mock as _mock

PdfPages = _mock.MagicMock()

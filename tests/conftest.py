import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture
def fake_kernels(monkeypatch):
    """Swap the C-ABI kernel wrappers for the torch stand-ins in tests/fake_kernels.py (CPU host-logic tests)."""
    from sin_inn_b200 import engine
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fake_kernels as FK
    monkeypatch.setattr(engine, "K", FK)
    monkeypatch.setattr(engine, "require_cuda", lambda t, what="tensor": None)
    import contextlib
    monkeypatch.setattr(engine, "_device_ctx", lambda t: contextlib.nullcontext())
    engine._pack_cache.clear()
    yield FK
    engine._pack_cache.clear()

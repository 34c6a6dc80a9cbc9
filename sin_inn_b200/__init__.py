"""sin_inn_b200 -- B200-native (sm_100a) implementation of the invertible-network hot path of
paramhanji/sin-inn: forward / inverse / backward of the archs.py INNs behind the reference's own
Python API.  (The directory is named with an underscore so it is importable; the project name is
sin-inn_b200.)

    from sin_inn_b200 import archs            # drop-in for the reference's archs.py
    from sin_inn_b200.freia import framework as Ff, modules as Fm   # FrEIA-protocol operators

All compute goes through libsininn.so (C ABI in include/sininn.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .engine import EngineConfig, default_config  # noqa: F401

__all__ = ["archs", "freia", "engine", "kernels", "train", "EngineConfig", "default_config"]

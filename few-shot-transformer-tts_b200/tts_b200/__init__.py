"""B200-native Transformer-TTS mel path: host side (ctypes over libtts_b200.so)."""
from . import _native  # noqa: F401

__all__ = ["_native", "ops", "engine"]

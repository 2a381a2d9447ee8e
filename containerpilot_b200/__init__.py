"""cpbus — a B200-native event bus behind ContainerPilot's `events` API.

Package layout (only what the hot path needs):
  csrc/cpbus_kernels.cuh   sm_100a kernels (fan-out, admission, digest fold)
  csrc/cpbus.cu            C-ABI implementation (include/cpbus.h) -> libcpbus.so
  _native.py               ctypes binding of the C-ABI (no fallback)
  bus.py                   numpy-friendly `Bus` wrapper, 1:1 with cpbus_*
  events.py                mirror of the Go `events` package API
"""
from . import _native  # noqa: F401

__all__ = ["_native", "bus", "events"]

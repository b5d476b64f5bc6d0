"""B200-native dense recall for the Photo Search Engine (flat exact search behind
``utils/vector_store.py::VectorStore``).  See DESIGN.md / INTEGRATION.md.

Importing the package is cheap; the native library is loaded when ``VectorStore`` (or
anything from ``_native``) is first used and its absence is an ``ImportError`` -- there is no
CPU fallback.
"""
from __future__ import annotations

__version__ = "0.1.0"
__all__ = ["VectorStore", "build_native"]


def __getattr__(name: str):
    if name == "VectorStore":
        from .vector_store import VectorStore

        return VectorStore
    if name == "build_native":
        from .build import build_native

        return build_native
    raise AttributeError(name)

"""Mirror of pydiskann/pq/__init__.py:8-22 (re-exports)."""
from .fast_pq import DiskANNPQ, FastPQ

__all__ = ["DiskANNPQ", "FastPQ"]

"""pangnn_b200 — B200-native (sm_100a) implementation of panGNN's message-passing hot path.

Layout: ``csrc/`` CUDA kernels + C ABI (``include/pangnn_b200.h``), ``_abi.py`` ctypes binding,
``ops.py`` tensor wrappers + autograd, ``gnn.py`` / ``dataset.py`` / ``preprocessing.py`` /
``simulate.py`` / ``setup.py`` / ``data.py`` host-side mirror of the reference's module API.
"""
from . import _abi  # noqa: F401

__all__ = ["_abi"]

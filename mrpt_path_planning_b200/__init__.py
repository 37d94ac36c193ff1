"""mrpt_path_planning_b200 — B200-native hot path of MRPT/mrpt_path_planning.

The product is ``lib/libmpp_b200.so`` (hand-written CUDA for sm_100a + the C++ host mirror of
the mpp:: planner API, behind the C ABI of ``include/mppb.h``). This package only loads it.
There is no CPU fallback: using the library without a CUDA device raises ``MppbError``.
"""
from .capi import Context, MppbError, PTG, lib, lib_path  # noqa: F401

__all__ = ["Context", "MppbError", "PTG", "lib", "lib_path"]

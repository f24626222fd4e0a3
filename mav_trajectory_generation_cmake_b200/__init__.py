"""B200-native batched minimum-snap trajectory solver (drop-in for the hot path of
magrimm/mav_trajectory_generation_cmake).

The product is the C-ABI shared library ``lib/libminsnap_b200.so`` (CUDA kernels for sm_100a,
declared in ``include/minsnap_b200.h``) and the C++ mirror of the reference's class API in
``include/mav_trajectory_generation/``.  This Python package is plumbing: it builds the
library, binds it with ctypes and passes torch device pointers / numpy host buffers through.
"""
from . import api, capi  # noqa: F401
from .api import *  # noqa: F401,F403
from .capi import MinsnapError, load  # noqa: F401

__all__ = [n for n in dir(api) if not n.startswith("_")] + ["MinsnapError", "load", "api", "capi"]

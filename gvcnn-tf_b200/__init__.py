"""gvcnn-tf_b200: B200-native (sm_100a) implementation of GVCNN's view-grouping + fusion hot path.

The directory name carries the reference's hyphen, so import it through the
``gvcnn_tf_b200`` shim package at the repo root::

    from gvcnn_tf_b200 import model          # drop-in for the reference's nets/model.py (this path only)

Everything computes in ``libgvcnn_sm100.so`` (``csrc/``, C ABI in ``include/gvcnn_b200.h``);
there is no CPU or pure-PyTorch fallback.
"""
from . import _cabi, model  # noqa: F401
from . import parallel, records  # noqa: F401

__version__ = "0.1.0"

"""npdr_b200 -- B200-native engine for NPDR's hot path behind the reference's operator API.

Host-side mirror of the reference's exported R functions (insilico/npdr, NAMESPACE:11-24,29-30)
over the C ABI of libnpdr_b200.so.  No CPU fallback: compute calls raise if the CUDA library
or an sm_100 device is missing.
"""
from ._lib import NpdrB200Error, Context, Session, context  # noqa: F401
from .api import (npdr, npdrDiff, npdrDistances, nearestNeighbors, nearestNeighborsSeparateHitMiss,  # noqa: F401
                  uniqueNeighbors, knnVec, knnSURF, diffRegression, uniReg, npdr_k_grid, vwok,
                  npdrDistances2, nearestNeighbors2, npdrLearner, npdr_stats)
from . import rstats  # noqa: F401

__version__ = "0.1.0"

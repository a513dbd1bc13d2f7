"""
Index plumbing of the EKF (reference: slam_system/util.py:75-96).  get_overlap_index is host logic in the
reference too (a two-pointer merge over two short ascending index arrays); it is vectorised here, not ported.
"""
import numpy as np


def get_overlap_index(index1, index2):
    """Positions (in index1, in index2) of the values both ascending arrays share (util.py:75-96)."""
    index1 = np.asarray(index1)
    index2 = np.asarray(index2)
    _, i1, i2 = np.intersect1d(index1, index2, assume_unique=True, return_indices=True)
    return i1.astype(np.int64), i2.astype(np.int64)

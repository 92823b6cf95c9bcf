"""Small rigid-body helpers (reference: python/gym_ignition/rbd/utils.py:8-92)."""
import numpy as np


def wedge(vector3: np.ndarray) -> np.ndarray:
    """Skew-symmetric matrix of a 3-vector."""
    x, y, z = np.asarray(vector3, float).reshape(3)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def vee(matrix3x3: np.ndarray) -> np.ndarray:
    """Inverse of :py:func:`wedge` (uses the skew-symmetric part)."""
    m = np.asarray(matrix3x3, float)
    if m.shape != (3, 3):
        raise ValueError(m.shape)
    s = (m - m.T) / 2
    return np.array([s[2, 1], s[0, 2], s[1, 0]])


def extract_skew(matrix: np.ndarray) -> np.ndarray:
    m = np.asarray(matrix, float)
    if m.shape != (3, 3):
        raise ValueError(m.shape)
    return 0.5 * (m - m.T)


def extract_symm(matrix: np.ndarray) -> np.ndarray:
    m = np.asarray(matrix, float)
    if m.shape != (3, 3):
        raise ValueError(m.shape)
    return 0.5 * (m + m.T)

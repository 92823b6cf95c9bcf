"""Rotation / transform conversions (reference: python/gym_ignition/rbd/conversions.py:7-94). wxyz quaternions."""
import numpy as np


class Quaternion:
    @staticmethod
    def to_wxyz(xyzw: np.ndarray) -> np.ndarray:
        q = np.asarray(xyzw, float)
        if q.shape != (4,):
            raise ValueError(q)
        return q[[3, 0, 1, 2]]

    @staticmethod
    def to_xyzw(wxyz: np.ndarray) -> np.ndarray:
        q = np.asarray(wxyz, float)
        if q.shape != (4,):
            raise ValueError(q)
        return q[[1, 2, 3, 0]]

    @staticmethod
    def to_rotation(quaternion: np.ndarray) -> np.ndarray:
        w, x, y, z = np.asarray(quaternion, float) / np.linalg.norm(quaternion)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    @staticmethod
    def from_matrix(matrix: np.ndarray) -> np.ndarray:
        R = np.asarray(matrix, float)
        if R.shape != (3, 3):
            raise ValueError(R)
        tr = np.trace(R)
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
        else:
            i = int(np.argmax(np.diag(R)))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
            q = np.zeros(4)
            q[0] = (R[k, j] - R[j, k]) / s
            q[1 + i] = 0.25 * s
            q[1 + j] = (R[j, i] + R[i, j]) / s
            q[1 + k] = (R[k, i] + R[i, k]) / s
        return q if q[0] >= 0 else -q


class Transform:
    @staticmethod
    def from_position_and_quaternion(position: np.ndarray, quaternion: np.ndarray) -> np.ndarray:
        if np.asarray(position).size != 3 or np.asarray(quaternion).size != 4:
            raise ValueError("wrong size of position or quaternion")
        H = np.eye(4)
        H[:3, :3] = Quaternion.to_rotation(quaternion)
        H[:3, 3] = np.asarray(position, float)
        return H

    @staticmethod
    def from_position_and_rotation(position: np.ndarray, rotation: np.ndarray) -> np.ndarray:
        if np.asarray(position).size != 3 or np.asarray(rotation).shape != (3, 3):
            raise ValueError("wrong size of position or rotation")
        H = np.eye(4)
        H[:3, :3] = rotation
        H[:3, 3] = np.asarray(position, float)
        return H

    @staticmethod
    def to_position_and_rotation(transform: np.ndarray):
        H = np.asarray(transform, float)
        if H.shape != (4, 4):
            raise ValueError(H.shape)
        return H[:3, 3].copy(), H[:3, :3].copy()

    @staticmethod
    def to_position_and_quaternion(transform: np.ndarray):
        p, R = Transform.to_position_and_rotation(transform)
        return p, Quaternion.from_matrix(R)

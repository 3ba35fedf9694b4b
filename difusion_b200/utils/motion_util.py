"""SE(3) poses with the interface the map/tracker API passes around (reference utils/motion_util.py:162-339):
``Isometry(q, t)`` with ``.q.rotation_matrix``, ``.t``, ``.dot``, ``.inv``, ``@``, ``.rotation``, ``from_twist``.
The reference builds on pyquaternion (not available here); ``Rotation`` below is a small stand-in that stores the
rotation matrix and offers the handful of members the hot path touches.
"""
from __future__ import annotations

import numpy as np


def so3_wedge(phi):
    x, y, z = np.asarray(phi, dtype=float)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def so3_left_jacobian(phi):                                   # reference motion_util.py:45-57
    phi = np.asarray(phi, dtype=float)
    angle = np.linalg.norm(phi)
    if np.isclose(angle, 0.):
        return np.identity(3) + 0.5 * so3_wedge(phi)
    axis = phi / angle
    s, c = np.sin(angle), np.cos(angle)
    return (s / angle) * np.identity(3) + (1 - s / angle) * np.outer(axis, axis) + ((1 - c) / angle) * so3_wedge(axis)


def _orthonormalize(R):
    u, _, vt = np.linalg.svd(R)
    R = u @ vt
    if np.linalg.det(R) < 0:
        u[:, -1] *= -1
        R = u @ vt
    return R


class Rotation:
    """Rotation with the pyquaternion members used by the reference's Isometry."""

    def __init__(self, matrix=None, axis=None, degrees=None, radians=None):
        if matrix is not None:
            self._R = _orthonormalize(np.asarray(matrix, dtype=float)[:3, :3])
        elif axis is not None:
            ang = radians if radians is not None else np.deg2rad(degrees or 0.0)
            a = np.asarray(axis, dtype=float)
            a = a / np.linalg.norm(a)
            K = so3_wedge(a)
            self._R = np.identity(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        else:
            self._R = np.identity(3)

    @property
    def rotation_matrix(self):
        return self._R.copy()

    @property
    def transformation_matrix(self):
        m = np.identity(4)
        m[:3, :3] = self._R
        return m

    @property
    def inverse(self):
        return Rotation(matrix=self._R.T)

    def rotate(self, v):
        return self._R @ np.asarray(v, dtype=float)

    def __mul__(self, other):
        return Rotation(matrix=self._R @ other._R)

    def __repr__(self):
        return f"Rotation({np.array2string(self._R, precision=4)})"


Quaternion = Rotation       # name used by the reference's call sites


class Isometry:
    def __init__(self, q=None, t=None):
        self.q = q if q is not None else Rotation()
        t = np.zeros(3) if t is None else np.asarray(t, dtype=float)
        assert t.shape == (3,)
        self.t = t

    def __repr__(self):
        return f"Isometry: t = {self.t}, q = {self.q}"

    @property
    def rotation(self):
        return Isometry(q=self.q)

    @property
    def matrix(self):
        m = self.q.transformation_matrix
        m[0:3, 3] = self.t
        return m

    @staticmethod
    def from_matrix(mat, t_component=None, ortho=False):
        mat = np.asarray(mat, dtype=float)
        if t_component is None:
            assert mat.shape == (4, 4)
            return Isometry(q=Rotation(matrix=mat[:3, :3]), t=mat[:3, 3])
        assert mat.shape == (3, 3)
        return Isometry(q=Rotation(matrix=mat), t=np.asarray(t_component, dtype=float))

    @staticmethod
    def from_so3_exp(phi):                                    # reference :213-229
        phi = np.asarray(phi, dtype=float)
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.):
            return Isometry(q=Rotation(matrix=np.identity(3) + so3_wedge(phi)))
        return Isometry(q=Rotation(axis=phi / angle, radians=angle))

    @staticmethod
    def from_twist(xi):                                       # reference :205-210: xi = [rho, phi], t = J_l(phi) rho
        xi = np.asarray(xi, dtype=float)
        iso = Isometry.from_so3_exp(xi[3:6])
        iso.t = so3_left_jacobian(xi[3:6]) @ xi[:3]
        return iso

    def inv(self):                                            # :273-275
        qinv = self.q.inverse
        return Isometry(q=qinv, t=-(qinv.rotate(self.t)))

    def dot(self, right):                                     # :277-278
        return Isometry(q=(self.q * right.q), t=(self.q.rotate(right.t) + self.t))

    def torch_matrices(self, device):
        import torch
        return torch.from_numpy(self.q.rotation_matrix).to(device).float(), torch.from_numpy(self.t).to(device).float()

    def __matmul__(self, other):                              # :322-333
        if type(other).__module__.startswith("torch"):       # (numpy >= 2 arrays also have .device)
            assert other.ndim == 2 and other.size(1) == 3
            th_R, th_t = self.torch_matrices(other.device)
            return other @ th_R.t() + th_t.unsqueeze(0)
        if isinstance(other, Isometry):
            return self.dot(other)
        other = np.asarray(other)
        if other.ndim == 1:
            return self.q.rotate(other) + self.t
        return other @ self.q.rotation_matrix.T + self.t[np.newaxis, :]
